"""CPU-side checks of the C-ABI library and the host mirror (no GPU needed):
the .so loads, exports every symbol include/nerfactor_b200.h declares, refuses to
create a context without a GPU, and packs weights on the host."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200.networks import mlp
from nerfactor_b200.networks.embedder import Embedder
from nerfactor_b200.brdf.renderer import gen_light_xyz

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'nerfactor_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nf_[A-Za-z0-9_]+)\s*\(', src)))


def test_library_exports_every_header_symbol():
    lib = _lib.load_library()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), 'missing export %s' % s
    assert sorted(_lib.EXPORTS) == syms
    assert lib.nf_version() == 100


def test_no_gpu_means_no_context():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    lib = _lib.load_library()
    h = C.c_void_p()
    assert lib.nf_ctx_create(C.byref(h), -1) == -4          # NF_ERR_NO_DEVICE
    with pytest.raises(_lib.NfError):
        _lib.Context()


def _desc(kind, layers, skip_at, out_act, **kw):
    d = _lib.MlpDesc()
    Ws = [np.ascontiguousarray(w) for w, _ in layers]
    bs = [np.ascontiguousarray(b) for _, b in layers]
    depth = len(layers) - 1
    d.kind, d.in_dim, d.width = _lib.KIND[kind], Ws[0].shape[0], Ws[0].shape[1]
    d.depth, d.skip_at, d.out_dim = depth, skip_at, Ws[-1].shape[1]
    d.out_act = _lib.ACT[out_act]
    d.n_freqs_a, d.n_freqs_b, d.z_dim = kw.get('a', 0), kw.get('b', 0), kw.get('z', 0)
    d.W = (C.c_void_p * (depth + 1))(*[w.ctypes.data for w in Ws])
    d.b = (C.c_void_p * (depth + 1))(*[b.ctypes.data for b in bs])
    return d, (Ws, bs)


def test_host_weight_packing_sizes_and_errors():
    lib = _lib.load_library()
    p = synth.make_stage_b_params(0, 'learned')
    lv = p['lvis_mlp']['layers'] + p['lvis_out']['layers']
    d, keep = _desc('lvis', lv, 2, 'sigmoid', a=10, b=4)
    h = C.c_void_p()
    assert lib.nf_mlp_create(None, C.byref(d), C.byref(h)) == 0
    nbytes = lib.nf_mlp_device_bytes(h)
    fp32 = sum(w.size + b.size for w, b in lv) * 4
    tc = 2 * (32 + 384 + 32 + 32) * 128 * 2 + (644 + 2 * 64 * 128) * 4     # + bias blocks (K2 v2)
    assert fp32 + tc <= nbytes <= fp32 + tc + 16 * 256
    lib.nf_mlp_destroy(h)
    # wrong embedding spec -> invalid argument, no handle
    d2, keep2 = _desc('lvis', lv, 2, 'sigmoid', a=9, b=4)
    assert lib.nf_mlp_create(None, C.byref(d2), C.byref(h)) == -1
    # sigma net 8x256 packs too (fp32 image only for now)
    nerf = synth.make_nerf_params(0)
    sg = nerf['coarse_enc']['layers'] + nerf['coarse_sigma_out']['layers']
    d3, keep3 = _desc('sigma', sg, 4, None, a=10)
    assert lib.nf_mlp_create(None, C.byref(d3), C.byref(h)) == 0
    assert lib.nf_mlp_device_bytes(h) >= sum(w.size + b.size for w, b in sg) * 4
    lib.nf_mlp_destroy(h)


def test_workspace_queries_are_host_only_and_consistent():
    """`*_workspace_bytes()` of the composite ops (one query per op that needs scratch):
    callable without a GPU, sized by the chunk the op processes at a time."""
    lib = _lib.load_library()
    dn = lib.nf_raymarch_depth_normal_workspace_bytes
    # one chunk of 32768 rays: 3 [c, Sc] + 2 [c, S] + [c, S, 3] fp32 buffers
    full = 32768 * (3 * 128 + 5 * 320) * 4
    assert dn(640000, 128, 192) == dn(32768, 128, 192) and full <= dn(32768, 128, 192) <= full + 6 * 256
    assert dn(100, 128, 192) < dn(32768, 128, 192)
    lv = lib.nf_raymarch_lvis_workspace_bytes
    assert lv(640000, 512, 128, 192) == lv(1024, 512, 128, 192)          # capped at 2^19 pairs
    assert lv(10, 16, 128, 192) < lv(1024, 512, 128, 192)
    a = _lib.StageBArgs()
    a.n, a.n_lights, a.n_envmaps, a.brdf_kind = 640000, 512, 1, 0
    one = lib.nf_stageB_fused_workspace_bytes(C.byref(a), _lib.PREC['f16'])
    assert 0 < one <= 48 << 20                                           # one chunk of lvis rows
    a.brdf_kind = 1
    assert lib.nf_stageB_fused_workspace_bytes(C.byref(a), _lib.PREC['f16']) == 2 * one
    a.lvis_d = 1                                                         # caller keeps lvis: no scratch for it
    assert lib.nf_stageB_fused_workspace_bytes(C.byref(a), _lib.PREC['f16']) == one
    assert lib.nf_dense_fwd_workspace_bytes(128, 128, 0, _lib.PREC['bf16']) > 0
    # whole-network train calls: 16-bit activations of every hidden layer + backward scratch
    def chain(in_dim, widths, skip):
        ch = _lib.MlpChain()
        ch.depth, ch.in_dim, ch.skip_layer = len(widths), in_dim, skip
        for l, w in enumerate(widths):
            ch.width[l] = w
        return ch
    rows = 524288
    wsb = lib.nf_mlp_chain_workspace_bytes
    full = wsb(C.byref(chain(92, (128, 128, 128, 128, 4), 3)), rows)
    saved = rows * 2 * (96 + 3 * 128 + (128 + 96))               # x0 + four hidden outputs (one with [h | x])
    assert saved < full < saved + rows * (2 * 256 * 2 + 16 * 2 + 92 * 4) + (64 << 20)
    assert wsb(C.byref(chain(92, (128, 128, 128, 128, 4), 0)), rows) < full          # no skip: narrower buffer
    assert wsb(C.byref(chain(90, (128, 128, 4), 0)), rows) == 0                      # in_dim not a multiple of 4
    assert wsb(C.byref(chain(92, (100, 128, 4), 0)), rows) == 0                      # hidden width not a multiple of 16
    assert wsb(C.byref(chain(92, (128, 128, 3), 0)), rows) == 0                      # head not padded to 4
    assert wsb(C.byref(chain(92, (128,), 0)), rows) == 0                             # needs >= 2 layers
    assert _lib.mlp_chain_supported(92, [128, 128, 128, 128, 4], [92, 128, 128, 224, 128])
    assert not _lib.mlp_chain_supported(92, [128, 512, 4], [92, 128, 512])


def test_network_mirror_shapes():
    net = mlp.Network([128] * 4, act=['relu'] * 4, skip_at=[2]).build(90)
    ks = [l.kernel.shape for l in net.layers]
    assert ks == [(90, 128), (128, 128), (128, 128), (218, 128)]     # mlp.py:39-50
    enc = mlp.Network([256] * 8, act=['relu'] * 8, skip_at=[4]).build(63)
    assert enc.layers[5].kernel.shape == (319, 256)                  # nerf.py:53-59
    assert Embedder(in_dims=3, log2_max_freq=9, n_freqs=10).out_dims == 63
    assert Embedder(in_dims=3, log2_max_freq=3, n_freqs=4).out_dims == 27
    with pytest.raises(TypeError):           # no CPU path: a Network runs on CUDA tensors only
        net(np.zeros((1, 90), np.float32))


def test_gen_light_xyz_mirror_matches_pinned_reference(golden_dir):
    ref = np.load(os.path.join(golden_dir, 'ref_pinned.npz'))
    for h, w in ((16, 32), (2, 8), (16, 64)):
        xyz, areas = gen_light_xyz(h, w)
        assert np.array_equal(xyz, ref['lxyz_%dx%d' % (h, w)])
        assert np.array_equal(areas, ref['lareas_%dx%d' % (h, w)])


def test_default_configs_match_reference_ini_values():
    c = nfconfig.default_config('nerfactor')
    assert c.getint('DEFAULT', 'mlp_width') == 128 and c.getint('DEFAULT', 'mlp_skip_at') == 2
    assert c.getfloat('DEFAULT', 'albedo_slope') == 0.77
    m = nfconfig.default_config('nerfactor_microfacet')
    assert m.getfloat('DEFAULT', 'fresnel_f0') == 0.04
    n = nfconfig.default_config('nerf')
    assert n.getint('DEFAULT', 'enc_depth') == 8 and n.getint('DEFAULT', 'mlp_width') == 256


def test_sample_rays_semantics():
    """nerf_shape.py:84-121: train batches come only from alpha > 0.9 pixels, with
    replacement; test batches are every ray in row-major order."""
    import torch
    from nerfactor_b200.datasets.nerf_shape import sample_rays
    h, w, L = 6, 5, 4
    g = torch.Generator().manual_seed(0)
    idx = torch.arange(h * w, dtype=torch.float32)
    maps = [idx.reshape(h, w, 1).expand(h, w, 3).clone() for _ in range(3)]
    alpha = torch.zeros((h, w))
    alpha[1, 2] = 1.0
    alpha[4, 0] = 0.95
    alpha[0, 0] = 0.9                       # not above the threshold
    xyz, nrm = maps[0].clone(), maps[0].clone()
    lvis = idx.reshape(h, w, 1).expand(h, w, L).clone()
    out = sample_rays(maps[0], maps[1], maps[2], alpha, xyz, nrm, lvis, 'train', bs=64, generator=g)
    picked = set(out[0][:, 0].long().tolist())
    assert picked == {1 * w + 2, 4 * w + 0} and out[3].shape == (64, 1) and out[6].shape == (64, L)
    allr = sample_rays(maps[0], maps[1], maps[2], alpha, xyz, nrm, lvis, 'test')
    assert torch.equal(allr[0][:, 0], idx) and allr[6].shape == (h * w, L)
    with pytest.raises(ValueError):
        sample_rays(maps[0], maps[1], maps[2], torch.zeros((h, w)), xyz, nrm, lvis, 'train')


def test_product_never_imports_test_infrastructure():
    """The package has no CPU path: nothing under nerfactor_b200/ may import the oracle, the test
    double or the TensorFlow shim (only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
    legs may)."""
    import ast
    import glob
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'nerfactor_b200')
    banned = ('oracle', 'cpu_backend', 'tensorflow', 'e2e_scenario')
    for path in glob.glob(os.path.join(root, '**', '*.py'), recursive=True):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom) and node.level == 0:
                mods = [node.module or '']
            for m in mods:
                assert m.split('.')[0] not in banned, (path, m)
