"""Training path (BASELINE configs[3] semantics at test size): Dense forward / backward
kernels, AMSGrad kernel and the full train-step gradient against the oracle's autograd."""
import numpy as np
import pytest
import torch

from oracle import stage_b, brdf as obrdf
from nerfactor_b200 import synth, config as nfconfig

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope='module')
def ctx():
    from nerfactor_b200 import _lib
    return _lib.default_context()


@pytest.mark.parametrize('shape', [(1000, 92, 0, 128, 'relu'), (777, 128, 92, 128, 'relu'),
                                   (513, 128, 0, 4, 'sigmoid'), (300, 256, 64, 256, 'relu'),
                                   (130, 20, 0, 128, 'softplus'), (64, 128, 0, 4, None)])
def test_dense_fwd_bwd_vs_torch(ctx, shape):
    from nerfactor_b200 import autodiff as ad
    m, k1, k2, n, act = shape
    g = torch.Generator(device='cpu').manual_seed(m)
    x1 = torch.randn((m, k1), generator=g).cuda().requires_grad_(True)
    x2 = torch.randn((m, k2), generator=g).cuda().requires_grad_(True) if k2 else None
    w = (torch.randn((k1 + k2, n), generator=g) * 0.1).cuda().requires_grad_(True)
    b = (torch.randn((n,), generator=g) * 0.1).cuda().requires_grad_(True)
    dy = torch.randn((m, n), generator=g).cuda()
    y = ad.DenseFn.apply(x1, x2, w, b, act)
    ins = [x1, w, b] + ([x2] if k2 else [])
    grads = torch.autograd.grad(y, ins, dy)
    xcat = (x1 if x2 is None else torch.cat((x1, x2), 1)).detach().double()
    wd = w.detach().double()
    pre = xcat @ wd + b.detach().double()
    f = {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'softplus': torch.nn.functional.softplus,
         None: lambda t: t}[act]
    yr = f(pre)
    assert rel_l2(y.detach().cpu(), yr.cpu()) < 1e-5
    # reference adjoint with act' evaluated on the kernel's own output (a pre-activation within
    # rounding of 0 may flip the ReLU mask between fp32 and fp64; that is not a kernel error)
    yk = y.detach().double()
    dact = {'relu': (yk > 0).double(), 'sigmoid': yk * (1 - yk), 'softplus': 1 - torch.exp(-yk),
            None: torch.ones_like(yk)}[act]
    dz = dy.double() * dact
    dx = dz @ wd.t()
    ref = [dx[:, :k1], xcat.t() @ dz, dz.sum(0)] + ([dx[:, k1:]] if k2 else [])
    for a, r in zip(grads, ref):
        assert rel_l2(a.cpu(), r.cpu()) < 2e-5


@pytest.mark.parametrize('prec', ['bf16', 'f16'])
@pytest.mark.parametrize('shape', [(1000, 92, 0, 128, 'relu'), (777, 128, 92, 128, 'relu'),
                                   (513, 128, 0, 4, 'sigmoid'), (40000, 128, 0, 128, 'relu'),
                                   (130, 20, 0, 128, 'softplus'), (64, 128, 0, 4, None),
                                   (20000, 128, 20, 128, 'relu'), (300, 64, 0, 16, None)])
def test_dense_tcgen05_fwd_bwd(ctx, shape, prec):
    """The tensor-core Dense kernels against fp64 matmuls of the SAME 16-bit-rounded operands
    (products of 16-bit values are exact in fp32, so only the accumulation order differs)."""
    from nerfactor_b200 import autodiff as ad
    m, k1, k2, n, act = shape
    dt = torch.bfloat16 if prec == 'bf16' else torch.float16
    q = lambda t: t.detach().to(dt).double()
    g = torch.Generator(device='cpu').manual_seed(m + 1)
    x1 = torch.randn((m, k1), generator=g).cuda().requires_grad_(True)
    x2 = torch.randn((m, k2), generator=g).cuda().requires_grad_(True) if k2 else None
    w = (torch.randn((k1 + k2, n), generator=g) * 0.1).cuda().requires_grad_(True)
    b = (torch.randn((n,), generator=g) * 0.1).cuda().requires_grad_(True)
    dy = torch.randn((m, n), generator=g).cuda()
    y = ad.DenseFn.apply(x1, x2, w, b, act, prec)
    ins = [x1, w, b] + ([x2] if k2 else [])
    grads = torch.autograd.grad(y, ins, dy)
    xcat = x1 if x2 is None else torch.cat((x1, x2), 1)
    pre = q(xcat) @ q(w) + b.detach().double()
    f = {'relu': torch.relu, 'sigmoid': torch.sigmoid, 'softplus': torch.nn.functional.softplus,
         None: lambda t: t}[act]
    assert rel_l2(y.detach().cpu(), f(pre).cpu()) < 1e-5
    yk = y.detach().double()
    dact = {'relu': (yk > 0).double(), 'sigmoid': yk * (1 - yk), 'softplus': 1 - torch.exp(-yk),
            None: torch.ones_like(yk)}[act]
    dz = dy.double() * dact                                  # fp32 product in the kernel
    dzq = q(dz.float())
    dx = dzq @ q(w).t()
    ref = [dx[:, :k1], q(xcat).t() @ dzq, dz.sum(0)] + ([dx[:, k1:]] if k2 else [])
    for a, r, name in zip(grads, ref, ('dx1', 'dw', 'db', 'dx2')):
        assert rel_l2(a.cpu(), r.cpu()) < 2e-5, name
    # and against the exact fp32 layer: operand rounding only
    y32 = ad.DenseFn.apply(x1, x2, w, b, act, 'fp32')
    assert rel_l2(y.detach().cpu(), y32.detach().cpu()) < (2e-2 if prec == 'bf16' else 3e-3)


def test_amsgrad_kernel_vs_numpy(ctx):
    from nerfactor_b200 import _lib
    rng = np.random.default_rng(0)
    n = 10007
    p = rng.standard_normal(n).astype(np.float32)
    pt = torch.tensor(p).cuda()
    mt, vt, vh = [torch.zeros(n, device='cuda') for _ in range(3)]
    pm, m, v, vhat = p.astype(np.float64), np.zeros(n), np.zeros(n), np.zeros(n)
    b1, b2, eps = 0.9, 0.999, 1e-7
    for step in range(1, 6):
        g = rng.standard_normal(n).astype(np.float32) * (1.0 if step != 3 else 10.0)
        lr = 5e-3 * 0.1 ** ((step - 1) / 500000)
        _lib.adam_amsgrad_step(ctx, pt, torch.tensor(g).cuda(), mt, vt, vh, lr, step)
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g.astype(np.float64) ** 2
        vhat = np.maximum(vhat, v)
        lr_t = lr * np.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        pm = pm - lr_t * m / (np.sqrt(vhat) + eps)
    assert np.allclose(pt.cpu().numpy(), pm, atol=2e-6)


def _models(ctx, brdf, lh=2, lw=8, seed=7, shape_mode='finetune'):
    from importlib import import_module
    name = 'nerfactor_microfacet' if brdf == 'microfacet' else 'nerfactor'
    Model = import_module('nerfactor_b200.models.' + name).Model
    params = synth.make_stage_b_params(seed, brdf, light_hw=(lh, lw))
    cfg = nfconfig.default_config(name, light_h=lh, shape_mode=shape_mode)
    m = Model(cfg, params=params, ctx=ctx, precision='fp32')
    lxyz, lareas = obrdf.gen_light_xyz(lh, lw)
    m.set_lights(lxyz.reshape(-1, 3), lareas.reshape(-1))
    m.light_res = (lh, lw)
    return m, params, (lxyz, lareas)


def _oracle_grads(params, brdf, lights, batch, noise, shape_mode='finetune'):
    tp = {}
    leaves = {}
    for k, v in params.items():
        if k == 'light':
            t = torch.tensor(v, requires_grad=True)
            tp[k] = t
            leaves[('light', 0, 'light')] = t
        else:
            layers = []
            for li, (w, b) in enumerate(v['layers']):
                wt, bt = torch.tensor(w, requires_grad=True), torch.tensor(b, requires_grad=True)
                layers.append((wt, bt))
                leaves[(k, li, 'kernel')], leaves[(k, li, 'bias')] = wt, bt
            tp[k] = dict(v, layers=layers)
    om = stage_b.StageB(tp, {'brdf': brdf, 'shape_mode': shape_mode}, lxyz=lights[0],
                        lareas=lights[1])
    pred, gt, lk = om.call(batch, 'train', xyz_noise=noise)
    # nerfactor_microfacet.ini:71 sets brdf_smooth_weight = 0 (nerfactor.ini: 0.01)
    wts = {'brdf_smooth_weight': 0.} if brdf == 'microfacet' else None
    loss = om.compute_loss(pred, gt, weights=wts, **lk)
    (loss.sum() / loss.shape[0]).backward()
    return loss.detach(), leaves


@pytest.mark.parametrize('brdf', ['microfacet', 'learned'])
def test_train_step_gradient_vs_oracle_autograd(ctx, brdf):
    from nerfactor_b200.trainvali import Trainer
    m, params, lights = _models(ctx, brdf)
    batch = synth.make_stage_b_batch(11, 48, 16)
    nfg = int((batch[5][:, 0] > 0).sum())
    noise = (0.01 * np.random.default_rng(2).standard_normal((nfg, 3))).astype(np.float32)
    tr = Trainer(m)
    loss, grad = tr.loss_and_grad(batch, xyz_noise=noise)
    oloss, leaves = _oracle_grads(params, brdf, lights, batch, noise)
    assert np.allclose(loss.cpu().numpy(), oloss.numpy(), atol=1e-5, rtol=1e-4)
    gv = tr.views(grad)
    checked = 0
    for key, t in leaves.items():
        if key not in gv:
            assert key[0].startswith('brdf_mlp') or key[0].startswith('brdf_out')   # frozen prior
            continue
        ref = t.grad.numpy()
        got = gv[key].cpu().numpy()
        scale = max(np.abs(ref).max(), 1e-8)
        assert np.abs(got - ref).max() / scale < 2e-3, key
        checked += 1
    assert checked >= 17


@pytest.mark.parametrize('brdf', ['microfacet', 'learned'])
def test_train_step_gradient_bf16_tensor_cores(ctx, brdf):
    """configs[3]: the same step with every Dense on the tcgen05 kernels (bf16 operands, fp32
    accumulation / master weights).  The gradient must stay aligned with the fp32 oracle's:
    per-tensor cosine > 0.99 and norm within 10 % (bf16 operand rounding is ~4e-3 per value)."""
    from nerfactor_b200.trainvali import Trainer
    m, params, lights = _models(ctx, brdf)
    batch = synth.make_stage_b_batch(11, 48, 16)
    nfg = int((batch[5][:, 0] > 0).sum())
    noise = (0.01 * np.random.default_rng(2).standard_normal((nfg, 3))).astype(np.float32)
    tr = Trainer(m, precision='bf16')
    loss, grad = tr.loss_and_grad(batch, xyz_noise=noise)
    oloss, leaves = _oracle_grads(params, brdf, lights, batch, noise)
    assert np.allclose(loss.cpu().numpy(), oloss.numpy(), atol=3e-3, rtol=5e-2)
    gv = tr.views(grad)
    cos_all, rel_all = [], []
    for key, t in leaves.items():
        if key not in gv:
            continue
        ref = t.grad.numpy().reshape(-1).astype(np.float64)
        got = gv[key].cpu().numpy().reshape(-1).astype(np.float64)
        if np.linalg.norm(ref) < 1e-7:
            continue
        cos = float(ref @ got / (np.linalg.norm(ref) * np.linalg.norm(got)))
        rel = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
        cos_all.append(cos)
        rel_all.append(rel)
        assert cos > 0.99, (key, cos)
        assert 0.9 < np.linalg.norm(got) / np.linalg.norm(ref) < 1.1, key
    assert len(cos_all) >= 17
    # stated tolerance in relative L2 per tensor (bf16 operands: 8 mantissa bits, fp32 accumulate).
    # Measured on B200: median 3.1e-2 / 3.4e-2, worst 1.07e-1 / 8.1e-2 (microfacet / learned).
    print('bf16 train step (%s): gradient rel-L2 per tensor median %.2e, worst %.2e; worst cosine %.5f'
          % (brdf, float(np.median(rel_all)), max(rel_all), min(cos_all)))
    assert float(np.median(rel_all)) < 5e-2 and max(rel_all) < 1.5e-1


@pytest.mark.parametrize('prec', ['fp32', 'bf16'])
def test_training_reduces_loss_and_syncs_back(ctx, prec):
    from nerfactor_b200.trainvali import Trainer
    m, params, lights = _models(ctx, 'microfacet')
    batch = synth.make_stage_b_batch(5, 64, 16, fg_frac=1.0)
    tr = Trainer(m, precision=prec)
    tr.lr0 = 1e-3
    noise = np.zeros((64, 3), np.float32)
    losses = [float(tr.train_step(batch, xyz_noise=noise)) for _ in range(25)]
    assert losses[-1] < 0.9 * losses[0]
    assert tr.iterations == 25
    before = m.net['albedo_mlp'].layers[0].kernel.copy()
    tr.sync_to_model()
    assert not np.array_equal(before, m.net['albedo_mlp'].layers[0].kernel)
    pred, _, _, _ = m.call(batch, 'test')          # fused kernels pick up the new weights
    assert torch.isfinite(pred['rgb']).all()


@pytest.mark.parametrize('brdf', ['microfacet', 'learned'])
def test_graph_replay_equals_eager_step(ctx, brdf):
    """train_step(graph=True) replays forward + backward as one CUDA graph; parameters after
    three steps must equal the eager path's bit for bit (same kernels, same order)."""
    from nerfactor_b200.trainvali import Trainer
    batch = synth.make_stage_b_batch(5, 64, 16, fg_frac=1.0)
    noise = (0.01 * np.random.default_rng(4).standard_normal((64, 3))).astype(np.float32)
    flats = []
    for graph in (False, True):
        m, _, _ = _models(ctx, brdf)
        tr = Trainer(m, precision='bf16')
        for _ in range(3):
            tr.train_step(batch, xyz_noise=noise, graph=graph)
        if graph:
            assert len(tr._graphs) == 1 and next(iter(tr._graphs.values()))['graph'] is not None
        flats.append(tr.flat.clone())
    assert torch.equal(flats[0], flats[1])
    # a batch with background rays falls back to the eager path
    m, _, _ = _models(ctx, brdf)
    tr = Trainer(m, precision='bf16')
    tr.train_step(synth.make_stage_b_batch(5, 64, 16, fg_frac=0.5))
    assert not tr._graphs


def test_shape_model_train_step(ctx):
    """shape.py pre-training (normal + visibility MLPs, shape.py:239-277)."""
    from nerfactor_b200.models.shape import Model
    from nerfactor_b200.trainvali import Trainer
    params = synth.make_stage_b_params(3, 'learned', light_hw=(2, 8))
    m = Model(nfconfig.default_config('shape', light_h=2), params=params, ctx=ctx, precision='fp32')
    batch = synth.make_stage_b_batch(9, 40, 8)          # light_h = 2 -> 2 x 4 lights
    tr = Trainer(m)
    noise = (0.01 * np.random.default_rng(1).standard_normal((40, 3))).astype(np.float32)
    l0 = float(tr.train_step(batch, xyz_noise=noise))
    for _ in range(15):
        l1 = float(tr.train_step(batch, xyz_noise=noise))
    assert np.isfinite(l0) and l1 < l0


def test_checkpoint_save_resume_and_restore_model(ctx, tmp_path):
    """trainvali.py:134-141 / util/io.py:36-45 without TensorFlow: a trainer checkpoints in the
    tensor-bundle format under the reference's variable names, a fresh trainer resumes from it
    and continues bit-identically, and `restore_model` loads the weights into a fresh model."""
    from nerfactor_b200.trainvali import Trainer
    from nerfactor_b200.util import io as ioutil, tfckpt
    batch = synth.make_stage_b_batch(5, 64, 16, fg_frac=1.0)
    noise = (0.01 * np.random.default_rng(4).standard_normal((64, 3))).astype(np.float32)
    m0, _, _ = _models(ctx, 'microfacet')
    tr0 = Trainer(m0, precision='bf16')
    for _ in range(3):
        tr0.train_step(batch, xyz_noise=noise)
    prefix = tr0.save_checkpoint(str(tmp_path / 'checkpoints'), step=3)
    assert ioutil.latest_checkpoint(str(tmp_path / 'checkpoints')) == prefix
    names = tfckpt.read_checkpoint(prefix)
    assert 'net/net_albedo_mlp_layer0/kernel/.ATTRIBUTES/VARIABLE_VALUE' in names
    assert 'net/net_albedo_mlp_layer0/kernel/.OPTIMIZER_SLOT/optimizer/vhat/.ATTRIBUTES/VARIABLE_VALUE' in names
    assert 'net/_light/.ATTRIBUTES/VARIABLE_VALUE' in names
    for _ in range(2):
        tr0.train_step(batch, xyz_noise=noise)
    # resume in a fresh trainer (different initial weights) and replay the same two steps
    m1, _, _ = _models(ctx, 'microfacet', seed=99)
    tr1 = Trainer(m1, precision='bf16')
    assert tr1.restore_checkpoint(prefix) == 3 and tr1.iterations == 3
    for _ in range(2):
        tr1.train_step(batch, xyz_noise=noise)
    assert torch.equal(tr0.flat, tr1.flat) and torch.equal(tr0.vhat, tr1.vhat)
    # inference model restored from the checkpoint == the trainer's synced model
    tr1.restore_checkpoint(prefix)
    m2, _, _ = _models(ctx, 'microfacet', seed=5)
    found = ioutil.restore_model(m2, prefix)
    assert {'albedo_mlp', 'lvis_out', 'light'} <= found
    p1, _, _, _ = m1.call(batch, 'test')
    p2, _, _, _ = m2.call(batch, 'test')
    assert torch.equal(p1['rgb'], p2['rgb'])


@pytest.mark.parametrize('case', [(3000, 90, (128, 128, 128, 128, 1), (2,), 'bf16', True),
                                  (1029, 18, (128, 128, 128, 128, 1), (2,), 'bf16', True),
                                  (700, 63, (128, 128, 3), None, 'f16', False),
                                  (260, 24, (64, 256, 128, 4), (0,), 'bf16', True)])
def test_mlp_chain_matches_layer_by_layer(ctx, case, monkeypatch):
    """nf_mlp_chain_fwd / _bwd (one call per network, 16-bit activations in a private workspace)
    against the layer-by-layer nf_dense_fwd / nf_dense_bwd path (fp32 activations in HBM) it
    replaces: the operand images round the activations to 16 bit in both, so the forward values
    agree to the last fp32 bit or two and the gradients to ~1e-3 (bias-gradient sums are taken over
    16-bit instead of fp32 products); ragged row counts, input widths that need padding, a skip
    connection, input gradient on and off."""
    from nerfactor_b200 import autodiff as ad
    rows, in_dim, widths, skip_at, prec, need_dx = case
    g = torch.Generator(device='cpu').manual_seed(rows)
    x = torch.randn((rows, in_dim), generator=g).cuda().requires_grad_(need_dx)
    layers, k = [], in_dim
    for i, n in enumerate(widths):
        kin = k + (in_dim if (skip_at and i - 1 in skip_at) else 0)
        w = (torch.randn((kin, n), generator=g) * (1.5 / np.sqrt(kin))).cuda().requires_grad_(True)
        b = (torch.randn((n,), generator=g) * 0.1).cuda().requires_grad_(True)
        layers.append((w, b))
        k = n
    acts = ['relu'] * (len(widths) - 1) + ['sigmoid']
    dy = torch.randn((rows, widths[-1]), generator=g).cuda()
    outs = {}
    for chain in (True, False):
        monkeypatch.setattr(ad, 'CHAIN', chain)
        y = ad.mlp_apply(x, layers, acts, skip_at, prec)
        ins = [t for wb in layers for t in wb] + ([x] if need_dx else [])
        outs[chain] = (y.detach().clone(), [t.clone() for t in torch.autograd.grad(y, ins, dy)])
    ya, ga = outs[True]
    yb, gb = outs[False]
    assert ya.shape == yb.shape == (rows, widths[-1])
    assert float((ya - yb).abs().max()) <= 2e-6
    worst = max(rel_l2(a.cpu(), b.cpu()) for a, b in zip(ga, gb))
    print('mlp chain vs layer-by-layer: worst gradient rel-L2 %.2e' % worst)
    assert worst < 3e-3
