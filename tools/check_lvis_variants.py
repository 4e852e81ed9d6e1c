"""Compares and times the two generations of the light-visibility / learned-BRDF tcgen05 kernel
(csrc/nf_mlp_tc.cu): v3 = default for the visibility network (per-point work on prefetch warps),
v2 = NF_LVIS_V2=1 (bias inside the MMA; default for the BRDF network), v1 = NF_LVIS_V1=1 (bias in the
epilogue).

    python tools/check_lvis_variants.py            # on a B200 (gpurun)

The switch is read once per process, so each variant runs in its own subprocess on the same
seeded inputs (640 k points x 512 lights for the timing, a ragged 203 x 200 case for edge tiles);
prints one JSON object: max |difference| (the head's 128-term dot product is summed in two halves
in the variant, so ~1e-7 is expected, not 0) and the two times."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import json, sys, numpy as np, torch
sys.path.insert(0, %r)
from nerfactor_b200 import _lib, synth, config as nfconfig
from nerfactor_b200.models.nerfactor import Model
from nerfactor_b200.brdf.renderer import gen_light_xyz
ctx = _lib.default_context()
out = {}
for tag, n, lh, lw in (('ragged', 203, 10, 20), ('full', 640000, 16, 32)):
    params = synth.make_stage_b_params(21, 'learned', light_hw=(lh, lw))
    m = Model(nfconfig.default_config('nerfactor', light_h=lh), params=params, ctx=ctx, precision='f16')
    lxyz, lareas = gen_light_xyz(lh, lw)
    m.set_lights(lxyz.reshape(-1, 3), lareas.reshape(-1)); m.light_res = (lh, lw)
    b = synth.make_stage_b_batch(22, n, 1, fg_frac=1.0)
    xyz = torch.as_tensor(b[6]).cuda(); nrm = torch.as_tensor(b[7]).cuda(); cam = torch.as_tensor(b[2]).cuda()
    z = m._pred_brdf_at(xyz)
    def t(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    lv = m._pred_lvis_at(xyz)
    sp = m._eval_brdf_at(None, None, nrm, None, z, pts=xyz, cam=cam)['spec']
    out[tag] = {'lvis_ms': t(lambda: m._pred_lvis_at(xyz)),
                'brdf_ms': t(lambda: m._eval_brdf_at(None, None, nrm, None, z, pts=xyz, cam=cam))}
    sel = slice(0, min(n, 4096))
    np.save(sys.argv[1] + '_%%s_lvis.npy' %% tag, lv[sel].cpu().numpy())
    np.save(sys.argv[1] + '_%%s_spec.npy' %% tag, sp[sel].cpu().numpy())
print(json.dumps(out))
''' % ROOT


def run(flag, prefix, **extra):
    env = dict(os.environ, NF_LVIS_V1=flag, **extra)
    r = subprocess.run([sys.executable, '-c', WORKER, prefix], env=env, capture_output=True,
                       text=True, timeout=200)
    if r.returncode != 0:
        return {'error': r.stderr[-600:]}
    return json.loads([l for l in r.stdout.splitlines() if l.startswith('{')][-1])


def main():
    import numpy as np
    import tempfile
    d = tempfile.mkdtemp()
    res = {'v3': run('0', os.path.join(d, 'c')),
           'v2': run('0', os.path.join(d, 'a'), NF_LVIS_V2='1'), 'v1': run('1', os.path.join(d, 'b'))}
    res['brdf_v2'] = run('0', os.path.join(d, 'e'), NF_BRDF_V3='0')
    if 'error' not in res['brdf_v2'] and 'error' not in res['v3']:
        for tag in ('ragged', 'full'):
            c = np.load(os.path.join(d, 'c_%s_spec.npy' % tag))
            e = np.load(os.path.join(d, 'e_%s_spec.npy' % tag))
            res['maxdiff_brdf_v3_v2_%s_spec' % tag] = float(np.abs(c - e).max())
    if all('error' not in res[k] for k in ('v1', 'v2', 'v3')):
        for tag in ('ragged', 'full'):
            for k in ('lvis', 'spec'):
                a = np.load(os.path.join(d, 'a_%s_%s.npy' % (tag, k)))
                b = np.load(os.path.join(d, 'b_%s_%s.npy' % (tag, k)))
                res['maxdiff_v2_v1_%s_%s' % (tag, k)] = float(np.abs(a - b).max())
            a = np.load(os.path.join(d, 'a_%s_lvis.npy' % tag))
            c = np.load(os.path.join(d, 'c_%s_lvis.npy' % tag))
            res['maxdiff_v3_v2_%s_lvis' % tag] = float(np.abs(a - c).max())
    print(json.dumps(res))


if __name__ == '__main__':
    main()
