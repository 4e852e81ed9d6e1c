"""Generates the committed golden fixtures under tests/golden/.

Run ONLY in the build container (it imports the reference from /root/reference,
which does not exist on the GPU box):

    python tests/golden/make_golden.py

Two kinds of fixture:

* ref_pinned.npz   outputs of the pieces of the reference that import without
                   TensorFlow, run here on seeded inputs.  These
                   pin the oracle (tests/test_oracle_pinning.py).
* ref_sphere_renderer.npz   the reference's NumPy light-stage renderer
                   (brdf/renderer.py SphereRenderer: gen_light_xyz, light directions,
                   cosines, front-lit visibility, `calc_light_contrib`, `render`) run here on
                   its own sphere scene with a seeded env-map and a Lambertian BRDF: pins the
                   rendering-equation estimator (oracle AND, on the GPU, nf_integrate_fwd).
* oracle_*.npz     outputs of the oracle itself on seeded synthetic inputs
                   (the reference ships no golden vectors), so
                   the CUDA parity tests have frozen vectors that do not drift
                   with the oracle.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def make_ref_pinned():
    sys.path.insert(0, '/root/reference')
    import refpin
    refpin.pin()      # the reference's namespace package `brdf`, not the repo-root stub
    from brdf.renderer import gen_light_xyz                      # noqa
    from third_party.nielsen2015on.coordinateFunctions import \
        DirectionsToRusink                                       # noqa
    from third_party.xiuminglib import xiuminglib as xm          # noqa
    rng = np.random.default_rng(1234)
    out = {}
    for h, w in ((16, 32), (2, 8), (16, 64)):
        xyz, areas = gen_light_xyz(h, w)
        out['lxyz_%dx%d' % (h, w)] = xyz
        out['lareas_%dx%d' % (h, w)] = areas
    a = rng.standard_normal((512, 3))
    b = rng.standard_normal((512, 3))
    # keep both in the upper hemisphere like real (light, view) local directions
    a[:, 2] = np.abs(a[:, 2])
    b[:, 2] = np.abs(b[:, 2])
    out['rusink_a'], out['rusink_b'] = a, b
    out['rusink_out'] = DirectionsToRusink(a, b)
    sph = np.stack((rng.uniform(0.5, 100., 64), rng.uniform(-1.5, 1.5, 64),
                    rng.uniform(-3.1, 3.1, 64)), axis=1)
    out['sph_in'] = sph
    out['sph_out'] = xm.geometry.sph.sph2cart(sph)
    lin = np.concatenate((np.linspace(0., 0.01, 64), rng.uniform(0., 1., 192)))
    out['srgb_in'] = lin
    out['srgb_out'] = xm.img.linear2srgb(lin)
    nrm = rng.standard_normal((128, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    out['w2l_normal'] = nrm
    out['w2l_out'] = xm.geometry.normal.gen_world2local(nrm)
    np.savez_compressed(os.path.join(HERE, 'ref_pinned.npz'), **out)
    print('ref_pinned.npz:', sorted(out))


def make_ref_sphere_renderer():
    """brdf/renderer.py:23-183 as it is (NumPy, fp64)."""
    sys.path.insert(0, '/root/reference')
    import refpin
    refpin.pin()      # the reference's namespace package `brdf`, not the repo-root stub
    from brdf.renderer import SphereRenderer                     # noqa
    rng = np.random.default_rng(4321)
    h, ims = 8, 24
    r = SphereRenderer('point', '/tmp/nf_sphere_renderer', envmap_h=h, ims=ims, spp=1)
    envmap = rng.uniform(0., 0.4, size=(h, 2 * h, 3))           # render stays below the clip
    envmap[2, 5] = 3.0                                          # one bright texel
    lcontrib = r.calc_light_contrib(envmap)                     # H x W x L x 3
    albedo = rng.uniform(0.1, 0.9, size=(ims, ims, 3))
    brdf = np.tile((albedo / np.pi)[:, :, None, :], (1, 1, lcontrib.shape[2], 1))
    r.lcontrib = lcontrib
    render = r.render(brdf, white_bg=True)
    out = dict(xyz=r.xyz, normal=r.normal, is_fg=r.is_fg, lvis=r.lvis.astype(np.uint8),
               lcos=r.lcos, lxyz=r.lxyz, lareas=r.lareas, envmap=envmap, albedo=albedo,
               render=render, cam_loc=np.asarray(r.cam.loc, float))
    np.savez_compressed(os.path.join(HERE, 'ref_sphere_renderer.npz'), **out)
    print('ref_sphere_renderer.npz: fg pixels', int(r.is_fg.sum()), 'render max',
          float(render[r.is_fg].max()))


def make_oracle_goldens():
    from oracle import stage_a, stage_b, brdf as obrdf
    from nerfactor_b200 import synth
    # ---- Stage B, config-1-like: 96 rays, L = 16 (2x8), both BRDF variants
    for brdf in ('microfacet', 'learned'):
        lh, lw = 2, 8
        lxyz, lareas = obrdf.gen_light_xyz(lh, lw)
        params = synth.make_stage_b_params(7, brdf, light_hw=(lh, lw))
        batch = synth.make_stage_b_batch(11, 96, lh * lw)
        m = stage_b.StageB(params, {'brdf': brdf}, lxyz=lxyz, lareas=lareas)
        probes = synth.make_probes(5, 3, (lh, lw))
        pred, _, _ = m.call(batch, 'test', relight_lights=[p for p in probes])
        np.savez_compressed(
            os.path.join(HERE, 'oracle_stage_b_%s.npz' % brdf),
            seed_params=7, seed_batch=11, seed_probes=5, n_rays=96, lh=lh, lw=lw,
            **{k: v.numpy() for k, v in pred.items()})
        print('oracle_stage_b_%s.npz' % brdf, {k: tuple(v.shape) for k, v in pred.items()})
    # ---- Stage A, config-1-like: 8x8 view, S = 32 single pass + hierarchical
    nerf = synth.make_nerf_params(3)
    c2w = synth.look_at_c2w()
    rayo, rayd = stage_a.gen_rays(c2w, synth.CAM_ANGLE_X, 8, 8)
    ro = torch.tensor(rayo.reshape(-1, 3))
    rd = stage_a.l2_normalize(torch.tensor(rayd.reshape(-1, 3)), 1)
    sp = stage_a.march_single_pass(nerf, ro, rd, 2., 6., 32)
    occu, depth, normal = stage_a.compute_depth_and_normal(
        nerf, ro, rd, 2., 6., n_samples_coarse=-32, n_samples_fine=-16)  # 32 + 48
    lx, _ = obrdf.gen_light_xyz(2, 4)
    surf = ro + rd * depth[:, None]
    lvis = stage_a.compute_light_visibility(
        nerf, surf[:16], normal[:16], lx, n_samples_coarse=-32, n_samples_fine=-16)
    np.savez_compressed(
        os.path.join(HERE, 'oracle_stage_a.npz'), seed_nerf=3, rayo=rayo, rayd=rayd,
        sp_sigma=sp['sigma'].numpy(), sp_weights=sp['weights'].numpy(),
        sp_occu=sp['occu'].numpy(), sp_depth=sp['depth'].numpy(),
        sp_surf=sp['surf'].numpy(), h_occu=occu.numpy(), h_depth=depth.numpy(),
        h_normal=normal.numpy(), h_lvis=lvis.numpy())
    print('oracle_stage_a.npz')


if __name__ == '__main__':
    if 'sphere' in sys.argv[1:]:
        make_ref_sphere_renderer()
        sys.exit(0)
    make_ref_pinned()
    make_ref_sphere_renderer()
    make_oracle_goldens()
