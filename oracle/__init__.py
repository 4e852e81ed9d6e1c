"""CPU oracle for the NeRFactor render-and-relight hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is an op-for-op CPU restatement (PyTorch-CPU, fp32 by default, fp64 on
request) of the reference's TensorFlow-2.2 algorithm for the path named by
BASELINE.json `north_star`:

  Stage A  nerfactor/geometry_from_nerf.py + nerfactor/models/nerf.py
  Stage B  nerfactor/models/{nerfactor,nerfactor_microfacet,shape,brdf}.py,
           nerfactor/networks/{mlp,seq,embedder}.py,
           brdf/microfacet/microfacet.py, brdf/renderer.py,
           nerfactor/util/{math,geom,img,tensor}.py

Every function cites the reference file:line it follows.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it; the product package `nerfactor_b200` never does.

PARITY PINNING.  The reference ships no tests, golden vectors or checkpoints
 and its arithmetic lives in TensorFlow 2.2
(environment.yml:19, an un-vendored pip dependency that cannot be installed in
this image), so the TF-dependent parts of this oracle are **parity unpinned**:
they are a line-by-line restatement reviewed against the cited ranges.  The
pieces of the reference that ARE importable without TensorFlow are pinned
against the reference run in the build container (tests/golden/make_golden.py
generates the fixtures, tests/test_oracle_pinning.py checks them):

  brdf.renderer.gen_light_xyz                      -> oracle.brdf.gen_light_xyz
  third_party.nielsen2015on DirectionsToRusink     -> oracle.brdf.dir2rusink
  third_party.xiuminglib ... sph2cart              -> oracle.brdf.sph2cart
  third_party.xiuminglib ... img.linear2srgb       -> oracle.stage_b.linear2srgb
  third_party.xiuminglib ... normal.gen_world2local (no-eps twin)
  brdf.renderer.SphereRenderer (NumPy light-stage estimator) -> oracle.stage_b.StageB.calc_ldir
                                                   + render (tests/golden/ref_sphere_renderer.npz)
"""
