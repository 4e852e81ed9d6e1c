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
this image).  The oracle is pinned in two ways, both by running the reference
itself in the build container (generators committed under tests/golden/):

(1) THE REFERENCE'S OWN MODEL CODE, unmodified, executed op by op through a
    small eager TensorFlow look-alike (tests/golden/tfshim: each function is
    the documented semantics of the one TF op of that name, on PyTorch-CPU):
    nerfactor/models/{shape,nerfactor,nerfactor_microfacet,brdf,nerf}.py
    Model.call / compute_loss, networks/*, brdf/microfacet/microfacet.py,
    util/{math,geom,img,tensor,light}.py and geometry_from_nerf.py
    compute_depth_and_normal / compute_light_visibility / eval_sigma_mlp
    -> tests/golden/ref_tfshim_*.npz (make_golden_tfshim.py).  The oracle
    reproduces them bit-for-bit (Stage B: forward, jitter, losses, edits, OLAT,
    probes; Stage A coarse pass, sigma) or to <= 2e-6 (Stage A fine pass).
    What this leaves to trust is the per-op semantics of the shim (l2_normalize,
    divide_no_nan, exclusive cumprod, searchsorted side, floormod, Dense,
    LinSpace, antialiased resize, scatter / gather), not the algorithm.
(2) The pieces of the reference that import without TensorFlow, run as they are
    (tests/golden/make_golden.py, tests/test_oracle_pinning.py):

  brdf.renderer.gen_light_xyz                      -> oracle.brdf.gen_light_xyz
  third_party.nielsen2015on DirectionsToRusink     -> oracle.brdf.dir2rusink
  third_party.xiuminglib ... sph2cart              -> oracle.brdf.sph2cart
  third_party.xiuminglib ... img.linear2srgb       -> oracle.stage_b.linear2srgb
  third_party.xiuminglib ... normal.gen_world2local (no-eps twin)
  brdf.renderer.SphereRenderer (NumPy light-stage estimator) -> oracle.stage_b.StageB.calc_ldir
                                                   + render (tests/golden/ref_sphere_renderer.npz)
"""
