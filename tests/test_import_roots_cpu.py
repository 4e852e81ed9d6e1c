"""Import-root drop-in (SURVEY.md 8b "must resolve both spellings"): the reference's own import
statements -- `nerfactor.models.X`, `brdf.renderer`, and with $REPO/nerfactor on sys.path the
bare `models.X` / `datasets.X` / `networks` / `util` / `losses` -- resolve to this package, to the
SAME module objects.  Run in fresh interpreters (other tests of this suite import the real
reference under those names)."""
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(code, extra_path=()):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, *extra_path]))
    r = subprocess.run([sys.executable, '-c', code], env=env, capture_output=True, text=True,
                       timeout=300, cwd='/')
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_spellings_resolve_to_the_same_modules():
    out = _run('''
import nerfactor_b200
# every import statement of the reference's hot-path files (nerfactor/models/nerfactor.py:17-27,
# shape.py:17-24, nerfactor_microfacet.py:15-19, trainvali.py:26-31, test.py:24-31)
from nerfactor.networks import mlp
from nerfactor.networks.embedder import Embedder
from nerfactor.networks.layers import LatentCode
from nerfactor.models.base import Model as BaseModel
from nerfactor.models.shape import Model as ShapeModel
from nerfactor.models.nerfactor import Model as NeRFactorModel
from nerfactor.models.brdf import Model as BRDFModel
from nerfactor.datasets.nerf_shape import Dataset
from nerfactor import models, datasets
from nerfactor.util import logging as logutil, io as ioutil, tensor as tutil, \\
    math as mathutil, img as imgutil, config as configutil, vis as visutil, light as lightutil, \\
    geom as geomutil
from brdf.renderer import gen_light_xyz
from brdf.microfacet.microfacet import Microfacet
import nerfactor_b200.models.nerfactor as twin, nerfactor_b200.brdf.microfacet.microfacet as twin_m
import sys
assert NeRFactorModel is twin.Model and Microfacet is twin_m.Microfacet
assert sys.modules['nerfactor.models.nerfactor'] is twin
assert models.get_model_class('nerfactor_microfacet').__module__ == 'nerfactor_b200.models.nerfactor_microfacet'
assert datasets.get_dataset_class('nerf_shape') is Dataset
xyz, areas = gen_light_xyz(16, 32)
assert xyz.shape == (16, 32, 3) and abs(areas.sum() - 12.566370614359172) < 1e-9
logutil.Logger(loggee="test").info("aliases %s", "ok")
print("OK")
''')
    assert out.strip().endswith('OK')


def test_bare_spellings_with_the_package_directory_on_the_path():
    """`python $REPO/nerfactor/trainvali.py` puts $REPO/nerfactor first on sys.path: the reference's
    registries then import `models.<name>` / `datasets.<name>` (models/__init__.py:19) and
    models/base.py imports bare `losses`, `networks`, `util`."""
    out = _run('''
import nerfactor                      # installs the bare names because .../nerfactor is on sys.path
from importlib import import_module
import losses
from networks import base as basenet
from util import logging as logutil
mod = import_module('models.nerfactor')
import nerfactor_b200.models.nerfactor as twin
assert mod is twin and import_module('datasets.nerf').Dataset.__module__ == 'nerfactor_b200.datasets.nerf'
assert losses.L2 is import_module('nerfactor_b200.losses').L2
print("OK")
''', extra_path=[os.path.join(ROOT, 'nerfactor')])
    assert out.strip().endswith('OK')


def test_script_stubs_parse_the_reference_flags():
    for script in ('trainvali.py', 'test.py', 'geometry_from_nerf.py'):
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'nerfactor', script), '--help'],
                           capture_output=True, text=True, timeout=300, cwd='/',
                           env=dict(os.environ, PYTHONPATH=ROOT))
        assert r.returncode == 0 and 'usage' in r.stdout.lower(), r.stderr[-1000:]


def test_helper_modules():
    from nerfactor_b200 import losses
    from nerfactor_b200.networks.embedder import Embedder
    from nerfactor_b200.networks.layers import LatentCode
    from nerfactor_b200.util import tensor as tutil
    rng = np.random.default_rng(0)
    gt, pr = torch.tensor(rng.uniform(size=(5, 7, 3))), torch.tensor(rng.uniform(size=(5, 7, 3)))
    w = torch.tensor(rng.uniform(size=(5, 7)))
    # Keras MeanSquaredError(reduction='none')(gt, pred, sample_weight): mean over the last axis,
    # times the weight; then reduce_mean
    assert torch.allclose(losses.L2()(gt, pr), ((gt - pr) ** 2).mean())
    assert torch.allclose(losses.L2()(gt, pr, weights=w), (((gt - pr) ** 2).mean(-1) * w).mean())
    assert losses.L2()(gt, pr, keep_batch=True).shape == (5,)
    assert torch.allclose(losses.L1()(gt, pr), (gt - pr).abs().mean())
    # embedder.py:23-47, general configurations
    x = torch.tensor(rng.uniform(-1, 1, (4, 3)), dtype=torch.float32)
    e = Embedder(incl_input=False, in_dims=3, log2_max_freq=4, n_freqs=3, log_sampling=True)
    assert e.out_dims == 18 and not e.fused_ok
    want = torch.cat([f(x * b) for b in (1., 4., 16.) for f in (torch.sin, torch.cos)], -1)
    assert torch.allclose(e(x), want)
    e = Embedder(incl_input=True, in_dims=3, log2_max_freq=3, n_freqs=4, log_sampling=False)
    assert np.allclose(e.freq_bands, [1., 1. + 7. / 3, 1. + 14. / 3, 8.])
    assert Embedder(log2_max_freq=9, n_freqs=10).fused_ok
    # layers.py:58-67: slerp of normalised codes stays on the sphere and hits the end points
    lc = LatentCode(3, 4, normalize=True, rng=rng)
    z = lc.interp(0.25, 0, 0.75, 2)
    assert abs(np.linalg.norm(z) - 1.) < 1e-5
    assert np.allclose(lc.interp(1., 0, 0., 2), lc(0), atol=1e-6)
    assert tutil.make_nhwc(torch.zeros(2, 4, 5), 3).shape == (2, 4, 5, 3)
    assert float(tutil.one_hot_img(2, 3, 3, 1, 2).sum()) == 3.
