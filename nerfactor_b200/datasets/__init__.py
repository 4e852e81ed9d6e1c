"""Mirror of nerfactor/datasets/__init__.py:18-20 (name -> class registry).  `nerf`,
`nerf_shape` and `mvs_shape` are the data formats either side of the hot path;
`brdf_merl` (training data of the BRDF prior) is out of scope."""
from importlib import import_module

from . import nerf_shape  # noqa: F401


def get_dataset_class(name):
    return import_module('nerfactor_b200.datasets.' + name).Dataset
