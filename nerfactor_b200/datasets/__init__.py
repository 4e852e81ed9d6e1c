"""Mirror of the hot-path parts of nerfactor/datasets (ray generation lives in the CUDA
library: nf_gen_rays; file I/O is out of scope)."""
from . import nerf_shape  # noqa: F401
