"""A tiny eager `tensorflow` look-alike on top of PyTorch-CPU -- TEST INFRASTRUCTURE.

Purpose: TensorFlow 2.2 cannot be installed in this image, so the reference's model
code cannot run as shipped.  With this package first on sys.path, `import tensorflow as tf` inside
the UNMODIFIED reference files (/root/reference/nerfactor/models/*.py, networks/*.py, util/*.py,
brdf/microfacet/microfacet.py, geometry_from_nerf.py) resolves here, and the reference's own
algorithm runs op by op: every function below is the documented semantics of the one TF op of the
same name, nothing more (no fusing, no re-ordering).  tests/golden/make_golden.py uses it, in the
build container only, to produce `tests/golden/ref_tfshim_*.npz`: outputs of the reference's code
on seeded inputs, against which the oracle (CPU) and the CUDA kernels (GPU) are then checked.

Only the ops on the render-and-relight path are provided; anything else raises AttributeError.
tf.Tensor IS torch.Tensor (so `isinstance(x, tf.Tensor)` and `.numpy()` work as in eager TF).
"""
import builtins

import numpy as np
import torch

from . import keras, math, linalg, debugging, random, image, nn, train  # noqa: F401

__version__ = '2.2.0-shim'

# Eager TF tensors are plain values: `.numpy()` works on anything, also on results computed under
# a GradientTape.  torch refuses that for tensors with autograd history, so while the shim is
# loaded (golden generator / build-container-only tests) `.numpy()` and `np.asarray` detach first.
if not getattr(torch.Tensor, '_shim_numpy', False):
    _orig_numpy = torch.Tensor.numpy

    def _numpy(self, *a, **k):
        return _orig_numpy(self.detach(), *a, **k)

    def _array(self, dtype=None, *a, **k):
        arr = _orig_numpy(self.detach())
        return arr if dtype is None else arr.astype(dtype, copy=False)

    torch.Tensor.numpy = _numpy
    torch.Tensor.__array__ = _array
    torch.Tensor._shim_numpy = True
Tensor = torch.Tensor
float16, float32, float64 = torch.float16, torch.float32, torch.float64
int32, int64, uint8 = torch.int32, torch.int64, torch.uint8
bool = torch.bool
string = 'string'
newaxis = None

_DT = {'float32': torch.float32, 'float64': torch.float64, 'int32': torch.int32,
       'int64': torch.int64, 'bool': torch.bool, 'uint8': torch.uint8,
       builtins.float: torch.float32, builtins.int: torch.int32, builtins.bool: torch.bool}


def as_dtype(d):
    return _DT.get(d, d)


class _Shape(tuple):
    """What tf.shape returns here: a tuple of Python ints that ops below accept as a tensor."""

    def numpy(self):
        return np.asarray(self, np.int32)


def _t(x, dtype=None):
    """Anything -> torch tensor (Python floats become float32, ints int32, like TF)."""
    dtype = as_dtype(dtype) if dtype is not None else None
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, _Shape):
        return torch.tensor(list(x), dtype=dtype or torch.int32)
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        return t if dtype is None else t.to(dtype)
    if isinstance(x, (list, tuple)) and len(x) and any(isinstance(e, torch.Tensor) for e in x):
        return torch.stack([_t(e, dtype) for e in x])
    arr = np.asarray(x)
    if dtype is None:
        if arr.dtype == np.float64:
            dtype = torch.float32
        elif arr.dtype == np.int64:
            dtype = torch.int32
    return torch.as_tensor(arr).to(dtype) if dtype is not None else torch.as_tensor(arr)


def convert_to_tensor(value, dtype=None, **_):
    return _t(value, dtype)


def constant(value, dtype=None, **_):
    return _t(value, dtype)


_TRAINING = [False]


def shim_set_training(flag):
    """Shim control (not a TF symbol): when on, Variables / Dense weights created or set from now
    on are autograd leaves, so GradientTape.gradient works like TF's automatic variable watching.
    Off by default so that forward-only runs keep plain tensors (`.numpy()` works everywhere)."""
    _TRAINING[0] = builtins.bool(flag)


def Variable(initial_value=None, trainable=True, dtype=None, **_):
    v = _t(initial_value, dtype).detach().clone()
    v._shim_variable = True
    if trainable and _TRAINING[0] and v.is_floating_point():
        v.requires_grad_(True)
    return v


def identity(x):
    return x


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def stop_gradient(x):
    return x.detach()


def cast(x, dtype):
    return _t(x).to(as_dtype(dtype))


def shape(x):
    """1-D int32 tensor (its elements are 0-D tensors with a dtype, as in TF)."""
    return torch.tensor([int(s) for s in _t(x).shape], dtype=torch.int32)


def rank(x):
    return _t(x).dim()


def _dims(shape_):
    if isinstance(shape_, torch.Tensor):
        return [int(v) for v in shape_.reshape(-1).tolist()]
    if isinstance(shape_, (int, np.integer)):
        return [int(shape_)]
    return [int(v) for v in shape_]


def reshape(x, shape_):
    return _t(x).reshape(_dims(shape_))


def zeros(shape_, dtype=float32):
    return torch.zeros(_dims(shape_), dtype=as_dtype(dtype))


def ones(shape_, dtype=float32):
    return torch.ones(_dims(shape_), dtype=as_dtype(dtype))


def zeros_like(x, dtype=None):
    return torch.zeros_like(_t(x), dtype=as_dtype(dtype) if dtype else None)


def ones_like(x, dtype=None):
    return torch.ones_like(_t(x), dtype=as_dtype(dtype) if dtype else None)


def concat(values, axis):
    return torch.cat([_t(v) for v in values], dim=axis)


def stack(values, axis=0):
    vals = [_t(v) for v in values]
    return torch.stack(vals, dim=axis)


def tile(x, multiples):
    return _t(x).repeat(*_dims(multiples))


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*perm) if perm is not None else x.permute(*reversed(range(x.dim())))


def broadcast_to(x, shape_):
    return _t(x).broadcast_to(_dims(shape_))


def range(*args, dtype=None, **_):  # noqa: A001
    out = torch.arange(*[int(a) for a in args], dtype=torch.int32)
    return out if dtype is None else out.to(as_dtype(dtype))


def linspace(start, stop, num):
    """tf.linspace (LinSpace CPU kernel): float32; flat(i) = start + step * i with
    step = (stop - start) / (num - 1), the last element set to `stop` exactly."""
    num = int(num)
    start, stop = np.float32(start), np.float32(stop)
    if num == 1:
        return torch.tensor([start], dtype=torch.float32)
    step = np.float32((stop - start) / np.float32(num - 1))
    out = (start + step * np.arange(num, dtype=np.float32)).astype(np.float32)
    out[-1] = stop
    return torch.from_numpy(out)


def meshgrid(*args, indexing='xy'):
    return list(torch.meshgrid(*[_t(a) for a in args], indexing=indexing))


def where(condition, x=None, y=None):
    if x is None:
        return torch.nonzero(condition).to(torch.int64)            # [n, rank]
    return torch.where(condition, _t(x), _t(y))


def boolean_mask(tensor, mask, axis=None):
    assert axis in (None, 0)
    tensor, mask = _t(tensor), _t(mask)
    return tensor[mask]


def gather(params, indices, axis=None, batch_dims=0):
    params, indices = _t(params), _t(indices).long()
    if batch_dims == 0:
        return torch.index_select(params, axis or 0, indices.reshape(-1)).reshape(
            params.shape[:axis or 0] + indices.shape + params.shape[(axis or 0) + 1:])
    # the one batched use on the path (util/math.py:86-87): [n, m] gathered with [n, s, 2] on axis -1
    assert batch_dims == indices.dim() - 2 == 1 and axis in (-1, params.dim() - 1)
    flat = indices.reshape(indices.shape[0], -1)
    return torch.gather(params, 1, flat).reshape(indices.shape)


def gather_nd(params, indices):
    params, indices = _t(params), _t(indices).long()
    k = indices.shape[-1]
    idx = tuple(indices[..., i] for i in builtins.range(k))
    return params[idx]


def scatter_nd(indices, updates, shape_):
    indices, updates = _t(indices).long(), _t(updates)
    out = torch.zeros(_dims(shape_), dtype=updates.dtype)
    k = indices.shape[-1]
    idx = tuple(indices[..., i] for i in builtins.range(k))
    return out.index_put(idx, updates, accumulate=True)              # duplicates add, like TF


def tensor_scatter_nd_update(tensor, indices, updates):
    tensor, indices, updates = _t(tensor).clone(), _t(indices).long(), _t(updates)
    idx = tuple(indices[..., i] for i in builtins.range(indices.shape[-1]))
    tensor[idx] = updates
    return tensor


def _no_axes(axis):
    """axis=() / []: TF reduces over NO dimension (torch would reduce over all of them)."""
    return isinstance(axis, (tuple, list)) and len(axis) == 0


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    if _no_axes(axis):
        return x
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    if isinstance(x, torch.Tensor) and not x.is_floating_point():
        x = x.float()
    if _no_axes(axis):
        return x
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False):
    x = _t(x)
    return x.max() if axis is None else x.amax(dim=axis, keepdim=keepdims)


def reduce_min(x, axis=None, keepdims=False):
    x = _t(x)
    return x.min() if axis is None else x.amin(dim=axis, keepdim=keepdims)


def cumsum(x, axis=0):
    return torch.cumsum(_t(x), dim=axis)


def einsum(eq, *ops):
    return torch.einsum(eq, *[_t(o) for o in ops])


def matmul(a, b):
    return torch.matmul(_t(a), _t(b))


def multiply(a, b):
    return _t(a) * _t(b)


def clip_by_value(x, clip_value_min, clip_value_max):
    x = _t(x)
    lo = clip_value_min if isinstance(clip_value_min, torch.Tensor) else float(clip_value_min)
    hi = clip_value_max if isinstance(clip_value_max, torch.Tensor) else float(clip_value_max)
    return torch.clamp(x, min=lo, max=hi)


def maximum(a, b):
    a, b = _t(a), _t(b)
    return torch.maximum(a, b.to(a.dtype))


def minimum(a, b):
    a, b = _t(a), _t(b)
    return torch.minimum(a, b.to(a.dtype))


def logical_and(a, b):
    return torch.logical_and(a, b)


def logical_or(a, b):
    return torch.logical_or(a, b)


def equal(a, b):
    return _t(a) == _t(b)


def roll(x, shift, axis):
    return torch.roll(_t(x), shifts=shift, dims=axis)


def sort(x, axis=-1, direction='ASCENDING'):
    return torch.sort(_t(x), dim=axis, descending=direction != 'ASCENDING').values


def searchsorted(sorted_sequence, values, side='left'):
    return torch.searchsorted(_t(sorted_sequence).contiguous(), _t(values).contiguous(),
                              right=(side == 'right')).to(torch.int32)


def ensure_shape(x, shape_):
    return x


class _CustomGradient(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f, n_args, *args):
        with torch.no_grad():
            y, grad_fn = f(*[a.detach() if isinstance(a, torch.Tensor) else a for a in args])
        ctx.grad_fn_, ctx.n = grad_fn, n_args
        return y

    @staticmethod
    def backward(ctx, dy):
        g = ctx.grad_fn_(dy)
        g = g if isinstance(g, (tuple, list)) else (g,)
        return (None, None) + tuple(g) + (None,) * (ctx.n - len(g))


def custom_gradient(f):
    """y, grad = f(*args): forward value y, backward through the function's own `grad`."""
    def wrapped(*args, **kwargs):
        assert not kwargs
        args = [a if isinstance(a, torch.Tensor) else _t(a) for a in args]
        return _CustomGradient.apply(f, len(args), *args)
    return wrapped


def random_normal_initializer(mean=0.0, stddev=0.05, seed=None):
    def init(shape, dtype=None):
        return random.normal(shape, mean=mean, stddev=stddev)
    return init


def function(f=None, **_):
    return f if f is not None else (lambda g: g)


def py_function(func, inp, Tout):
    return func(*inp)


class GradientTape:
    """watch / batch_jacobian for the one use on the path (geometry_from_nerf.py:289-297:
    d sigma / d xyz, rows independent)."""

    def __init__(self, persistent=False, **_):
        self._grad_was = None

    def __enter__(self):
        self._grad_was = torch.is_grad_enabled()
        torch.set_grad_enabled(True)
        return self

    def __exit__(self, *a):
        torch.set_grad_enabled(self._grad_was)

    def watch(self, x):
        x.requires_grad_(True)

    def batch_jacobian(self, target, source):
        assert target.dim() == 2 and source.dim() == 2
        cols = []
        for j in builtins.range(target.shape[1]):
            (g,) = torch.autograd.grad(target[:, j].sum(), source, retain_graph=True)
            cols.append(g)
        return torch.stack(cols, dim=1).detach()                       # [N, out, in]

    def gradient(self, target, sources):
        return torch.autograd.grad(target, sources, allow_unused=True)


# element-wise functions also exported at top level by TF
sin, cos, exp, sqrt, abs, square = torch.sin, torch.cos, torch.exp, torch.sqrt, torch.abs, torch.square  # noqa: A001
acos, atan2 = torch.acos, (lambda y, x: torch.atan2(_t(y), _t(x)))
rsqrt = torch.rsqrt
