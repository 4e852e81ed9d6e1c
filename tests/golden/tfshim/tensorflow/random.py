"""tf.random.*: seeded torch draws (the golden inputs never depend on them: jitter and stratified
perturbation are disabled or overridden where the reference would draw)."""
import torch

_gen = torch.Generator().manual_seed(0)


def set_seed(seed):
    _gen.manual_seed(int(seed))


def normal(shape, mean=0.0, stddev=1.0, dtype=torch.float32, **_):
    return torch.randn([int(s) for s in shape], generator=_gen, dtype=dtype) * stddev + mean


def uniform(shape, minval=0, maxval=None, dtype=torch.float32, **_):
    shape = [int(s) for s in shape]
    if dtype in (torch.int32, torch.int64):
        return torch.randint(int(minval), int(maxval), shape, generator=_gen).to(dtype)
    maxval = 1.0 if maxval is None else maxval
    return torch.rand(shape, generator=_gen, dtype=dtype) * (maxval - minval) + minval
