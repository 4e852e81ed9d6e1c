"""tf.debugging.*: the checks are real checks."""
import torch


def check_numerics(tensor, message):
    if not bool(torch.isfinite(tensor).all()):
        raise FloatingPointError(message + ': tensor had NaN / Inf values')
    return tensor


def assert_greater(x, y, message=None, **_):
    if not bool((torch.as_tensor(x) > torch.as_tensor(y)).all()):
        raise AssertionError(message or 'assert_greater failed')


def Assert(condition, data, **_):
    if not bool(condition):
        raise AssertionError(data)
