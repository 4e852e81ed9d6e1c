"""Mirror of nerfactor/util/tensor.py:25-64 on torch tensors."""
import torch


def shape_as_list(x):
    return list(x.shape)


def make_nhwc(batch, c=3):
    """N x H x W (x 1) -> N x H x W x C by repeating the single channel (tensor.py:29-48)."""
    if batch.dim() == 3:
        batch = batch[..., None]
    if batch.dim() != 4 or batch.shape[3] != 1:
        raise ValueError("expected N x H x W or N x H x W x 1, got %s" % (tuple(batch.shape),))
    return batch.expand(-1, -1, -1, c).contiguous()


def eager_tensor_to_str(x):
    if isinstance(x, str):
        return x
    if isinstance(x, bytes):
        return x.decode()
    return x.item().decode() if hasattr(x, 'item') else str(x)


def one_hot_img(h, w, c, i, j):
    """float32 H x W x C with ones at (i, j, :) (tensor.py:57-64)."""
    out = torch.zeros((h, w, c), dtype=torch.float32)
    out[i, j, :] = 1.
    return out
