"""Mirror of nerfactor/losses.py:20-80 on torch tensors (Keras `reduction='none'` semantics:
the per-element loss is first averaged over the LAST axis, `weights` multiply that result)."""
import torch


def _per_sample(elem, weights):
    loss = elem.mean(dim=-1)
    if weights is not None:
        loss = loss * torch.as_tensor(weights, dtype=loss.dtype, device=loss.device)
    return loss


class L1:
    def __call__(self, gt, pred, weights=None):
        return _per_sample((gt - pred).abs(), weights).mean()


class L2:
    def __call__(self, gt, pred, keep_batch=False, weights=None):
        loss = _per_sample((gt - pred) ** 2, weights)
        if keep_batch:
            return loss.reshape(loss.shape[0], -1).mean(dim=1) if loss.dim() > 1 else loss
        return loss.mean()


class UVL2:
    """L2 on the chroma channels of YUV (losses.py:49-65; tf.image.rgb_to_yuv matrix)."""
    _M = ((0.299, -0.14714119, 0.61497538), (0.587, -0.28886916, -0.51496512),
          (0.114, 0.43601035, -0.10001026))

    def __call__(self, gt, pred, weights=None):
        m = torch.tensor(self._M, dtype=gt.dtype, device=gt.device)
        uv = lambda x: (torch.clamp(x, 0., 1.) @ m)[..., 1:]
        return _per_sample((uv(gt) - uv(pred)) ** 2, weights).mean()


class SSIM:
    def __init__(self, dynamic_range):
        self.dynamic_range = dynamic_range

    def __call__(self, gt, pred, weights=None):
        raise NotImplementedError("SSIM loss has no caller in the NeRFactor pipeline "
                                  "(nerfactor/losses.py:68-80 is unused by the models)")
