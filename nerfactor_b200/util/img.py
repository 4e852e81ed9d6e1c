"""Mirror of nerfactor/util/img.py (alpha_blend :76-95, resize :98-137, linear2srgb :140-163,
frame_image / hconcat / vconcat) plus the xiuminglib image helpers the hot path's callers use
(third_party/xiuminglib/xiuminglib/img.py: normalize_uint :10-25, denormalize_float :28-45,
resize(method='cv2') :77-121, rgb2lum :597-611, tonemap :690-719; io/img.py write_float
:122-154; metric.py PSNR :103-149).  torch tensors are handled where the models call these on the
device ([N,3]-sized loss glue); everything image-sized is host-side numpy.
"""
import os

import numpy as np
import torch


# ------------------------------------------------------------------ blending / tone curves
def alpha_blend(tensor1, alpha, tensor2=None):
    """util/img.py:76-95.  [H,W,C] with an [H,W] alpha broadcasts the alpha over channels."""
    if isinstance(tensor1, torch.Tensor):
        if tensor2 is None:
            tensor2 = torch.zeros_like(tensor1)
        if tensor1.dim() == 3 and alpha.dim() == 2:
            alpha = alpha[:, :, None]
        return tensor1 * alpha + tensor2 * (1. - alpha)
    tensor1 = np.asarray(tensor1)
    alpha = np.asarray(alpha)
    if tensor2 is None:
        tensor2 = np.zeros_like(tensor1)
    if tensor1.ndim == 3 and alpha.ndim == 2:
        alpha = np.tile(alpha[:, :, None], (1, 1, tensor1.shape[2]))
    return tensor1 * alpha + tensor2 * (1. - alpha)


def linear2srgb(tensor_0to1):
    """util/img.py:140-163 (clip, then pow on every element, then select)."""
    if isinstance(tensor_0to1, torch.Tensor):
        x = torch.clamp(tensor_0to1, 0., 1.)
        return torch.where(x <= 0.0031308, x * 12.92, 1.055 * torch.pow(x, 1 / 2.4) - 0.055)
    x = np.clip(np.asarray(tensor_0to1), 0., 1.)
    return np.where(x <= 0.0031308, x * 12.92, 1.055 * np.power(x, 1 / 2.4) - 0.055)


def tonemap(hdr, method='gamma', gamma=2.2):
    """xiuminglib img.tonemap (gamma only; 'reinhard' needs cv2's tone mapper and is unused
    by the reference's hot-path callers)."""
    if method != 'gamma':
        raise ValueError(method)
    hdr = np.asarray(hdr)
    return np.clip((hdr / hdr.max()) ** (1 / gamma), 0, 1)


def rgb2lum(im):
    """xiuminglib img.rgb2lum."""
    assert im.shape[-1] == 3, "Input's last dimension must hold RGB"
    return 0.2126 * im[..., 0] + 0.7152 * im[..., 1] + 0.0722 * im[..., 2]


# ------------------------------------------------------------------------ integer <-> float
def normalize_uint(arr):
    """xiuminglib img.normalize_uint."""
    if arr.dtype not in (np.uint8, np.uint16):
        raise TypeError(arr.dtype)
    return arr.astype(float) / np.iinfo(arr.dtype).max


def denormalize_float(arr, uint_type='uint8'):
    """xiuminglib img.denormalize_float: truncation, like write_float."""
    arr = np.asarray(arr)
    if arr.dtype.kind != 'f':
        raise TypeError("Input must be float (is %s)" % arr.dtype)
    if (arr < 0).any() or (arr > 1).any():
        raise ValueError("Input image has pixels outside [0, 1]")
    if uint_type not in ('uint8', 'uint16'):
        raise TypeError(uint_type)
    return (arr * np.iinfo(uint_type).max).astype(uint_type)


# ----------------------------------------------------------------------------- file I/O
def read(path):
    """xiuminglib io.img.read / load (whatever Pillow reads) -> array."""
    from PIL import Image
    if path.endswith(('.exr', '.hdr')):
        raise ValueError("Use util.light.read_hdr for .hdr (EXR is not supported here)")
    with open(path, 'rb') as h:
        img = Image.open(h)
        img.load()
    return np.array(img)


def write_uint(arr_uint, outpath):
    """xiuminglib io.img.write_uint (:99-119): single-channel 3D arrays are tiled to RGB."""
    from PIL import Image
    if arr_uint.ndim == 3 and arr_uint.shape[2] == 1:
        arr_uint = np.dstack([arr_uint] * 3)
    os.makedirs(os.path.dirname(os.path.abspath(outpath)), exist_ok=True)
    with open(outpath, 'wb') as h:
        Image.fromarray(arr_uint).save(h, format=_pil_format(outpath))


def _pil_format(path):
    ext = os.path.splitext(path)[1].lower()
    return {'.png': 'PNG', '.jpg': 'JPEG', '.jpeg': 'JPEG', '.apng': 'PNG', '.gif': 'GIF'}.get(
        ext, 'PNG')


def write_arr(arr_0to1, outpath, img_dtype='uint8', clip=False):
    """xiuminglib io.img.write_float / write_arr (:122-154) -> the uint array written."""
    arr_0to1 = np.asarray(arr_0to1)
    if clip:
        arr_0to1 = np.clip(arr_0to1, 0, 1)
    elif arr_0to1.size and (arr_0to1.min() < 0 or arr_0to1.max() > 1):
        raise AssertionError("Input should be in [0, 1], or allow it to be clipped")
    img_arr = (arr_0to1 * np.iinfo(img_dtype).max).astype(img_dtype)
    write_uint(img_arr, outpath)
    return img_arr


# ------------------------------------------------------------------------------ resizing
def _new_hw(h, w, new_h, new_w):
    if new_h is None and new_w is None:
        raise ValueError("At least one of new height or width must be given")
    if new_h is None:
        new_h = int(h / w * new_w)
    elif new_w is None:
        new_w = int(w / h * new_h)
    return int(new_h), int(new_w)


def _aa_weights(in_size, out_size):
    """Row-stochastic [out, in] matrix of TensorFlow's ScaleAndTranslate spans for
    `tf.image.resize(method='bilinear', antialias=True)`: half-pixel centres, triangle kernel
    of radius 1 stretched by max(in/out, 1), weights renormalised per output sample."""
    inv_scale = in_size / out_size
    kscale = max(inv_scale, 1.)
    mat = np.zeros((out_size, in_size), np.float64)
    for x in range(out_size):
        sample = (x + 0.5) * inv_scale
        lo = int(np.ceil(sample - kscale - 0.5))
        hi = int(np.floor(sample + kscale - 0.5))
        lo, hi = min(max(lo, 0), in_size - 1), min(max(hi, 0), in_size - 1)
        src = np.arange(lo, hi + 1)
        wgt = np.maximum(0., 1. - np.abs((src + 0.5 - sample) / kscale))
        tot = wgt.sum()
        if abs(tot) >= 1000. * np.finfo(np.float32).tiny:
            wgt = wgt / tot
        mat[x, src] = wgt
    return mat


def resize(img, new_h=None, new_w=None):
    """util/img.py:98-137: `tf.image.resize(bilinear, antialias=True)` restated in numpy
    (separable; rows then columns, float32 intermediate like the TF kernel).  This is the
    resampling `Model._load_light` applies to every HDR probe (nerfactor.py:169-179), so it
    is on the relighting path.  Accepts [H,W] or [H,W,C]; torch tensors are returned as tensors.
    The original dtype is restored for arrays (util/img.py:133-137)."""
    is_tensor = isinstance(img, torch.Tensor)
    arr = img.detach().cpu().numpy() if is_tensor else np.asarray(img)
    h, w = arr.shape[:2]
    new_h, new_w = _new_hw(h, w, new_h, new_w)
    x = arr.astype(np.float32)
    squeeze = x.ndim == 2
    if squeeze:
        x = x[:, :, None]
    wy = _aa_weights(h, new_h).astype(np.float32)
    wx = _aa_weights(w, new_w).astype(np.float32)
    x = np.einsum('oh,hwc->owc', wy, x, optimize=True).astype(np.float32)
    x = np.einsum('pw,owc->opc', wx, x, optimize=True).astype(np.float32)
    if squeeze:
        x = x[:, :, 0]
    if is_tensor:
        return torch.from_numpy(x).to(img.device)
    return x.astype(arr.dtype)


def resize_cv2(arr, new_h=None, new_w=None):
    """xiuminglib img.resize(method='cv2') (:77-121): INTER_AREA when shrinking, INTER_LINEAR when
    enlarging; what the datasets use for buffers and RGBA images (nerf_shape.py:174-179,
    nerf.py:160-161).  cv2 takes at most 512 channels per call."""
    import cv2
    arr = np.asarray(arr)
    h, w = arr.shape[:2]
    new_h, new_w = _new_hw(h, w, new_h, new_w)
    interp = cv2.INTER_LINEAR if new_h > h else cv2.INTER_AREA
    if arr.ndim == 3 and arr.shape[2] > 512:
        parts = [cv2.resize(np.ascontiguousarray(arr[:, :, i:i + 512]), (new_w, new_h),
                            interpolation=interp) for i in range(0, arr.shape[2], 512)]
        parts = [p[:, :, None] if p.ndim == 2 else p for p in parts]
        return np.concatenate(parts, axis=2)
    return cv2.resize(arr, (new_w, new_h), interpolation=interp)


# ----------------------------------------------------------------------------- collages
def to_uint(tensor_0to1, target_type='uint8'):
    """util/img.py:166-178 (clips, then truncates)."""
    a = np.clip(np.asarray(tensor_0to1), 0, 1)
    return (np.iinfo(target_type).max * a).astype(target_type)


def frame_image(img, rgb=None, width=4):
    """util/img.py:226-243: paints a border IN PLACE (like the reference)."""
    kind = str(img.dtype)
    if kind.startswith('float'):
        dtype_max = 1.
    elif kind.startswith('uint'):
        dtype_max = np.iinfo(img.dtype).max
    else:
        raise NotImplementedError(kind)
    if rgb is None:
        rgb = (0, 0, 1)
    rgb = np.array(rgb, dtype=img.dtype) * dtype_max
    img[:width, :, :] = rgb
    img[-width:, :, :] = rgb
    img[:, :width, :] = rgb
    img[:, -width:, :] = rgb


def _concat(img_list, horizontal, out_size):
    total = []
    for img in img_list:
        if img.ndim == 2:
            img = np.dstack([img] * 3)
        if total:                      # match the previous tile's height (width)
            prev = total[-1]
            if horizontal and img.shape[0] != prev.shape[0]:
                img = resize(img, new_h=prev.shape[0])
            elif not horizontal and img.shape[1] != prev.shape[1]:
                img = resize(img, new_w=prev.shape[1])
        total.append(img)
    total = np.hstack(total) if horizontal else np.vstack(total)
    if out_size is not None:
        total = resize(total, new_w=out_size) if horizontal else resize(total, new_h=out_size)
    return total


def hconcat(img_list, out_w=None):
    """util/img.py:200-211."""
    return _concat(img_list, True, out_w)


def vconcat(img_list, out_h=None):
    """util/img.py:214-225."""
    return _concat(img_list, False, out_h)


# ------------------------------------------------------------------------------- metric
class PSNR:
    """xiuminglib metric.PSNR: PSNR in dB on the luma of two same-typed images."""

    def __init__(self, dtype='uint8'):
        self.dtype = np.dtype(dtype)
        self.drange = float(np.iinfo(self.dtype).max) if self.dtype.kind == 'u' else 1.

    def __call__(self, im1, im2, mask=None):
        for im in (im1, im2):
            assert im.dtype == self.dtype, "Input data type must be %s" % self.dtype
        im1, im2 = im1.astype(float), im2.astype(float)
        if im1.ndim == 2:
            im1, im2 = im1[:, :, None], im2[:, :, None]
        assert im1.shape == im2.shape, "The two images are not even of the same shape"
        if im1.shape[2] == 3:
            im1, im2 = rgb2lum(im1)[..., None], rgb2lum(im2)[..., None]
        if mask is None:
            mask = np.ones(im1.shape)
        elif mask.ndim == 2:
            mask = mask[:, :, None]
        mask = mask.astype(bool)
        mse = np.sum(np.square(im1[mask] - im2[mask])) / np.sum(mask)
        return 10 * np.log10((self.drange ** 2) / mse)
