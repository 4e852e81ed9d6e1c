"""Mirror of nerfactor/util/vis.py:27-114 (`make_frame`, `get_nearest_input`) and of the
xiuminglib visualisation helpers `vis_batch` / `compile_batch_vis` call
(third_party/xiuminglib/xiuminglib/vis/{text.py:14-59, anim.py:13-60, html.py, video.py:13-111}).
Host-side only: nothing here touches the GPU.
"""
import os

import numpy as np

from . import img as imgutil
from .io import read_json


def _font(font_ttf, font_size):
    from PIL import ImageFont
    if font_ttf is not None and os.path.exists(font_ttf):
        return ImageFont.truetype(font_ttf, font_size)
    try:
        return ImageFont.load_default(size=font_size)        # Pillow >= 10.1
    except TypeError:
        return ImageFont.load_default()


def put_text(img, text, label_top_left_xy=None, font_size=None, font_color=(1, 0, 0),
             font_ttf=None):
    """xiuminglib vis.text.put_text: uint image in, RGB uint image with the label out.  The
    reference's Open Sans .ttf ships with xiuminglib; without it Pillow's built-in font is used."""
    from PIL import Image, ImageDraw
    assert img.dtype.kind == 'u', "Input image must be `uint` (i.e., an actual image)"
    if font_size is None:
        font_size = int(0.1 * img.shape[0])
    if label_top_left_xy is None:
        label_top_left_xy = (int(0.1 * img.shape[1]), int(0.05 * img.shape[0]))
    dtype_max = np.iinfo(img.dtype).max
    color = tuple(int(x * dtype_max) for x in font_color)
    pil = Image.fromarray(img).convert('RGB')
    ImageDraw.Draw(pil).text(label_top_left_xy, text, fill=color,
                             font=_font(font_ttf, max(1, font_size)))
    return np.array(pil)


_put_text = put_text          # `make_frame` has a boolean argument of the same name


def make_anim(imgs, duration=1, outpath=None):
    """xiuminglib vis.anim.make_anim: list of uint arrays / paths -> .apng or .gif."""
    from PIL import Image
    assert outpath is not None
    if not outpath.endswith(('.apng', '.gif')):
        outpath += '.gif'
    os.makedirs(os.path.dirname(os.path.abspath(outpath)), exist_ok=True)
    frames = []
    for im in imgs:
        if isinstance(im, str):
            frames.append(Image.open(im))
            continue
        assert im.dtype.kind == 'u', "If image is provided as an array, it has to be `uint`"
        if im.ndim == 2 or (im.ndim == 3 and im.shape[2] == 1):
            im = np.dstack([im.reshape(im.shape[:2])] * 3)
        frames.append(Image.fromarray(im))
    with open(outpath, 'wb') as h:
        frames[0].save(h, format='PNG' if outpath.endswith('.apng') else 'GIF', save_all=True,
                       append_images=frames[1:], duration=duration * 1000, loop=0)


def make_video(imgs, fps=24, outpath=None):
    """xiuminglib vis.video.make_video: uint8 RGB frames -> .mp4 (cv2 + FFMPEG, 'mp4v').
    Frames of different sizes are resized to the first frame's size."""
    import cv2
    assert outpath is not None and imgs, "need frames and an output path"
    os.makedirs(os.path.dirname(os.path.abspath(outpath)), exist_ok=True)
    h, w = imgs[0].shape[:2]
    writer = cv2.VideoWriter(outpath, cv2.VideoWriter_fourcc(*'mp4v'), fps, (w, h))
    if not writer.isOpened():
        raise IOError("cannot open %s for writing" % outpath)
    for im in imgs:
        if im.ndim == 2:
            im = np.dstack([im] * 3)
        if im.shape[:2] != (h, w):
            im = cv2.resize(im, (w, h))
        writer.write(np.ascontiguousarray(im[:, :, 2::-1]))
    writer.release()


class HTML:
    """xiuminglib vis.html.HTML, reduced to what `_compile_into_webpage` uses: an optional header
    and one table whose cells are text or images with captions."""

    def __init__(self, title="Results", bgcolor='black', text_font='roboto', text_color='white'):
        self.title, self.bgcolor, self.font, self.color = title, bgcolor, text_font, text_color
        self.body = ''
        self.rows = []

    def add_header(self, text, level=1):
        self.body += '\n    <h%d>%s</h%d>\n' % (level, text, level)

    def add_table(self):
        return self

    def add_row(self, cells, types, captions=None):
        self.rows.append((cells, types, captions or [''] * len(cells)))

    def save(self, out_html):
        out_dir = os.path.dirname(os.path.abspath(out_html))
        os.makedirs(out_dir, exist_ok=True)
        s = ['<!DOCTYPE html>\n<html>\n<head>\n    <title>%s</title>\n</head>\n'
             '<body bgcolor="%s">\n<font face="%s" color="%s">\n'
             % (self.title, self.bgcolor, self.font, self.color), self.body,
             '<table width="100%" border="6">\n']
        for cells, types, caps in self.rows:
            s.append('  <tr>\n')
            for c, t, cap in zip(cells, types, caps):
                if t == 'image':
                    rel = os.path.relpath(c, out_dir)
                    cell = '<img src="%s" style="max-width:100%%"><br>%s' % (rel, cap)
                else:
                    cell = '%s<br>%s' % (c, cap)
                s.append('    <td align="center">%s</td>\n' % cell)
            s.append('  </tr>\n')
        s.append('</table>\n</font>\n</body>\n</html>\n')
        with open(out_html, 'w') as h:
            h.write(''.join(s))


def get_nearest_input(view_dir, data_root):
    """util/vis.py:108-114."""
    id_ = read_json(os.path.join(view_dir, 'metadata.json'))['id']
    return os.path.join(data_root, id_, 'nn.png')


_LABELS = {'normal': "Normals", 'normals': "Normals", 'lvis': "Visibility (mean)",
           'brdf': "BRDF", 'albedo': "Albedo"}


def make_frame(view_dir, layout, put_text=True, put_text_param=None, data_root=None,
               rgb_embed_light=None):
    """util/vis.py:27-105: collage of a view's `pred_<name>.png` files laid out as `layout`
    (1D or 2D list of names; 'nn' = the nearest input view).  Returns None when a file is
    missing (the caller skips the frame)."""
    param = dict(put_text_param or {})
    param.setdefault('text_loc_ratio', 0.05)
    param.setdefault('text_size_ratio', 0.05)
    param.setdefault('font_path', None)
    layout = np.array(layout)
    if layout.ndim == 1:
        layout = layout.reshape(1, -1)
    elif layout.ndim != 2:
        raise ValueError(layout.ndim)
    rows = []
    for row_names in layout:
        row = []
        for name in row_names:
            name = str(name)
            is_render, is_nn = name.startswith('rgb'), name == 'nn'
            if is_nn:
                assert data_root is not None, "When including NN, you must provide `data_root`"
                path = get_nearest_input(view_dir, data_root)
            else:
                path = os.path.join(view_dir, 'pred_%s.png' % name)
            if not os.path.exists(path):
                return None
            im = imgutil.read(path)
            if im.ndim == 2:
                im = np.dstack([im] * 3)
            im = np.ascontiguousarray(im[:, :, :3])
            hw = im.shape[:2]
            if is_render and rgb_embed_light is not None:
                light = rgb_embed_light
                imgutil.frame_image(light, rgb=(1, 1, 1),
                                    width=int(max(1 / 16 * light.shape[0], 1)))
                light = imgutil.resize_cv2(light, new_h=max(1, int(32 / 256 * hw[0])))
                im[:light.shape[0], -light.shape[1]:] = light
            if put_text:
                if is_nn:
                    label = "Nearest Input"
                elif is_render:
                    label = "Rendering"
                elif name.startswith('lvis_olat_'):
                    label = "Visibility"
                elif name in _LABELS:
                    label = _LABELS[name]
                else:
                    raise NotImplementedError(name)
                im = _put_text(
                    im, label,
                    label_top_left_xy=(int(param['text_loc_ratio'] * hw[1]),
                                       int(param['text_loc_ratio'] * hw[0])),
                    font_size=int(param['text_size_ratio'] * hw[0]),
                    font_color=(1, 1, 1) if is_render or is_nn else (0, 0, 0),
                    font_ttf=param['font_path'])
            row.append(im)
        rows.append(imgutil.hconcat(row))
    return imgutil.vconcat(rows)
