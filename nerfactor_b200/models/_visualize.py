"""`vis_batch` / `compile_batch_vis` of the reference models, host-side:

  shape      nerfactor/models/shape.py:279-401
  nerfactor  nerfactor/models/nerfactor.py:543-879  (also nerfactor_microfacet)
  nerf       nerfactor/models/nerf.py:338-480

Input is the `to_vis` dict `Model.call` returns (device tensors, full-N row-major); output is
the reference's per-view directory: `pred_*.png` / `gt_*.png`, `pred-vs-gt_*.apng`,
`metadata.json` (view id, PSNR), and per run an HTML table (validation) or an .mp4 (test).
Mixed into the model classes; nothing here touches the GPU beyond the device->host copy.
"""
from os.path import basename, dirname, exists, join

import numpy as np

from ..util import img as imgutil, io as ioutil, light as lightutil, vis as visutil


def _to_numpy(v):
    if v is None:
        return None
    if hasattr(v, 'detach'):
        return v.detach().cpu().numpy()
    return np.asarray(v)


def _pop_hw_id(data_dict):
    """`hw` / `id` are per-view values here; the reference's per-ray tiled copies
    (nerf_shape.py:79-81) are accepted too."""
    hw = data_dict.pop('hw')
    hw = _to_numpy(hw) if not isinstance(hw, tuple) else np.asarray(hw)
    hw = tuple(int(x) for x in (hw[0] if hw.ndim == 2 else hw))
    id_ = data_dict.pop('id')
    if not isinstance(id_, (str, bytes)):
        id_ = np.asarray(id_).reshape(-1)[0]
    if isinstance(id_, bytes):
        id_ = id_.decode()
    return hw, str(id_)


class VisMixin:
    put_text_param = {'text_loc_ratio': 0.05, 'text_size_ratio': 0.05, 'font_path': None}

    def _put_text_kwargs(self, hw):
        p = self.put_text_param
        return {'label_top_left_xy': (int(p['text_loc_ratio'] * hw[1]),
                                      int(p['text_loc_ratio'] * hw[0])),
                'font_size': int(p['text_size_ratio'] * hw[0]),
                'font_color': (0, 0, 0) if self.white_bg else (1, 1, 1),
                'font_ttf': p['font_path']}

    def _apng(self, img_dict, key, outdir, hw, first_label):
        if 'gt_' + key not in img_dict or 'pred_' + key not in img_dict:
            return                    # e.g. a batch without Stage-A visibility
        kw = self._put_text_kwargs(hw)
        im1 = visutil.put_text(img_dict['gt_' + key], first_label, **kw)
        im2 = visutil.put_text(img_dict['pred_' + key], "Prediction", **kw)
        visutil.make_anim((im1, im2), outpath=join(outdir, 'pred-vs-gt_%s.apng' % key))

    def _viewer_prefix(self):
        return self.config.get('DEFAULT', 'viewer_prefix', fallback='')

    def _html(self, rows, caps, types, out_html, header=None):
        assert rows, "No row"
        html = visutil.HTML(bgcolor='white' if self.white_bg else 'black',
                            text_color='black' if self.white_bg else 'white')
        if header:
            html.add_header(header)
        table = html.add_table()
        for r, rcaps, rtypes in zip(rows, caps, types):
            table.add_row(r, rtypes, captions=rcaps)
        html.save(out_html)


class ShapeVis(VisMixin):
    def vis_batch(self, data_dict, outdir, mode='train', dump_raw_to=None):
        """shape.py:279-353."""
        self._validate_mode(mode)
        if mode == 'train':           # random rays do not form an image
            return
        data_dict = dict(data_dict)
        hw, id_ = _pop_hw_id(data_dict)
        for k, v in list(data_dict.items()):
            v_ = _to_numpy(v)
            if k.endswith('normal'):
                v_ = v_.reshape(hw + (3,))
            elif k.endswith(('occu', 'alpha')):
                v_ = v_.reshape(hw)
            elif k.endswith('lvis'):
                v_ = v_.reshape(hw + (v_.shape[1],))
            else:
                raise NotImplementedError(k)
            data_dict[k] = v_
        img_dict = {}
        alpha = data_dict['gt_alpha']
        for k, v in data_dict.items():
            if k.endswith('normal'):
                v = (v + 1) / 2
            elif k.endswith('lvis'):
                v = np.mean(v, axis=2)            # average across all lights
            elif k.endswith(('occu', 'alpha')):
                img_dict[k] = imgutil.write_arr(v, join(outdir, k + '.png'), clip=True)
                continue
            bg = np.ones_like(v) if self.white_bg else np.zeros_like(v)
            img_dict[k] = imgutil.write_arr(imgutil.alpha_blend(v, alpha, bg),
                                            join(outdir, k + '.png'), clip=True)
        if mode == 'test':
            ioutil.write_json({'id': id_}, join(outdir, 'metadata.json'))
            return
        self._apng(img_dict, 'normal', outdir, hw, "Initial")
        self._apng(img_dict, 'lvis', outdir, hw, "Initial")
        ioutil.write_json({'id': id_}, join(outdir, 'metadata.json'))

    def compile_batch_vis(self, batch_vis_dirs, outpref, mode='train'):
        """shape.py:355-369."""
        self._validate_mode(mode)
        if mode == 'train':
            return None
        if mode != 'vali':
            raise NotImplementedError(mode)
        outpath = outpref + '.html'
        self._compile_into_webpage(batch_vis_dirs, outpath)
        return self._viewer_prefix() + outpath

    def _compile_into_webpage(self, batch_dirs, out_html):
        """shape.py:371-401."""
        rows, caps, types = [], [], []
        for batch_dir in batch_dirs:
            metadata = str(ioutil.read_json(join(batch_dir, 'metadata.json')))
            rows.append([metadata, join(batch_dir, 'pred-vs-gt_normal.apng'),
                         join(batch_dir, 'pred-vs-gt_lvis.apng')])
            caps.append(["Metadata", "Normal", "Light Visibility"])
            types.append(['text', 'image', 'image'])
        self._html(rows, caps, types, out_html,
                   header="Refining and Caching Geometry Initialization")


class NeRFactorVis(VisMixin):
    # ---- lighting thumbnails (nerfactor.py:93-104), made on first use
    @property
    def embed_light_h(self):
        return self.config.getint('DEFAULT', 'embed_light_h', fallback=32)

    @property
    def novel_probes_uint(self):
        cache = self.__dict__.setdefault('_probes_uint', {})
        for k, v in self.novel_probes.items():
            if k not in cache:
                cache[k] = lightutil.vis_light(v, h=self.embed_light_h)
        return cache

    @property
    def novel_olat_uint(self):
        return _LazyOlatUint(self)

    @property
    def psnr(self):
        return imgutil.PSNR('uint8')

    def _brdf_prop_as_img(self, brdf_prop):
        """nerfactor.py:543-560: z clipped to the range of the prior's latent codes."""
        seen_z = np.asarray(self.brdf_model.latent_code.z)[:, :3]
        min_, max_ = seen_z.min(), seen_z.max()
        range_ = max_ - min_
        assert range_ > 0, "Range of seen BRDF Zs is 0"
        return (np.clip(brdf_prop[:, :, :3], min_, max_) - min_) / range_

    def vis_batch(self, data_dict, outdir, mode='train', dump_raw_to=None, light_vis_h=256,
                  olat_vis=False, alpha_thres=0.8):
        """nerfactor.py:562-739."""
        if mode == 'vali':            # the estimated light, once per epoch directory
            light_vis_path = join(dirname(outdir), 'pred_light.png')
            if not exists(light_vis_path):
                lightutil.vis_light(self.light, outpath=light_vis_path, h=light_vis_h)
        self._validate_mode(mode)
        if mode == 'train':
            return
        data_dict = dict(data_dict)
        hw, id_ = _pop_hw_id(data_dict)
        for k, v in list(data_dict.items()):
            if v is None:
                continue
            v_ = _to_numpy(v)
            if k in ('pred_rgb_olat', 'pred_rgb_probes'):
                v_ = v_.reshape(hw + (v_.shape[1], 3))
            elif k.endswith(('rgb', 'albedo', 'normal')):
                v_ = v_.reshape(hw + (3,))
            elif k.endswith(('occu', 'depth', 'disp', 'alpha')):
                v_ = v_.reshape(hw)
            elif k.endswith('brdf'):
                v_ = v_.reshape(hw + (-1,))
            elif k.endswith('lvis'):
                v_ = v_.reshape(hw + (v_.shape[1],))
            else:
                raise NotImplementedError(k)
            data_dict[k] = v_
        alpha = np.array(data_dict['gt_alpha'])
        alpha[alpha < alpha_thres] = 0            # stricter compositing
        lareas = _to_numpy(self.lareas).reshape(self.light_res)

        def composite_on_avg_light(render, light_uint):
            """nerfactor.py:600-616: background = area-weighted mean of the upper hemisphere."""
            lareas_upper = lareas[:(lareas.shape[0] // 2), :]
            weights = np.dstack([lareas_upper] * 3)
            light = imgutil.normalize_uint(light_uint)
            light = imgutil.resize_cv2(light, new_h=lareas.shape[0])
            light_upper = light[:(light.shape[0] // 2), :, :]
            avg_light = np.average(light_upper, axis=(0, 1), weights=weights)
            bg = np.tile(avg_light[None, None, :], render.shape[:2] + (1,))
            return imgutil.alpha_blend(render, alpha, bg)

        def blend_write(v, key):
            bg = np.ones_like(v) if self.white_bg else np.zeros_like(v)
            return imgutil.write_arr(imgutil.alpha_blend(v, alpha, bg),
                                     join(outdir, key + '.png'), clip=True)

        img_dict = {}
        for k, v in data_dict.items():
            if v is None:
                continue
            if k == 'pred_rgb_olat':              # H x W x L x 3, top half of the sphere only
                names = list(self.novel_olat.keys())
                for i, lname in enumerate(names[:int(np.prod(self.light_res)) // 2]):
                    img = composite_on_avg_light(v[:, :, i, :], self.novel_olat_uint[lname])
                    img_dict[k + '_' + lname] = imgutil.write_arr(
                        img, join(outdir, k + '_' + lname + '.png'), clip=True)
            elif k == 'pred_rgb_probes':
                for i, lname in enumerate(self.novel_probes.keys()):
                    img = composite_on_avg_light(v[:, :, i, :], self.novel_probes_uint[lname])
                    img_dict[k + '_' + lname] = imgutil.write_arr(
                        img, join(outdir, k + '_' + lname + '.png'), clip=True)
            elif k.endswith('rgb'):
                img_dict[k] = blend_write(v, k)
            elif k.endswith('normal'):
                img_dict[k] = blend_write((v + 1) / 2, k)
            elif k.endswith('albedo'):
                img_dict[k] = blend_write(v ** (1 / 2.2), k)
            elif k.endswith('lvis'):
                img_dict[k] = blend_write(np.mean(v, axis=2), k)
                if olat_vis:                       # per-light visibility, first half
                    for i in range(4 if self.debug else v.shape[2] // 2):
                        ij = np.unravel_index(i, self.light_res)
                        k_olat = k + '_olat_%04d-%04d' % ij
                        img_dict[k_olat] = blend_write(v[:, :, i], k_olat)
            elif k.endswith('brdf'):
                img_dict[k] = blend_write(self._brdf_prop_as_img(v), k)
            else:
                img_dict[k] = imgutil.write_arr(v, join(outdir, k + '.png'), clip=True)
        if mode == 'test':
            ioutil.write_json({'id': id_}, join(outdir, 'metadata.json'))
            return
        self._apng(img_dict, 'rgb', outdir, hw, "Ground Truth")
        if self.shape_mode != 'nerf':
            self._apng(img_dict, 'normal', outdir, hw, "Initial")
            self._apng(img_dict, 'lvis', outdir, hw, "Initial")
        psnr = float(self.psnr(img_dict['gt_rgb'], img_dict['pred_rgb']))
        ioutil.write_json({'id': id_, 'psnr': psnr}, join(outdir, 'metadata.json'))

    def compile_batch_vis(self, batch_vis_dirs, outpref, mode='train', fps=12):
        """nerfactor.py:741-755."""
        self._validate_mode(mode)
        if mode == 'train':
            return None
        if mode == 'vali':
            outpath = outpref + '.html'
            self._compile_into_webpage(batch_vis_dirs, outpath)
        else:
            outpath = outpref + '.mp4'
            self._compile_into_video(batch_vis_dirs, outpath, fps=fps)
        return self._viewer_prefix() + outpath

    def _compile_into_webpage(self, batch_dirs, out_html):
        """nerfactor.py:757-807."""
        rows, caps, types = [], [], []
        nerf = self.shape_mode == 'nerf'
        for batch_dir in batch_dirs:
            metadata = str(ioutil.read_json(join(batch_dir, 'metadata.json')))
            row = [metadata, join(batch_dir, 'pred-vs-gt_rgb.apng'),
                   join(batch_dir, 'pred_rgb.png'), join(batch_dir, 'pred_albedo.png'),
                   join(batch_dir, 'pred_brdf.png')]
            rowcaps = ["Metadata", "RGB", "RGB (pred.)", "Albedo (pred.)", "BRDF (pred.)"]
            for key, cap in (('normal', "Normal"), ('lvis', "Light Visibility")):
                if nerf:
                    row.append(join(batch_dir, 'gt_%s.png' % key))
                    rowcaps.append(cap + " (initial)")
                else:
                    row += [join(batch_dir, 'pred-vs-gt_%s.apng' % key),
                            join(batch_dir, 'pred_%s.png' % key)]
                    rowcaps += [cap, cap + " (pred.)"]
            rows.append(row)
            caps.append(rowcaps)
            types.append(['text'] + ['image'] * (len(row) - 1))
        self._html(rows, caps, types, out_html)

    def _compile_into_video(self, batch_dirs, out_mp4, fps=12):
        """nerfactor.py:809-879: view synthesis, OLAT sweep on the final view, then a view
        round trip under the light probes."""
        data_root = self.config.get(
            'DEFAULT', 'mvs_root' if self.config.get('DEFAULT', 'dataset', fallback='') ==
            'mvs_shape' else 'data_root', fallback=None)
        batch_dirs = sorted(batch_dirs)
        if self.debug:
            batch_dirs = batch_dirs[:10]
        have_nn = data_root is not None and all(
            exists(visutil.get_nearest_input(d, data_root)) for d in batch_dirs)
        orig_light_uint = lightutil.vis_light(self.light, h=self.embed_light_h)
        frames = []

        def add(view_dir, lvis_name, rgb_name, light_uint):
            if have_nn:                     # the reference's 2 x 3 collage
                layout = (('normal', lvis_name, 'nn'), ('brdf', 'albedo', rgb_name))
            else:                           # no nearest-input image on disk: one row
                layout = ('normal', lvis_name, 'brdf', 'albedo', rgb_name)
            frame = visutil.make_frame(
                view_dir, layout, data_root=data_root, put_text_param=self.put_text_param,
                rgb_embed_light=np.array(light_uint))
            if frame is not None:
                frames.append(frame)

        for batch_dir in batch_dirs:
            add(batch_dir, 'lvis', 'rgb', orig_light_uint)
        relight_view_dir = batch_dirs[-1]
        for lvis_path in ioutil.sortglob(relight_view_dir, 'pred_lvis_olat*', ext='png'):
            olat_id = basename(lvis_path)[len('pred_lvis_olat_'):-len('.png')]
            add(relight_view_dir, 'lvis_olat_%s' % olat_id, 'rgb_olat_%s' % olat_id,
                self.novel_olat_uint[olat_id])
        envmap_names = list(self.novel_probes.keys())
        if envmap_names:
            roundtrip = list(reversed(batch_dirs)) + batch_dirs
            roundtrip += roundtrip
            per_map = len(roundtrip) / len(envmap_names)
            map_i = 0
            for view_i, batch_dir in enumerate(roundtrip):
                name = envmap_names[min(map_i, len(envmap_names) - 1)]
                add(batch_dir, 'lvis', 'rgb_probes_%s' % name, self.novel_probes_uint[name])
                if (view_i + 1) > per_map * (map_i + 1):
                    map_i += 1
        assert frames, "no frame to compile (missing pred_*.png files?)"
        visutil.make_video(frames, outpath=out_mp4, fps=fps)


class _LazyOlatUint:
    """name -> tonemapped thumbnail of the OLAT env-map (nerfactor.py:96-99), on demand."""

    def __init__(self, model):
        self.m, self.cache = model, {}

    def __getitem__(self, key):
        if key not in self.cache:
            self.cache[key] = lightutil.vis_light(self.m.novel_olat[key],
                                                  h=self.m.embed_light_h)
        return self.cache[key]


class NerfVis(VisMixin):
    def vis_batch(self, data_dict, outdir, mode='train', dump_raw_to=None, text_loc_ratio=0.05,
                  text_size_ratio=0.05):
        """nerf.py:302-395: gt_rgb, {coarse,fine}_{rgb,occu,depth,disp}."""
        self._validate_mode(mode)
        if mode == 'train':
            return
        data_dict = dict(data_dict)
        hw, id_ = _pop_hw_id(data_dict)
        for k, v in list(data_dict.items()):
            if v is None:
                del data_dict[k]
            elif k.endswith('rgb'):
                data_dict[k] = _to_numpy(v).reshape(hw + (3,))
            elif k.endswith(('occu', 'depth', 'disp')):
                data_dict[k] = _to_numpy(v).reshape(hw)
            else:
                raise NotImplementedError(k)
        img_dict = {}
        for k, v in data_dict.items():
            if k.endswith(('depth', 'disp')):
                if k.endswith('depth'):
                    img = (v - self.near) / (self.far - self.near)
                else:
                    min_disp, max_disp = 1 / self.far, 1 / self.near
                    img = (v - min_disp) / (max_disp - min_disp)
                alpha = data_dict[k.replace('depth', 'occu').replace('disp', 'occu')]
                bg = np.ones_like(img) if self.white_bg else np.zeros_like(img)
                img = imgutil.alpha_blend(img, alpha, bg)
            elif k in ('coarse_occu', 'fine_occu'):
                img = 1 - v if self.white_bg else v
            else:                                  # RGB: already composited onto the background
                img = v
            img_dict[k] = imgutil.write_arr(img, join(outdir, k + '.png'), clip=True)
        if mode == 'test':
            ioutil.write_json({'id': id_}, join(outdir, 'metadata.json'))
            return
        kw = self._put_text_kwargs(hw)

        def anim(a, la, b, lb, name):
            if a in img_dict and b in img_dict:
                visutil.make_anim((visutil.put_text(img_dict[a], la, **kw),
                                   visutil.put_text(img_dict[b], lb, **kw)),
                                  outpath=join(outdir, name))

        anim('gt_rgb', "Ground Truth", 'fine_rgb', "Prediction (fine)", 'fine-vs-gt_rgb.apng')
        for key in ('rgb', 'depth', 'disp', 'occu'):
            anim('fine_' + key, "Prediction (fine)", 'coarse_' + key, "Prediction (coarse)",
                 'fine-vs-coarse_%s.apng' % key)
        psnr = float(imgutil.PSNR('uint8')(img_dict['gt_rgb'], img_dict['fine_rgb']))
        ioutil.write_json({'id': id_, 'psnr': psnr}, join(outdir, 'metadata.json'))

    def compile_batch_vis(self, batch_vis_dirs, outpref, mode='train', fps=12):
        """nerf.py:397-480."""
        self._validate_mode(mode)
        if mode == 'train':
            return None
        if mode == 'vali':
            outpath = outpref + '.html'
            rows, caps, types = [], [], []
            for batch_dir in batch_vis_dirs:
                metadata = str(ioutil.read_json(join(batch_dir, 'metadata.json')))
                rows.append([metadata, join(batch_dir, 'fine-vs-gt_rgb.apng')] + [
                    join(batch_dir, 'fine-vs-coarse_%s.apng' % k)
                    for k in ('rgb', 'depth', 'disp', 'occu')])
                caps.append(["Metadata", "RGB", "RGB", "Depth", "Disparity", "Occupancy"])
                types.append(['text'] + ['image'] * 5)
            self._html(rows, caps, types, outpath, header="NeRF")
        else:
            outpath = outpref + '.mp4'
            data_root = self.config.get('DEFAULT', 'data_root', fallback='')
            frames = {}
            for batch_dir in batch_vis_dirs:
                json_path, pred_path = join(batch_dir, 'metadata.json'), join(batch_dir, 'fine_rgb.png')
                if not exists(json_path) or not exists(pred_path):
                    continue
                id_ = ioutil.read_json(json_path)['id']
                frame = imgutil.read(pred_path)
                nn_paths = ioutil.sortglob(join(data_root, 'test_phys_nn'), id_ + '_nn_*', ext='png')
                if len(nn_paths) == 1:
                    frame = imgutil.hconcat((frame, imgutil.read(nn_paths[0])))
                elif len(nn_paths) > 1:
                    raise RuntimeError(
                        "There must be either zero or one nearest neighbor for each test "
                        "camera, but found %d" % len(nn_paths))
                frames[id_] = frame
            visutil.make_video([frames[k] for k in sorted(frames)], fps=fps, outpath=outpath)
        return self._viewer_prefix() + outpath
