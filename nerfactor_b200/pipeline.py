"""View-level public API: camera pose in, relit image out.

`ViewRenderer.render` chains what the reference runs as two programs:
  Stage A  geometry_from_nerf.process_view (geometry_from_nerf.py:93-149): rays ->
           sigma march -> occupancy / expected depth -> alpha-premultiplied surface
           points (xyz buffer), and
  Stage B  nerfactor Model.call (nerfactor.py:181-313) on the 9-tuple batch the
           dataset would have built from those buffers (nerf_shape.py:72-95).
Both stages stay separately callable / testable (nerfactor_b200.geometry_from_nerf,
nerfactor_b200.models.*)."""
import numpy as np
import torch

from . import _lib
from . import geometry_from_nerf as gfn


class ViewRenderer:
    def __init__(self, nerf_model, nerfactor_model, n_samples=128, use_fine=True,
                 occu_thres=0., scene_bbox=None):
        self.nerf = nerf_model
        self.model = nerfactor_model
        self.ctx = nerfactor_model.ctx
        self.n_samples = n_samples
        self.use_fine = use_fine
        self.occu_thres = occu_thres
        self.scene_bbox = scene_bbox

    def stage_a(self, c2w, cam_angle_x, h, w, ray_range=None):
        """-> dict(rayo, rayd, alpha [n,1], xyz [n,3] premultiplied, surf, occu, depth)."""
        ctx = self.ctx
        rayo, rayd = _lib.gen_rays(ctx, c2w, cam_angle_x, h, w, normalize=True)
        if ray_range is not None:                       # ray sharding (multi-GPU)
            a, b = ray_range
            rayo, rayd = rayo[a:b].contiguous(), rayd[a:b].contiguous()
        out = gfn.march_single_pass(self.nerf, rayo, rayd, self.n_samples,
                                    use_fine=self.use_fine, scene_bbox=self.scene_bbox)
        occu = out['occu']
        if self.occu_thres > 0:                         # gfn.py:123-125
            occu = torch.where(occu < self.occu_thres, torch.zeros_like(occu), occu)
        alpha = torch.clamp(occu, 0., 1.)[:, None]      # gfn.py:128-130
        xyz = out['surf'] * alpha                       # gfn.py:137 (alpha_blend, zero bg)
        return {'rayo': rayo, 'rayd': rayd, 'alpha': alpha.contiguous(),
                'xyz': xyz.contiguous(), 'surf': out['surf'], 'occu': out['occu'],
                'depth': out['depth']}

    def stage_b(self, a, relight_olat=False, relight_probes=False, fused=True, all_lights=False):
        """Stage B on Stage A's buffers.  `fused` (default): Model.render_rgb -- the per-point
        networks, then nf_stageB_fused_fwd (light visibility -> BRDF -> rendering equation in one
        call, no [N, L] tensor in HBM); the image-level outputs are the same as `Model.call`'s.
        OLAT relighting and fused=False go through Model.call (which also returns pred['lvis']).
        In the fused path the visibility network skips the lights facing away from the shading
        normal (zero weight in the renderer, nerfactor.py:329-330) unless `all_lights`."""
        n = a['xyz'].shape[0]
        zeros3 = torch.zeros((n, 3), device=self.ctx.device)
        batch = (None, None, a['rayo'], a['rayd'], zeros3, a['alpha'], a['xyz'], zeros3,
                 None)
        if fused and not relight_olat and hasattr(self.model, 'render_rgb'):
            return self.model.render_rgb(batch, relight_probes=relight_probes,
                                         all_lights=all_lights)
        pred, _, _, _ = self.model.call(batch, 'test', relight_olat=relight_olat,
                                        relight_probes=relight_probes)
        return pred

    def render(self, c2w, cam_angle_x, h, w, ray_range=None, relight_olat=False,
               relight_probes=False, fused=True, all_lights=False):
        a = self.stage_a(c2w, cam_angle_x, h, w, ray_range)
        pred = self.stage_b(a, relight_olat, relight_probes, fused, all_lights)
        pred['alpha'] = a['alpha']
        pred['xyz'] = a['xyz']
        return pred

    def render_to_host(self, c2w, cam_angle_x, h, w, light_host, out_rgb_host,
                       out_alpha_host, ray_range=None):
        """End-to-end call with HOST buffers: env-map from pinned host memory in,
        rgb / alpha into pinned host memory out (copies on the current stream)."""
        self.model._light = light_host.to(self.ctx.device, non_blocking=True)
        pred = self.render(c2w, cam_angle_x, h, w, ray_range)
        out_rgb_host.copy_(pred['rgb'], non_blocking=True)
        out_alpha_host.copy_(pred['alpha'], non_blocking=True)
        return pred


def shard_range(n_total, rank, world):
    """Contiguous ray range of `rank` (ceil split)."""
    per = (n_total + world - 1) // world
    a = min(n_total, rank * per)
    return a, min(n_total, a + per)


def gather_image(local_rgb, n_total, rank, world):
    """Assembles the full image from per-rank ray shards with one all_gather
    (north star: NCCL all-gather only to assemble the final image)."""
    import torch.distributed as dist
    per = (n_total + world - 1) // world
    pad = torch.zeros((per,) + tuple(local_rgb.shape[1:]), dtype=local_rgb.dtype,
                      device=local_rgb.device)
    pad[:local_rgb.shape[0]] = local_rgb
    out = torch.empty((world * per,) + tuple(local_rgb.shape[1:]), dtype=local_rgb.dtype,
                      device=local_rgb.device)
    dist.all_gather_into_tensor(out, pad)
    return out[:n_total]


class PeerImageGather:
    """Image all-gather by ONE-SIDED pushes over NVLink peer memory: every rank copies its image
    straight into its slot of every peer's symmetric buffer (`torch.distributed._symmetric_memory`:
    peer-mapped allocations; the copies are device-to-device memcpys on the copy engines).  No SM
    kernel competes with the persistent one-CTA-per-SM tensor kernels of the next step and no rank
    waits for another inside a step; `finish()` is the only rendezvous (a device-side barrier on
    the signal pads).  Two buffers alternate, so the pushes of step k + 1 never overwrite the image
    of step k while a consumer may still read it.  Measured at N = 2, 800 x 800 (tools/scale_probe.py):
    111.2 ms / step vs 111.5 (NCCL all_gather, async) vs 110.3-110.6 with no exchange at all."""

    def __init__(self, n_rows, row_shape, world, rank, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.world, self.rank, self.n = world, rank, n_rows
        self.row_shape = tuple(row_shape)
        self.row_elems = int(np.prod(self.row_shape)) if self.row_shape else 1
        self.buf = symm.empty((2 * world * n_rows,) + self.row_shape, dtype=torch.float32,
                              device=device)
        self.hdl = symm.rendezvous(self.buf, group or dist.group.WORLD)
        self.step = 0

    def push(self, local):
        """local [n_rows, *row_shape] (fp32, this rank's image) -> its slot on every rank."""
        slot = self.step & 1
        self.step += 1
        off = (slot * self.world * self.n + self.rank * self.n) * self.row_elems
        local = local.contiguous()
        for pr in range(self.world):
            dst = self.hdl.get_buffer(pr, (self.n,) + self.row_shape, torch.float32, off)
            dst.copy_(local, non_blocking=True)
        return slot

    def finish(self):
        """All pushes of all ranks have landed -> [world, n_rows, *row_shape] of the last step."""
        self.hdl.barrier()
        slot = (self.step - 1) & 1
        return self.buf[slot * self.world * self.n:(slot + 1) * self.world * self.n].reshape(
            (self.world, self.n) + self.row_shape)
