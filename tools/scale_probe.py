"""Where does the per-step cost of N >= 2 come from?  (a flat ~4.4 ms at N = 2 / 4 / 8.)
Run under torchrun; every rank renders the benchmark view and the variants differ only in how the
image all-gather is issued.  Rank 0 prints one JSON object (ms per step, max over ranks):
   none        no collective at all (N independent replicas)
   async       all_gather_into_tensor(async_op=True), waited before the next gather (bench.py)
   sync        all_gather_into_tensor on the compute stream
   side_event  gather on a side stream after an event, joined at the end of the step
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    os.environ.setdefault('NCCL_DEBUG', 'WARN')
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from nerfactor_b200 import _lib, synth, config as nfconfig
    from nerfactor_b200.models.nerfactor_microfacet import Model
    from nerfactor_b200.models.nerf import Model as NerfModel
    from nerfactor_b200.pipeline import ViewRenderer
    ctx = _lib.Context(local)
    nerf = NerfModel(nfconfig.default_config('nerf'), params=synth.make_nerf_params(0), ctx=ctx,
                     precision='f16e')
    model = Model(nfconfig.default_config('nerfactor_microfacet'),
                  params=synth.make_stage_b_params(0, 'microfacet'), ctx=ctx, precision='f16')
    vr = ViewRenderer(nerf, model, n_samples=128)
    c2w = synth.look_at_c2w(4.0, 30.0, 30.0)
    n = 800 * 800
    gathered = torch.empty((world * n, 3), device=ctx.device)
    side = torch.cuda.Stream()
    pending = [None]

    def render():
        return vr.render(c2w, synth.CAM_ANGLE_X, 800, 800)['rgb'].contiguous()

    def v_none():
        render()

    def v_async():
        rgb = render()
        if pending[0] is not None:
            pending[0].wait()
        pending[0] = dist.all_gather_into_tensor(gathered, rgb, async_op=True)

    def v_sync():
        dist.all_gather_into_tensor(gathered, render())

    def v_side():
        rgb = render()
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(side):
            side.wait_event(ev)
            rgb.record_stream(side)
            dist.all_gather_into_tensor(gathered, rgb)
        if pending[0] is not None:
            torch.cuda.current_stream().wait_event(pending[0])
        done = torch.cuda.Event()
        done.record(side)
        pending[0] = done

    # one-sided pushes into the peers' symmetric buffers (copy engines, no rendezvous per step)
    symm_ok, symm_err = True, None
    try:
        import torch.distributed._symmetric_memory as symm
        sbuf = symm.empty((2 * world * n, 3), dtype=torch.float32, device=ctx.device)
        hdl = symm.rendezvous(sbuf, dist.group.WORLD)
        step_i = [0]
    except Exception as e:                       # API not available in this build
        symm_ok, symm_err = False, repr(e)

    def v_symm():
        rgb = render()
        slot = step_i[0] & 1
        step_i[0] += 1
        for pr in range(world):
            dst = hdl.get_buffer(pr, (n, 3), torch.float32, (slot * world * n + rank * n) * 3)
            dst.copy_(rgb, non_blocking=True)

    def drain_symm():
        hdl.barrier()

    def timed(fn, steps=6, warmup=3, drain=None):
        for _ in range(warmup):
            fn()
        if drain:
            drain()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if drain:
            drain()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / steps], device=ctx.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def drain_work():
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None

    def drain_event():
        if pending[0] is not None:
            torch.cuda.current_stream().wait_event(pending[0])
            pending[0] = None

    out = {'world': world, 'NCCL_MAX_CTAS': os.environ.get('NCCL_MAX_CTAS'),
           'none': timed(v_none), 'async': timed(v_async, drain=drain_work),
           'sync': timed(v_sync), 'side_event': timed(v_side, drain=drain_event),
           'none_again': timed(v_none)}
    if symm_ok:
        try:
            out['symm_push'] = timed(v_symm, drain=drain_symm)
            # every rank's image must be in every rank's buffer (slot of the last step)
            torch.cuda.synchronize()
            dist.barrier()
            last = (step_i[0] - 1) & 1
            mine = render()
            got = sbuf[last * world * n:(last + 1) * world * n].reshape(world, n, 3)
            out['symm_push_correct'] = bool(all(torch.equal(got[r], mine) for r in range(world)))
            out['none_third'] = timed(v_none)
        except Exception as e:
            out['symm_push_error'] = repr(e)
    else:
        out['symm_unavailable'] = symm_err
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
