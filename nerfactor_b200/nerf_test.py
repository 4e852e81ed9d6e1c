"""Mirror of nerfactor/nerf_test.py: novel-view rendering with a trained NeRF.

    python -m nerfactor_b200.nerf_test --ckpt <nerf run>/checkpoints/ckpt-N [--debug]

Every test camera goes through `Model.call(batch, 'test')` -- coarse and fine passes of the fused
tcgen05 NeRF kernel (trunk + bottleneck + view-dependent colour), compositing kernel -- and
`vis_batch`; the frames are compiled into `<run>/vis_test/ckpt-N.mp4`.  Views are split
round-robin over the ranks under torchrun."""
import argparse
import os
from os.path import basename, join

import torch

from . import datasets, models
from .util import config as configutil, io as ioutil


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--ckpt', default='/path/to/ckpt-100', help="path to checkpoint (prefix only)")
    ap.add_argument('--debug', action='store_true')
    ap.add_argument('--precision', default='f16', choices=['f16', 'bf16', 'fp32'])
    return ap.parse_args(argv)


def main(argv=None):
    FLAGS = parse_args(argv)
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    if torch.cuda.is_available():
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    config_ini = configutil.get_config_ini(FLAGS.ckpt)
    config = ioutil.read_config(config_ini)
    outroot = join(config_ini[:-4], 'vis_test', basename(FLAGS.ckpt))
    Dataset = datasets.get_dataset_class(config.get('DEFAULT', 'dataset'))
    dataset = Dataset(config, 'test', debug=FLAGS.debug)
    datapipe = dataset.build_pipeline(no_batch=True, no_shuffle=True)
    Model = models.get_model_class(config.get('DEFAULT', 'model'))
    model = Model(config, debug=FLAGS.debug, precision=FLAGS.precision)
    ioutil.restore_model(model, FLAGS.ckpt)
    for batch_i, batch in enumerate(datapipe):
        if batch_i % world == rank:
            with torch.no_grad():
                _, _, _, to_vis = model.call(batch, mode='test')
            model.vis_batch(to_vis, join(outroot, 'batch{i:09d}'.format(i=batch_i)), mode='test')
        if FLAGS.debug:
            break
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group('nccl' if torch.cuda.is_available() else 'gloo')
        dist.barrier()
    view_at = None
    if rank == 0:
        view_at = model.compile_batch_vis(ioutil.sortglob(outroot, 'batch?????????'), outroot,
                                          mode='test')
        print("Compilation available for viewing at\n\t%s" % view_at)
    return outroot, view_at


if __name__ == '__main__':
    main()
