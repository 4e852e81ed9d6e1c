"""Mirror of nerfactor/explore_brdf_space.py: the trained BRDF prior evaluated on the shared test
coordinates for every seen material and every interpolated identity
(`<k>_<w1>_<mat1>_<w2>_<mat2>`, datasets/brdf_merl.py).

    python -m nerfactor_b200.explore_brdf_space --ckpt <prior run>/checkpoints/ckpt-N

Per identity `<run>/vis_test/ckpt-N/batch?????????/` gets `metadata.json`, `z.npy` and
`log10_brdf.npy` (the numbers behind the reference's z / log10-BRDF bar plots); the reference's
sphere render and MERL characteristic slice (matplotlib, brdf/merl) are not produced."""
import argparse
import os
from os.path import basename, exists, join

from . import datasets, models
from .util import config as configutil, io as ioutil


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split('\n')[0])
    ap.add_argument('--ckpt', default='/path/to/ckpt-100', help="path to checkpoint (prefix only)")
    ap.add_argument('--debug', action='store_true')
    return ap.parse_args(argv)


def main(argv=None):
    FLAGS = parse_args(argv)
    config_ini = configutil.get_config_ini(FLAGS.ckpt)
    config = ioutil.read_config(config_ini)
    outroot = join(config_ini[:-4], 'vis_test', basename(FLAGS.ckpt))
    Dataset = datasets.get_dataset_class(config.get('DEFAULT', 'dataset'))
    dataset = Dataset(config, 'test', debug=FLAGS.debug)
    datapipe = dataset.build_pipeline(no_batch=True, no_shuffle=True)
    Model = models.get_model_class(config.get('DEFAULT', 'model'))
    model = Model(config, debug=FLAGS.debug)
    ioutil.restore_model(model, FLAGS.ckpt)
    done = 0
    for batch_i, batch in enumerate(dataset.files):
        outdir = join(outroot, f'batch{batch_i:09d}')
        expects = [join(outdir, f) for f in ('metadata.json', 'z.npy', 'log10_brdf.npy')]
        if all(exists(x) for x in expects):           # explore_brdf_space.py:65-71
            continue
        element = dataset._process_example_postcache(*dataset._process_example_precache(batch))
        _, _, _, to_vis = model.call(element, mode='test')
        model.vis_batch(to_vis, outdir, mode='test')
        done += 1
        if FLAGS.debug:
            break
    return outroot, done


if __name__ == '__main__':
    main()
