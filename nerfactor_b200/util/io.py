"""Mirror of nerfactor/util/io.py:36-45 (`restore_model`) on top of the TensorFlow-free
checkpoint reader (`util/tfckpt.py`), plus `latest_checkpoint`."""
from . import tfckpt


def restore_model(model, ckpt_path, submodel=''):
    """util/io.py:36-45: registers the trainable layers and restores every variable the
    checkpoint has for them (`expect_partial`: optimizer slots and anything the model does not
    own are ignored).  `ckpt_path` is the checkpoint prefix (`.../checkpoints/ckpt-100`).
    Returns the set of parameter groups that were found."""
    if hasattr(model, 'register_trainable') and not getattr(model, 'trainable_registered', False):
        model.register_trainable()
    params = tfckpt.params_from_checkpoint(ckpt_path, submodel)
    own = {k: v for k, v in params.items() if k in model.net or k in ('light', 'z')}
    nets = {k: v for k, v in own.items() if k in model.net}
    for k, v in nets.items():
        have = [tuple(w.shape) for w, _ in model.net[k].weights()] if all(
            l.built for l in model.net[k].layers) else None
        got = [tuple(w.shape) for w, _ in v['layers']]
        if have is not None and have != got:
            raise ValueError("checkpoint %s: network %s has kernel shapes %s, the model expects %s"
                             % (ckpt_path, k, got, have))
    if hasattr(model, 'load_params'):
        model.load_params(own)
    else:
        for k, v in nets.items():
            model.net[k].load(v)
    if 'z' in own and hasattr(model, 'latent_code'):
        model.latent_code.z = own['z']
    return set(own)


latest_checkpoint = tfckpt.latest_checkpoint


# ---------------------------------------------------------------- nerfactor/util/io.py:29-120
def all_exist(path_dict):
    """util/io.py:29-33."""
    import os
    return all(os.path.exists(v) for v in path_dict.values())


def read_config(path):
    """util/io.py:48-52: the flat `[DEFAULT]` .ini files of nerfactor/config/."""
    from configparser import ConfigParser
    config = ConfigParser()
    with open(path, 'r') as h:
        config.read_file(h)
    return config


def write_config(config, path):
    """util/io.py:55-57."""
    import os
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, 'w') as h:
        config.write(h)


def prepare_outdir(outdir, overwrite=False, quiet=True):
    """util/io.py:60-74: wipe when `overwrite`, otherwise keep what is there."""
    import os
    from shutil import rmtree
    if os.path.isdir(outdir):
        if not overwrite:
            return
        rmtree(outdir)
    os.makedirs(outdir)


def read_json(path):
    """util/io.py:96-99."""
    import json
    with open(path, 'r') as h:
        return json.load(h)


def write_json(data, path):
    """util/io.py:102-108."""
    import json
    import os
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, 'w') as h:
        json.dump(data, h, indent=4, sort_keys=True)


def load_np(np_f):
    """util/io.py:111-120."""
    import numpy as np
    if np_f.endswith('.npy'):
        with open(np_f, 'rb') as h:
            return np.load(h)
    with open(np_f, 'rb') as h:
        return dict(np.load(h, allow_pickle=True))


def sortglob(directory, filename='*', ext=None):
    """xiuminglib os.sortglob (local paths): sorted glob of `filename` + each extension."""
    import os
    from glob import glob
    if ext is None:
        ext = ()
    elif isinstance(ext, str):
        ext = (ext,)
    if isinstance(filename, str):
        filename = (filename,)
    exts = [x if x.startswith('.') else '.' + x for x in ext]
    files = []
    for f in filename:
        if exts:
            for e in exts:
                files += glob(os.path.join(directory, f + e))
        else:
            files += glob(os.path.join(directory, f))
    return sorted(files)
