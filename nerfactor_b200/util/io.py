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
