"""Mirror of nerfactor/util/config.py:15-26 (flat-.ini helpers)."""
import posixpath


def config2dict(config):
    """`[DEFAULT]` section -> plain dict (util/config.py:15-22; the reference's configs have no
    other section)."""
    return {key: value for key, value in config.items('DEFAULT')}


def get_config_ini(ckpt_path):
    """The .ini written next to a run directory, from one of its checkpoint prefixes:
    `<outroot>/<xname>/checkpoints/ckpt-N` -> `<outroot>/<xname>.ini` (util/config.py:25-26)."""
    run_dir = posixpath.dirname(posixpath.dirname(ckpt_path))
    return run_dir + '.ini'
