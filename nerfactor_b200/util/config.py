"""Mirror of nerfactor/util/config.py:15-26."""


def config2dict(config):
    """Assumes the configuration .ini has only the default section (util/config.py:15-22)."""
    config_dict = {}
    for k, v in config.items('DEFAULT'):
        assert k not in config_dict, "Duplicate flags not allowed"
        config_dict[k] = v
    return config_dict


def get_config_ini(ckpt_path):
    """`<outroot>/<xname>/checkpoints/ckpt-N` -> `<outroot>/<xname>.ini` (util/config.py:25-26)."""
    return '/'.join(ckpt_path.split('/')[:-2]) + '.ini'
