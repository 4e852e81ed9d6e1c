"""Mirror of nerfactor/models/base.py:25-143 (the Keras-Model shell): `net` dict of
trainable networks, `register_trainable()` aliases, mode validation."""


class Model:
    def __init__(self, config, debug=False):
        self.config = config
        self.debug = debug
        self.net = {}
        self.trainable_registered = False

    def register_trainable(self):
        """base.py:81-104: exposes every Dense layer as self.net_<name>_layer<i>
        (the names TF checkpoints use)."""
        registered = []
        for net_name, net in self.net.items():
            attr_name = 'net_' + net_name
            assert attr_name.isidentifier()
            for layer_i, layer in enumerate(net.layers):
                if layer.trainable:
                    full = attr_name + '_layer%d' % layer_i
                    assert not hasattr(self, full), \
                        "Can't register `{}` because it is already an attribute".format(full)
                    setattr(self, full, layer)
                    registered.append(full)
        self.trainable_registered = True
        return registered

    @property
    def trainable_variables(self):
        out = []
        for net in self.net.values():
            for layer in net.layers:
                if layer.trainable and layer.built:
                    out += [layer.kernel, layer.bias]
        return out

    @staticmethod
    def _validate_mode(mode):
        if mode not in ('train', 'vali', 'test'):
            raise ValueError(mode)

    def call(self, batch, mode='train'):
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)

    def compute_loss(self, pred, gt, **kwargs):
        raise NotImplementedError
