"""Mirror of nerfactor/networks/base.py:21-26: the common base of `mlp.Network` and `seq.Network`
-- an ordered list of layers (index order = checkpoint order, models/base.py:81-104) -- plus the
weight plumbing both subclasses share here (NumPy weights instead of Keras variables)."""


class Network:
    def __init__(self):
        self.layers = []

    # ---- weights (Keras layout: kernel [in, out], bias [out]) --------------------------------
    @property
    def built(self):
        return bool(self.layers) and all(layer.built for layer in self.layers)

    def weights(self):
        """[(kernel, bias), ...] in layer order."""
        if not self.built:
            raise AssertionError("Some layers not built")
        return [(layer.kernel, layer.bias) for layer in self.layers]

    def load(self, mlp_dict):
        """mlp_dict: {'layers': [(W, b), ...]} (nerfactor_b200.synth / checkpoints)."""
        if len(mlp_dict['layers']) != len(self.layers):
            raise ValueError("%d weight pairs for %d layers" % (len(mlp_dict['layers']), len(self.layers)))
        for layer, (w, b) in zip(self.layers, mlp_dict['layers']):
            layer.set_weights(w, b)
        return self

    def n_params(self):
        return sum(int(k.size) + int(b.size) for k, b in self.weights())

    def __call__(self, x):
        raise NotImplementedError

    def __repr__(self):
        dims = ' -> '.join(str(getattr(layer, 'units', '?')) for layer in self.layers)
        return '%s(%s)' % (type(self).__name__, dims)
