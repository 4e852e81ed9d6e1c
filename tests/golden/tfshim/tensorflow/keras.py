"""tf.keras pieces the reference models use: Model (only its attribute tracking is used,
models/base.py:25-27), layers.Dense / Activation, losses.MSE / MAE / Mean*Error."""
import numpy as np
import torch


def _act(name):
    if name is None or callable(name):
        return name
    return {'relu': torch.relu, 'sigmoid': torch.sigmoid,
            'softplus': torch.nn.functional.softplus, 'linear': None}[name]


class Layer:
    def __init__(self, *a, **k):
        self.trainable = True
        self.built = False

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class Activation(Layer):
    def __init__(self, activation):
        super().__init__()
        self.fn = _act(activation)

    def call(self, x):
        return self.fn(x)


class Dense(Layer):
    """y = activation(x @ kernel + bias); kernel [in, units] built on first call (glorot-uniform
    / zeros like Keras -- the golden generator overwrites both)."""

    def __init__(self, units, activation=None, **_):
        super().__init__()
        self.units = int(units)
        self.activation = activation if isinstance(activation, Layer) else _act(activation)
        self.kernel = self.bias = None

    def build(self, input_shape):
        fan_in = int(input_shape[-1])
        lim = np.sqrt(6.0 / (fan_in + self.units))
        self.kernel = (torch.rand((fan_in, self.units)) * 2 - 1) * lim
        self.bias = torch.zeros((self.units,))
        self.built = True

    def set_weights(self, weights):
        from . import _TRAINING
        self.kernel = torch.as_tensor(np.asarray(weights[0], np.float32)).clone()
        self.bias = torch.as_tensor(np.asarray(weights[1], np.float32)).clone()
        if _TRAINING[0] and self.trainable:
            self.kernel.requires_grad_(True)
            self.bias.requires_grad_(True)
        self.built = True

    def get_weights(self):
        return [self.kernel.detach().numpy(), self.bias.detach().numpy()]

    def call(self, x):
        if not self.built:
            self.build(x.shape)
        y = torch.matmul(x, self.kernel.to(x.dtype)) + self.bias.to(x.dtype)
        return self.activation(y) if self.activation is not None else y


class Sequential(Layer):
    def __init__(self, layers=None):
        super().__init__()
        self.layers = list(layers or [])

    def build(self, input_shape):
        shape = tuple(input_shape)
        for layer in self.layers:
            if hasattr(layer, 'build') and not layer.built:
                layer.build(shape)
            if hasattr(layer, 'units'):
                shape = shape[:-1] + (layer.units,)

    def call(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


class Model(Layer):
    def __init__(self, *a, **k):
        super().__init__()

    @property
    def trainable_variables(self):
        """Dense layers and Variables directly under `self` (what Keras tracks once
        models/base.py register_trainable has aliased the layers), in attribute order."""
        out = []
        for v in vars(self).values():
            if isinstance(v, Dense) and v.built and v.trainable:
                out += [v.kernel, v.bias]
            elif isinstance(v, torch.Tensor) and getattr(v, '_shim_variable', False) \
                    and v.requires_grad:
                out.append(v)
            elif isinstance(v, Layer) and not isinstance(v, (Dense, Model)):
                for vv in vars(v).values():              # e.g. networks/layers.py LatentCode._z
                    if isinstance(vv, torch.Tensor) and getattr(vv, '_shim_variable', False) \
                            and vv.requires_grad:
                        out.append(vv)
        return out


class _Layers:
    Layer, Dense, Activation = Layer, Dense, Activation

    @staticmethod
    def Lambda(f):
        layer = Layer()
        layer.call = f
        return layer


layers = _Layers()


class _Losses:
    @staticmethod
    def MSE(y_true, y_pred):
        return torch.mean((y_pred - y_true) ** 2, dim=-1)

    @staticmethod
    def MAE(y_true, y_pred):
        return torch.mean(torch.abs(y_pred - y_true), dim=-1)

    class MeanSquaredError:
        def __init__(self, reduction='auto'):
            self.reduction = reduction

        def __call__(self, y_true, y_pred, sample_weight=None):
            loss = torch.mean((y_pred - y_true) ** 2, dim=-1)
            if sample_weight is not None:
                loss = loss * sample_weight
            return loss if self.reduction == 'none' else loss.mean()

    class MeanAbsoluteError:
        def __init__(self, reduction='auto'):
            self.reduction = reduction

        def __call__(self, y_true, y_pred, sample_weight=None):
            loss = torch.mean(torch.abs(y_pred - y_true), dim=-1)
            if sample_weight is not None:
                loss = loss * sample_weight
            return loss if self.reduction == 'none' else loss.mean()


losses = _Losses()
