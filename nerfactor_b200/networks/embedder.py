"""Mirror of nerfactor/networks/embedder.py:23-47.  Inside the models the encoding is fused
into the network kernels (which support the configuration the reference models use: include the
input, log-spaced octaves 2^0 .. 2^(F-1), [sin, cos]: `fused_ok`); called directly it evaluates
any configuration of the reference class with torch ops on the tensor's device."""
import torch

_FUNCS = {'sin': torch.sin, 'cos': torch.cos}


class Embedder:
    def __init__(self, incl_input=True, in_dims=3, log2_max_freq=3, n_freqs=4,
                 log_sampling=True, periodic_func=None):
        if periodic_func in (None, 'sincos'):
            periodic_func = ['sin', 'cos']
        names = []
        for f in periodic_func:
            name = f if isinstance(f, str) else getattr(f, '__name__', None)
            if name not in _FUNCS:
                raise NotImplementedError("periodic function %r (sin / cos only)" % (f,))
            names.append(name)
        self.incl_input = bool(incl_input)
        self.in_dims = in_dims
        self.n_freqs = n_freqs
        self.log2_max_freq = log2_max_freq
        self.log_sampling = bool(log_sampling)
        self.periodic_func = names
        if n_freqs <= 0:
            self.freq_bands = []
        elif log_sampling:          # 2 ** linspace(0, log2_max_freq, n_freqs)   (embedder.py:33-34)
            self.freq_bands = [2. ** (log2_max_freq * k / max(n_freqs - 1, 1)) for k in range(n_freqs)]
        else:                       # linspace(2 ** 0, 2 ** log2_max_freq, n_freqs) (:35-37)
            hi = 2. ** log2_max_freq
            self.freq_bands = [1. + (hi - 1.) * k / max(n_freqs - 1, 1) for k in range(n_freqs)]
        self.out_dims = in_dims * ((1 if incl_input else 0) + len(names) * n_freqs)

    @property
    def fused_ok(self):
        """True for the configuration the fused kernels implement."""
        return (self.incl_input and self.log_sampling and self.periodic_func == ['sin', 'cos']
                and (self.n_freqs <= 1 or self.log2_max_freq == self.n_freqs - 1))

    def __call__(self, x):
        out = [x] if self.incl_input else []
        for freq in self.freq_bands:
            for name in self.periodic_func:
                out.append(_FUNCS[name](x * float(freq)))
        return torch.cat(out, -1)
