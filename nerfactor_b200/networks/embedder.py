"""Mirror of nerfactor/networks/embedder.py:23-47.  The encoding itself is fused
into the network kernels; this class carries the spec (`n_freqs`, `out_dims`) and
offers a torch implementation for host-side inspection."""
import torch


class Embedder:
    def __init__(self, incl_input=True, in_dims=3, log2_max_freq=3, n_freqs=4,
                 log_sampling=True, periodic_func=None):
        if not (incl_input and log_sampling and log2_max_freq == n_freqs - 1
                and periodic_func in (None, 'sincos')):
            raise NotImplementedError(
                "only the configuration the reference models use is supported: "
                "incl_input, log sampling with log2_max_freq = n_freqs - 1, [sin, cos]")
        self.in_dims = in_dims
        self.n_freqs = n_freqs
        self.out_dims = in_dims * (1 + 2 * n_freqs)

    def __call__(self, x):
        out = [x]
        for k in range(self.n_freqs):
            out += [torch.sin(x * float(2 ** k)), torch.cos(x * float(2 ** k))]
        return torch.cat(out, -1)
