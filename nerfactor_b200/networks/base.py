"""Mirror of nerfactor/networks/base.py:21-26."""


class Network:
    def __init__(self):
        self.layers = []

    def __call__(self, x):
        raise NotImplementedError
