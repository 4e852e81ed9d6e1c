"""tf.train.Checkpoint stand-in: restoring is a no-op (the golden generator sets the weights)."""


class _Status:
    def expect_partial(self):
        return self

    def assert_consumed(self):
        return self


class Checkpoint:
    def __init__(self, **kwargs):
        self.objects = kwargs

    def restore(self, path):
        return _Status()
