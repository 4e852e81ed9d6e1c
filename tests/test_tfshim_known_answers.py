"""Known answers for the TensorFlow shim (tests/golden/tfshim) that carries the reference's model
code when the golden fixtures are generated.  Each case is the documented behaviour of the TF op
(the examples of the TF 2.x API reference, or the op's defining formula) -- the list of semantics
the reference-code fixtures rest on.  Pure CPU; runs everywhere."""
import os
import sys

import numpy as np
import pytest
import torch

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'tfshim')


@pytest.fixture(scope='module')
def tf():
    had = sys.modules.get('tensorflow')
    sys.path.insert(0, SHIM)
    try:
        if had is not None and not getattr(had, '__version__', '').endswith('shim'):
            pytest.skip('a real tensorflow is importable')
        import tensorflow as tf_
        assert tf_.__version__.endswith('shim')
        yield tf_
    finally:
        sys.path.remove(SHIM)


def _np(x):
    return x.detach().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def test_elementwise_and_normalisation(tf):
    x = tf.constant([[3., 4.], [0., 0.]])
    # l2_normalize: x * rsqrt(max(sum(x^2), epsilon)) -- epsilon bounds the SQUARED norm
    assert np.allclose(_np(tf.linalg.l2_normalize(x, axis=1)), [[.6, .8], [0., 0.]])
    assert np.allclose(_np(tf.math.l2_normalize(tf.constant([[3., 4.]]), axis=1, epsilon=100.)),
                       [[.3, .4]])
    assert np.allclose(_np(tf.math.divide_no_nan(tf.constant([3., 3.]), tf.constant([2., 0.]))),
                       [1.5, 0.])
    assert np.allclose(_np(tf.math.floormod(tf.constant([-5., 5., -0.5]), 3.)), [1., 2., 2.5])
    assert np.allclose(_np(tf.clip_by_value(tf.constant([-1., .5, 7.]), 0., np.inf)), [0., .5, 7.])
    assert np.allclose(_np(tf.nn.softplus(tf.constant([0.]))), [np.log(2.)])
    assert np.allclose(_np(tf.nn.relu(tf.constant([-1., 2.]))), [0., 2.])


def test_scans_sorting_sampling(tf):
    x = tf.constant([[2., 3., 4.]])
    assert np.allclose(_np(tf.math.cumprod(x, axis=-1, exclusive=True)), [[1., 2., 6.]])
    assert np.allclose(_np(tf.math.cumprod(x, axis=-1)), [[2., 6., 24.]])
    assert np.allclose(_np(tf.cumsum(x, -1)), [[2., 5., 9.]])
    seq = tf.constant([[1., 2., 3., 3., 5.]])
    assert _np(tf.searchsorted(seq, tf.constant([[3., 0., 6.]]), side='right')).tolist() == [[4, 0, 5]]
    assert _np(tf.searchsorted(seq, tf.constant([[3.]]), side='left')).tolist() == [[2]]
    assert np.allclose(_np(tf.sort(tf.constant([[3., 1., 2.]]), -1)), [[1., 2., 3.]])
    assert np.allclose(_np(tf.roll(tf.constant([0, 1, 2, 3, 4]), shift=2, axis=0)), [3, 4, 0, 1, 2])
    ls = _np(tf.linspace(10., 12., 3))
    assert ls.dtype == np.float32 and ls.tolist() == [10., 11., 12.]
    assert _np(tf.linspace(0., 1., 128))[-1] == 1.0 and len(_np(tf.linspace(0., 9., 10))) == 10
    assert np.allclose(_np(2. ** tf.linspace(0., 9., 10)), 2. ** np.arange(10))


def test_gather_scatter_mask(tf):
    # tf.scatter_nd API example
    out = tf.scatter_nd(tf.constant([[4], [3], [1], [7]]), tf.constant([9., 10., 11., 12.]), [8])
    assert _np(out).tolist() == [0., 11., 0., 10., 9., 0., 0., 12.]
    dup = tf.scatter_nd(tf.constant([[1], [1]]), tf.constant([2., 3.]), [3])    # duplicates add
    assert _np(dup).tolist() == [0., 5., 0.]
    upd = tf.tensor_scatter_nd_update(tf.zeros((3, 2)), tf.constant([[2], [0]]),
                                      tf.constant([[1., 1.], [7., 8.]]))
    assert _np(upd).tolist() == [[7., 8.], [0., 0.], [1., 1.]]
    p = tf.constant([[1., 2.], [3., 4.]])
    assert _np(tf.gather_nd(p, tf.constant([[0, 0], [1, 1]]))).tolist() == [1., 4.]
    assert _np(tf.gather_nd(p, tf.constant([[1], [0]]))).tolist() == [[3., 4.], [1., 2.]]
    m = tf.boolean_mask(tf.constant([[1, 2], [3, 4], [5, 6]]), tf.constant([True, False, True]))
    assert _np(m).tolist() == [[1, 2], [5, 6]]
    # the batched gather of util/math.py:86-87
    cdf = tf.constant([[0., .5, 1.], [0., .2, 1.]])
    ind = tf.constant([[[0, 1], [1, 2]], [[0, 0], [2, 2]]])
    gg = tf.gather(cdf, ind, axis=-1, batch_dims=1)
    assert _np(gg).tolist() == [[[0., .5], [.5, 1.]], [[0., 0.], [1., 1.]]]
    w = tf.where(tf.constant([True, False, True]))
    assert _np(w).tolist() == [[0], [2]]
    assert _np(tf.where(tf.constant([True, False]), tf.constant([1., 1.]),
                        tf.constant([5., 5.]))).tolist() == [1., 5.]


def test_shapes_and_reductions(tf):
    x = tf.reshape(tf.range(6), (2, 3))
    assert _np(tf.shape(x)).tolist() == [2, 3] and tf.shape(x)[0].dtype == tf.int32
    assert _np(tf.tile(tf.constant([[1, 2]]), (2, 2))).tolist() == [[1, 2, 1, 2], [1, 2, 1, 2]]
    assert _np(tf.concat((x, x), -1)).shape == (2, 6) and _np(tf.stack((x, x), 0)).shape == (2, 2, 3)
    assert _np(tf.expand_dims(x, 0)).shape == (1, 2, 3)
    xf = tf.cast(x, tf.float32)
    assert float(tf.reduce_sum(xf)) == 15. and _np(tf.reduce_sum(xf, axis=1)).tolist() == [3., 12.]
    assert _np(tf.reduce_mean(xf, axis=0)).tolist() == [1.5, 2.5, 3.5]
    assert _np(tf.reduce_mean(xf, axis=())).shape == (2, 3)          # no axis -> no reduction
    assert np.allclose(_np(tf.einsum('ijk,ik->ij', tf.ones((2, 3, 4)), tf.ones((2, 4)))), 4.)
    assert _np(tf.broadcast_to(tf.constant([1., 2.]), (3, 2))).shape == (3, 2)
    a, b = tf.meshgrid(tf.range(2), tf.range(3), indexing='ij')
    assert _np(a).tolist() == [[0, 0, 0], [1, 1, 1]] and _np(b).tolist() == [[0, 1, 2]] * 2
    assert _np(tf.transpose(xf)).shape == (3, 2)
    assert np.allclose(_np(tf.linalg.norm(tf.constant([[3., 4.]]), axis=-1)), [5.])
    assert np.allclose(_np(tf.linalg.cross(tf.constant([[1., 0., 0.]]), tf.constant([[0., 1., 0.]]))),
                       [[0., 0., 1.]])


def test_keras_pieces(tf):
    d = tf.keras.layers.Dense(2, activation=tf.keras.layers.Activation('relu'))
    d.set_weights([np.array([[1., -1.], [2., 0.5]], np.float32), np.array([0.5, -3.], np.float32)])
    y = d(tf.constant([[1., 2.]]))
    assert np.allclose(_np(y), [[5.5, 0.]])                          # relu(x @ W + b)
    yt, yp = tf.constant([[0., 0.], [1., 1.]]), tf.constant([[1., 3.], [1., 0.]])
    assert _np(tf.keras.losses.MSE(yt, yp)).tolist() == [5., .5]     # mean over the LAST axis
    assert _np(tf.keras.losses.MAE(yt, yp)).tolist() == [2., .5]
    mse = tf.keras.losses.MeanSquaredError(reduction='none')
    assert _np(mse(yt, yp)).tolist() == [5., .5]
    assert float(tf.nn.compute_average_loss(tf.constant([2., 4.]), global_batch_size=4)) == 1.5


def test_gradient_tape_and_custom_gradient(tf):
    x = tf.constant([[1., 2., 3.], [0., -1., 2.]])
    with tf.GradientTape() as g:
        g.watch(x)
        y = tf.reduce_sum(x * x, axis=1, keepdims=True)              # [N, 1]
    jac = g.batch_jacobian(y, x)                                     # [N, 1, 3] = 2 x
    assert _np(jac).shape == (2, 1, 3) and np.allclose(_np(jac)[:, 0], 2 * _np(x))

    @tf.custom_gradient
    def clipped_double(v):
        def grad(dy):
            return dy * 7.                                            # NOT the true derivative
        return 2. * v, grad
    tf.shim_set_training(True)
    try:
        v = tf.Variable(tf.constant([1., 2.]))
        with tf.GradientTape() as tape:
            out = tf.reduce_sum(clipped_double(v))
        (gv,) = tape.gradient(out, [v])
    finally:
        tf.shim_set_training(False)
    assert _np(out).tolist() == 6. and _np(gv).tolist() == [7., 7.]


def test_antialiased_resize_matches_pillow(tf):
    from PIL import Image
    a = (np.random.default_rng(0).random((64, 128)) * 10).astype(np.float32)
    r = _np(tf.image.resize(a[:, :, None], (16, 32), method='bilinear', antialias=True))[:, :, 0]
    p = np.array(Image.fromarray(a, mode='F').resize((32, 16), Image.BILINEAR))
    assert np.abs(r - p).max() < 5e-6
