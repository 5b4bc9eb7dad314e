"""Pins of the oracle's [TF1] semantics against an INDEPENDENT engine that was written to reproduce TensorFlow graphs: OpenCV's
TensorFlow importer (`cv2.dnn.readNetFromTensorflow`).  TensorFlow itself is not installable here (no network, not in the
wheelhouse), so these tests do not pin against TensorFlow; they pin the restatement against a third party's reading of the same
ops, on hand-encoded frozen GraphDefs (tests/tf_graphdef.py) of the reference's own call sites:

  * tf.nn.conv2d(strides [1,2,2,1] / [1,1,1,1], padding 'SAME')          main.py:133-135  - asymmetric SAME padding
  * tf.nn.max_pool(2x2, stride 2, 'SAME') on odd extents (45 -> 23)       main.py:172-174
  * tf.image.resize_images (legacy bilinear, align_corners False)         main.py:51,58,60,67,89 - every size pair the graph uses
  * fused batch norm in inference mode, epsilon 1e-3                      main.py:128-130 (the epsilon DEFAULT of
    tf.contrib.layers.batch_norm is a Python-side constant of TF 1.x and stays [TF1]-unpinned; the formula is pinned)
  * the conv1 layer chain conv2d + bias + relu + batch_norm + max_pool    main.py:156-169,44-45

What stays pinned only by the restatement's reading of TF 1.x: the training-mode moving-average update (unbiased variance, decay),
the Adam / clip_by_global_norm / piecewise_constant forms, truncated_normal, and the softmax-cross-entropy gradient convention.
"""
import numpy as np
import pytest
import torch

import jcm_oracle as orc
import tf_graphdef as tfg

cv2 = pytest.importorskip('cv2')
if not hasattr(cv2, 'dnn'):
    pytest.skip('OpenCV without the dnn module', allow_module_level=True)


def t64(a):
    return torch.from_numpy(np.asarray(a)).double()


@pytest.mark.parametrize('shape,k,stride', [((1, 12, 20, 3), 5, 2), ((2, 13, 21, 3), 5, 2), ((1, 16, 24, 4), 5, 1), ((1, 10, 14, 4), 9, 1),
                                            ((1, 11, 16, 2), 3, 2)])
def test_conv2d_same_padding_matches_opencv_tf_importer(shape, k, stride):
    rng = np.random.default_rng(sum(shape) + k)
    x = rng.standard_normal(shape).astype(np.float32)
    w = rng.standard_normal((k, k, shape[3], 6)).astype(np.float32)
    g = tfg.placeholder('x', shape) + tfg.const('w', w) + tfg.conv2d('conv', 'x', 'w', stride)
    got = tfg.run_opencv(g, x)
    ref = orc.conv2d(t64(x), t64(w), stride).numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 2e-5 * np.abs(ref).max()
    if stride == 2 and shape[1] % 2 == 0:
        # the symmetric padding a PyTorch port would use (padding = k // 2) is a DIFFERENT function
        sym = torch.nn.functional.conv2d(t64(x).permute(0, 3, 1, 2), t64(w).permute(3, 2, 0, 1), stride=2, padding=k // 2).permute(0, 2, 3, 1).numpy()
        assert np.abs(got - sym).max() > 0.1 * np.abs(ref).max()


@pytest.mark.parametrize('shape', [(2, 45, 31, 4), (1, 30, 44, 3), (1, 15, 23, 2)])
def test_max_pool_same_matches_opencv_tf_importer(shape):
    x = np.random.default_rng(1).standard_normal(shape).astype(np.float32)
    got = tfg.run_opencv(tfg.placeholder('x', shape) + tfg.max_pool('pool', 'x'), x)
    ref = orc.max_pool_layer(t64(x)).numpy()
    assert got.shape == ref.shape and np.array_equal(got, ref.astype(np.float32))


@pytest.mark.parametrize('hi,wi,ho,wo', [(30, 45, 60, 90), (15, 23, 60, 90), (61, 91, 60, 90), (48, 72, 24, 36), (48, 72, 12, 18), (97, 129, 96, 128)])
def test_legacy_bilinear_resize_matches_opencv_tf_importer(hi, wi, ho, wo):
    """Every size pair of the reference graph: the two up-samplings of the bank outputs, conv_mrf's 61x91 -> 60x90 (and its 96x128
    analogue), and the two input down-samplings (exact sub-sampling)."""
    x = np.random.default_rng(2).standard_normal((1, hi, wi, 3)).astype(np.float32)
    g = tfg.placeholder('x', x.shape) + tfg.const('size', np.array([ho, wo], np.int32)) + tfg.resize_bilinear('rs', 'x', 'size')
    got = tfg.run_opencv(g, x)
    ref = orc.resize_images(t64(x), ho, wo).numpy()
    assert np.abs(got - ref).max() < 2e-6 * max(1.0, np.abs(ref).max())
    if hi == 2 * ho:
        assert np.array_equal(got, x[:, ::2, ::2])


def test_batch_norm_inference_formula_matches_opencv_tf_importer():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 6, 7, 5)).astype(np.float32)
    gam, bet, mu, var = [rng.random(5).astype(np.float32) + 0.5 for _ in range(4)]
    g = tfg.placeholder('x', x.shape) + tfg.const('gamma', gam) + tfg.const('beta', bet) + tfg.const('mean', mu) + tfg.const('var', var) \
        + tfg.fused_batch_norm('bn', 'x', 'gamma', 'beta', 'mean', 'var', orc.BN_EPS)
    got = tfg.run_opencv(g, x)
    bn = {'gamma': t64(gam), 'beta': t64(bet), 'moving_mean': t64(mu), 'moving_variance': t64(var)}
    ref = orc.batch_norm(t64(x), bn, False).numpy()
    assert np.abs(got - ref).max() < 2e-6 * np.abs(ref).max()


def test_conv1_layer_chain_matches_opencv_tf_importer():
    """conv_layer(x, 5, 2, 3, 16, 'conv1_fullres') + max_pool_layer in inference mode (main.py:44-45,156-169): conv SAME stride 2 +
    bias + ReLU + batch norm (AFTER the ReLU) + 2x2 SAME max-pool, as one frozen graph."""
    rng = np.random.default_rng(4)
    x = rng.random((2, 46, 70, 3)).astype(np.float32)                 # 46 / 2 = 23: odd extent into the pool
    p = {'conv1_fullres/weights': rng.standard_normal((5, 5, 3, 16)).astype(np.float32) * 0.2,
         'conv1_fullres/biases': rng.standard_normal(16).astype(np.float32) * 0.1,
         'conv1_fullres/BatchNorm/gamma': rng.random(16).astype(np.float32) + 0.5,
         'conv1_fullres/BatchNorm/beta': rng.standard_normal(16).astype(np.float32) * 0.1,
         'conv1_fullres/BatchNorm/moving_mean': rng.random(16).astype(np.float32) * 0.3,
         'conv1_fullres/BatchNorm/moving_variance': rng.random(16).astype(np.float32) + 0.5}
    n = 'conv1_fullres'
    g = tfg.placeholder('x', x.shape) + tfg.const('w', p[n + '/weights']) + tfg.const('b', p[n + '/biases'])
    g += tfg.conv2d('conv', 'x', 'w', 2) + tfg.bias_add('pre', 'conv', 'b') + tfg.relu('act', 'pre')
    for k in ('gamma', 'beta', 'moving_mean', 'moving_variance'):
        g += tfg.const(k, p[n + '/BatchNorm/' + k])
    g += tfg.fused_batch_norm('bn', 'act', 'gamma', 'beta', 'moving_mean', 'moving_variance', orc.BN_EPS) + tfg.max_pool('pool', 'bn')
    got = tfg.run_opencv(g, x)
    ref = orc.max_pool_layer(orc.conv_layer(t64(x), {k: t64(v) for k, v in p.items()}, 5, 2, n, False)).numpy()
    assert got.shape == ref.shape == (2, 12, 18, 16)
    assert np.abs(got - ref).max() < 2e-5 * np.abs(ref).max()
