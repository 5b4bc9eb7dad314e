"""multi-scale test-time inference (main.py:326-425) as `jcm.multiscale` builds it on the host.
`resize` restates skimage.transform.resize (0.13 defaults); scikit-image is not installable here, so it is pinned against
scipy.ndimage.map_coordinates(order=1, mode='grid-constant') - an independent implementation of the same algorithm.  CPU only:
the forward pass is injected (on a GPU box it is `jcm.multiscale.gpu_forward`, a thin wrapper of the tested `tower_forward`)."""
import os
import sys

import numpy as np
import pytest
from scipy import ndimage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'joint-cnn-mrf_b200'))


@pytest.fixture(scope='module')
def ms(built_lib):
    from jcm import multiscale
    return multiscale


def ndimage_resize(img, rows, cols):
    h, w = img.shape[:2]
    r = (np.arange(rows) + 0.5) * (h / rows) - 0.5
    c = (np.arange(cols) + 0.5) * (w / cols) - 0.5
    rr, cc = np.meshgrid(r, c, indexing='ij')
    if img.ndim == 2:
        return ndimage.map_coordinates(img.astype(np.float64), [rr, cc], order=1, mode='grid-constant', cval=0.0)
    return np.stack([ndimage.map_coordinates(img[..., k].astype(np.float64), [rr, cc], order=1, mode='grid-constant', cval=0.0)
                     for k in range(img.shape[2])], axis=2)


@pytest.mark.parametrize('shape,out', [((48, 72, 3), (40, 60)), ((37, 53, 2), (60, 90)), ((60, 90), (60, 90)), ((66, 100, 1), (60, 90)),
                                       ((5, 7, 3), (1, 1)), ((480, 720, 3), (480, 720))])
def test_resize_is_half_pixel_bilinear_with_zero_outside(ms, shape, out):
    rng = np.random.default_rng(1)
    img = rng.random(shape).astype(np.float32)
    got = ms.resize(img, out, clip=False)
    assert got.dtype == np.float64 and got.shape[:2] == out
    np.testing.assert_allclose(got, ndimage_resize(img, *out), rtol=0, atol=1e-12)
    if shape[:2] == out:
        np.testing.assert_array_equal(got, img.astype(np.float64))       # same size: the identity


def test_resize_clip_rule_and_range_check(ms):
    """warp(clip=True): clipped to the input's range, but exact cval pixels survive when cval is outside that range."""
    img = np.full([4, 4], 0.5)
    padded = np.pad(img, 2, 'constant')
    out = ms.resize(padded, (4, 4))                                       # input range [0, 0.5] contains cval: plain clip
    assert out.min() >= 0 and out.max() <= 0.5
    hm = np.full([6, 6, 1], 0.25)
    up = ms.resize(hm, (12, 12))          # enlarging samples outside the image at the border: blends with 0 -> clipped up to 0.25
    np.testing.assert_allclose(up, 0.25)
    raw = ms.resize(hm, (12, 12), clip=False)
    assert raw.min() < 0.25
    far = ms.resize(np.pad(hm, ((6, 6), (6, 6), (0, 0)), 'constant') + 0.0, (18, 18))     # range [0, .25]: zeros stay zeros
    assert far[0, 0, 0] == 0.0
    with pytest.raises(ValueError):
        ms.resize(np.full([4, 4], 1.5), (2, 2))                           # img_as_float: float images must be in [-1, 1]


def test_scales_geometry_and_inverse(ms):
    """get_different_scales / scale_hm_back are inverse geometries: a blob seen through every scale and a stand-in network
    (8x8 mean pooling = the /8 heat-map grid) comes back at its own location in all 8 maps."""
    H, W = 480, 720
    yy, xx = np.mgrid[0:H, 0:W]
    cy, cx = 260.0, 300.0
    img = np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 20.0 ** 2))[..., None].repeat(3, axis=2)
    xs = ms.get_different_scales(img)
    assert xs.shape == (8, H, W, 3) and xs.dtype == np.float64
    np.testing.assert_array_equal(xs[7], img)                              # crop 1.0 is the identity
    assert xs[3][:60].max() < 1e-3                                         # pad 1.4: the frame's outer band is the zero padding
    hms = xs.reshape(8, 60, 8, 90, 8, 3).mean(axis=(2, 4))[..., :1]        # stand-in network
    back = ms.scale_hm_back(hms)
    assert back.shape == (8, 60, 90, 1)
    for i in range(8):
        r, c = np.unravel_index(back[i, :, :, 0].argmax(), (60, 90))
        # cell r of the /8 grid is centred on pixel 8r + 3.5; the rounded crop / pad extents make the round trip exact to ~1 cell
        assert abs(r - (cy - 3.5) / 8) <= 1.25 and abs(c - (cx - 3.5) / 8) <= 1.25, (i, r, c)
    # the reference's rounding of the crop / pad extents (Python round: half to even)
    assert round(480 * (1.3 - 1) / 2) == 72 and round((1 - 1 / 1.1) / 2 * 60) == 3 and round(60 * (1 / 0.7 - 1) / 2) == 13


def test_get_predictions_composes_like_main_py(ms):
    rng = np.random.default_rng(2)
    K, N = 3, 2
    X = rng.random([N, 96, 144, 3]).astype(np.float32)
    Y = rng.random([N, 12, 18, K + 1]).astype(np.float32)
    calls = []

    def forward(x8, y8):
        assert x8.shape == (8, 96, 144, 3) and y8.shape == (8, 12, 18, K + 1)
        np.testing.assert_array_equal(y8[0], y8[7])
        calls.append(1)
        pd = x8.reshape(8, 12, 8, 18, 8, 3).mean(axis=(2, 4))
        return pd.astype(np.float32), (pd ** 2).astype(np.float32)

    det = lambda hm, y: float(hm.shape == (1, 12, 18, K) and y.shape == (1, 12, 18, K + 1))
    cpd, csm, dpd, dsm = ms.get_predictions(X, Y, forward, det_rate=det, n=1100)
    assert len(calls) == N and cpd.shape == (2, K, N) and csm.shape == (2, K, N) and dpd == 1.0 and dsm == 1.0
    # by hand for image 1
    xs = ms.get_different_scales(X[1])
    pd, sm_ = forward(xs, np.repeat(Y[1][None], 8, axis=0))
    avg = np.average(ms.scale_hm_back(pd), axis=0)
    flat = avg.reshape(12 * 18, K).argmax(axis=0)
    np.testing.assert_array_equal(cpd[:, :, 1], np.stack([flat // 18, flat % 18]))
    assert ms.get_predictions(X, Y, forward, n=1)[0].shape == (2, K, 1)


def test_argmax_hm_takes_the_first_maximum(ms):
    hm = np.zeros([1, 4, 5, 2])
    hm[0, 1, 2, 0] = hm[0, 3, 4, 0] = 1.0      # tie: the first in row-major order wins (np.argmax, main.py:392)
    hm[0, 3, 0, 1] = 2.0
    np.testing.assert_array_equal(ms.argmax_hm(hm), [[1, 3], [2, 0]])


@pytest.mark.parametrize('tag', ['a', 'b', 'c'])
def test_geometry_matches_the_reference_functions(ms, tag):
    """get_different_scales / scale_hm_back against the outputs of the reference's own two functions (main.py:326-379, executed by
    tests/golden/make_multiscale_golden.py with skimage's resize bound to jcm.multiscale.resize): same windows, same margins, same
    rounding for every scale, on three map sizes including the odd 15x23."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'multiscale_reference.npz'))
    x, hms = z[tag + '_x'].astype(np.float64), z[tag + '_hms'].astype(np.float64)
    got = ms.get_different_scales(x)
    assert got.shape == z[tag + '_scales'].shape
    assert np.abs(got - z[tag + '_scales']).max() < 1e-6
    back = ms.scale_hm_back(hms)
    assert np.abs(back - z[tag + '_back']).max() < 1e-6
