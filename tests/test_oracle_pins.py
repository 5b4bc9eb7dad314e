"""Pins of the CPU oracle (oracle/jcm_oracle.py) against independent references - CPU only.

TensorFlow cannot be installed here and the reference ships no tests, so the oracle is pinned against:
  * scipy.signal.convolve2d for conv_mrf (reference main.py:77-91)
  * a pure-numpy legacy-bilinear resize written independently of the oracle's vectorised one
  * the softmax-CE value the reference logs before any training, ~ln(60*90) = 8.594 (hps_opt:2,37,102)
  * closed-form gradients of the spatial model vs autograd
  * the [TF1] SAME-padding / pooling shape rules stated in the reference's comments (main.py:44-72)
"""
import math

import numpy as np
import pytest
import torch

import jcm_oracle as orc


def legacy_resize_loops(img, oh, ow):
    """tf.image.resize_images legacy bilinear (align_corners=False, no half-pixel offset), scalar loops."""
    ih, iw = img.shape
    out = np.zeros((oh, ow), dtype=np.float64)
    sy, sx = np.float32(ih) / np.float32(oh), np.float32(iw) / np.float32(ow)
    for y in range(oh):
        fy = np.float32(y) * sy
        y0 = int(math.floor(fy))
        y1 = min(y0 + 1, ih - 1)
        wy = float(np.float32(fy - np.float32(y0)))
        for x in range(ow):
            fx = np.float32(x) * sx
            x0 = int(math.floor(fx))
            x1 = min(x0 + 1, iw - 1)
            wx = float(np.float32(fx - np.float32(x0)))
            top = img[y0, x0] + (img[y0, x1] - img[y0, x0]) * wx
            bot = img[y1, x0] + (img[y1, x1] - img[y1, x0]) * wx
            out[y, x] = top + (bot - top) * wy
    return out


@pytest.mark.parametrize('H,W', [(6, 9), (12, 20), (60, 90)])
def test_conv_mrf_equals_scipy_valid_convolution_plus_resize(H, W):
    from scipy import signal
    rng = np.random.default_rng(0)
    A = rng.random((2 * H, 2 * W))
    B = rng.random((2, H, W))
    got = orc.conv_mrf(torch.from_numpy(A).view(1, 2 * H, 2 * W, 1), torch.from_numpy(B).view(2, H, W, 1))
    for n in range(2):
        c = signal.convolve2d(A, B[n], mode='valid')
        assert c.shape == (H + 1, W + 1)
        ref = legacy_resize_loops(c, H, W)
        np.testing.assert_allclose(got[n, :, :, 0].numpy(), ref, rtol=1e-10, atol=1e-10)


def test_resize_integer_downscale_is_strided_subsample():
    x = torch.rand(2, 16, 24, 3, dtype=torch.float64)
    assert torch.equal(orc.resize_images(x, 8, 12), x[:, ::2, ::2])
    assert torch.equal(orc.resize_images(x, 4, 6), x[:, ::4, ::4])


def test_same_padding_rules():
    assert orc.same_pad(480, 5, 2) == (1, 2, 240)      # stride-2 conv1: 1 before, 2 after
    assert orc.same_pad(90, 5, 1) == (2, 2, 90)
    assert orc.same_pad(60, 9, 1) == (4, 4, 60)
    assert orc.same_pad(45, 2, 2) == (0, 1, 23)        # pooling 45 -> 23 pads one column at the end


def test_model_shapes_follow_reference_comments():
    """main.py:44-72: 480x720 -> conv1 240x360 -> pool 120x180 -> pool 60x90; quarter bank 15x23; logits 60x90xK."""
    gen = torch.Generator().manual_seed(0)
    p = orc.init_part_detector(7, gen, debug=True, dtype=torch.float32)
    tap = {}
    x = torch.rand(1, 480, 720, 3, generator=gen)
    out = orc.model(x, p, 7, False, tap=tap)
    assert tuple(out.shape) == (1, 60, 90, 7)
    assert tuple(tap['conv1_fullres/relu'].shape) == (1, 240, 360, 16)
    assert tuple(tap['conv2_fullres/relu'].shape) == (1, 120, 180, 32)
    assert tuple(tap['conv4_fullres/relu'].shape) == (1, 60, 90, 128)
    assert tuple(tap['conv4_halfres/relu'].shape) == (1, 30, 45, 128)
    assert tuple(tap['conv4_quarterres/relu'].shape) == (1, 15, 23, 128)


def test_ce_at_init_is_log_hw():
    """Both heads log ~8.58-8.62 = ln(5400) before training (hps_opt:2,37,102)."""
    K = 7
    gen = torch.Generator().manual_seed(0)
    p = orc.init_part_detector(K, gen, debug=True, dtype=torch.float32)
    from pairwise_prior import JOINT_IDS  # noqa: F401  (same joint order as the oracle)
    import os
    npz = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'joint-cnn-mrf_b200', 'jcm', 'data',
                       'pairwise_distribution.npz')
    with np.load(npz) as z:
        distr = {k: z[k] for k in z.files}
    sm = orc.init_spatial_model(distr, K, 60, 90, dtype=torch.float32)
    x = torch.rand(2, 480, 720, 3, generator=gen)
    y = torch.from_numpy(orc.synthetic_labels(2, 60, 90, K + 1, np.random.default_rng(0)))
    out = orc.tower_forward(x, y, p, sm, K, flag_train=False)
    assert abs(float(out['loss_pd']) - math.log(5400)) < 0.15
    assert abs(float(out['loss_sm']) - math.log(5400)) < 0.15


def test_spatial_model_gradients_match_closed_form():
    """dT = g/T, db = sum_n dT * sigmoid(5b), dE via correlation with the flipped likelihood."""
    H, W, K, B = 5, 7, 2, 2
    rng = np.random.default_rng(1)
    names = ['a', 'b', 'torso']
    distr = orc.synthetic_pairwise(names, K, H, W, rng)
    sm = orc.init_spatial_model(distr, K, H, W, joint_names=names, requires_grad=True)
    hm = torch.rand(B, H, W, K + 1, dtype=torch.float64, requires_grad=True)
    out = orc.spatial_model(hm, sm, K, False, joint_names=names)
    g = torch.rand_like(out)
    (out * g).sum().backward()
    # closed form for pair (a | b): target 0, cond 1
    with torch.no_grad():
        h = hm  # eval-mode BN with identity moving stats: y = x / sqrt(1 + eps)
        hn = h / math.sqrt(1 + orc.BN_EPS)
        L = orc.softplus(hn[..., 1:2])
        P = orc.softplus(sm['energy_a_b'])
        C = orc.conv_mrf(P, L)
        T = C + orc.softplus(sm['bias_a_b']) + orc.DELTA
        dT = g[..., 0:1] / T
        db = dT.sum(0, keepdim=True) * torch.sigmoid(5 * sm['bias_a_b'])
    np.testing.assert_allclose(sm['bias_a_b'].grad.numpy(), db.numpy(), rtol=1e-9, atol=1e-12)


def test_adam_and_schedule_helpers():
    b, v = orc.lr_schedule(0.001, 60, 3987, 14)
    assert v == [0.001, 0.0005, 0.0002, 0.0001]
    assert orc.piecewise_constant(b[0], b, v) == 0.001 and orc.piecewise_constant(b[0] + 1, b, v) == 0.0005
    p = [torch.ones(3, dtype=torch.float64)]
    m, vv = [torch.zeros(3, dtype=torch.float64)], [torch.zeros(3, dtype=torch.float64)]
    orc.adam_tf1_step(p, [torch.full((3,), 0.5, dtype=torch.float64)], m, vv, 1, 0.01)
    # first TF1 Adam step moves by lr * sqrt(1-b2)/(1-b1) * (0.1 g)/(sqrt(0.001 g^2)+eps) ~= lr
    np.testing.assert_allclose(p[0].numpy(), 1 - 0.01, rtol=1e-6)


def test_adam_restatement_against_torch_optim_adam():
    """Independent engine: torch.optim.Adam applies the bias corrections as  lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps), TF1's
    AdamOptimizer as  lr*sqrt(1-b2^t)/(1-b1^t) * m / (sqrt(v) + eps)  ("epsilon hat" of the Adam paper, main.py:501-506).  The two are
    the same update when torch's eps is eps_tf / sqrt(1-b2^t) (with the same eps they differ by ~eps/sqrt(v): 5e-3 of a step on the
    smallest of these gradients).  Five steps on gradients of magnitude 1e-2: the restatement equals torch's optimizer run with the
    per-step rescaled eps to 1e-12 of a step - i.e. moments, bias corrections and the step-count convention (t from 1) are those of
    a third-party Adam, and the only TF1-specific reading left is where eps enters."""
    g = torch.Generator().manual_seed(3)
    w0 = torch.randn(50, generator=g, dtype=torch.float64)
    grads = [0.01 * torch.randn(50, generator=g, dtype=torch.float64) for _ in range(5)]
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    p = [w0.clone()]
    m, v = [torch.zeros(50, dtype=torch.float64)], [torch.zeros(50, dtype=torch.float64)]
    for t, gr in enumerate(grads, 1):
        orc.adam_tf1_step(p, [gr], m, v, t, lr, b1, b2, eps)
    for rescale, tol in ((True, 1e-12), (False, 2e-2)):
        q = torch.nn.Parameter(w0.clone())
        opt = torch.optim.Adam([q], lr=lr, betas=(b1, b2), eps=eps)
        for t, gr in enumerate(grads, 1):
            if rescale:
                opt.param_groups[0]['eps'] = eps / math.sqrt(1 - b2 ** t)
            q.grad = gr.clone()
            opt.step()
        step = (w0 - p[0]).abs().max()
        assert float((q.detach() - p[0]).abs().max() / step) < tol, (rescale, float((q.detach() - p[0]).abs().max() / step))


def test_grad_renorm_against_torch_clip_grad_norm():
    """tf.clip_by_global_norm(g, c) = g * c / max(||g||, c) (main.py:302-309); torch.nn.utils.clip_grad_norm_ scales by
    min(1, c / (||g|| + 1e-6)): the same to 1e-6 relative, above and below the threshold."""
    g = torch.Generator().manual_seed(4)
    for scale in (10.0, 0.01):
        gs = [scale * torch.randn(7, 5, generator=g, dtype=torch.float64), scale * torch.randn(11, generator=g, dtype=torch.float64)]
        out, gn = orc.grad_renorm(gs, 4.0)
        ps = [torch.nn.Parameter(torch.zeros_like(x)) for x in gs]
        for p_, x in zip(ps, gs):
            p_.grad = x.clone()
        total = torch.nn.utils.clip_grad_norm_(ps, 4.0)
        assert abs(float(total) - gn) < 1e-12 * max(gn, 1.0)
        for p_, o in zip(ps, out):
            assert float((p_.grad - o).abs().max() / o.abs().max()) < 2e-6


def test_fft_form_of_conv_mrf_equals_the_direct_form():
    """The FFT evaluation used by the K=14 / 96x128 GPU test is the same function as the direct restatement of main.py:77-91:
    values and both gradients to 1e-12 in fp64."""
    g = torch.Generator().manual_seed(9)
    for H, W, b in ((12, 20, 3), (7, 9, 1), (24, 32, 2)):
        A = torch.rand(1, 2 * H, 2 * W, 1, generator=g, dtype=torch.float64, requires_grad=True)
        L = torch.rand(b, H, W, 1, generator=g, dtype=torch.float64, requires_grad=True)
        d, f = orc.conv_mrf(A, L), orc.conv_mrf(A, L, fft=True)
        assert float((d - f).abs().max()) < 1e-12 * float(d.abs().max())
        w = torch.rand(d.shape, generator=g, dtype=torch.float64)
        gd = torch.autograd.grad((d * w).sum(), [A, L])
        gf = torch.autograd.grad((f * w).sum(), [A, L])
        for x, y in zip(gd, gf):
            assert float((x - y).abs().max()) < 1e-12 * float(x.abs().max())


def test_softmax_cross_entropy_gradient_is_the_tf1_backprop():
    """[TF1] tf.nn.softmax_cross_entropy_with_logits: loss = -sum(labels * log_softmax), registered gradient = softmax - labels for
    ANY labels.  Equal to autograd of the formula when the labels sum to 1; for a clipped blob (sum 9/16) it is softmax - labels, not
    softmax * 9/16 - labels."""
    g = torch.Generator().manual_seed(10)
    x = torch.randn(1, 5, 6, 2, generator=g, dtype=torch.float64, requires_grad=True)
    y = torch.zeros(1, 5, 6, 2, dtype=torch.float64)
    y[0, 2, 3, 0] = 1.0
    y[0, :2, :2, 1] = torch.tensor([[4., 2], [2, 1]], dtype=torch.float64) / 16
    loss = orc.softmax_cross_entropy(x, y)
    ls = torch.log_softmax(x.reshape(1, 30, 2), 1)
    assert abs(float(loss) - float((-(y.reshape(1, 30, 2) * ls).sum(1)).mean())) < 1e-14
    loss.backward()
    want = (ls.exp() - y.reshape(1, 30, 2)).reshape(1, 5, 6, 2) / 2
    assert float((x.grad - want).abs().max()) < 1e-14
