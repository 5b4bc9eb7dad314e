"""CPU ORACLE for the joint-cnn-mrf hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this.
The product path (`joint-cnn-mrf_b200/jcm`) never does, and raises if its CUDA library is missing.

What this is: the reference TensorFlow-1.x graph of `/root/reference/main.py` restated operation by operation
on torch-CPU tensors (fp64 by default = ground truth, fp32 = "TF-CPU stand-in" for timing), keeping the
reference's layouts (activations NHWC, conv kernels HWIO, energies [1,2H,2W,1], biases [1,H,W,1]) and the
TF1 semantics that are not visible in the reference tree ([TF1] tags):

  * SAME padding is asymmetric: total = max((out-1)*s + k - in, 0), before = total // 2        [TF1]
  * `tf.image.resize_images` = legacy bilinear: src = dst * (in/out), no half-pixel offset     [TF1]
  * `tf.contrib.layers.batch_norm`: eps 1e-3, decay 0.9, biased batch variance for the
    normalisation, unbiased variance into the moving average, applied AFTER the ReLU            [TF1]
  * max_pool SAME pads at the end with -inf                                                     [TF1]
  * Adam in the TF1 form, clip_by_global_norm, piecewise_constant                               [TF1]
  * softmax_cross_entropy_with_logits: gradient = softmax - labels even for labels not summing to 1 [TF1]

PARITY PINNING STATUS: **parity unpinned against TensorFlow itself** - TensorFlow 1.x cannot be installed in
this image (no network) and the reference holds no tests, golden vectors or fixtures for this path.
What the oracle IS pinned against (tests/test_oracle_pins.py, tests/test_tf_semantics_opencv.py, tests/test_cpu_suite.py):
  * an INDEPENDENT engine written to reproduce TensorFlow graphs - OpenCV's TensorFlow importer (cv2.dnn), fed
    hand-encoded frozen GraphDefs of the reference's call sites: conv2d 'SAME' with strides 1 and 2 (the asymmetric
    padding), max_pool 2x2 'SAME' on odd extents, tf.image.resize_images for every size pair the graph uses (incl.
    conv_mrf's 61x91 -> 60x90), the inference-mode batch-norm formula, and the whole conv1 layer chain
    (conv + bias + ReLU + BN-after-ReLU + pool).  This is a third party's reading of TF, not TF.
  * `conv_mrf` == scipy.signal.convolve2d(prior, likelihood, 'valid') + legacy resize (independent library)
  * softmax-CE of both heads at initialisation == ln(H*W) = 8.594, the value the reference logs
    (`hps_opt:2,37,102`)
  * the pairwise initialisation regenerated from `data_FLIC.mat` reproduces the zero-run structure of all 90
    arrays that survives in the corrupt shipped pickle
  * shapes stated in the reference's comments (`main.py:44-72,79-87`)
  * autograd gradients of the spatial model == the closed forms of SURVEY.md Appendix D

Every function cites the reference lines it follows.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

JOINT_NAMES = ['lsho', 'lelb', 'lwri', 'rsho', 'relb', 'rwri', 'lhip', 'rhip', 'nose', 'torso']  # main.py:18
BN_EPS = 1e-3      # [TF1] tf.contrib.layers.batch_norm default epsilon
BN_DECAY = 0.9     # main.py:113,129
SOFTPLUS_ALPHA = 5.0  # main.py:107
DELTA = 10 ** -6   # main.py:110


# ----------------------------------------------------------------------------------------------------------
# layer helpers (main.py:128-174)
# ----------------------------------------------------------------------------------------------------------
def same_pad(in_size, k, s):
    """[TF1] SAME padding: (before, after, out)."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return total // 2, total - total // 2, out


def conv2d(x, W, stride):
    """main.py:133-135  tf.nn.conv2d(x NHWC, W HWIO, SAME). Cross-correlation."""
    kh, kw = W.shape[0], W.shape[1]
    pt, pb, _ = same_pad(x.shape[1], kh, stride)
    pl, pr, _ = same_pad(x.shape[2], kw, stride)
    xc = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    y = F.conv2d(xc, W.permute(3, 2, 0, 1), stride=stride)
    return y.permute(0, 2, 3, 1)


def batch_norm(x, bn, flag_train, update=True):
    """main.py:128-130 (+ :113 for bn_sm).  bn = dict(gamma, beta, moving_mean, moving_variance) over the last axis.
    Training: biased batch variance; moving stats <- 0.9*old + 0.1*batch (unbiased variance)   [TF1]."""
    C = x.shape[-1]
    if flag_train:
        xf = x.reshape(-1, C)
        n = xf.shape[0]
        mean = xf.mean(0)
        var = ((xf - mean) ** 2).mean(0)
        if update:
            with torch.no_grad():
                bn['moving_mean'].mul_(BN_DECAY).add_((1 - BN_DECAY) * mean)
                bn['moving_variance'].mul_(BN_DECAY).add_((1 - BN_DECAY) * var * (n / max(n - 1, 1)))
    else:
        mean, var = bn['moving_mean'], bn['moving_variance']
    return (x - mean) * torch.rsqrt(var + BN_EPS) * bn['gamma'] + bn['beta']


def max_pool_layer(x, size=2, stride=2, select=None):
    """main.py:172-174  2x2 s2 SAME max-pool (45 -> 23, the padded column never wins)   [TF1].
    select (tests only): int64 [B,Ho,Wo,C] with the window element (2*dy + dx) to take instead of the arg-max, so that a
    gradient comparison does not depend on how a near-tie between two window elements was rounded (see conv_layer)."""
    pt, pb, _ = same_pad(x.shape[1], size, stride)
    pl, pr, _ = same_pad(x.shape[2], size, stride)
    xc = F.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb), value=float('-inf'))
    if select is None:
        return F.max_pool2d(xc, size, stride).permute(0, 2, 3, 1)
    assert size == 2 and stride == 2
    B, C, Hp, Wp = xc.shape
    win = xc.reshape(B, C, Hp // 2, 2, Wp // 2, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, Hp // 2, Wp // 2, 4, C)
    return torch.gather(win, 3, select.unsqueeze(3)).squeeze(3)


def resize_images(x, out_h, out_w):
    """[TF1] tf.image.resize_images(x NHWC, [out_h, out_w]) = legacy bilinear, align_corners=False:
    scale = in/out (float32), src = dst*scale (float32), lo = floor(src), hi = min(lo+1, in-1), w = src-lo;
    value = top + (bottom-top)*wy with top = tl + (tr-tl)*wx.   Used at main.py:51,58,60,67,89."""
    in_h, in_w = x.shape[1], x.shape[2]

    def taps(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        src = np.arange(n_out, dtype=np.float32) * scale
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, n_in - 1)
        w = (src - lo.astype(np.float32)).astype(np.float32)
        return torch.from_numpy(lo), torch.from_numpy(hi), torch.from_numpy(w.astype(np.float64)).to(x.dtype)

    ylo, yhi, wy = taps(in_h, out_h)
    xlo, xhi, wx = taps(in_w, out_w)
    top_l, top_r = x[:, ylo][:, :, xlo], x[:, ylo][:, :, xhi]
    bot_l, bot_r = x[:, yhi][:, :, xlo], x[:, yhi][:, :, xhi]
    wx_ = wx.view(1, 1, -1, 1)
    wy_ = wy.view(1, -1, 1, 1)
    top = top_l + (top_r - top_l) * wx_
    bot = bot_l + (bot_r - bot_l) * wx_
    return top + (bot - top) * wy_


def weight_variable(shape, gen, dtype=torch.float64):
    """main.py:138-147  He init, tf.truncated_normal (re-drawn outside +-2 sigma)   [TF1]."""
    n_in = shape[0] * shape[1] * shape[2]
    std = math.sqrt(2.0 / n_in)
    w = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
    return (w * std).to(dtype)


def conv_layer(x, p, size, stride, name, flag_train, last_layer=False, tap=None, relu_masks=None):
    """main.py:156-169  relu(conv_SAME(x,w)+b) then BN (BN after ReLU); last layer: bias only.
    relu_masks (tests only): dict name -> bool tensor; the ReLU is then evaluated as pre * mask, i.e. with the on/off
    pattern decided elsewhere (the fp32 GPU forward).  The two agree except where |pre| is below the fp32 forward error,
    so the forward value is unchanged to ~1e-6 while the GRADIENT comparison no longer depends on which side of zero
    such elements were rounded to (one flipped element moves a bias gradient by several per cent)."""
    pre = conv2d(x, p[name + '/weights'], stride) + p[name + '/biases']
    if last_layer:
        return pre
    if relu_masks is not None and name in relu_masks:
        act = pre * relu_masks[name].to(pre.dtype)
    else:
        act = torch.relu(pre)
    if tap is not None:
        tap[name + '/relu'] = act
    bn = {k: p[name + '/BatchNorm/' + k] for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')}
    return batch_norm(act, bn, flag_train)


# ----------------------------------------------------------------------------------------------------------
# part detector (main.py:29-74)
# ----------------------------------------------------------------------------------------------------------
def n_filters(debug=False):
    f = np.array([64, 128, 256, 512, 512])  # main.py:38
    return f // 4 if debug else f           # main.py:40-41


def model(x, p, n_joints, flag_train, tap=None, relu_masks=None, pool_select=None):
    """main.py:29-74. x [B,H,W,3] -> logits [B,H/8,W/8,n_joints]."""
    kw = dict(tap=tap, relu_masks=relu_masks)
    ps = pool_select or {}

    def bank(xin, sfx):
        h = conv_layer(xin, p, 5, 2, 'conv1_' + sfx, flag_train, **kw)
        h = max_pool_layer(h, select=ps.get('conv1_' + sfx))
        h = conv_layer(h, p, 5, 1, 'conv2_' + sfx, flag_train, **kw)
        h = max_pool_layer(h, select=ps.get('conv2_' + sfx))
        h = conv_layer(h, p, 5, 1, 'conv3_' + sfx, flag_train, **kw)
        h = conv_layer(h, p, 9, 1, 'conv4_' + sfx, flag_train, **kw)
        return h

    H, W = x.shape[1], x.shape[2]
    x1 = bank(x, 'fullres')
    x2 = bank(resize_images(x, H // 2, W // 2), 'halfres')
    x2 = resize_images(x2, x1.shape[1], x1.shape[2])
    x3 = bank(resize_images(x, H // 4, W // 4), 'quarterres')
    x3 = resize_images(x3, x1.shape[1], x1.shape[2])
    h = (x1 + x2 + x3) / 3
    if tap is not None:
        tap['merge'] = h
    h = conv_layer(h, p, 9, 1, 'conv5', flag_train, **kw)
    return conv_layer(h, p, 9, 1, 'conv6', flag_train, last_layer=True)


def init_part_detector(n_joints, gen, debug=False, dtype=torch.float64, requires_grad=False):
    """Variables of `model` under the TF names (main.py:138-159; BN: gamma 1, beta 0, moving (0,1))."""
    f = n_filters(debug)
    p = {}
    specs = []
    for sfx in ('fullres', 'halfres', 'quarterres'):
        specs += [('conv1_' + sfx, 5, 3, f[0]), ('conv2_' + sfx, 5, f[0], f[1]),
                  ('conv3_' + sfx, 5, f[1], f[2]), ('conv4_' + sfx, 9, f[2], f[3])]
    specs += [('conv5', 9, f[3], f[4]), ('conv6', 9, f[4], n_joints)]
    for name, k, cin, cout in specs:
        cin, cout = int(cin), int(cout)
        p[name + '/weights'] = weight_variable([k, k, cin, cout], gen, dtype)
        p[name + '/biases'] = torch.zeros(cout, dtype=dtype)
        if name != 'conv6':
            p[name + '/BatchNorm/gamma'] = torch.ones(cout, dtype=dtype)
            p[name + '/BatchNorm/beta'] = torch.zeros(cout, dtype=dtype)
            p[name + '/BatchNorm/moving_mean'] = torch.zeros(cout, dtype=dtype)
            p[name + '/BatchNorm/moving_variance'] = torch.ones(cout, dtype=dtype)
    if requires_grad:
        for k, v in p.items():
            if 'moving_' not in k:
                v.requires_grad_(True)
    return p


# ----------------------------------------------------------------------------------------------------------
# spatial model (main.py:77-125)
# ----------------------------------------------------------------------------------------------------------
def softplus(x):
    """main.py:106-108  1/alpha * softplus(alpha*x)."""
    return F.softplus(SOFTPLUS_ALPHA * x) / SOFTPLUS_ALPHA


def conv_mrf(A, B, fft=False):
    """main.py:77-91.  A [1,2H,2W,1] prior, B [b,H,W,1] likelihood -> [b,H,W,1].
    transpose+reverse+conv2d VALID == true 2-D convolution 'valid' ([1,H+1,W+1,b]); then NOT cropped but
    legacy-bilinear resized to [H,W] (main.py:88-89).
    fft=True (tests at K=14 / 96x128, where the direct form takes seconds per pair): the same 'valid' convolution through
    torch.fft in the input precision - equal to the direct form to ~1e-13 in fp64 (tests/test_oracle_pins.py), differentiable."""
    hm_h, hm_w = B.shape[1], B.shape[2]
    if fft:
        a = A[0, :, :, 0]
        s = (3 * hm_h - 1, 3 * hm_w - 1)                             # full linear convolution size
        full = torch.fft.irfft2(torch.fft.rfft2(a, s=s).unsqueeze(0) * torch.fft.rfft2(B[:, :, :, 0], s=s), s=s)
        C = full[:, hm_h - 1:2 * hm_h, hm_w - 1:2 * hm_w].unsqueeze(3)   # 'valid' part [b, H+1, W+1, 1]
        return resize_images(C, hm_h, hm_w)
    Bf = torch.flip(B.permute(1, 2, 3, 0), dims=[0, 1])            # [h, w, 1, b]  (main.py:83-84)
    C = F.conv2d(A.permute(0, 3, 1, 2), Bf.permute(3, 2, 0, 1))    # [1, b, H+1, W+1] (main.py:87)
    C = resize_images(C.permute(0, 2, 3, 1), hm_h, hm_w)           # main.py:89
    return C.permute(3, 1, 2, 0)                                   # main.py:90


def spatial_model(heat_map, sm, n_joints, flag_train, joint_names=None, fft=False):
    """main.py:94-125. heat_map [B,H,W,K+1]; sm = dict with 'bn_sm/BatchNorm/*', 'energy_<a>_<b>' [1,2H,2W,1],
    'bias_<a>_<b>' [1,H,W,1].  Sum order: unary first, then cond joints ascending in joint_names (main.py:117-123).
    fft: see conv_mrf."""
    names = list(joint_names if joint_names is not None else JOINT_NAMES[:n_joints] + ['torso'])
    bn = {k: sm['bn_sm/BatchNorm/' + k] for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')}
    h = batch_norm(heat_map, bn, flag_train)                       # main.py:112-113
    out = []
    for i, jn in enumerate(names[:n_joints]):
        m = torch.log(softplus(h[:, :, :, i:i + 1]) + DELTA)       # main.py:117
        for j, cn in enumerate(names):                             # joint_dependence: all others (main.py:24-26)
            if cn == jn:
                continue
            prior = softplus(sm['energy_' + jn + '_' + cn])        # main.py:120
            lik = softplus(h[:, :, :, j:j + 1])                    # main.py:121
            bias = softplus(sm['bias_' + jn + '_' + cn])           # main.py:122
            m = m + torch.log(conv_mrf(prior, lik, fft=fft) + bias + DELTA)  # main.py:123
        out.append(m)
    return torch.stack(out, dim=3)[:, :, :, :, 0]                  # main.py:125


def init_spatial_model(pairwise_distr, n_joints, hm_h, hm_w, joint_names=None, dtype=torch.float64,
                       requires_grad=False):
    """main.py:477-487 + bn_sm variables (main.py:112-113)."""
    names = list(joint_names if joint_names is not None else JOINT_NAMES[:n_joints] + ['torso'])
    sm = {'bn_sm/BatchNorm/gamma': torch.ones(len(names), dtype=dtype),
          'bn_sm/BatchNorm/beta': torch.zeros(len(names), dtype=dtype),
          'bn_sm/BatchNorm/moving_mean': torch.zeros(len(names), dtype=dtype),
          'bn_sm/BatchNorm/moving_variance': torch.ones(len(names), dtype=dtype)}
    for jn in names[:n_joints]:
        for cn in names:
            if cn == jn:
                continue
            key = jn + '_' + cn
            e = torch.from_numpy(np.asarray(pairwise_distr[key], dtype=np.float32)).to(dtype)  # cast as main.py:482
            sm['energy_' + key] = e.reshape(1, e.shape[0], e.shape[1], 1).clone()
            sm['bias_' + key] = torch.full((1, hm_h, hm_w, 1), 0.00001, dtype=torch.float32).to(dtype)
    if requires_grad:
        for k, v in sm.items():
            if 'moving_' not in k:
                v.requires_grad_(True)
    return sm


# ----------------------------------------------------------------------------------------------------------
# softmax / loss / metric (main.py:195-240, evaluation.py)
# ----------------------------------------------------------------------------------------------------------
def spatial_softmax(hm):
    """main.py:212-217."""
    B, H, W, K = hm.shape
    return torch.softmax(hm.reshape(B, H * W, K), dim=1).reshape(B, H, W, K)


class _SoftmaxXentTF1(torch.autograd.Function):
    """[TF1] tf.nn.softmax_cross_entropy_with_logits(dim=1) on [B, S, K]: loss = -sum_s labels * log_softmax(logits) and - as the
    op's registered gradient - backprop = softmax(logits) - labels, WHATEVER sum(labels) is (TF's xent kernel emits `backprop`
    next to `loss` and the gradient function is grad_loss * backprop).  For label maps that sum to 1 this is the derivative of the
    loss; for the border-clipped blobs of data.py:180-186 (sum down to 0.25) it is not: autograd of the formula would give
    softmax * sum(labels) - labels.  No gradient flows into the labels."""

    @staticmethod
    def forward(ctx, logits, labels):
        ls = torch.log_softmax(logits, dim=1)
        ctx.save_for_backward(ls, labels)
        return -(labels * ls).sum(1)

    @staticmethod
    def backward(ctx, g):
        ls, labels = ctx.saved_tensors
        return g.unsqueeze(1) * (ls.exp() - labels), None


def softmax_cross_entropy(hm1, hm2):
    """main.py:220-240  mean over (n,k) of -sum_s labels*log_softmax(logits)."""
    B, H, W, K = hm1.shape
    return _SoftmaxXentTF1.apply(hm1.reshape(B, H * W, K), hm2.reshape(B, H * W, K)).mean()


def weight_decay(p, var_pattern='weights'):
    """main.py:195-205  sum of tf.nn.l2_loss = sum(w^2)/2 over variables whose name contains the pattern."""
    return sum((v ** 2).sum() / 2 for k, v in p.items() if var_pattern in k)


def get_joints_coords(hm):
    """evaluation.py:15-24 -> [B, 2, K] (row, col); first maximal index in row-major order."""
    B, H, W, K = hm.shape
    idx = torch.argmax(hm.reshape(B, H * W, K), dim=1)
    row = idx // W
    col = idx - row * W
    return torch.stack([row, col], dim=1)


def det_rate(heat_map_pred, heat_map_target, normalized_radius=10, joints='all'):
    """evaluation.py:4-37 (keeps the reference's lhip_idx, rsho_idx = 0, 7)."""
    lhip_idx, rsho_idx = 0, 7
    pred = get_joints_coords(heat_map_pred).to(torch.float32)
    true = get_joints_coords(heat_map_target).to(torch.float32)
    torso = torch.norm(true[:, :, lhip_idx] - true[:, :, rsho_idx], dim=1, keepdim=True)
    nd = torch.norm(pred - true, dim=1) * 100 / torso
    if joints != 'all':
        nd = torch.stack([nd[:, j] for j in joints], dim=1)
    return 100 * (nd <= normalized_radius).to(torch.float32).mean()


# ----------------------------------------------------------------------------------------------------------
# tower loss + DP step (main.py:243-267, 302-309, 491-577)
# ----------------------------------------------------------------------------------------------------------
def tower_forward(x, hm_target, p, sm, n_joints, flag_train, use_sm=True, lmbd=0.001, tap=None, relu_masks=None,
                  joint_names=None, pool_select=None):
    """main.py:522-541 for one tower. Returns dict(loss, loss_pd, loss_sm, logits..)."""
    logit_pd = model(x, p, n_joints, flag_train, tap=tap, relu_masks=relu_masks, pool_select=pool_select)
    hm_pd = spatial_softmax(logit_pd)
    if use_sm:
        cat = torch.cat([hm_pd, hm_target[:, :, :, n_joints:]], dim=3)          # main.py:528
        logit_sm = spatial_model(cat, sm, n_joints, flag_train, joint_names=joint_names)
        hm_sm = spatial_softmax(logit_sm)
    else:
        logit_sm, hm_sm = logit_pd, hm_pd
    loss_pd = softmax_cross_entropy(logit_pd, hm_target[:, :, :, :n_joints])   # main.py:539
    loss_sm = softmax_cross_entropy(logit_sm, hm_target[:, :, :, :n_joints])   # main.py:540
    loss = loss_pd + loss_sm + lmbd * weight_decay(p)                           # main.py:541
    return dict(loss=loss, loss_pd=loss_pd, loss_sm=loss_sm, logit_pd=logit_pd, logit_sm=logit_sm,
                hm_pd=hm_pd, hm_sm=hm_sm)


def average_gradients(tower_grads):
    """main.py:243-267: list (per tower) of lists of grads -> mean over towers."""
    return [torch.stack(gs, 0).mean(0) for gs in zip(*tower_grads)]


def grad_renorm(grads, norm):
    """main.py:302-309 tf.clip_by_global_norm: g * norm / max(||g||, norm)   [TF1]."""
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads))
    scale = norm / max(float(gn), norm)
    return [g * scale for g in grads], float(gn)


def adam_tf1_step(params, grads, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    """[TF1] tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps). t starts at 1."""
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    with torch.no_grad():
        for p_, g, m_, v_ in zip(params, grads, m, v):
            m_.mul_(b1).add_((1 - b1) * g)
            v_.mul_(b2).add_((1 - b2) * g * g)
            p_.sub_(lr_t * m_ / (torch.sqrt(v_) + eps))


def piecewise_constant(x, boundaries, values):
    """[TF1] tf.train.piecewise_constant: x <= b0 -> v0; b0 < x <= b1 -> v1; ... (main.py:492)."""
    for b, v in zip(boundaries, values):
        if x <= b:
            return v
    return values[-1]


def lr_schedule(lr, n_epochs, n_train, batch_size):
    """main.py:468-470."""
    n_updates_total = n_epochs * n_train // batch_size
    bounds = [round(0.7 * n_updates_total), round(0.8 * n_updates_total), round(0.9 * n_updates_total)]
    return bounds, [lr, lr / 2, lr / 5, lr / 10]


# ----------------------------------------------------------------------------------------------------------
# synthetic inputs - shared by tests and the cpu_baseline leg
# ----------------------------------------------------------------------------------------------------------
def synthetic_labels(B, H, W, n_ch, rng):
    """3x3 binomial blob at a uniformly random interior (row, col) per channel (data.py:112-114,180-188)."""
    k = np.outer([1, 2, 1], [1, 2, 1]).astype(np.float32) / 16
    y = np.zeros([B, H, W, n_ch], dtype=np.float32)
    for b in range(B):
        for c in range(n_ch):
            r, q = int(rng.integers(1, H - 1)), int(rng.integers(1, W - 1))
            y[b, r - 1:r + 2, q - 1:q + 2, c] = k
    return y


def synthetic_pairwise(names, n_joints, H, W, rng):
    """Non-negative, sum 1, 9x9-binomial-smoothed histogram of 4000 displacements ~ N(0,(H/6)^2) centred at (H,W)."""
    from scipy import signal
    c = np.array([[1, 8, 28, 56, 70, 56, 28, 8, 1]], dtype=np.float64) / 256
    kern = c.T @ c
    out = {}
    for jn in names[:n_joints]:
        for cn in names:
            if cn == jn:
                continue
            pd = np.zeros([2 * H, 2 * W])
            mu = rng.normal(0, H / 8, size=2)
            d = np.rint(rng.normal(mu, H / 6, size=(4000, 2))).astype(int)
            r = np.clip(H + d[:, 0], 0, 2 * H - 1)
            q = np.clip(W + d[:, 1], 0, 2 * W - 1)
            np.add.at(pd, (r, q), 1)
            pd /= pd.sum()
            out[jn + '_' + cn] = signal.convolve2d(pd, kern, mode='same')
    return out
