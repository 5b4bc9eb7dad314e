"""Thin torch-tensor wrappers over the C ABI (include/jcm.h).  PyTorch is only the buffer/stream provider here:
every function checks its arguments, allocates the outputs with torch.empty on the inputs' device and launches the
sm_100a kernels on torch's current stream.  Nothing in this file computes with torch ops."""
import ctypes

import torch

from ._lib import lib, check

BF16 = torch.bfloat16
F32 = torch.float32
BN_EPS = 1e-3     # [TF1] tf.contrib.layers.batch_norm default (reference main.py:113,129)
BN_DECAY = 0.9    # reference main.py:113,129


def _ptr(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _req_act(t, name):
    """An activation tensor: fp32, or bf16 in the bf16 training configuration."""
    if t is None or not t.is_cuda or t.dtype not in (F32, BF16) or not t.is_contiguous():
        raise ValueError('%s must be a contiguous fp32 or bf16 CUDA tensor (jcm has no CPU path)' % name)
    return int(t.dtype == BF16)


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, dtype, name):
    if t is None:
        raise ValueError('%s is None' % name)
    if not t.is_cuda:
        raise ValueError('%s must be a CUDA tensor (jcm has no CPU path)' % name)
    if t.dtype != dtype:
        raise ValueError('%s must be %s, got %s' % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise ValueError('%s must be contiguous' % name)
    return t


class _ConvProfile:
    """Optional CUDA-event timing of every conv launch (bench.py's roofline leg): algorithmic FLOPs and device time."""

    def __init__(self):
        self.enabled = False
        self.records = []

    def clear(self):
        self.records = []

    def add(self, kernel, flops, e0, e1, tag=None):
        self.records.append((kernel, flops, e0, e1, tag))

    def summary(self, steps=None, kernel=None):
        """Aggregate over the recorded launches of `kernel` (all if None): algorithmic FLOPs / summed CUDA-event time."""
        torch.cuda.synchronize()
        recs = [r for r in self.records if kernel is None or r[0] == kernel]
        ms = sum(r[2].elapsed_time(r[3]) for r in recs)
        fl = sum(r[1] for r in recs)
        n = len(recs)
        steps = steps or 1
        return {'launches': n, 'launches_per_step': n // steps if steps else n, 'ms_total': ms, 'ms_per_step': ms / steps,
                'tflops': (fl / (ms * 1e-3) / 1e12) if ms > 0 else 0.0, 'flops_per_step': fl / steps}

    def largest(self, kernel):
        """The launch shape with the most FLOPs: (flops per launch, mean ms per launch, launches)."""
        torch.cuda.synchronize()
        recs = [r for r in self.records if r[0] == kernel]
        if not recs:
            return None
        fmax = max(r[1] for r in recs)
        big = [r for r in recs if r[1] == fmax]
        ms = sum(r[2].elapsed_time(r[3]) for r in big) / len(big)
        return {'flops': fmax, 'ms': ms, 'launches': len(big), 'tflops': fmax / (ms * 1e-3) / 1e12}

    def by_shape(self, kernel, steps=1):
        """Per launch shape (the tag given at the call site): launches per step, mean ms per launch, algorithmic TFLOP/s."""
        torch.cuda.synchronize()
        groups = {}
        for r in self.records:
            if r[0] != kernel:
                continue
            g = groups.setdefault(r[4] or 'untagged', [0, 0.0, 0.0])
            g[0] += 1
            g[1] += r[2].elapsed_time(r[3])
            g[2] += r[1]
        out = []
        for tag, (n, ms, fl) in groups.items():
            out.append({'shape': tag, 'launches_per_step': n / max(steps, 1), 'ms_per_launch': ms / n, 'ms_per_step': ms / max(steps, 1),
                        'tflops': fl / (ms * 1e-3) / 1e12 if ms > 0 else 0.0})
        return sorted(out, key=lambda d: -d['ms_per_step'])


PROFILE = _ConvProfile()


class Planes:
    """bf16 operand planes (hi, lo) of an NHWC activation or packed weight; lo is None in bf16 mode."""
    __slots__ = ('hi', 'lo', 'shape')

    def __init__(self, hi, lo):
        self.hi, self.lo, self.shape = hi, lo, tuple(hi.shape)


def _new_planes(shape, device, split):
    hi = torch.empty(shape, dtype=BF16, device=device)
    lo = torch.empty(shape, dtype=BF16, device=device) if split else None
    return Planes(hi, lo)


# ------------------------------------------------------------------------------------------------ preparation
def _khw(ksize):
    """kernel extent: int k (square) or (kh, kw)."""
    return (int(ksize), int(ksize)) if isinstance(ksize, int) else (int(ksize[0]), int(ksize[1]))


S2D_KSIZE = (3, 1)   # conv1_* (5x5 stride 2) over the x-folded space-to-depth planes of prep_input: 3 vertical taps x 64 channels


def prep_input(x, split):
    """x [B,H,W,3] fp32 -> (full, half, quarter) x-folded s2d operand planes [B,H/2,W/2,64], [B,H/4,W/4,64], [B,H/8,W/8,64]."""
    _req(x, F32, 'x')
    B, H, W, C = x.shape
    if C != 3:
        raise ValueError('x must have 3 channels')
    outs = [_new_planes((B, H // (2 * s), W // (2 * s), 64), x.device, split) for s in (1, 2, 4)]
    check(lib().jcm_prep_input(_ptr(x), B, H, W, _ptr(outs[0].hi), _ptr(outs[0].lo), _ptr(outs[1].hi), _ptr(outs[1].lo),
                               _ptr(outs[2].hi), _ptr(outs[2].lo), _stream()), 'jcm_prep_input')
    return outs


def pad16(c):
    return (c + 15) // 16 * 16


def pack_weights(w, split, transpose=False):
    """HWIO fp32 [k,k,Cin,Cout] -> packed planes [k*k, Opad, Ipad]."""
    _req(w, F32, 'w')
    k, k2, cin, cout = w.shape
    if k != k2:
        raise ValueError('square kernels only')
    o, i = (cin, cout) if transpose else (cout, cin)
    opad = pad16(o)
    if opad > 256:
        opad = (opad + 255) // 256 * 256
    ipad = pad16(i)
    out = _new_planes((k * k, opad, ipad), w.device, split)
    check(lib().jcm_pack_weights(_ptr(w), k, cin, cout, opad, ipad, int(transpose), _ptr(out.hi), _ptr(out.lo), _stream()),
          'jcm_pack_weights')
    return out


class _PackDesc(ctypes.Structure):
    _fields_ = [('w', ctypes.c_void_p), ('fwd_hi', ctypes.c_void_p), ('fwd_lo', ctypes.c_void_p), ('dgrad_hi', ctypes.c_void_p),
                ('dgrad_lo', ctypes.c_void_p), ('ksize', ctypes.c_int), ('Cin', ctypes.c_int), ('Cout', ctypes.c_int)]


def batch_packable(w):
    """True when jcm_pack_weights_batch can write both layouts of this kernel without padding."""
    k, k2, cin, cout = w.shape
    ok = lambda c: c % 32 == 0 and (c <= 256 or c % 256 == 0)
    return k == k2 and ok(cin) and ok(cout)


def pack_weights_batch(entries, split):
    """entries: list of (w fp32 HWIO, fwd Planes [k*k,Cout,Cin], dgrad Planes [k*k,Cin,Cout]); re-packs all of them in place with
    launches of up to 16 layers."""
    for i in range(0, len(entries), 16):
        chunk = entries[i:i + 16]
        arr = (_PackDesc * len(chunk))()
        for d, (w, fwd, dg) in zip(arr, chunk):
            _req(w, F32, 'w')
            k, _, cin, cout = w.shape
            if tuple(fwd.shape) != (k * k, cout, cin) or tuple(dg.shape) != (k * k, cin, cout):
                raise ValueError('packed planes %s / %s do not match kernel %s' % (fwd.shape, dg.shape, tuple(w.shape)))
            d.w, d.fwd_hi, d.dgrad_hi = w.data_ptr(), fwd.hi.data_ptr(), dg.hi.data_ptr()
            d.fwd_lo = fwd.lo.data_ptr() if split else None
            d.dgrad_lo = dg.lo.data_ptr() if split else None
            d.ksize, d.Cin, d.Cout = k, cin, cout
        check(lib().jcm_pack_weights_batch(arr, len(chunk), _stream()), 'jcm_pack_weights_batch')


def pack_weights_s2d(w, split):
    """conv1 kernels [5,5,3,Cout] -> [3, Cout, 64] (use with ksize = S2D_KSIZE)."""
    _req(w, F32, 'w')
    if tuple(w.shape[:3]) != (5, 5, 3) or w.shape[3] % 16:
        raise ValueError('pack_weights_s2d expects [5,5,3,Cout] with Cout a multiple of 16')
    out = _new_planes((3, w.shape[3], 64), w.device, split)
    check(lib().jcm_pack_weights_s2d(_ptr(w), w.shape[3], _ptr(out.hi), _ptr(out.lo), _stream()), 'jcm_pack_weights_s2d')
    return out


def split_planes(x, split):
    _req(x, F32, 'x')
    out = _new_planes(tuple(x.shape), x.device, split)
    check(lib().jcm_split_planes(_ptr(x), x.numel(), _ptr(out.hi), _ptr(out.lo), _stream()), 'jcm_split_planes')
    return out


def pad_planes(x, cpad, split):
    """fp32 [..., C] -> operand planes [..., cpad] with zero-filled extra channels."""
    _req(x, F32, 'x')
    C = x.shape[-1]
    M = x.numel() // C
    out = _new_planes(tuple(x.shape[:-1]) + (cpad,), x.device, split)
    check(lib().jcm_pad_planes(_ptr(x), M, C, cpad, _ptr(out.hi), _ptr(out.lo), _stream()), 'jcm_pad_planes')
    return out


def subsample2(x, oy, ox):
    """x[:, oy::2, ox::2, :] (stride-2 SAME convolution = sampled stride-1 convolution, see graph.conv2d)."""
    _req(x, F32, 'x')
    B, H, W, C = x.shape
    y = torch.empty((B, (H - oy + 1) // 2, (W - ox + 1) // 2, C), dtype=F32, device=x.device)
    check(lib().jcm_subsample2(_ptr(x), B, H, W, C, oy, ox, _ptr(y), _stream()), 'jcm_subsample2')
    return y


def add_bias_relu(x, bias, relu):
    """[relu](x + b) per channel (main.py:160-162) on an fp32 [..., C] tensor."""
    _req(x, F32, 'x')
    _req(bias, F32, 'bias')
    C = x.shape[-1]
    if bias.numel() != C:
        raise ValueError('bias has %d elements for %d channels' % (bias.numel(), C))
    out = torch.empty_like(x)
    check(lib().jcm_bias_relu(_ptr(x), _ptr(bias), x.numel() // C, C, int(relu), _ptr(out), _stream()), 'jcm_bias_relu')
    return out


def l2_loss_sum(tensors):
    """sum over the tensors of sum(t^2)/2 -> 0-d fp32 CUDA tensor."""
    dev = tensors[0].device
    out = torch.empty((1,), dtype=F32, device=dev)
    nb = max(lib().jcm_optim_blocks(t.numel()) for t in tensors)
    partial = torch.empty((nb,), dtype=F32, device=dev)
    for i, t in enumerate(tensors):
        _req(t, F32, 'variable')
        check(lib().jcm_sumsq(_ptr(t), t.numel(), 0.5, int(i > 0), _ptr(partial), _ptr(out), _stream()), 'jcm_sumsq')
    return out[0]


def tower_mean(grads):
    """mean of same-shaped fp32 CUDA tensors, summed in list order."""
    g0 = _req(grads[0], F32, 'grad')
    for g in grads[1:]:
        _req(g, F32, 'grad')
        if g.shape != g0.shape or g.device != g0.device:
            raise ValueError('tower gradients must have one shape and live on one device')
    out = torch.empty_like(g0)
    arr = (ctypes.c_void_p * len(grads))(*[g.data_ptr() for g in grads])
    check(lib().jcm_tower_mean(arr, len(grads), g0.numel(), _ptr(out), _stream()), 'jcm_tower_mean')
    return out


def clip_by_global_norm(grads, clip):
    """[g * clip / max(global norm, clip)] for a list of fp32 CUDA tensors (tf.clip_by_global_norm)."""
    dev = grads[0].device
    sumsq = torch.empty((1,), dtype=F32, device=dev)
    nb = max(lib().jcm_optim_blocks(g.numel()) for g in grads)
    partial = torch.empty((nb,), dtype=F32, device=dev)
    for i, g in enumerate(grads):
        _req(g, F32, 'grad')
        check(lib().jcm_sumsq(_ptr(g), g.numel(), 1.0, int(i > 0), _ptr(partial), _ptr(sumsq), _stream()), 'jcm_sumsq')
    out = []
    for g in grads:
        o = torch.empty_like(g)
        check(lib().jcm_clip_scale(_ptr(g), g.numel(), _ptr(sumsq), float(clip), _ptr(o), _stream()), 'jcm_clip_scale')
        out.append(o)
    return out


# ------------------------------------------------------------------------------------------------ part detector
def conv2d_planes(xp, wp, bias, cout, ksize, relu, naive=False, alg_kdim=None, out_bf16=False, variant=None):
    """xp activation planes [B,H,W,Cin], wp packed weight planes [kh*kw,Cout_pad,Cin] -> fp32 [B,H,W,cout].
    ksize: int (square) or (kh, kw).  alg_kdim: contraction length of the ALGORITHMIC convolution (default kh*kw*Cin; 75 for
    the s2d conv1)."""
    B, H, W, cin = xp.shape
    taps, cout_pad, cin_w = wp.shape
    kh, kw = _khw(ksize)
    if cin_w != cin or taps != kh * kw:
        raise ValueError('weight planes %s do not match activation planes %s (ksize %d)' % (wp.shape, xp.shape, ksize))
    if (xp.lo is None) != (wp.lo is None):
        raise ValueError('activation and weight planes must use the same precision mode')
    if bias is not None:
        _req(bias, F32, 'bias')
    if out_bf16 and naive:
        raise ValueError('the naive test convolution writes fp32 only')
    y = torch.empty((B, H, W, cout), dtype=BF16 if out_bf16 else F32, device=xp.hi.device)
    prof = PROFILE.enabled and not naive
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    if naive:
        check(lib().jcm_debug_conv2d_naive(_ptr(xp.hi), _ptr(xp.lo), _ptr(wp.hi), _ptr(wp.lo), _ptr(bias), _ptr(y), B, H, W, cin, cout,
                                           cout_pad, kh, kw, int(relu), _stream()), 'jcm_debug_conv2d_naive')
    elif variant is not None:
        # tests / measurements: force the kernel variant (bit 0 single-CTA, bit 1 uniform tiles, bit 2 no N-split tail)
        check(lib().jcm_conv2d_fwd_variant(_ptr(xp.hi), _ptr(xp.lo), _ptr(wp.hi), _ptr(wp.lo), _ptr(bias), _ptr(y), int(out_bf16), B, H, W,
                                           cin, cout, cout_pad, kh, kw, int(relu), int(variant), _stream()), 'jcm_conv2d_fwd_variant')
    else:
        check(lib().jcm_conv2d_fwd(_ptr(xp.hi), _ptr(xp.lo), _ptr(wp.hi), _ptr(wp.lo), _ptr(bias), _ptr(y), int(out_bf16), B, H, W, cin,
                                   cout, cout_pad, kh, kw, int(relu), _stream()), 'jcm_conv2d_fwd')
    if prof:
        e1.record()
        PROFILE.add('conv_igemm_kernel', 2.0 * B * H * W * (alg_kdim or kh * kw * cin) * cout, e0, e1,
                    tag='%dx%d %d->%d k%dx%d' % (H, W, cin, cout, kh, kw))
    return y


# ---- tap-expanded path for convolutions with very few output channels (conv6; csrc/taps.cu explains the re-association)
def tap_layout(ksize, cout):
    """(KP, ZC, Npad): channels per tap (cout padded to 4), used Z channels k*k*KP, and the GEMM N extent (16-/256-padded)."""
    kp = (cout + 3) // 4 * 4
    zc = ksize * ksize * kp
    npad = pad16(zc) if zc <= 256 else (zc + 255) // 256 * 256
    return kp, zc, npad


def pack_weights_taps(w, split, transpose=False):
    """HWIO fp32 [k,k,Cin,Cout] -> planes [1, Npad, Cin] (forward) or [1, Cin, Npad] (data gradient), rows n = tap*KP + co."""
    _req(w, F32, 'w')
    k, k2, cin, cout = w.shape
    if k != k2 or cin % 16:
        raise ValueError('pack_weights_taps expects square kernels and Cin a multiple of 16')
    kp, zc, npad = tap_layout(k, cout)
    out = _new_planes((1, cin, npad) if transpose else (1, npad, cin), w.device, split)
    check(lib().jcm_pack_weights_taps(_ptr(w), k, cin, cout, kp, npad, int(transpose), _ptr(out.hi), _ptr(out.lo), _stream()),
          'jcm_pack_weights_taps')
    return out


def conv2d_taps(xp, wz, bias, cout, ksize):
    """[B,H,W,Cin] planes (*) [k,k,Cin,cout] (+ bias), SAME, stride 1, for small cout: 1x1 tcgen05 GEMM to Z [B,H,W,k*k*KP], then
    the tap gather.  Same result as conv2d_planes(..., relu=False) on the directly packed weights."""
    B, H, W, cin = xp.shape
    kp, zc, npad = tap_layout(ksize, cout)
    if tuple(wz.shape) != (1, npad, cin):
        raise ValueError('tap-packed weight planes %s do not match (1, %d, %d)' % (wz.shape, npad, cin))
    z = conv2d_planes(xp, wz, None, zc, 1, relu=False, alg_kdim=ksize * ksize * cin * cout / zc)
    if bias is not None:
        _req(bias, F32, 'bias')
    y = torch.empty((B, H, W, cout), dtype=F32, device=z.device)
    check(lib().jcm_tap_gather(_ptr(z), _ptr(bias), B, H, W, ksize, kp, zc, cout, _ptr(y), _stream()), 'jcm_tap_gather')
    return y


def tap_scatter_planes(g, ksize, split):
    """g fp32 [B,H,W,cout] (gradient w.r.t. the conv output) -> planes Gt [B,H,W,Npad], Gt[q, tap*KP+co] = g[q-(tap-pad), co]."""
    _req(g, F32, 'g')
    B, H, W, cout = g.shape
    kp, zc, npad = tap_layout(ksize, cout)
    out = _new_planes((B, H, W, npad), g.device, split)
    check(lib().jcm_tap_scatter_planes(_ptr(g), B, H, W, ksize, cout, kp, npad, _ptr(out.hi), _ptr(out.lo), _stream()),
          'jcm_tap_scatter_planes')
    return out


def unpack_tap_grad(dwz, ksize, cin, cout, dw):
    kp, zc, npad = tap_layout(ksize, cout)
    check(lib().jcm_unpack_tap_grad(_ptr(dwz), ksize, cin, cout, kp, zc, _ptr(dw), _stream()), 'jcm_unpack_tap_grad')
    return dw


def bn_scale_shift(a, gamma, beta, moving_mean, moving_var, train, update_moving=True, save=False):
    """Per-channel (scale, shift) of tf.contrib.layers.batch_norm for a [.., C] fp32 (or bf16) tensor (batch stats if train)."""
    a_bf16 = _req_act(a, 'a')
    C = a.shape[-1]
    M = a.numel() // C
    for t, n in ((gamma, 'gamma'), (beta, 'beta'), (moving_mean, 'moving_mean'), (moving_var, 'moving_variance')):
        _req(t, F32, n)
    dev = a.device
    ss = torch.empty((2, C), dtype=F32, device=dev)
    saved = torch.empty((2, C), dtype=F32, device=dev) if save else None
    partial = None
    if train:
        nb = lib().jcm_bn_stats_blocks(M, C)
        partial = torch.empty((nb, 2, C), dtype=F32, device=dev)
        check(lib().jcm_bn_stats(_ptr(a), a_bf16, M, C, _ptr(partial), _stream()), 'jcm_bn_stats')
    check(lib().jcm_bn_finalize(_ptr(partial), M, C, _ptr(gamma), _ptr(beta), _ptr(moving_mean), _ptr(moving_var), BN_EPS, BN_DECAY,
                                int(train), int(update_moving), _ptr(ss[0]), _ptr(ss[1]), _ptr(saved[0] if save else None),
                                _ptr(saved[1] if save else None), _stream()), 'jcm_bn_finalize')
    return (ss, saved) if save else ss


def bn_apply_pool(a, ss, pool, split, want_planes=True, want_f32=False):
    a_bf16 = _req_act(a, 'a')
    B, H, W, C = a.shape
    Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if pool else (H, W)
    planes = _new_planes((B, Ho, Wo, C), a.device, split) if want_planes else None
    f32 = torch.empty((B, Ho, Wo, C), dtype=F32, device=a.device) if want_f32 else None
    check(lib().jcm_bn_apply_pool(_ptr(a), a_bf16, _ptr(ss[0]), _ptr(ss[1]), B, H, W, C, int(pool), _ptr(planes.hi if planes else None),
                                  _ptr(planes.lo if planes else None), _ptr(f32), _stream()), 'jcm_bn_apply_pool')
    if want_planes and want_f32:
        return planes, f32
    return planes if want_planes else f32


def upsample_avg3(a1, a2, a3, ss6, split, want_planes=True, want_f32=False):
    flags = {_req_act(t, 'a') for t in (a1, a2, a3)}
    if len(flags) != 1:
        raise ValueError('the three bank outputs must have the same dtype')
    a_bf16 = flags.pop()
    _req(ss6, F32, 'scale_shift')
    B, H, W, C = a1.shape
    planes = _new_planes((B, H, W, C), a1.device, split) if want_planes else None
    f32 = torch.empty((B, H, W, C), dtype=F32, device=a1.device) if want_f32 else None
    check(lib().jcm_upsample_avg3(_ptr(a1), _ptr(a2), _ptr(a3), a_bf16, _ptr(ss6), B, H, W, a2.shape[1], a2.shape[2], a3.shape[1], a3.shape[2],
                                  C, _ptr(planes.hi if planes else None), _ptr(planes.lo if planes else None), _ptr(f32), _stream()),
          'jcm_upsample_avg3')
    if want_planes and want_f32:
        return planes, f32
    return planes if want_planes else f32


# ------------------------------------------------------------------------------------------------ heads
def spatial_softmax(logits):
    _req(logits, F32, 'logits')
    B, H, W, K = logits.shape
    out = torch.empty_like(logits)
    check(lib().jcm_spatial_softmax(_ptr(logits), B, H * W, K, _ptr(out), _stream()), 'jcm_spatial_softmax')
    return out


def softmax_ce(logits, labels, want_lse=False):
    """Returns (loss scalar tensor, per-(n,k) losses[, lse])."""
    _req(logits, F32, 'logits')
    _req(labels, F32, 'labels')
    B, H, W, K = logits.shape
    if tuple(labels.shape[:3]) != (B, H, W) or labels.shape[3] < K:
        raise ValueError('labels %s do not match logits %s' % (tuple(labels.shape), tuple(logits.shape)))
    per = torch.empty((B, K), dtype=F32, device=logits.device)
    lse = torch.empty((B, K), dtype=F32, device=logits.device) if want_lse else None
    loss = torch.empty((1,), dtype=F32, device=logits.device)
    check(lib().jcm_softmax_ce(_ptr(logits), _ptr(labels), B, H * W, K, labels.shape[3], _ptr(per), _ptr(lse), _ptr(loss), _stream()),
          'jcm_softmax_ce')
    return (loss, per, lse) if want_lse else (loss, per)


def argmax_hw(hm):
    _req(hm, F32, 'hm')
    B, H, W, K = hm.shape
    out = torch.empty((B, 2, K), dtype=torch.int32, device=hm.device)
    check(lib().jcm_argmax_hw(_ptr(hm), B, H, W, K, _ptr(out), _stream()), 'jcm_argmax_hw')
    return out


# ------------------------------------------------------------------------------------------------ spatial model
def spatial_model_fwd(heat_map, ss, energies, biases, pair_target, pair_cond, n_joints, keep_workspace=False, tensor_core=False):
    """tensor_core=True: the pairwise convolutions run as grouped Toeplitz GEMMs on the tensor cores with bf16 operands
    (jcm_spatial_model_tc_fwd, the bf16 configuration); False: the fp32 FFMA kernels."""
    _req(heat_map, F32, 'heat_map')
    _req(energies, F32, 'energies')
    _req(biases, F32, 'biases')
    B, H, W, KC = heat_map.shape
    P = energies.shape[0]
    if KC != n_joints + 1:
        raise ValueError('heat_map must have n_joints + 1 channels')
    if tuple(energies.shape) != (P, 2 * H, 2 * W) or tuple(biases.shape) != (P, H, W):
        raise ValueError('energies/biases shapes %s %s do not match heat maps %dx%d' % (tuple(energies.shape), tuple(biases.shape), H, W))
    fn_ws, fn = (lib().jcm_spatial_model_tc_workspace, lib().jcm_spatial_model_tc_fwd) if tensor_core else \
                (lib().jcm_spatial_model_workspace, lib().jcm_spatial_model_fwd)
    nbytes = fn_ws(B, H, W, n_joints, P)
    if nbytes < 0:
        check(-1, 'jcm_spatial_model_tc_workspace')
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=heat_map.device)
    out = torch.empty((B, H, W, n_joints), dtype=F32, device=heat_map.device)
    if PROFILE.enabled:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(fn(_ptr(heat_map), _ptr(ss[0]), _ptr(ss[1]), _ptr(energies), _ptr(biases), _ptr(pair_target),
             _ptr(pair_cond), _ptr(out), _ptr(ws), nbytes, B, H, W, n_joints, P, _stream()),
          'jcm_spatial_model_tc_fwd' if tensor_core else 'jcm_spatial_model_fwd')
    if PROFILE.enabled:
        e1.record()
        PROFILE.add('spatial_model_fwd', 2.0 * P * B * (H + 1) * (W + 1) * H * W, e0, e1)   # algorithmic: P (H+1)(W+1)HW MAC per image
    return (out, ws) if keep_workspace else out


def conv_mrf_fwd(A, Bm):
    """A [2H,2W] fp32, Bm [b,H,W] fp32 -> [b,H,W]."""
    _req(A, F32, 'A')
    _req(Bm, F32, 'B')
    b, H, W = Bm.shape
    if tuple(A.shape) != (2 * H, 2 * W):
        raise ValueError('A must be [2H,2W]')
    nbytes = lib().jcm_spatial_model_workspace(b, H, W, 0, 1)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=A.device)
    out = torch.empty((b, H, W), dtype=F32, device=A.device)
    check(lib().jcm_conv_mrf_fwd(_ptr(A), _ptr(Bm), _ptr(out), _ptr(ws), nbytes, b, H, W, _stream()), 'jcm_conv_mrf_fwd')
    return out


def fma_peak(blocks, iters, packed):
    """Launches the FMA loop; returns FLOPs per launch (time it with CUDA events)."""
    scratch = torch.rand((4096 + blocks * 640,), dtype=F32, device='cuda')
    flops = ctypes.c_double(0.0)
    check(lib().jcm_fma_peak(_ptr(scratch), blocks, iters, int(packed), ctypes.byref(flops), _stream()), 'jcm_fma_peak')
    return flops.value
