"""Joint training step of the part detector + spatial model (reference main.py:474-577) on the sm_100a kernels.

One process per GPU.  Per step: forward in training mode (batch-norm batch statistics, activations kept), hand-written
backward (tcgen05 data- and weight-gradient GEMMs, fused BN/ReLU/pool backward, spatial-model backward), ONE NCCL
all-reduce of the flat gradient buffer (the reference averages tower gradients on the CPU, main.py:243-267), then global-norm
clipping (main.py:302-309) and TF1-style Adam / Momentum (main.py:501-506,577) as two fused kernels over the flat buffers.

All trainable variables are views into one flat fp32 buffer laid out as
    [conv kernels ('weights': weight-decayed, main.py:195-205)] [conv biases] [BN gamma/beta] [energies] [pairwise biases] [bn_sm gamma/beta]
so weight decay is a prefix, the all-reduce is one call and the optimizer is shape-agnostic.
"""
import ctypes
import math

import torch

from . import ops
from ._lib import lib, check
from .ops import _ptr, _stream, F32, Planes

ADAM_B1, ADAM_B2, ADAM_EPS = 0.9, 0.999, 1e-8   # tf.train.AdamOptimizer defaults (main.py:502)
CLIP_NORM = 4.0                                  # main.py:576


# ------------------------------------------------------------------------------------------------ backward op wrappers
def softmax_ce_bwd(logits, labels, lse, scale):
    B, H, W, K = logits.shape
    d = torch.empty_like(logits)
    check(lib().jcm_softmax_ce_bwd(_ptr(logits), _ptr(labels), _ptr(lse), B, H * W, K, labels.shape[3], float(scale), _ptr(d), _stream()),
          'jcm_softmax_ce_bwd')
    return d


def spatial_softmax_bwd(y, dy, dx, accumulate):
    B, H, W, K = y.shape
    check(lib().jcm_spatial_softmax_bwd(_ptr(y), _ptr(dy), B, H * W, K, dy.shape[3], int(accumulate), _ptr(dx), _stream()),
          'jcm_spatial_softmax_bwd')
    return dx


def bn_relu_bwd(a, dout, ss, saved, dy_scale, pool, split, dgamma, dbeta, dbias, want_f32=False):
    B, H, W, C = a.shape
    Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if pool else (H, W)
    if tuple(dout.shape) != (B, Ho, Wo, C):
        raise ValueError('dout %s does not match layer output %s' % (tuple(dout.shape), (B, Ho, Wo, C)))
    planes = ops._new_planes((B, H, W, C), a.device, split)
    f32 = torch.empty((B, H, W, C), dtype=F32, device=a.device) if want_f32 else None
    nb = lib().jcm_bn_relu_bwd_blocks(B * Ho * Wo, C)
    ws = torch.empty(((4 * nb + 2) * C,), dtype=F32, device=a.device)
    check(lib().jcm_bn_relu_bwd(_ptr(a), ops._req_act(a, 'a'), _ptr(dout), ops._req_act(dout, 'dout'), _ptr(ss[0]), _ptr(ss[1]), _ptr(saved[0]), _ptr(saved[1]), float(dy_scale), B, H, W, C,
                                int(pool), _ptr(planes.hi), _ptr(planes.lo), _ptr(f32), _ptr(dgamma), _ptr(dbeta), _ptr(dbias), _ptr(ws),
                                _stream()), 'jcm_bn_relu_bwd')
    return (planes, f32) if want_f32 else planes


def colsum(x, out):
    C = x.shape[-1]
    M = x.numel() // C
    nb = lib().jcm_bn_stats_blocks(M, C)
    partial = torch.empty((nb, 2, C), dtype=F32, device=x.device)
    check(lib().jcm_colsum(_ptr(x), M, C, _ptr(partial), _ptr(out), _stream()), 'jcm_colsum')
    return out


def upsample_avg3_bwd(dm, shape2, shape3):
    B, H, W, C = dm.shape
    d2 = torch.empty((B, shape2[0], shape2[1], C), dtype=F32, device=dm.device)
    d3 = torch.empty((B, shape3[0], shape3[1], C), dtype=F32, device=dm.device)
    check(lib().jcm_upsample_avg3_bwd(_ptr(dm), ops._req_act(dm, 'dm'), B, H, W, shape2[0], shape2[1], shape3[0], shape3[1], C, _ptr(d2), _ptr(d3), _stream()),
          'jcm_upsample_avg3_bwd')
    return d2, d3


pad_planes = ops.pad_planes


def conv2d_wgrad(xp, gp, dw, cout, ksize, alg_flops=None):
    """xp input planes [B,H,W,Cin], gp output-gradient planes [B,H,W,Gc] -> dw (contiguous fp32 [kh*kw, Cin, cout]) in place.
    ksize: int (square) or (kh, kw)."""
    B, H, W, cin = xp.shape
    gc = gp.shape[3]
    kh, kw = ops._khw(ksize)
    if tuple(gp.shape[:3]) != (B, H, W):
        raise ValueError('gradient planes %s do not match input planes %s' % (gp.shape, xp.shape))
    if dw.numel() != kh * kw * cin * cout or not dw.is_contiguous():
        raise ValueError('dw must be a contiguous [%d,%d,%d] tensor' % (kh * kw, cin, cout))
    nbytes = lib().jcm_conv2d_wgrad_workspace(B, H, W, cin, gc, kh, kw)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dw.device)
    prof = ops.PROFILE.enabled
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib().jcm_conv2d_wgrad(_ptr(xp.hi), _ptr(xp.lo), _ptr(gp.hi), _ptr(gp.lo), _ptr(dw), _ptr(ws), nbytes, B, H, W, cin, gc, cout, cout,
                                 kh, kw, _stream()), 'jcm_conv2d_wgrad')
    if prof:
        e1.record()   # includes the (small) deterministic split reduction kernel that follows the GEMM
        ops.PROFILE.add('conv_wgrad_kernel', alg_flops if alg_flops is not None else 2.0 * B * H * W * kh * kw * cin * cout, e0, e1,
                        tag='%dx%d %d->%d k%dx%d' % (H, W, cin, cout, kh, kw))
    return dw


def unpack_s2d_grad(g9, dw):
    check(lib().jcm_unpack_s2d_grad(_ptr(g9), dw.shape[3], _ptr(dw), _stream()), 'jcm_unpack_s2d_grad')
    return dw


def spatial_model_bwd(g, heat_map, ss, saved, train, sm, fwd_ws, d_energies, d_biases, dgamma, dbeta, tensor_core=False):
    """tensor_core must match the forward call that filled fwd_ws (the two paths keep different workspaces)."""
    B, H, W, KC = heat_map.shape
    K, P = KC - 1, sm.energies.shape[0]
    fn_ws, fn = (lib().jcm_spatial_model_tc_bwd_workspace, lib().jcm_spatial_model_tc_bwd) if tensor_core else \
                (lib().jcm_spatial_model_bwd_workspace, lib().jcm_spatial_model_bwd)
    nbytes = fn_ws(B, H, W, K, P)
    if nbytes < 0:
        check(-1, 'jcm_spatial_model_tc_bwd_workspace')
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=g.device)
    d_hm = torch.empty_like(heat_map)
    prof = ops.PROFILE.enabled
    if prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(fn(_ptr(g), _ptr(heat_map), _ptr(ss[0]), _ptr(ss[1]), _ptr(saved[0] if saved is not None else None),
                                      _ptr(saved[1] if saved is not None else None), int(train), _ptr(sm.energies), _ptr(sm.biases),
                                      _ptr(sm.pair_target), _ptr(sm.pair_cond), _ptr(fwd_ws), _ptr(ws), nbytes, _ptr(d_hm), _ptr(d_energies),
                                      _ptr(d_biases), _ptr(dgamma), _ptr(dbeta), B, H, W, K, P, _stream()),
          'jcm_spatial_model_tc_bwd' if tensor_core else 'jcm_spatial_model_bwd')
    if prof:
        e1.record()
        ops.PROFILE.add('spatial_model_bwd', 4.0 * P * B * (H + 1) * (W + 1) * H * W, e0, e1)   # dL + dP: twice the forward MACs
    return d_hm


# ------------------------------------------------------------------------------------------------ trainer
class Trainer:
    """Owns the flat parameter / gradient / optimizer-state buffers and runs the training step."""

    def __init__(self, p, sm, ctx, world_size=1, lr=0.001, optimizer='adam', n_updates_total=None, bn_moving='mean'):
        """bn_moving: how the BatchNorm moving statistics of the replicas are combined each step.  The reference's towers all
        update ONE shared variable (main.py:555-560), so every replica must end the step with the same values:
          'mean'   - 0.9 * old + 0.1 * mean over replicas of the batch statistic (one tiny all-reduce; equals the reference for
                     one tower and keeps the decay at 0.9 per step whatever the replica count),
          'towers' - the reference's literal behaviour with its N per-tower assign_moving_average ops executed in tower order:
                     0.9^N * old + 0.1 * sum_i 0.9^(N-1-i) * batch_i (the decay compounds with the number of towers)."""
        if optimizer not in ('adam', 'momentum'):
            raise Exception('wrong optimizer')            # same message as main.py:506
        if bn_moving not in ('mean', 'towers'):
            raise ValueError("bn_moving must be 'mean' or 'towers'")
        if not ctx.train_pd and not ctx.use_sm:
            raise ValueError('train_pd=False without the spatial model leaves no trainable variable')
        self.p, self.sm, self.ctx = p, sm, ctx
        self.world_size, self.lr, self.optimizer = world_size, lr, optimizer
        self.bn_moving = bn_moving
        self.t = 0
        self.n_updates_total = n_updates_total
        dev = sm.energies.device
        # conv kernels first (the weight-decayed prefix), ordered by when the backward pass finishes them - LAST finished first - so
        # that the gradient all-reduce can be issued in contiguous buckets while earlier layers are still being differentiated.  The
        # backward pass finishes conv6, conv5, then bank by bank (full, half, quarter) conv4, conv3, conv2, conv1; conv5 (21 M) and the
        # three conv4_* (10.6 M each) hold 94 % of the parameters, so each of them closes a bucket:
        #   [q_small | conv4_q | h_small | conv4_h | f_small | conv4_f | conv5 conv6]      (x_small = conv1_x, conv2_x, conv3_x)
        #    bucket 4   \--- bucket 3 ---/ \--- bucket 2 ---/  bucket 1   bucket 0         (bucket 4 also carries everything else)
        def bank_of(k):
            return 3 if 'quarterres' in k else 2 if 'halfres' in k else 1 if 'fullres' in k else 0

        def slot(k):      # position in the layout above, left to right
            if bank_of(k) == 0:
                return 6
            return {3: 0, 2: 2, 1: 4}[bank_of(k)] + (1 if k.startswith('conv4_') else 0)
        names_w = sorted([k for k in p if k.endswith('/weights')], key=slot)     # stable: spec order inside a slot
        names_rest = [k for k in p if not k.endswith('/weights') and 'moving_' not in k]
        entries = [(k, p[k]) for k in names_w]
        self.n_decay = sum(t.numel() for _, t in entries)
        entries += [(k, p[k]) for k in names_rest]
        entries += [('sm/energies', sm.energies), ('sm/biases', sm.biases), ('sm/gamma', sm.bn['gamma']), ('sm/beta', sm.bn['beta'])]
        # 16-byte aligned offsets so that every view can be a TMA / float4 base
        offs, n = {}, 0
        for k, t in entries:
            offs[k] = n
            n += (t.numel() + 3) // 4 * 4
        if self.n_decay % 4 or any(p[k].numel() % 4 for k in names_w):
            raise ValueError('conv kernel sizes must be multiples of 4 elements')
        self.n = n
        # trainable range of the flat buffer: everything, or (train_pd=False, main.py:129,147,153,443: the part detector's conv
        # kernels, biases and BN gamma/beta are created with trainable=False) only the spatial model's variables
        self.opt_lo = 0 if ctx.train_pd else offs['sm/energies']
        # gradient buckets (element ranges of the flat buffer), in the order the backward pass completes them
        if ctx.train_pd:
            start = {sl: min(offs[k] for k in names_w if slot(k) == sl) for sl in range(7) if any(slot(k) == sl for k in names_w)}
            edge = lambda sl: start.get(sl, min([v for s2, v in start.items() if s2 > sl] + [self.n_decay]))
            self.buckets = [[(edge(6), self.n_decay)], [(edge(5), edge(6))], [(edge(3), edge(5))], [(edge(1), edge(3))],
                            [(0, edge(1)), (self.n_decay, n)]]
            # the layer whose weight gradient completes each bucket (Trainer.forward_backward issues the all-reduce right after it)
            self.bucket_after = {'conv5': 0, 'conv4_fullres': 1, 'conv4_halfres': 2, 'conv4_quarterres': 3}
        else:
            self.buckets = [[(self.opt_lo, n)]]
            self.bucket_after = {}
        self._packs = None          # resident packed operand planes of the regular conv kernels (filled at the first apply())
        self._pending = []          # async all-reduce handles of this step
        self._started = set()       # buckets whose all-reduce has been issued this step
        self._comm_stream = None
        self.flat = torch.zeros(n, dtype=F32, device=dev)
        self.grads = torch.zeros(n, dtype=F32, device=dev)
        self.m = torch.zeros(n, dtype=F32, device=dev)
        self.v = torch.zeros(n, dtype=F32, device=dev)
        self.g = {}
        for k, t in entries:
            view = self.flat[offs[k]:offs[k] + t.numel()].view(t.shape)
            view.copy_(t)
            self.g[k] = self.grads[offs[k]:offs[k] + t.numel()].view(t.shape)
            if k.startswith('sm/'):
                if k == 'sm/energies':
                    sm.energies = view
                elif k == 'sm/biases':
                    sm.biases = view
                else:
                    sm.bn[k[3:]] = view
            else:
                p[k] = view
        # BatchNorm moving statistics: views into ONE small buffer so that the replicas can combine them with one all-reduce
        mov = [(k, p[k]) for k in p if 'moving_' in k] + [('sm/' + k, sm.bn[k]) for k in ('moving_mean', 'moving_variance')]
        self.moving = torch.zeros(sum((t.numel() + 3) // 4 * 4 for _, t in mov), dtype=F32, device=dev)
        o = 0
        for k, t in mov:
            view = self.moving[o:o + t.numel()].view(t.shape)
            view.copy_(t)
            o += (t.numel() + 3) // 4 * 4
            if k.startswith('sm/'):
                sm.bn[k[3:]] = view
            else:
                p[k] = view
        self._moving_prev = torch.empty_like(self.moving) if world_size > 1 else None
        if world_size > 1:
            # the reference's towers share ONE set of variables (main.py:555): start every replica from rank 0's values
            torch.distributed.broadcast(self.flat, 0)
            torch.distributed.broadcast(self.moving, 0)
        nb = lib().jcm_optim_blocks(n)
        self.partial = torch.empty(2 * nb, dtype=F32, device=dev)
        self.stats = torch.zeros(2, dtype=F32, device=dev)

    # ---------------------------------------------------------------------------------------- learning-rate schedule
    def current_lr(self):
        """main.py:468-470,492: piecewise constant, lr/2, lr/5, lr/10 after 70/80/90 % of the updates."""
        if not self.n_updates_total:
            return self.lr
        b = [round(f * self.n_updates_total) for f in (0.7, 0.8, 0.9)]
        vals = [self.lr, self.lr / 2, self.lr / 5, self.lr / 10]
        # tf.train.piecewise_constant(n_iters_tf, ...) is evaluated with the global step BEFORE apply_gradients increments it
        # (main.py:492,577): the k-th update (k = 1, 2, ...) sees x = k - 1.  Call this before self.t is incremented.
        for bound, v in zip(b, vals):
            if self.t <= bound:
                return v
        return vals[-1]

    # ---------------------------------------------------------------------------------------- forward + backward
    def forward_backward(self, x, y, tap=None):
        """Fills self.grads with d(loss_pd + loss_sm)/d(variables) of THIS replica (weight decay is added in apply()).
        tap (optional dict) receives the ReLU output of every layer ('<name>/relu'), as graph.model does."""
        p, sm, ctx, g = self.p, self.sm, self.ctx, self.g
        self._started = set()
        K, split = ctx.n_joints, ctx.split
        B = x.shape[0]
        dev = x.device
        if self._moving_prev is not None:
            self._moving_prev.copy_(self.moving)          # the replicas' updates of this step are combined in apply()
        banks = ops.prep_input(x, split)
        saved = {}
        train_pd = ctx.train_pd

        def fwd_layer(xp, name, ksize, kind='fwd'):
            w, b = p[name + '/weights'], p[name + '/biases']
            a = ops.conv2d_planes(xp, ctx.packed(name, w, kind), b, w.shape[3], ksize, relu=True,
                                  alg_kdim=w.shape[0] * w.shape[1] * w.shape[2],
                                  out_bf16=ctx.act_bf16 and w.shape[3] % 64 == 0)   # narrow (--debug) layers keep fp32 activations
            ss, st = ops.bn_scale_shift(a, p[name + '/BatchNorm/gamma'], p[name + '/BatchNorm/beta'], p[name + '/BatchNorm/moving_mean'],
                                        p[name + '/BatchNorm/moving_variance'], train=True, save=True)
            if train_pd:
                saved[name] = (xp, a, ss, st)
            if tap is not None:
                tap[name + '/relu'] = a
            return a, ss

        outs = []
        sfxs = ('fullres', 'halfres', 'quarterres')
        for xp, sfx in zip(banks, sfxs):
            a, ss = fwd_layer(xp, 'conv1_' + sfx, ops.S2D_KSIZE, 's2d')
            h = ops.bn_apply_pool(a, ss, True, split)
            a, ss = fwd_layer(h, 'conv2_' + sfx, 5)
            h = ops.bn_apply_pool(a, ss, True, split)
            a, ss = fwd_layer(h, 'conv3_' + sfx, 5)
            h = ops.bn_apply_pool(a, ss, False, split)
            a, ss = fwd_layer(h, 'conv4_' + sfx, 9)
            outs.append((a, ss))
        ss6 = torch.cat([o[1] for o in outs], dim=0).contiguous()
        merged = ops.upsample_avg3(outs[0][0], outs[1][0], outs[2][0], ss6, split)
        a5, ss5 = fwd_layer(merged, 'conv5', 9)
        h5 = ops.bn_apply_pool(a5, ss5, False, split)
        w6 = p['conv6/weights']
        logit_pd = ops.conv2d_taps(h5, ctx.packed('conv6', w6, 'taps'), p['conv6/biases'], K, 9)
        hm_pd = ops.spatial_softmax(logit_pd)
        loss_pd, _, lse_pd = ops.softmax_ce(logit_pd, y, want_lse=True)
        inv_bk = 1.0 / (B * K)
        d_logit = softmax_ce_bwd(logit_pd, y, lse_pd, inv_bk)
        loss_sm = loss_pd
        if ctx.use_sm:
            cat = torch.cat([hm_pd, y[:, :, :, K:]], dim=3).contiguous()       # main.py:528
            ss_sm, st_sm = ops.bn_scale_shift(cat, sm.bn['gamma'], sm.bn['beta'], sm.bn['moving_mean'], sm.bn['moving_variance'],
                                              train=True, save=True)
            logit_sm, ws_sm = ops.spatial_model_fwd(cat, ss_sm, sm.energies, sm.biases, sm.pair_target, sm.pair_cond, K, keep_workspace=True,
                                                    tensor_core=ctx.sm_tc)
            loss_sm, _, lse_sm = ops.softmax_ce(logit_sm, y, want_lse=True)
            g_sm = softmax_ce_bwd(logit_sm, y, lse_sm, inv_bk)
            d_cat = spatial_model_bwd(g_sm, cat, ss_sm, st_sm, True, sm, ws_sm, g['sm/energies'], g['sm/biases'], g['sm/gamma'], g['sm/beta'],
                                      tensor_core=ctx.sm_tc)
            if train_pd:
                spatial_softmax_bwd(hm_pd, d_cat, d_logit, accumulate=True)
        else:
            d_logit.mul_(2.0)   # loss_sm == loss_pd when the spatial model is off (main.py:533-536)
        if not train_pd:
            # frozen part detector (train_pd = False, main.py:443): its variables are not trainable, so TF differentiates the loss
            # only w.r.t. the pairwise energies / biases and bn_sm's gamma / beta - nothing flows back through `model`
            return {'loss_pd': loss_pd, 'loss_sm': loss_sm, 'logit_pd': logit_pd}

        # ---- part-detector backward
        # conv6 in its tap-expanded form (csrc/taps.cu): one scatter of the K-channel gradient, then two 1x1 GEMMs
        c5 = w6.shape[2]
        kp, zc, npad = ops.tap_layout(9, K)
        gt6 = ops.tap_scatter_planes(d_logit, 9, split)
        dwz = torch.empty((1, c5, zc), dtype=F32, device=dev)
        conv2d_wgrad(h5, gt6, dwz, zc, 1, alg_flops=2.0 * B * h5.shape[1] * h5.shape[2] * 81 * c5 * K)
        ops.unpack_tap_grad(dwz, 9, c5, K, g['conv6/weights'])
        colsum(d_logit, g['conv6/biases'])
        # data gradients (inputs of the BN / ReLU backward kernels): bf16 in the bf16 configuration, like the activations
        dg_bf16 = lambda c: ctx.act_bf16 and c % 64 == 0
        dh = ops.conv2d_planes(gt6, ctx.packed('conv6', w6, 'taps_dgrad'), None, c5, 1, relu=False, alg_kdim=81 * K, out_bf16=dg_bf16(c5))

        def bwd_layer(name, dout, dy_scale, pool, ksize, need_dx):
            xp, a, ss, st = saved.pop(name)
            w = p[name + '/weights']
            cin, cout = w.shape[2], w.shape[3]
            d_pre = bn_relu_bwd(a, dout, ss, st, dy_scale, pool, split, g[name + '/BatchNorm/gamma'], g[name + '/BatchNorm/beta'],
                                g[name + '/biases'])
            if name.startswith('conv1_'):
                g9 = torch.empty((3, 64, cout), dtype=F32, device=dev)
                conv2d_wgrad(xp, d_pre, g9, cout, ops.S2D_KSIZE, alg_flops=2.0 * B * xp.shape[1] * xp.shape[2] * 75 * cout)
                unpack_s2d_grad(g9, g[name + '/weights'])
            else:
                conv2d_wgrad(xp, d_pre, g[name + '/weights'].view(ksize * ksize, cin, cout), cout, ksize)
            if name in self.bucket_after:
                self.reduce_bucket(self.bucket_after[name])       # this kernel's gradient closes a bucket: overlap its all-reduce
            if need_dx:
                return ops.conv2d_planes(d_pre, ctx.packed(name, w, 'dgrad'), None, cin, ksize, relu=False, out_bf16=dg_bf16(cin))
            return None

        dmerged = bwd_layer('conv5', dh, 1.0, False, 9, True)
        a4_2, a4_3 = outs[1][0], outs[2][0]
        d2, d3 = upsample_avg3_bwd(dmerged, a4_2.shape[1:3], a4_3.shape[1:3])
        for sfx, dout, sc in zip(sfxs, (dmerged, d2, d3), (1.0 / 3.0, 1.0, 1.0)):
            d = bwd_layer('conv4_' + sfx, dout, sc, False, 9, True)
            d = bwd_layer('conv3_' + sfx, d, 1.0, False, 5, True)
            d = bwd_layer('conv2_' + sfx, d, 1.0, True, 5, True)
            bwd_layer('conv1_' + sfx, d, 1.0, True, 5, False)
        return {'loss_pd': loss_pd, 'loss_sm': loss_sm, 'logit_pd': logit_pd}

    # ---------------------------------------------------------------------------------------- optimizer
    def reduce_bucket(self, b):
        """Starts the all-reduce (sum) of gradient bucket b.  On CUDA it runs asynchronously on a side stream that first waits for
        everything queued so far on the current stream (= the kernels that produced the bucket), so NCCL overlaps with the rest
        of the backward pass; reduce_gradients() waits for all of them.  No-op for a single replica."""
        if self.world_size == 1 or b in self._started:
            return
        self._started.add(b)
        dist = torch.distributed
        if self.grads.is_cuda:
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=self.grads.device)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.grads.device))
            self._comm_stream.wait_event(ev)
            with torch.cuda.stream(self._comm_stream):
                for lo, hi in self.buckets[b]:
                    if hi > lo:
                        self._pending.append(dist.all_reduce(self.grads[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        else:
            for lo, hi in self.buckets[b]:
                if hi > lo:
                    dist.all_reduce(self.grads[lo:hi], op=dist.ReduceOp.SUM)

    def reduce_gradients(self):
        """The one exchange step of the path (main.py:243-267 averages tower gradients on the CPU): an all-reduce (sum) of the flat
        gradient buffer over NCCL / NVLink, issued in five buckets as the backward pass completes them; the division by the
        replica count is folded into jcm_grad_prepare.  Buckets not started yet are started here; then all are awaited."""
        if self.world_size > 1:
            for b in range(len(self.buckets)):
                self.reduce_bucket(b)
            for w in self._pending:
                w.wait()                    # the current stream waits for the NCCL work
            self._pending = []
        self._started = set()

    def sync_moving_statistics(self):
        """Combines the replicas' BatchNorm moving-statistic updates of this step (see __init__, bn_moving).  One tiny all-reduce
        (8 K floats); no-op for a single replica."""
        if self.world_size == 1:
            return
        dist = torch.distributed
        N = self.world_size
        if self.bn_moving == 'mean':
            dist.all_reduce(self.moving, op=dist.ReduceOp.SUM)
            self.moving.mul_(1.0 / N)
        else:
            rank = dist.get_rank()
            d = ops.BN_DECAY
            self.moving.sub_(self._moving_prev, alpha=d).mul_(d ** (N - 1 - rank))      # (1 - d) * batch_i * d^(N-1-i)
            dist.all_reduce(self.moving, op=dist.ReduceOp.SUM)
            self.moving.add_(self._moving_prev, alpha=d ** N)

    def apply(self):
        """Gradient mean over replicas (one NCCL all-reduce), weight decay, global-norm clip, Adam / Momentum."""
        self.reduce_gradients()
        self.sync_moving_statistics()
        lr = self.current_lr()          # schedule evaluated at the pre-increment step count, as tf.train.piecewise_constant is
        self.t += 1
        lo, n = self.opt_lo, self.n - self.opt_lo
        n_decay = max(self.n_decay - lo, 0)
        off = ctypes.c_void_p
        ptr = lambda t: off(t.data_ptr() + 4 * lo)
        check(lib().jcm_grad_prepare(ptr(self.grads), ptr(self.flat), n, n_decay, 1.0 / self.world_size, float(self.ctx.lmbd),
                                     _ptr(self.partial), _ptr(self.stats), _stream()), 'jcm_grad_prepare')
        if lo > 0:
            # frozen part detector: its kernels still enter the loss value through lmbd * weight_decay (main.py:541 sums over
            # tf.global_variables(), trainable or not) but receive no gradient
            check(lib().jcm_sumsq(_ptr(self.flat), self.n_decay, 0.5, 0, _ptr(self.partial), _ptr(self.stats[1:2]), _stream()), 'jcm_sumsq')
        if self.optimizer == 'adam':
            lr_t = lr * math.sqrt(1.0 - ADAM_B2 ** self.t) / (1.0 - ADAM_B1 ** self.t)
            check(lib().jcm_clip_adam(ptr(self.flat), ptr(self.grads), ptr(self.m), ptr(self.v), n, _ptr(self.stats), CLIP_NORM,
                                      lr_t, ADAM_B1, ADAM_B2, ADAM_EPS, 0, _stream()), 'jcm_clip_adam')
        else:
            check(lib().jcm_clip_adam(ptr(self.flat), ptr(self.grads), ptr(self.m), ptr(self.v), n, _ptr(self.stats), CLIP_NORM,
                                      lr, 0.9, 0.0, 0.0, 1, _stream()), 'jcm_clip_adam')
        from .graph import params_updated, register_packed
        params_updated()    # the kernels changed the weights through raw pointers: packed operand planes are stale in every Context
        # ... and are re-packed right here, all regular kernels in one launch into resident planes (both layouts from one read of the
        # fp32 master copy); the three conv1_* and conv6 (special layouts, 0.1 % of the weights) re-pack lazily at their next use
        if self.flat.is_cuda and self.ctx.train_pd:
            if self._packs is None:
                self._packs = []
                for k in self.p:
                    w = self.p[k]
                    if k.endswith('/weights') and k != 'conv6/weights' and ops.batch_packable(w):
                        kk, _, cin, cout = w.shape
                        self._packs.append((w, ops._new_planes((kk * kk, cout, cin), w.device, self.ctx.split),
                                            ops._new_planes((kk * kk, cin, cout), w.device, self.ctx.split)))
            if self._packs:
                ops.pack_weights_batch(self._packs, self.ctx.split)
                for w, fwd, dg in self._packs:
                    register_packed(w, 'fwd', self.ctx.split, fwd)
                    register_packed(w, 'dgrad', self.ctx.split, dg)

    def step(self, x, y):
        out = self.forward_backward(x, y)
        self.apply()
        # loss_tower = loss_pd + loss_sm + lmbd * weight_decay (main.py:541); stats[1] holds sum w^2/2 of this step's weights
        out['loss'] = out['loss_pd'] + out['loss_sm'] + self.ctx.lmbd * self.stats[1:2]
        return out
