"""The reference's model-level function surface (main.py) on top of the sm_100a kernels.

Same names, argument order and tensor layouts as /root/reference/main.py (activations NHWC, conv kernels HWIO, pairwise
energies [1,2H,2W,1] / biases [1,H,W,1] keyed '<joint>_<cond>'), with the reference's module globals (`n_joints`,
`joint_names`, `flag_train`, `train_pd`, `hps`, `pairwise_energies`, ...) made explicit as a `Context` plus parameter
dictionaries that use the TensorFlow variable names ('conv1_fullres/weights', '.../BatchNorm/gamma', 'bn_sm/BatchNorm/*',
'energy_<a>_<b>', 'bias_<a>_<b>').  Eager execution on torch CUDA tensors; all arithmetic is in libjcm.so.
"""
import math
import weakref

import numpy as np
import torch

from . import ops

JOINT_NAMES = ['lsho', 'lelb', 'lwri', 'rsho', 'relb', 'rwri', 'lhip', 'rhip', 'nose', 'torso']  # main.py:18

# Packed bf16 operand planes of the conv kernels, shared by every Context (training and evaluation contexts see the same
# variables, as the reference's towers and its inference run share one set of tf.Variables, main.py:555).  An entry is valid
# while the parameter tensor is alive, was not modified through torch (`_version`) and no raw-pointer update happened since
# it was packed (`_PARAM_EPOCH`: Trainer.apply / load_checkpoint update the flat parameter buffer through the C ABI, which
# torch's version counter cannot see, and call params_updated()).
_PACKS = {}
_PARAM_EPOCH = 0


def params_updated():
    """Call after parameters were written through raw pointers (jcm_clip_adam, a checkpoint load): every cached packed
    operand plane is stale from now on, in every Context."""
    global _PARAM_EPOCH
    _PARAM_EPOCH += 1


def register_packed(w, kind, split, planes):
    """Stores externally packed operand planes of `w` as current (Trainer.apply re-packs every kernel in one launch right after the
    optimizer step, into planes that stay resident)."""
    _PACKS[(w.data_ptr(), tuple(w.shape), kind, split)] = (weakref.ref(w), w._version, _PARAM_EPOCH, planes)


def _purge_packs():
    for k in [k for k, e in _PACKS.items() if e[0]() is None]:
        del _PACKS[k]


class Context:
    """The reference's module-level globals (main.py:428-472) as one explicit object."""

    def __init__(self, n_joints=9, joint_names=None, flag_train=False, train_pd=True, precision='fp32', debug=False,
                 lmbd=0.001, use_sm=True, bf16_activations=None, sm_tensor_core=None):
        if precision not in ('fp32', 'bf16'):
            raise ValueError("precision must be 'fp32' (bf16x3 split products) or 'bf16'")
        self.n_joints = n_joints
        self.joint_names = list(joint_names) if joint_names is not None else JOINT_NAMES[:n_joints] + ['torso']
        if len(self.joint_names) != n_joints + 1:
            raise ValueError('joint_names must have n_joints + 1 entries (main.py:119-121 indexes channels by position)')
        self.flag_train = flag_train
        self.train_pd = train_pd
        self.precision = precision
        self.debug = debug
        self.lmbd = lmbd
        self.use_sm = use_sm
        self.bf16_activations = (precision == 'bf16') if bf16_activations is None else bool(bf16_activations)
        self.sm_tensor_core = (precision == 'bf16') if sm_tensor_core is None else bool(sm_tensor_core)

    @property
    def split(self):
        return self.precision == 'fp32'

    @property
    def act_bf16(self):
        """bf16 precision (default on, bf16_activations=False turns it off): the post-ReLU activations kept for batch norm and the
        backward pass are stored in bf16 as well - the conv epilogue writes half the bytes and the BN / backward kernels read them
        with 16-byte loads of 8 channels (measured +1.8 % images/s at batch 64; layers narrower than 64 channels keep fp32).
        fp32 precision: always fp32."""
        return self.bf16_activations and self.precision == 'bf16'

    @property
    def sm_tc(self):
        """bf16 precision (default on, sm_tensor_core=False turns it off): the spatial model's pairwise convolutions run as grouped
        Toeplitz GEMMs on the tensor cores with bf16 operands (prior centred per pair) instead of the fp32 FFMA kernels.
        fp32 precision: the FFMA kernels, unless sm_tensor_core=True is given AND the context is an inference one
        (flag_train=False): the centred tensor-core forward meets the fp32 bound (1e-4 on the logits, measured 2e-7 on the
        reference's priors, 3e-5 on rough ones) - its prior gradient (2.5e-3) does not, so fp32 training always stays on FFMA."""
        return self.sm_tensor_core and (self.precision == 'bf16' or not self.flag_train)

    def packed(self, name, w, kind='fwd'):
        """Packed bf16 operand planes of a conv kernel: resident, re-packed only after the variable changed (see _PACKS)."""
        key = (w.data_ptr(), tuple(w.shape), kind, self.split)
        ent = _PACKS.get(key)
        if ent is not None and ent[0]() is w and ent[1] == w._version and ent[2] == _PARAM_EPOCH:
            return ent[3]
        if kind == 's2d':
            p = ops.pack_weights_s2d(w, self.split)
        elif kind in ('taps', 'taps_dgrad'):
            p = ops.pack_weights_taps(w, self.split, transpose=(kind == 'taps_dgrad'))
        else:
            p = ops.pack_weights(w, self.split, transpose=(kind == 'dgrad'))
        if len(_PACKS) > 256:
            _purge_packs()
        _PACKS[key] = (weakref.ref(w), w._version, _PARAM_EPOCH, p)
        return p


# ----------------------------------------------------------------------------------------------------------
# variables (main.py:138-159, 477-487)
# ----------------------------------------------------------------------------------------------------------
def n_filters(debug=False):
    f = [64, 128, 256, 512, 512]                       # main.py:38
    return [v // 4 for v in f] if debug else f         # main.py:40-41


def weight_variable(shape, fc=False, gen=None, device='cuda'):
    """main.py:138-147: He init, truncated normal (re-drawn outside +-2 sigma).  fc=True: shape = [n_in, n_out] (main.py:143-145)."""
    n_in = shape[0] if fc else shape[0] * shape[1] * shape[2]
    w = torch.empty(shape, dtype=torch.float32)
    torch.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=gen)
    return (w * math.sqrt(2.0 / n_in)).to(device)


def bias_variable(shape, init=0.0, device='cuda'):
    """main.py:150-153."""
    return torch.full(tuple(shape), float(init), dtype=torch.float32, device=device)


def conv_specs(n_joints, debug=False):
    f = n_filters(debug)
    specs = []
    for sfx in ('fullres', 'halfres', 'quarterres'):
        specs += [('conv1_' + sfx, 5, 3, f[0]), ('conv2_' + sfx, 5, f[0], f[1]), ('conv3_' + sfx, 5, f[1], f[2]),
                  ('conv4_' + sfx, 9, f[2], f[3])]
    return specs + [('conv5', 9, f[3], f[4]), ('conv6', 9, f[4], n_joints)]


def init_part_detector(n_joints, gen=None, debug=False, device='cuda'):
    """All variables of `model` under the TF names."""
    p = {}
    for name, k, cin, cout in conv_specs(n_joints, debug):
        p[name + '/weights'] = weight_variable([k, k, cin, cout], gen=gen, device=device)
        p[name + '/biases'] = bias_variable([cout], device=device)
        if name != 'conv6':
            p[name + '/BatchNorm/gamma'] = torch.ones(cout, dtype=torch.float32, device=device)
            p[name + '/BatchNorm/beta'] = torch.zeros(cout, dtype=torch.float32, device=device)
            p[name + '/BatchNorm/moving_mean'] = torch.zeros(cout, dtype=torch.float32, device=device)
            p[name + '/BatchNorm/moving_variance'] = torch.ones(cout, dtype=torch.float32, device=device)
    return p


def load_params(np_dict, device='cuda'):
    """numpy / torch dict (TF names) -> fp32 CUDA parameter dict."""
    return {k: torch.as_tensor(np.asarray(v), dtype=torch.float32).contiguous().to(device) for k, v in np_dict.items()}


class PairwiseParams:
    """pairwise_energies / pairwise_biases of main.py:477-487 stored as two contiguous tensors [P,2H,2W] and [P,H,W]
    (what the fused kernel reads) with dictionary views under the reference names."""

    def __init__(self, energies, biases, keys, joint_names, n_joints, bn):
        self.energies, self.biases, self.keys = energies, biases, list(keys)
        self.joint_names, self.n_joints = list(joint_names), n_joints
        self.bn = bn  # 'gamma', 'beta', 'moving_mean', 'moving_variance' of bn_sm (main.py:112-113)
        tgt = [self.joint_names.index(k.split('_')[0]) for k in self.keys]
        cnd = [self.joint_names.index(k.split('_')[1]) for k in self.keys]
        if sorted(zip(tgt, cnd)) != list(zip(tgt, cnd)):
            raise ValueError('pairs must be sorted by (target, cond): that is the summation order of main.py:114-123')
        self.pair_target = torch.tensor(tgt, dtype=torch.int32, device=energies.device)
        self.pair_cond = torch.tensor(cnd, dtype=torch.int32, device=energies.device)

    @classmethod
    def from_distribution(cls, pairwise_distr, joint_names, n_joints, hm_h, hm_w, device='cuda'):
        """main.py:477-487: E <- float32(pairwise_distribution[key]) reshaped, b <- 1e-5; one pair per (joint, cond != joint)."""
        keys, e = [], []
        for jn in joint_names[:n_joints]:
            for cn in joint_names:
                if cn != jn:
                    keys.append(jn + '_' + cn)
                    a = np.asarray(pairwise_distr[jn + '_' + cn], dtype=np.float32)
                    if a.shape != (2 * hm_h, 2 * hm_w):
                        raise ValueError('pairwise distribution %s has shape %s, expected %s' % (keys[-1], a.shape, (2 * hm_h, 2 * hm_w)))
                    e.append(a)
        energies = torch.from_numpy(np.stack(e)).to(device)
        biases = torch.full((len(keys), hm_h, hm_w), 0.00001, dtype=torch.float32, device=device)
        nch = len(joint_names)
        bn = {'gamma': torch.ones(nch, device=device), 'beta': torch.zeros(nch, device=device),
              'moving_mean': torch.zeros(nch, device=device), 'moving_variance': torch.ones(nch, device=device)}
        return cls(energies, biases, keys, joint_names, n_joints, bn)

    @classmethod
    def from_dict(cls, sm, joint_names, n_joints, device='cuda'):
        """From a dict with the reference variable names (as the oracle's init_spatial_model builds)."""
        keys = [jn + '_' + cn for jn in joint_names[:n_joints] for cn in joint_names if cn != jn]
        f = lambda v: torch.as_tensor(np.asarray(v), dtype=torch.float32)
        energies = torch.stack([f(sm['energy_' + k]).reshape(f(sm['energy_' + k]).shape[1:3]) for k in keys]).contiguous().to(device)
        biases = torch.stack([f(sm['bias_' + k]).reshape(f(sm['bias_' + k]).shape[1:3]) for k in keys]).contiguous().to(device)
        bn = {k: f(sm['bn_sm/BatchNorm/' + k]).contiguous().to(device) for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')}
        return cls(energies, biases, keys, joint_names, n_joints, bn)

    def as_dict(self):
        d = {}
        P, H2, W2 = self.energies.shape
        for i, k in enumerate(self.keys):
            d['energy_' + k] = self.energies[i].view(1, H2, W2, 1)
            d['bias_' + k] = self.biases[i].view(1, H2 // 2, W2 // 2, 1)
        for k, v in self.bn.items():
            d['bn_sm/BatchNorm/' + k] = v
        return d


def get_pairwise_distr(path=None):
    """main.py:297-299.  The pickle shipped with the reference is corrupt; this package ships the
    regenerated table (oracle/pairwise_prior.py, variant 'shipped') as an .npz with the same keys and float64 arrays."""
    import os
    import pickle
    if path is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', 'pairwise_distribution.npz')
    if path.endswith('.npz'):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    with open(path, 'rb') as f:
        return pickle.load(f)


# ----------------------------------------------------------------------------------------------------------
# layers (main.py:128-174)
# ----------------------------------------------------------------------------------------------------------
def _bn_vars(p, name):
    return [p[name + '/BatchNorm/' + k] for k in ('gamma', 'beta', 'moving_mean', 'moving_variance')]


def conv2d(x, W, stride, ctx):
    """main.py:133-135: tf.nn.conv2d(x NHWC fp32, W HWIO, strides [1,stride,stride,1], padding 'SAME') -> fp32 NHWC.
    stride 1 or 2, any channel counts (input channels are zero-padded to a multiple of 16 for the tensor-core operand).
    Stride 2 is the stride-1 SAME convolution sampled at every second position - offset 1 where the input extent is even,
    0 where it is odd: TF's SAME rule pads (k-3)/2 resp. (k-1)/2 before [TF1].  The part detector's own stride-2 layers
    (conv1_*) take the space-to-depth path inside model() and do 1/4 of this work; this is the general stand-alone form."""
    if stride not in (1, 2):
        raise ValueError('conv2d supports strides 1 and 2 (the reference uses no other, main.py:44-72)')
    ops._req(x, torch.float32, 'x')
    ops._req(W, torch.float32, 'W')
    kh, kw, cin, cout = W.shape
    if x.shape[3] != cin:
        raise ValueError('x has %d channels, W expects %d' % (x.shape[3], cin))
    if kh != kw or kh % 2 == 0:
        raise ValueError('square odd kernels only')
    cpad = ops.pad16(cin)
    xp = ops.split_planes(x, ctx.split) if cpad == cin else ops.pad_planes(x, cpad, ctx.split)
    wp = ops.pack_weights(W, ctx.split)
    y = ops.conv2d_planes(xp, wp, None, cout, kh, relu=False, alg_kdim=kh * kw * cin)
    if stride == 2:
        y = ops.subsample2(y, 1 - x.shape[1] % 2, 1 - x.shape[2] % 2)
    return y


def conv_layer(x, size, stride, n_in, n_out, name, p, ctx, last_layer=False):
    """main.py:156-169: relu(conv2d(x, w, stride) + b) followed by batch_norm (BN AFTER the ReLU); last_layer: conv + bias only.
    x fp32 NHWC [B,H,W,n_in]; the variables are p[name + '/weights' | '/biases' | '/BatchNorm/*'] (the reference creates them
    inside the variable scope `name`).  Returns the fp32 NHWC activation.  The TensorBoard side effects (main.py:167-168) are
    not reproduced."""
    w, b = p[name + '/weights'], p[name + '/biases']
    if tuple(w.shape) != (size, size, n_in, n_out):
        raise ValueError('%s/weights has shape %s, expected %s' % (name, tuple(w.shape), (size, size, n_in, n_out)))
    pre = ops.add_bias_relu(conv2d(x, w, stride, ctx), b, relu=not last_layer)
    if last_layer:
        return pre
    return batch_norm(pre, _bn_vars(p, name), ctx)


def batch_norm(x, bn_vars, ctx):
    """main.py:128-130: returns the normalised fp32 tensor."""
    ss = ops.bn_scale_shift(x, *bn_vars, train=ctx.flag_train)
    return ops.bn_apply_pool(x, ss, pool=False, split=False, want_planes=False, want_f32=True)


def max_pool_layer(x, size=2, stride=2):
    """main.py:172-174 (2x2 s2 SAME)."""
    if size != 2 or stride != 2:
        raise ValueError('only the 2x2 stride-2 pooling of the reference is implemented')
    C = x.shape[3]
    one = torch.ones((2, C), dtype=torch.float32, device=x.device)
    one[1].zero_()
    return ops.bn_apply_pool(x, one, pool=True, split=False, want_planes=False, want_f32=True)


# ----------------------------------------------------------------------------------------------------------
# part detector (main.py:29-74)
# ----------------------------------------------------------------------------------------------------------
def model(x, n_joints, p, ctx, tap=None):
    """main.py:29-74.  x [B,H,W,3] fp32 in [0,1] -> logits [B,H/8,W/8,n_joints].
    tap (optional dict) receives the ReLU outputs of every layer ('<name>/relu') and the merged map ('merge')."""
    split, train = ctx.split, ctx.flag_train
    banks = ops.prep_input(x, split)

    def layer(xp, name, ksize, kind='fwd'):
        w, b = p[name + '/weights'], p[name + '/biases']
        a = ops.conv2d_planes(xp, ctx.packed(name, w, kind), b, w.shape[3], ksize, relu=True,
                              alg_kdim=w.shape[0] * w.shape[1] * w.shape[2],
                                  out_bf16=ctx.act_bf16 and w.shape[3] % 64 == 0)   # narrow (--debug) layers keep fp32 activations
        if tap is not None:
            tap[name + '/relu'] = a
        ss = ops.bn_scale_shift(a, *_bn_vars(p, name), train=train)
        return a, ss

    outs = []
    for xp, sfx in zip(banks, ('fullres', 'halfres', 'quarterres')):
        a, ss = layer(xp, 'conv1_' + sfx, ops.S2D_KSIZE, 's2d')
        h = ops.bn_apply_pool(a, ss, True, split)
        a, ss = layer(h, 'conv2_' + sfx, 5)
        h = ops.bn_apply_pool(a, ss, True, split)
        a, ss = layer(h, 'conv3_' + sfx, 5)
        h = ops.bn_apply_pool(a, ss, False, split)
        a, ss = layer(h, 'conv4_' + sfx, 9)
        outs.append((a, ss))
    ss6 = torch.cat([o[1] for o in outs], dim=0).contiguous()
    if tap is not None:
        merged, mf = ops.upsample_avg3(outs[0][0], outs[1][0], outs[2][0], ss6, split, want_f32=True)
        tap['merge'] = mf
    else:
        merged = ops.upsample_avg3(outs[0][0], outs[1][0], outs[2][0], ss6, split)
    a, ss = layer(merged, 'conv5', 9)
    h = ops.bn_apply_pool(a, ss, False, split)
    w6 = p['conv6/weights']
    return ops.conv2d_taps(h, ctx.packed('conv6', w6, 'taps'), p['conv6/biases'], w6.shape[3], 9)


# ----------------------------------------------------------------------------------------------------------
# spatial model (main.py:77-125)
# ----------------------------------------------------------------------------------------------------------
def conv_mrf(A, B):
    """main.py:77-91.  A [1,2H,2W,1], B [b,H,W,1] -> [b,H,W,1]."""
    b, H, W, _ = B.shape
    out = ops.conv_mrf_fwd(A.reshape(2 * H, 2 * W).contiguous(), B.reshape(b, H, W).contiguous())
    return out.view(b, H, W, 1)


def spatial_model(heat_map, sm, ctx):
    """main.py:94-125.  heat_map [B,H,W,K+1] (K soft-maxed part-detector maps + the conditioning channel)."""
    bn = sm.bn
    ss = ops.bn_scale_shift(heat_map, bn['gamma'], bn['beta'], bn['moving_mean'], bn['moving_variance'], train=ctx.flag_train)
    return ops.spatial_model_fwd(heat_map, ss, sm.energies, sm.biases, sm.pair_target, sm.pair_cond, sm.n_joints, tensor_core=ctx.sm_tc)


# ----------------------------------------------------------------------------------------------------------
# heads (main.py:195-240, evaluation.py)
# ----------------------------------------------------------------------------------------------------------
def spatial_softmax(hm):
    """main.py:212-217."""
    return ops.spatial_softmax(hm)


def softmax_cross_entropy(hm1, hm2):
    """main.py:220-240: logits hm1 [B,H,W,K], labels hm2 [B,H,W,>=K] -> scalar (0-d tensor)."""
    return ops.softmax_ce(hm1, hm2)[0][0]


def weight_decay(var_pattern, variables):
    """main.py:195-205: sum of tf.nn.l2_loss(v) = sum(v^2)/2 over the variables whose name contains var_pattern.
    variables: dict name -> fp32 CUDA tensor (the reference walks tf.global_variables()).  Returns a 0-d CUDA tensor."""
    sel = [v for k, v in variables.items() if var_pattern in k]
    if not sel:
        raise ValueError('no variable name contains %r' % (var_pattern,))    # tf.add_n([]) raises as well
    return ops.l2_loss_sum(sel)


def average_gradients(tower_grads):
    """main.py:243-267: tower_grads = [[(grad, var), ...] per tower] (towers of THIS process, tensors on one device) ->
    [(mean grad, var of the first tower), ...].  Across processes (one per GPU) the same mean is the NCCL all-reduce of
    jcm.train.Trainer.reduce_gradients."""
    out = []
    for grad_and_vars in zip(*tower_grads):
        out.append((ops.tower_mean([g for g, _ in grad_and_vars]), grad_and_vars[0][1]))
    return out


def grad_renorm(gs_vs, norm):
    """main.py:302-309: tf.clip_by_global_norm over the gradients of the (grad, var) list -> list of (clipped grad, var)."""
    gs_vs = list(gs_vs)
    clipped = ops.clip_by_global_norm([g for g, _ in gs_vs], norm)
    return [(c, v) for c, (_, v) in zip(clipped, gs_vs)]


def get_joints_coords(hm):
    """evaluation.py:15-24 -> int64 [B,2,K] (row, col)."""
    return ops.argmax_hw(hm).to(torch.int64)


def det_rate(heat_map_pred, heat_map_target, normalized_radius=10, joints='all'):
    """evaluation.py:4-37 (keeps the reference's lhip_idx, rsho_idx = 0, 7); argmax on the GPU kernel, the 2xK distance
    arithmetic on the tiny coordinate tensors."""
    lhip_idx, rsho_idx = 0, 7
    pred = get_joints_coords(heat_map_pred).to(torch.float32)
    true = get_joints_coords(heat_map_target).to(torch.float32)
    torso = torch.norm(true[:, :, lhip_idx] - true[:, :, rsho_idx], dim=1, keepdim=True)
    nd = torch.norm(pred - true, dim=1) * 100 / torso
    if joints != 'all':
        nd = torch.stack([nd[:, j] for j in joints], dim=1)
    return 100 * (nd <= normalized_radius).to(torch.float32).mean()


# ----------------------------------------------------------------------------------------------------------
# tower forward (main.py:522-541)
# ----------------------------------------------------------------------------------------------------------
def tower_forward(x, hm_target, p, sm, ctx, tap=None):
    """Forward of one tower: part detector -> spatial softmax -> (concat GT conditioning channel) -> spatial model ->
    spatial softmax; both cross-entropies.  Returns a dict of tensors."""
    K = ctx.n_joints
    logit_pd = model(x, K, p, ctx, tap=tap)
    hm_pd = spatial_softmax(logit_pd)
    out = dict(logit_pd=logit_pd, hm_pd=hm_pd)
    if ctx.use_sm:
        cat = torch.cat([hm_pd, hm_target[:, :, :, K:]], dim=3).contiguous()    # main.py:528
        logit_sm = spatial_model(cat, sm, ctx)
        out.update(logit_sm=logit_sm, hm_sm=spatial_softmax(logit_sm), hm_cat=cat)
    else:
        out.update(logit_sm=logit_pd, hm_sm=hm_pd)
    if hm_target is not None:
        out['loss_pd'] = softmax_cross_entropy(logit_pd, hm_target)
        out['loss_sm'] = softmax_cross_entropy(out['logit_sm'], hm_target)
    return out


def eval_error(X, Y, p, sm, ctx, batch_size, det_radius=10, joints_to_eval=(2,)):
    """main.py:275-283: loss_pd, loss_sm, det_rate_pd, det_rate_sm averaged over the full batches of a dataset in inference mode
    (flag_train False).  X [n,H,W,3], Y [n,h,w,K+1]: CPU (pinned or not) or CUDA tensors; incomplete last batches are dropped as
    `get_next_batch` does (main.py:184-192).  joints_to_eval / det_radius: main.py:455-456 (wrist, r = 10)."""
    if ctx.flag_train:
        raise ValueError('eval_error runs in inference mode: pass a Context with flag_train=False')
    n_batches = len(X) // batch_size
    if n_batches == 0:
        raise ValueError('dataset smaller than one batch')
    tot = torch.zeros(4, dtype=torch.float64)
    dev = sm.energies.device
    K = ctx.n_joints
    for i in range(n_batches):
        xb = X[i * batch_size:(i + 1) * batch_size].to(dev, non_blocking=True).contiguous()
        yb = Y[i * batch_size:(i + 1) * batch_size].to(dev, non_blocking=True).contiguous()
        out = tower_forward(xb, yb, p, sm, ctx)
        dr_pd = det_rate(out['hm_pd'], yb[..., :K].contiguous(), normalized_radius=det_radius, joints=list(joints_to_eval))
        dr_sm = det_rate(out['hm_sm'], yb[..., :K].contiguous(), normalized_radius=det_radius, joints=list(joints_to_eval))
        tot += torch.stack([out['loss_pd'].double().cpu(), out['loss_sm'].double().cpu(), dr_pd.double().cpu(), dr_sm.double().cpu()])
    return tuple(float(v) for v in tot / n_batches)
