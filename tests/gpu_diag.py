"""GPU diagnostic script (run under gpurun): exercises every kernel against the oracle and prints per-stage errors and
timings.  Not a pytest file - the pytest parity tests are tests/test_gpu_*.py; this one is for fast triage when a
kernel is wrong (it localises the first diverging stage) and for quick timing sweeps.

    python tests/gpu_diag.py [sections...]     sections: peak conv glue sm model time grad step wgtime convsweep smtime smtc smtck14 (K=14 / 96x128 tensor-core spatial model timing)
"""
import os
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'joint-cnn-mrf_b200'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)

import numpy as np
import torch

import jcm
if os.environ.get('JCM_DIAG_EXP'):
    # measurement build (python joint-cnn-mrf_b200/build.py --experiments): the JCM_CONV_* environment switches and the CTA-pair
    # kernel exist only there; the product library ignores them
    import jcm._lib as _jl
    _jl.LIB_PATH = os.path.join(os.path.dirname(_jl.LIB_PATH), 'libjcm_exp.so')
from jcm import ops
import jcm_oracle as orc
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import pins

dev = 'cuda'


def relerr(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def sec_peak():
    sms = jcm.lib().jcm_sm_count()
    for packed in (0, 1, 2, 3, 4, 5, 6, 7):
        fl = [0.0]

        def run():
            fl[0] = ops.fma_peak(sms * 2 if packed < 2 else sms, 2000, packed)
        best, med = timeit(run)
        print('PEAK packed=%d  %.1f TFLOP/s best, %.1f median (%.3f ms)' % (packed, fl[0] / best / 1e9, fl[0] / med / 1e9, best))


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def sec_conv():
    g = torch.Generator().manual_seed(0)
    cases = [  # B, H, W, Cin, Cout, k
        (1, 16, 24, 64, 64, 5), (2, 20, 33, 16, 64, 3), (1, 16, 24, 32, 32, 5), (1, 15, 23, 128, 256, 9),
        (2, 60, 90, 64, 128, 5), (1, 30, 45, 256, 512, 9), (1, 60, 90, 128, 7, 9), (1, 9, 200, 64, 16, 3)]
    for (B, H, W, Cin, Cout, k) in cases:
        for split in (False, True):
            try:
                x = torch.randn(B, H, W, Cin, generator=g).to(dev)
                w = (torch.randn(k, k, Cin, Cout, generator=g) / np.sqrt(k * k * Cin)).to(dev)
                b = torch.randn(Cout, generator=g).to(dev)
                xp = ops.split_planes(x, split)
                wp = ops.pack_weights(w, split)
                y = ops.conv2d_planes(xp, wp, b, Cout, k, relu=True)
                yn = ops.conv2d_planes(xp, wp, b, Cout, k, relu=True, naive=True)
                torch.cuda.synchronize()
                if split:
                    xr, wr = x.double().cpu(), w.double().cpu()
                else:
                    xr, wr = bf16_round(x).double().cpu(), bf16_round(w).double().cpu()
                ref = torch.relu(orc.conv2d(xr, wr, 1) + b.double().cpu())
                print('CONV B%d %dx%d Cin%d Cout%d k%d split=%d: tc-vs-naive %.2e  tc-vs-oracle %.2e  naive-vs-oracle %.2e' % (
                    B, H, W, Cin, Cout, k, split, relerr(y, yn), relerr(y, ref), relerr(yn, ref)))
            except Exception as e:
                print('CONV case', (B, H, W, Cin, Cout, k, split), 'FAILED:', repr(e))
                traceback.print_exc()
    # the space-to-depth conv1
    try:
        for split in (False, True):
            x = torch.rand(2, 48, 80, 3, generator=g).to(dev)
            w = (torch.randn(5, 5, 3, 64, generator=g) / np.sqrt(75)).to(dev)
            b = torch.randn(64, generator=g).to(dev)
            banks = ops.prep_input(x, split)
            wp = ops.pack_weights_s2d(w, split)
            xd = x.double().cpu() if split else bf16_round(x).double().cpu()
            wd = w.double().cpu() if split else bf16_round(w).double().cpu()
            for bi, step in enumerate((1, 2, 4)):
                y = ops.conv2d_planes(banks[bi], wp, b, 64, ops.S2D_KSIZE, relu=True)
                ref = torch.relu(orc.conv2d(xd[:, ::step, ::step], wd, 2) + b.double().cpu())
                print('CONV1 s2d bank %d split=%d: err %.2e  shape %s' % (bi, split, relerr(y, ref), tuple(y.shape)))
    except Exception as e:
        print('CONV1 FAILED', repr(e))
        traceback.print_exc()


def sec_glue():
    g = torch.Generator().manual_seed(1)
    for (B, H, W, C) in [(2, 45, 31, 64), (1, 15, 23, 512), (3, 10, 12, 8)]:
        a = torch.relu(torch.randn(B, H, W, C, generator=g)).to(dev)
        gamma = (torch.rand(C, generator=g) + 0.5).to(dev)
        beta = torch.randn(C, generator=g).to(dev)
        for train in (True, False):
            mm = torch.randn(C, generator=g).to(dev) * 0.1
            mv = (torch.rand(C, generator=g) + 0.5).to(dev)
            bn = {'gamma': gamma.double().cpu(), 'beta': beta.double().cpu(), 'moving_mean': mm.double().cpu().clone(),
                  'moving_variance': mv.double().cpu().clone()}
            ref = orc.batch_norm(a.double().cpu(), bn, train)
            ss = ops.bn_scale_shift(a, gamma, beta, mm, mv, train=train)
            out = ops.bn_apply_pool(a, ss, False, True, want_planes=True, want_f32=True)
            planes, f32 = out
            rec = planes.hi.float() + planes.lo.float()
            refp = orc.max_pool_layer(ref)
            outp = ops.bn_apply_pool(a, ss, True, False, want_planes=False, want_f32=True)
            print('BN %s train=%d: f32 %.2e planes %.2e pool %.2e moving_mean %.2e moving_var %.2e' % (
                (B, H, W, C), train, relerr(f32, ref), relerr(rec, ref), relerr(outp, refp), relerr(mm, bn['moving_mean']),
                relerr(mv, bn['moving_variance'])))
    # upsample + average
    B, C = 2, 64
    a1 = torch.randn(B, 60, 90, C, generator=g).to(dev)
    a2 = torch.randn(B, 30, 45, C, generator=g).to(dev)
    a3 = torch.randn(B, 15, 23, C, generator=g).to(dev)
    ss6 = torch.randn(6, C, generator=g).to(dev)
    d = lambda t: t.double().cpu()
    r1 = d(a1) * d(ss6[0]) + d(ss6[1])
    r2 = orc.resize_images(d(a2) * d(ss6[2]) + d(ss6[3]), 60, 90)
    r3 = orc.resize_images(d(a3) * d(ss6[4]) + d(ss6[5]), 60, 90)
    ref = (r1 + r2 + r3) / 3
    out = ops.upsample_avg3(a1, a2, a3, ss6, False, want_planes=False, want_f32=True)
    print('UPSAMPLE_AVG3 err %.2e' % relerr(out, ref))
    # softmax / CE / argmax
    B, H, W, K = 3, 60, 90, 7
    logits = (torch.randn(B, H, W, K, generator=g) * 3).to(dev)
    labels = torch.from_numpy(orc.synthetic_labels(B, H, W, K + 1, np.random.default_rng(0))).to(dev)
    sm = ops.spatial_softmax(logits)
    print('SOFTMAX err %.2e' % relerr(sm, orc.spatial_softmax(d(logits))))
    loss, per = ops.softmax_ce(logits, labels)
    print('CE gpu %.6f oracle %.6f' % (float(loss), float(orc.softmax_cross_entropy(d(logits), d(labels)[..., :K]))))
    am = ops.argmax_hw(sm).cpu().long()
    print('ARGMAX equal:', bool((am == orc.get_joints_coords(d(sm))).all()))


def make_sm_inputs(B, K, H, W, seed=0):
    rng = np.random.default_rng(seed)
    names = (orc.JOINT_NAMES[:K] if K <= len(orc.JOINT_NAMES) - 1 else ['j%02d' % i for i in range(K)]) + ['torso']
    if (H, W) == (60, 90):
        distr = jcm.get_pairwise_distr()
    else:
        distr = orc.synthetic_pairwise(names, K, H, W, rng)
    g = torch.Generator().manual_seed(seed)
    hm = torch.softmax((3 * torch.randn(B, H * W, K, generator=g)), dim=1).reshape(B, H, W, K)
    torso = torch.from_numpy(orc.synthetic_labels(B, H, W, 1, rng))
    cat = torch.cat([hm, torso], dim=3).contiguous()
    sm64 = orc.init_spatial_model(distr, K, H, W, joint_names=names)
    # perturb so that nothing is at its symmetric init
    for k, v in sm64.items():
        if k.startswith('bias_'):
            v.add_(torch.rand(v.shape, generator=g).double() * 0.01)
        if k.startswith('bn_sm') and ('gamma' in k or 'beta' in k):
            v.add_(torch.randn(v.shape, generator=g).double() * 0.1)
    return names, cat, sm64


def sec_sm():
    for (B, K, H, W) in [(2, 7, 60, 90), (5, 7, 60, 90), (3, 4, 12, 20), (2, 9, 60, 90)]:
        try:
            names, cat, sm64 = make_sm_inputs(B, K, H, W)
            for train in (False, True):
                ref = orc.spatial_model(cat.double(), {k: v.clone() for k, v in sm64.items()}, K, train, joint_names=names)
                smp = jcm.PairwiseParams.from_dict(sm64, names, K)
                ctx = jcm.Context(n_joints=K, joint_names=names, flag_train=train)
                out = jcm.spatial_model(cat.to(dev), smp, ctx)
                torch.cuda.synchronize()
                am_ref = orc.get_joints_coords(orc.spatial_softmax(ref))
                am = jcm.get_joints_coords(jcm.spatial_softmax(out)).cpu()
                print('SM B%d K%d %dx%d train=%d: err %.2e (max |ref| %.3f) argmax equal %s' % (
                    B, K, H, W, train, relerr(out, ref), float(ref.abs().max()), bool((am == am_ref).all())))
            # conv_mrf alone vs scipy
            from scipy import signal
            A = torch.rand(2 * H, 2 * W, dtype=torch.float64)
            Bm = torch.rand(3, H, W, dtype=torch.float64)
            got = jcm.conv_mrf(A.float().view(1, 2 * H, 2 * W, 1).to(dev), Bm.float().view(3, H, W, 1).to(dev)).cpu()
            refc = orc.conv_mrf(A.view(1, 2 * H, 2 * W, 1), Bm.view(3, H, W, 1))
            sc = signal.convolve2d(A.numpy(), Bm[1].numpy(), 'valid')
            print('CONV_MRF %dx%d err vs oracle %.2e ; oracle-vs-scipy(valid conv, pre-resize row0) %.2e' % (
                H, W, relerr(got, refc), float(np.abs(sc[0, 0] - refc[1, 0, 0, 0].item()))))
        except Exception as e:
            print('SM case', (B, K, H, W), 'FAILED', repr(e))
            traceback.print_exc()


def sec_model():
    for (B, H, W, K, debug, precision) in [(1, 96, 160, 7, True, 'fp32'), (1, 96, 160, 7, True, 'bf16'), (1, 480, 720, 7, False, 'fp32'),
                                           (2, 480, 720, 7, False, 'bf16')]:
        try:
            gen = torch.Generator().manual_seed(3)
            p64 = orc.init_part_detector(K, gen, debug=debug)
            for k, v in p64.items():   # non-trivial BN parameters / moving stats
                if 'gamma' in k or 'moving_variance' in k:
                    v.add_(torch.rand(v.shape, generator=gen).double() * 0.5)
                if 'beta' in k or 'moving_mean' in k or 'biases' in k:
                    v.add_(torch.randn(v.shape, generator=gen).double() * 0.1)
            x = torch.rand(B, H, W, 3, generator=gen)
            for train in (False, True):
                if train and H == 480 and B == 1 and precision == 'fp32':
                    pass
                tap_ref, tap = {}, {}
                t0 = time.time()
                ref = orc.model(x.double(), {k: v.clone() for k, v in p64.items()}, K, train, tap=tap_ref)
                t_or = time.time() - t0
                p = jcm.load_params(p64)
                ctx = jcm.Context(n_joints=K, flag_train=train, precision=precision, debug=debug)
                out = jcm.model(x.to(dev), K, p, ctx, tap=tap)
                torch.cuda.synchronize()
                print('MODEL B%d %dx%d debug=%d %s train=%d: logits err %.2e (oracle %.1fs)' % (B, H, W, debug, precision, train, relerr(out, ref), t_or))
                for name in sorted(tap_ref):
                    if name in tap:
                        print('    %-24s err %.2e' % (name, relerr(tap[name], tap_ref[name])))
                am_ref = orc.get_joints_coords(orc.spatial_softmax(ref))
                am = jcm.get_joints_coords(jcm.spatial_softmax(out)).cpu()
                print('    argmax equal: %s' % bool((am == am_ref).all()))
        except Exception as e:
            print('MODEL case FAILED', repr(e))
            traceback.print_exc()


def sec_time():
    K = 7
    gen = torch.Generator().manual_seed(5)
    p = jcm.init_part_detector(K, gen)
    distr = jcm.get_pairwise_distr()
    names = jcm.JOINT_NAMES[:K] + ['torso']
    smp = jcm.PairwiseParams.from_distribution(distr, names, K, 60, 90)
    for precision in ('bf16', 'fp32'):
        for B in (4, 16):
            ctx = jcm.Context(n_joints=K, flag_train=False, precision=precision)
            x = torch.rand(B, 480, 720, 3, generator=gen).to(dev)
            y = torch.from_numpy(orc.synthetic_labels(B, 60, 90, K + 1, np.random.default_rng(0))).to(dev)
            best, med = timeit(lambda: jcm.model(x, K, p, ctx), n=3, warm=1)
            fl = 2 * 203.718e9 * B
            print('TIME model fwd %s B=%d: %.2f ms  %.1f img/s  %.1f TFLOP/s (algorithmic)' % (precision, B, best, B / best * 1e3, fl / best / 1e9))
            logit = jcm.model(x, K, p, ctx)
            hm = jcm.spatial_softmax(logit)
            cat = torch.cat([hm, y[..., K:]], dim=3).contiguous()
            best, med = timeit(lambda: jcm.spatial_model(cat, smp, ctx), n=5, warm=2)
            fl = 2 * 1.469e9 * B
            print('TIME spatial model fwd B=%d: %.3f ms  %.1f TFLOP/s (algorithmic)' % (B, best, fl / best / 1e9))
            # single layers
            h = ops.split_planes(torch.randn(B, 60, 90, 512, generator=gen).to(dev), ctx.split)
            w5 = ctx.packed('conv5', p['conv5/weights'])
            best, med = timeit(lambda: ops.conv2d_planes(h, w5, p['conv5/biases'], 512, 9, True), n=3, warm=1)
            fl = 2 * 114.6618e9 * B
            print('TIME conv5 %s B=%d: %.2f ms  %.1f TFLOP/s' % (precision, B, best, fl / best / 1e9))


def grad_case(B, H, W, K, debug, precision, use_sm=True, seed=4, perturb=True, pin_relu=True):
    """Returns (oracle grads dict, gpu grads dict, oracle losses, gpu losses) for one training-mode forward+backward."""
    from jcm import train as jtrain
    gen = torch.Generator().manual_seed(seed)
    hm_h, hm_w = H // 8, W // 8
    names = orc.JOINT_NAMES[:K] + ['torso']
    p64 = orc.init_part_detector(K, gen, debug=debug)
    for k, v in p64.items():
        if perturb and 'gamma' in k:
            v.add_(torch.rand(v.shape, generator=gen).double() * 0.5)
        if perturb and ('beta' in k or 'biases' in k):
            v.add_(torch.randn(v.shape, generator=gen).double() * 0.1)
    rng = np.random.default_rng(seed)
    distr = jcm.get_pairwise_distr() if (hm_h, hm_w) == (60, 90) else orc.synthetic_pairwise(names, K, hm_h, hm_w, rng)
    sm64 = orc.init_spatial_model(distr, K, hm_h, hm_w, joint_names=names)
    for k, v in sm64.items():
        if k.startswith('bias_'):
            v.add_(torch.rand(v.shape, generator=gen).double() * 0.01)
        if 'gamma' in k or 'beta' in k:
            v.add_(torch.randn(v.shape, generator=gen).double() * 0.1)
    p32 = {k: v.float() for k, v in p64.items()}
    sm32 = {k: v.float() for k, v in sm64.items()}
    x = torch.rand(B, H, W, 3, generator=gen)
    y = torch.from_numpy(orc.synthetic_labels(B, hm_h, hm_w, K + 1, rng))
    # gpu
    p = jcm.load_params(p32)
    smp = jcm.PairwiseParams.from_dict(sm32, names, K)
    ctx = jcm.Context(n_joints=K, joint_names=names, flag_train=True, precision=precision, debug=debug, use_sm=use_sm)
    tr = jtrain.Trainer(p, smp, ctx)
    tap = {}
    res = tr.forward_backward(x.to(dev), y.to(dev), tap=tap)
    torch.cuda.synchronize()
    # oracle (ReLU on/off pattern pinned to the GPU forward's unless pin_relu=False)
    masks = pins.relu_masks(tap) if pin_relu else None
    psel = pins.pool_select(jcm, tap, tr.p) if pin_relu else None
    po = {k: v.double().clone().requires_grad_('moving_' not in k) for k, v in p32.items()}
    so = {k: v.double().clone().requires_grad_('moving_' not in k) for k, v in sm32.items()}
    out = orc.tower_forward(x.double(), y.double(), po, so, K, True, use_sm=use_sm, lmbd=0.0, relu_masks=masks, joint_names=names,
                            pool_select=psel)
    (out['loss_pd'] + out['loss_sm']).backward()
    ref = {k: v.grad for k, v in po.items() if v.requires_grad}
    ref.update({k: v.grad for k, v in so.items() if v.requires_grad})
    got = {}
    for k in p32:
        if 'moving_' not in k:
            got[k] = tr.g[k]
    if use_sm:
        P, H2, W2 = smp.energies.shape
        for i, key in enumerate(smp.keys):
            got['energy_' + key] = tr.g['sm/energies'][i].view(1, H2, W2, 1)
            got['bias_' + key] = tr.g['sm/biases'][i].view(1, H2 // 2, W2 // 2, 1)
        got['bn_sm/BatchNorm/gamma'] = tr.g['sm/gamma']
        got['bn_sm/BatchNorm/beta'] = tr.g['sm/beta']
    return ref, got, (float(out['loss_pd']), float(out['loss_sm'])), (float(res['loss_pd']), float(res['loss_sm']))


def sec_grad():
    cases = [(2, 96, 160, 4, True, 'fp32', True, 4, True), (2, 96, 160, 4, True, 'bf16', True, 4, True),
             (1, 128, 192, 7, False, 'fp32', False, 4, True), (2, 64, 96, 3, True, 'fp32', True, 5, False)]
    for (B, H, W, K, debug, precision, use_sm, seed, perturb) in cases:
        try:
            ref, got, lo, lg = grad_case(B, H, W, K, debug, precision, use_sm, seed=seed, perturb=perturb)
            print('GRAD B%d %dx%d K%d debug=%d %s use_sm=%d seed=%d perturb=%d: losses oracle %.5f %.5f gpu %.5f %.5f' % (
                (B, H, W, K, debug, precision, use_sm, seed, perturb) + lo + lg))
            worst = 0.0
            agg = {}
            for k in sorted(ref):
                if ref[k] is None:
                    continue
                if k not in got:
                    print('    missing', k)
                    continue
                e = relerr(got[k], ref[k])
                worst = max(worst, e)
                if k.startswith('energy_') or k.startswith('bias_'):
                    kk = k.split('_')[0] + '_*'
                    agg[kk] = max(agg.get(kk, 0.0), e)
                else:
                    print('    %-40s err %.2e  |ref|max %.2e' % (k, e, float(ref[k].abs().max())))
            for kk, e in agg.items():
                print('    %-40s err %.2e (max over pairs)' % (kk, e))
            print('    WORST %.2e' % worst)
        except Exception as e:
            print('GRAD case FAILED', repr(e))
            traceback.print_exc()


def sec_step():
    """One optimizer step at debug width: gradient, clip norm and updated parameters vs the oracle, per variable."""
    from jcm import train as jtrain
    B, H, W, K = 2, 64, 96, 3
    gen = torch.Generator().manual_seed(5)
    names = orc.JOINT_NAMES[:K] + ['torso']
    rng = np.random.default_rng(5)
    p32 = {k: v.float() for k, v in orc.init_part_detector(K, gen, debug=True).items()}
    sm32 = {k: v.float() for k, v in orc.init_spatial_model(orc.synthetic_pairwise(names, K, H // 8, W // 8, rng), K, H // 8, W // 8,
                                                           joint_names=names).items()}
    x = torch.rand(B, H, W, 3, generator=gen)
    y = torch.from_numpy(orc.synthetic_labels(B, H // 8, W // 8, K + 1, rng))
    lmbd, lr = 0.01, 2e-2
    p = jcm.load_params(p32)
    smp = jcm.PairwiseParams.from_dict(sm32, names, K)
    ctx = jcm.Context(n_joints=K, joint_names=names, flag_train=True, precision='fp32', debug=True, lmbd=lmbd)
    tr = jtrain.Trainer(p, smp, ctx, lr=lr, optimizer='momentum')
    po = {k: v.double().clone().requires_grad_('moving_' not in k) for k, v in p32.items()}
    so = {k: v.double().clone().requires_grad_('moving_' not in k) for k, v in sm32.items()}
    tap = {}
    res = tr.forward_backward(x.to(dev), y.to(dev), tap=tap)
    graw = {k: tr.g[k].clone() for k in p32 if 'moving_' not in k}
    masks, psel = pins.relu_masks(tap), pins.pool_select(jcm, tap, tr.p)
    tr.apply()
    out = orc.tower_forward(x.double(), y.double(), po, so, K, True, lmbd=lmbd, relu_masks=masks, joint_names=names, pool_select=psel)
    keys = [k for k, v in po.items() if v.requires_grad]
    g_nowd = torch.autograd.grad(out['loss_pd'] + out['loss_sm'], [po[k] for k in keys], retain_graph=True)
    allv = [v for v in list(po.values()) + list(so.values()) if v.requires_grad]
    g_all = torch.autograd.grad(out['loss'], allv)
    _, gn = orc.grad_renorm(list(g_all), 4.0)
    print('STEP norm oracle %.6f gpu %.6f ; wd term oracle %.6f gpu %.6f' % (gn, float(tr.stats[0]), float(orc.weight_decay(po)), float(tr.stats[1])))
    scale = 4.0 / max(gn, 4.0)
    for k, g0 in zip(keys, g_nowd):
        if k in ('conv2_fullres/weights', 'conv1_fullres/weights'):
            d = (graw[k].double().cpu() - g0).abs()
            print('   ', k, 'max err per cin', [float('%.1e' % v) for v in d.amax((0, 1, 3))])
            print('   ', k, 'max err per cout', [float('%.1e' % v) for v in d.amax((0, 1, 2))])
            print('   ', k, 'max err per tap', [float('%.1e' % v) for v in d.amax((2, 3)).flatten()])
    a1 = tap['conv1_fullres/relu']
    a64 = a1.double().cpu()
    print('    conv1_fullres relu: per-channel fraction > 0', [float('%.3f' % v) for v in (a64 > 0).double().mean((0, 1, 2))])
    bn1 = {'gamma': torch.ones(16, dtype=torch.float64), 'beta': torch.zeros(16, dtype=torch.float64)}
    hb = orc.batch_norm(a64, bn1, True, update=False)
    # ties inside 2x2 pooling windows among positive values
    hp = hb.permute(0, 3, 1, 2)
    win = torch.nn.functional.unfold(hp.reshape(-1, 1, hp.shape[2], hp.shape[3]), 2, stride=2)   # [B*C, 4, L]
    srt = win.sort(dim=1, descending=True).values
    gap = (srt[:, 0] - srt[:, 1])
    apos = torch.nn.functional.unfold(a64.permute(0, 3, 1, 2).reshape(-1, 1, hp.shape[2], hp.shape[3]), 2, stride=2).amax(1) > 0
    print('    pooling windows with top-2 gap < 1e-6 and a positive max: %d of %d; exact ties: %d' % (
        int(((gap < 1e-6) & apos).sum()), gap.numel(), int(((gap == 0) & apos).sum())))
    for k, g0 in zip(keys, g_nowd):
        gfull = g0 + (lmbd * po[k].detach() if 'weights' in k else 0)
        new = po[k].detach() - lr * scale * gfull
        print('    %-36s grad err %.2e  param err %.2e  (|update|max %.2e, |param| max %.2e)' % (
            k, relerr(graw[k], g0), relerr(tr.p[k], new), float((lr * scale * gfull).abs().max()), float(new.abs().max())))


def sec_wgtime():
    """Weight-gradient / forward kernel timings on the dominant layers (bf16 operands)."""
    from jcm import train as jtrain
    gen = torch.Generator().manual_seed(6)
    for (B, H, W, Cin, Cout, k, name) in [(32, 60, 90, 512, 512, 9, 'conv5'), (32, 60, 90, 256, 512, 9, 'conv4_fullres'),
                                          (32, 120, 180, 64, 128, 5, 'conv2_fullres'), (32, 240, 360, 16, 64, 3, 'conv1_fullres(s2d)'),
                                          (32, 30, 45, 256, 512, 9, 'conv4_halfres')]:
        x = ops.split_planes(torch.randn(B, H, W, Cin, generator=gen).to(dev), False)
        g = ops.split_planes(torch.randn(B, H, W, Cout, generator=gen).to(dev), False)
        w = ops.pack_weights((torch.randn(k, k, Cin, Cout, generator=gen) / 30).to(dev), False)
        dw = torch.empty(k * k, Cin, Cout, device=dev)
        fl = 2.0 * B * H * W * k * k * Cin * Cout
        bw, _ = timeit(lambda: jtrain.conv2d_wgrad(x, g, dw, Cout, k), n=3, warm=1)
        bf, _ = timeit(lambda: ops.conv2d_planes(x, w, None, Cout, k, False), n=3, warm=1)
        print('WGTIME %-20s B=%d: wgrad %.3f ms %.0f TFLOP/s | fwd %.3f ms %.0f TFLOP/s' % (name, B, bw, fl / bw / 1e9, bf, fl / bf / 1e9))


def sec_convsweep():
    """Forward conv time per k-block for small-N layers (what bounds them?)."""
    gen = torch.Generator().manual_seed(7)
    for (B, H, W, Cin, Cout, k) in [(32, 120, 180, 64, 64, 5), (32, 120, 180, 64, 128, 5), (32, 120, 180, 64, 256, 5), (32, 120, 180, 128, 64, 5),
                                    (32, 60, 90, 256, 128, 5), (32, 60, 90, 128, 256, 5), (16, 60, 90, 512, 512, 9)]:
        x = ops.split_planes(torch.randn(B, H, W, Cin, generator=gen).to(dev), False)
        w = ops.pack_weights((torch.randn(k, k, Cin, Cout, generator=gen) / 30).to(dev), False)
        fl = 2.0 * B * H * W * k * k * Cin * Cout
        bf, _ = timeit(lambda: ops.conv2d_planes(x, w, None, Cout, k, False), n=3, warm=1)
        tiles = B * ((H + 7) // 8) * ((W + 15) // 16) * max(1, Cout // 256)
        kb = tiles * k * k * (Cin // 64)
        print('CONVSWEEP B%d %dx%d Cin%d Cout%d k%d: %.3f ms %.0f TFLOP/s  ~%.0f clk/k-block (at 1.9 GHz, %d k-blocks/SM)' % (
            B, H, W, Cin, Cout, k, bf, fl / bf / 1e9, bf * 1e-3 * 1.9e9 / (kb / 148.0), kb // 148))


def sec_smtime():
    K = 7
    names = jcm.JOINT_NAMES[:K] + ['torso']
    smp = jcm.PairwiseParams.from_distribution(jcm.get_pairwise_distr(), names, K, 60, 90)
    ctx = jcm.Context(n_joints=K, flag_train=False)
    for B in (16, 64):
        _, cat, _ = make_sm_inputs(B, K, 60, 90)
        cat = cat.to(dev)
        best, med = timeit(lambda: jcm.spatial_model(cat, smp, ctx), n=10, warm=3)
        fl = 2 * 1.469e9 * B
        print('SMTIME spatial model fwd (bn + prep + conv + finish) B=%d: %.3f ms best %.3f median -> %.1f TFLOP/s algorithmic' % (
            B, best, med, fl / best / 1e9))


def sec_smtc():
    """tensor-core spatial model (jcm_spatial_model_tc_*) against the fp32 FFMA kernels on the same inputs, forward and backward."""
    from jcm import train as jt
    for (B, K, H, W) in [(2, 4, 12, 20), (3, 7, 60, 90), (64, 7, 60, 90)]:
        try:
            names, cat, sm64 = make_sm_inputs(B, K, H, W)
            smp = jcm.PairwiseParams.from_dict(sm64, names, K)
            catd = cat.to(dev)
            ss, saved = ops.bn_scale_shift(catd, smp.bn['gamma'], smp.bn['beta'], smp.bn['moving_mean'], smp.bn['moving_variance'],
                                           train=True, save=True)
            g = torch.randn(B, H, W, K, generator=torch.Generator().manual_seed(5)).to(dev) / (B * K)
            res = {}
            for tc in (False, True):
                out, ws = ops.spatial_model_fwd(catd, ss, smp.energies, smp.biases, smp.pair_target, smp.pair_cond, K, keep_workspace=True,
                                                tensor_core=tc)
                dE, db = torch.zeros_like(smp.energies), torch.zeros_like(smp.biases)
                dg, dbt = torch.zeros(K + 1, device=dev), torch.zeros(K + 1, device=dev)
                dhm = jt.spatial_model_bwd(g, catd, ss, saved, True, smp, ws, dE, db, dg, dbt, tensor_core=tc)
                torch.cuda.synchronize()
                res[tc] = dict(out=out, dhm=dhm, dE=dE, db=db, dgamma=dg, dbeta=dbt)
            print('SMTC B%d K%d %dx%d  ' % (B, K, H, W) + '  '.join('%s %.2e' % (k, relerr(res[True][k], res[False][k])) for k in res[True]))
            if B <= 3:
                ref = orc.spatial_model(cat.double(), {k: v.clone() for k, v in sm64.items()}, K, True, joint_names=names)
                print('   vs oracle: ffma %.2e  tc %.2e' % (relerr(res[False]['out'], ref), relerr(res[True]['out'], ref)))
            if B == 64:
                for tc in (False, True):
                    f = lambda: ops.spatial_model_fwd(catd, ss, smp.energies, smp.biases, smp.pair_target, smp.pair_cond, K, keep_workspace=True,
                                                      tensor_core=tc)
                    _, ws = f()
                    dE, db = torch.zeros_like(smp.energies), torch.zeros_like(smp.biases)
                    dg, dbt = torch.zeros(K + 1, device=dev), torch.zeros(K + 1, device=dev)
                    b = lambda: jt.spatial_model_bwd(g, catd, ss, saved, True, smp, ws, dE, db, dg, dbt, tensor_core=tc)
                    print('   tensor_core=%d: fwd %.3f ms  bwd %.3f ms (best of 10)' % (tc, timeit(f, 10, 3)[0], timeit(b, 10, 3)[0]))
        except Exception as e:
            print('SMTC case', (B, K, H, W), 'FAILED', repr(e))
            traceback.print_exc()
        sys.stdout.flush()


def sec_smtck14():
    """timing of the tensor-core spatial model at the K=14 / 96x128 configuration (BASELINE configs[4] per-GPU shape: batch 32)"""
    from jcm import train as jt
    B, K, H, W = 32, 14, 96, 128
    names, cat, sm64 = make_sm_inputs(B, K, H, W)
    smp = jcm.PairwiseParams.from_dict(sm64, names, K)
    catd = cat.to(dev)
    ss, saved = ops.bn_scale_shift(catd, smp.bn['gamma'], smp.bn['beta'], smp.bn['moving_mean'], smp.bn['moving_variance'], train=True, save=True)
    g = torch.randn(B, H, W, K, generator=torch.Generator().manual_seed(5)).to(dev) / (B * K)
    f = lambda: ops.spatial_model_fwd(catd, ss, smp.energies, smp.biases, smp.pair_target, smp.pair_cond, K, keep_workspace=True, tensor_core=True)
    _, ws = f()
    dE, db = torch.zeros_like(smp.energies), torch.zeros_like(smp.biases)
    dg, dbt = torch.zeros(K + 1, device=dev), torch.zeros(K + 1, device=dev)
    b = lambda: jt.spatial_model_bwd(g, catd, ss, saved, True, smp, ws, dE, db, dg, dbt, tensor_core=True)
    print('SMTC K=14 96x128 B=32: fwd %.3f ms  bwd %.3f ms (best of 10)' % (timeit(f, 10, 3)[0], timeit(b, 10, 3)[0]))


def sec_smtc64():
    """one forward + backward of the tensor-core spatial model at the bench shape (for ncu captures of its three GEMM launches)"""
    from jcm import train as jt
    B, K, H, W = 64, 7, 60, 90
    names, cat, sm64 = make_sm_inputs(B, K, H, W)
    smp = jcm.PairwiseParams.from_dict(sm64, names, K)
    catd = cat.to(dev)
    ss, saved = ops.bn_scale_shift(catd, smp.bn['gamma'], smp.bn['beta'], smp.bn['moving_mean'], smp.bn['moving_variance'], train=True, save=True)
    g = torch.randn(B, H, W, K, generator=torch.Generator().manual_seed(5)).to(dev) / (B * K)
    out, ws = ops.spatial_model_fwd(catd, ss, smp.energies, smp.biases, smp.pair_target, smp.pair_cond, K, keep_workspace=True, tensor_core=True)
    dE, db = torch.zeros_like(smp.energies), torch.zeros_like(smp.biases)
    dg, dbt = torch.zeros(K + 1, device=dev), torch.zeros(K + 1, device=dev)
    jt.spatial_model_bwd(g, catd, ss, saved, True, smp, ws, dE, db, dg, dbt, tensor_core=True)
    torch.cuda.synchronize()
    print('SMTC64 done', float(out.abs().max()))


def _interleaved(fns, rounds=7, warm=2):
    """median ms of each callable, timed round-robin so that every candidate sees the same thermal / power-cap state"""
    for _ in range(warm):
        for f in fns:
            f()
    torch.cuda.synchronize()
    ts = [[] for _ in fns]
    for _ in range(rounds):
        for i, f in enumerate(fns):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f()
            e1.record()
            torch.cuda.synchronize()
            ts[i].append(e0.elapsed_time(e1))
    return [float(np.median(t)) for t in ts]


def sec_cta2():
    """N = 256 plain-mode layers: time per launch of the kernel variants (bit 0 single-CTA instead of the CTA pair, bit 1 uniform tile
    grid instead of the mixed-shape plan, bit 2 no N-split tail), batch 64, variants interleaved launch by launch."""
    variants = (0, 4, 6, 1, 7)
    for (B, H, W, Cin, Cout, k) in [(64, 60, 90, 512, 512, 9), (64, 60, 90, 256, 512, 9), (64, 60, 90, 512, 256, 9), (64, 30, 45, 256, 512, 9)]:
        xp = ops.Planes(torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16), None)
        wp = ops.Planes((torch.randn(k * k, Cout, Cin, device=dev) / np.sqrt(k * k * Cin)).to(torch.bfloat16), None)
        fl = 2.0 * B * H * W * k * k * Cin * Cout
        med = _interleaved([(lambda v=v: ops.conv2d_planes(xp, wp, None, Cout, k, relu=True, out_bf16=True, variant=v)) for v in variants])
        print('IGEMM %dx%d Cin%d Cout%d k%d: ' % (H, W, Cin, Cout, k) + '  '.join('v%d %.3f ms (%.0f TF)' % (v, m, fl / m / 1e9) for v, m in zip(variants, med)),
              flush=True)


def sec_halo2():
    """halo-mode layers: default (two M tiles per weight stage) vs one tile per stage (variant 8) vs one tile per stage with tap groups
    for N <= 64 only (variant 16), interleaved, batch 64"""
    for (B, H, W, Cin, Cout, k) in [(64, 120, 180, 64, 128, 5), (64, 60, 90, 256, 128, 5), (64, 120, 180, 128, 64, 5), (64, 60, 90, 128, 64, 5),
                                    (64, 240, 360, 64, 64, (3, 1))]:
        xp = ops.Planes(torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16), None)
        kh, kw = ops._khw(k)
        wp = ops.Planes((torch.randn(kh * kw, Cout, Cin, device=dev) / np.sqrt(kh * kw * Cin)).to(torch.bfloat16), None)
        fl = 2.0 * B * H * W * kh * kw * Cin * Cout
        med = _interleaved([(lambda v=v: ops.conv2d_planes(xp, wp, None, Cout, k, relu=True, out_bf16=True, variant=v)) for v in (0, 8, 16)])
        print('HALO %dx%d Cin%d Cout%d k%s: ' % (H, W, Cin, Cout, k) + '  '.join('v%d %.3f ms (%.0f TF)' % (v, m, fl / m / 1e9) for v, m in zip((0, 8, 16), med)),
              flush=True)


def sec_wgpair():
    """weight gradient: CTA-pair kernel / mixed-shape patch plan variants (bit 0 single-CTA, bit 1 uniform grid), batch 64, interleaved"""
    from jcm import train as jt
    for (B, H, W, Cin, Cout, k) in [(64, 60, 90, 512, 512, 9), (64, 60, 90, 256, 512, 9), (64, 120, 180, 64, 128, 5), (64, 60, 90, 64, 128, 5),
                                    (64, 240, 360, 64, 64, (3, 1))]:
        kh, kw = ops._khw(k)
        xp = ops.Planes(torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16), None)
        gp = ops.Planes(torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16), None)
        dw = torch.empty(kh * kw, Cin, Cout, device=dev)
        fl = 2.0 * B * H * W * kh * kw * Cin * Cout

        def run(v):
            jcm.lib().jcm_debug_set_wgrad_variant(v)
            jt.conv2d_wgrad(xp, gp, dw, Cout, k)
        med = _interleaved([(lambda v=v: run(v)) for v in (0, 1, 2, 4)])
        jcm.lib().jcm_debug_set_wgrad_variant(0)
        print('WGRAD %dx%d Cin%d Cout%d k%s: ' % (H, W, Cin, Cout, k) + '  '.join('v%d %.3f ms (%.0f TF)' % (v, m, fl / m / 1e9) for v, m in zip((0, 1, 2, 4), med)),
              flush=True)


if __name__ == '__main__':
    secs = sys.argv[1:] or ['peak', 'conv', 'glue', 'sm', 'model', 'time']
    print('device:', torch.cuda.get_device_name(0), 'SMs', jcm.lib().jcm_sm_count())
    for s in secs:
        print('=' * 30, s)
        try:
            globals()['sec_' + s]()
        except Exception as e:
            print('SECTION', s, 'FAILED', repr(e))
            traceback.print_exc()
        sys.stdout.flush()
