"""Small end-to-end invocation for compute-sanitizer (memcheck / racecheck): debug-width training step on a 64x96 image
plus a 60x90 spatial-model forward/backward.  Run: compute-sanitizer --tool memcheck python tests/sanitize_small.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'joint-cnn-mrf_b200'), os.path.join(ROOT, 'oracle')):
    sys.path.insert(0, p)
import numpy as np
import torch
import jcm
import jcm_oracle as orc

K = 3
gen = torch.Generator().manual_seed(0)
names = orc.JOINT_NAMES[:K] + ['torso']
rng = np.random.default_rng(0)
p = jcm.init_part_detector(K, gen, debug=True)
sm = jcm.PairwiseParams.from_distribution(orc.synthetic_pairwise(names, K, 8, 12, rng), names, K, 8, 12)
ctx = jcm.Context(n_joints=K, joint_names=names, flag_train=True, precision='bf16', debug=True)
tr = jcm.train.Trainer(p, sm, ctx)
x = torch.rand(2, 64, 96, 3, generator=gen).cuda()
y = torch.from_numpy(orc.synthetic_labels(2, 8, 12, K + 1, rng)).cuda()
out = tr.step(x, y)
torch.cuda.synchronize()
print('train step ok, loss', float(out['loss']))
ctx2 = jcm.Context(n_joints=K, joint_names=names, flag_train=False, precision='fp32', debug=True)
o = jcm.tower_forward(x, y, tr.p, tr.sm, ctx2)
torch.cuda.synchronize()
print('fp32 forward ok', float(o['loss_pd']), float(o['loss_sm']))

# ---- round 2: the kernels debug width never reaches.  Full-width channel counts on small maps: CTA-pair forward / data gradient
# (N = 256 tiles, mixed-shape tile plan on a 60x90 map, partial last wave -> N-split tail), two-tile halo mode (N = 128 and 64), CTA-pair
# and tap-group weight gradients, the prefetching glue kernels of the bf16 configuration, batched weight re-pack, fast tap scatter.
from jcm import ops, train as jt
g = torch.Generator().manual_seed(1)


def planes(*shape):
    return ops.Planes(torch.randn(*shape, generator=g).cuda().to(torch.bfloat16), None)


for (B, H, W, Cin, Cout, k) in [(2, 60, 90, 64, 512, 3), (3, 15, 23, 128, 256, 3), (2, 24, 32, 64, 128, 5), (3, 20, 33, 128, 64, 5)]:
    xp, wp = planes(B, H, W, Cin), planes(k * k, Cout, Cin)
    y1 = ops.conv2d_planes(xp, wp, None, Cout, k, relu=True, out_bf16=True)
    y2 = ops.conv2d_planes(xp, wp, None, Cout, k, relu=True, out_bf16=True, variant=15)
    assert torch.equal(y1, y2)
    gp = planes(B, H, W, Cout)
    dw = torch.empty(k * k, Cin, Cout, device='cuda')
    jt.conv2d_wgrad(xp, gp, dw, Cout, k)
torch.cuda.synchronize()
print('conv variants ok')
a = torch.relu(torch.randn(2, 45, 31, 64, generator=g)).cuda().to(torch.bfloat16)
mm, mv = torch.zeros(64, device='cuda'), torch.ones(64, device='cuda')
ss, st = ops.bn_scale_shift(a, torch.ones(64, device='cuda'), torch.zeros(64, device='cuda'), mm, mv, train=True, save=True)
for pool in (False, True):
    pl = ops.bn_apply_pool(a, ss, pool, False)
    dout = torch.randn(pl.shape, generator=g).cuda().to(torch.bfloat16)
    dg, db_, dbias = (torch.empty(64, device='cuda') for _ in range(3))
    jt.bn_relu_bwd(a, dout, ss, st, 1.0, pool, False, dg, db_, dbias)
b1, b2, b3 = (torch.randn(2, h, w, 64, generator=g).cuda().to(torch.bfloat16) for h, w in ((60, 90), (30, 45), (15, 23)))
ops.upsample_avg3(b1, b2, b3, torch.randn(6, 64, generator=g).cuda(), False)
jt.upsample_avg3_bwd(b1, (30, 45), (15, 23))
ops.tap_scatter_planes(torch.randn(2, 20, 30, 7, generator=g).cuda(), 9, False)
ws = [torch.randn(3, 3, 64, 128, generator=g).cuda(), torch.randn(5, 5, 32, 32, generator=g).cuda()]
ops.pack_weights_batch([(w, ops._new_planes((w.shape[0] ** 2, w.shape[3], w.shape[2]), 'cuda', False),
                         ops._new_planes((w.shape[0] ** 2, w.shape[2], w.shape[3]), 'cuda', False)) for w in ws], False)
torch.cuda.synchronize()
print('round-2 kernels ok')

# ---- tensor-core spatial model (centred prior operand, swapped dP GEMM, row-uniform diagonal reduce, staged Toeplitz pack) at the
# shapes that take its different paths: 60x90 (K extent 128, N = 96), a 20x136 map (W > 128: two M tiles in the dP GEMM, 24 column
# groups in the pack, more than 256 diagonals per reduce CTA) and a padded batch of 48 (48-row tile pass of smt_dc_kernel); FFMA form
# of the first shape next to it
for (B, Ks, H, W) in [(3, 3, 60, 90), (2, 2, 20, 136), (48, 2, 12, 20)]:
    nm = orc.JOINT_NAMES[:Ks] + ['torso']
    smp = jcm.PairwiseParams.from_distribution(orc.synthetic_pairwise(nm, Ks, H, W, rng), nm, Ks, H, W)
    hm = torch.softmax(3 * torch.randn(B, H * W, Ks + 1, generator=g), dim=1).reshape(B, H, W, Ks + 1).cuda()
    bn = smp.bn
    ss, st = ops.bn_scale_shift(hm, bn['gamma'], bn['beta'], bn['moving_mean'], bn['moving_variance'], train=True, save=True)
    gout = (torch.randn(B, H, W, Ks, generator=g) / (B * Ks)).cuda()
    for tc in ((True, False) if (H, W) == (60, 90) else (True,)):
        o, wsp = ops.spatial_model_fwd(hm, ss, smp.energies, smp.biases, smp.pair_target, smp.pair_cond, Ks, keep_workspace=True, tensor_core=tc)
        dE, db2 = torch.empty_like(smp.energies), torch.empty_like(smp.biases)
        dgm, dbt = torch.empty(Ks + 1, device='cuda'), torch.empty(Ks + 1, device='cuda')
        jt.spatial_model_bwd(gout, hm, ss, st, True, smp, wsp, dE, db2, dgm, dbt, tensor_core=tc)
        torch.cuda.synchronize()
        assert bool(torch.isfinite(o).all()) and bool(torch.isfinite(dE).all())
print('spatial model (tensor-core / FFMA) ok')
