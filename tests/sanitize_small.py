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
