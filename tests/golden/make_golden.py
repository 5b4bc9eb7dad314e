"""Generates the committed golden vectors from the CPU oracle (fp64).  Run from the repo root:

    python tests/golden/make_golden.py

Seeded inputs + parameters -> oracle outputs, stored as float32/float64 .npz (small).  tests/test_golden_cpu.py checks
that the oracle still reproduces them; tests/test_gpu_parity.py checks the CUDA path against them.
The reference itself cannot produce vectors (TensorFlow 1.x not installable), so these are oracle-made.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import jcm_oracle as orc  # noqa: E402


def sm_case(B, K, H, W, seed):
    rng = np.random.default_rng(seed)
    names = ['j%d' % i for i in range(K)] + ['torso']
    distr = orc.synthetic_pairwise(names, K, H, W, rng)
    g = torch.Generator().manual_seed(seed)
    hm = torch.softmax(3 * torch.randn(B, H * W, K, generator=g, dtype=torch.float64), dim=1).reshape(B, H, W, K)
    torso = torch.from_numpy(orc.synthetic_labels(B, H, W, 1, rng)).double()
    cat = torch.cat([hm, torso], dim=3).contiguous()
    sm = orc.init_spatial_model(distr, K, H, W, joint_names=names)
    for k, v in sm.items():
        if k.startswith('bias_'):
            v.add_(torch.rand(v.shape, generator=g, dtype=torch.float64) * 0.01)
        if k.startswith('bn_sm'):
            v.add_(torch.rand(v.shape, generator=g, dtype=torch.float64) * 0.2)
    out = {'names': np.array(names), 'heat_map': cat.numpy().astype(np.float32)}
    for k, v in sm.items():
        out['sm/' + k] = v.numpy().astype(np.float32)
    sm32 = {k: torch.from_numpy(out['sm/' + k]).double() for k in sm}   # outputs are for the float32-rounded parameters
    cat32 = torch.from_numpy(out['heat_map']).double()
    for train in (0, 1):
        o = orc.spatial_model(cat32, {k: v.clone() for k, v in sm32.items()}, K, bool(train), joint_names=names)
        out['out_train%d' % train] = o.numpy()
        out['argmax_train%d' % train] = orc.get_joints_coords(orc.spatial_softmax(o)).numpy()
    return out


def model_params(K, seed):
    """Seeded --debug-width part-detector parameters (float32 values), regenerated identically by the tests; the golden
    file stores their checksum instead of 13 MB of weights."""
    gen = torch.Generator().manual_seed(seed)
    p = orc.init_part_detector(K, gen, debug=True)
    for k, v in p.items():
        if 'gamma' in k or 'moving_variance' in k:
            v.add_(torch.rand(v.shape, generator=gen).double() * 0.5)
        if 'beta' in k or 'moving_mean' in k or 'biases' in k:
            v.add_(torch.randn(v.shape, generator=gen).double() * 0.1)
    return {k: v.float() for k, v in p.items()}, gen


def params_checksum(p):
    return np.float64(sum(float(v.double().abs().sum()) * (i + 1) for i, (k, v) in enumerate(sorted(p.items()))))


def model_case(B, H, W, K, seed):
    p, gen = model_params(K, seed)
    x = torch.rand(B, H, W, 3, generator=gen)
    out = {'x': x.numpy(), 'seed': np.int64(seed), 'K': np.int64(K), 'params_checksum': params_checksum(p)}
    p32 = {k: v.double() for k, v in p.items()}
    y = torch.from_numpy(orc.synthetic_labels(B, H // 8, W // 8, K + 1, np.random.default_rng(seed)))
    out['labels'] = y.numpy()
    for train in (0, 1):
        logits = orc.model(x.double(), {k: v.clone() for k, v in p32.items()}, K, bool(train))
        out['logits_train%d' % train] = logits.numpy()
        out['softmax_train%d' % train] = orc.spatial_softmax(logits).numpy()
        out['ce_train%d' % train] = np.float64(orc.softmax_cross_entropy(logits, y.double()[..., :K]))
        out['argmax_train%d' % train] = orc.get_joints_coords(orc.spatial_softmax(logits)).numpy()
    return out


if __name__ == '__main__':
    np.savez_compressed(os.path.join(HERE, 'sm_small.npz'), **sm_case(3, 4, 12, 20, 11))
    np.savez_compressed(os.path.join(HERE, 'sm_tiny_ragged.npz'), **sm_case(5, 2, 7, 9, 12))
    np.savez_compressed(os.path.join(HERE, 'model_debug_small.npz'), **model_case(1, 64, 96, 7, 13))
    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, 'KiB')
