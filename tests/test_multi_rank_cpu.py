"""World-size-2 tests of the N>1 host path on CPU (gloo): the flat gradient buffer layout and its single all-reduce
(reference main.py:243-267 gradient mean), replica consistency, and bench.py's rank handling in the reference arm.
No kernels are launched here (there is no GPU); the all-reduce + kernels together run in the -m gpu tests / bench."""
import json
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, os.path.join(ROOT, 'joint-cnn-mrf_b200'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import jcm
    K = 3
    names = jcm.JOINT_NAMES[:K] + ['torso']
    gen = torch.Generator().manual_seed(0)                       # identical parameters on every replica
    p = jcm.init_part_detector(K, gen, debug=True, device='cpu')
    distr = {a + '_' + b: torch.rand(16, 24, generator=gen).numpy() for a in names[:K] for b in names if a != b}
    sm = jcm.PairwiseParams.from_distribution(distr, names, K, 8, 12, device='cpu')
    ctx = jcm.Context(n_joints=K, joint_names=names, flag_train=True, precision='bf16', debug=True)
    tr = jcm.train.Trainer(p, sm, ctx, world_size=world)
    # every trainable variable is a view into the flat buffer, offsets 16-byte aligned, weights first (decayed prefix)
    n_train = sum(v.numel() for k, v in p.items() if 'moving_' not in k) + sm.energies.numel() + sm.biases.numel() + 2 * (K + 1)
    assert tr.n >= n_train and tr.n - n_train < 4 * (len(tr.g) + 1)
    assert all(v.data_ptr() % 16 == 0 for v in tr.g.values())
    assert tr.n_decay == sum(v.numel() for k, v in p.items() if k.endswith('/weights'))
    for k in tr.g:
        tr.g[k].fill_(float(rank + 1))                           # "this replica's gradient"
    tr.g['conv1_fullres/weights'].view(-1)[0] = 10.0 * (rank + 1)
    # the buckets partition the flat buffer: every element is reduced exactly once, kernels of conv5/conv6 first
    cover = torch.zeros(tr.n, dtype=torch.int32)
    for ranges in tr.buckets:
        for lo, hi in ranges:
            cover[lo:hi] += 1
    assert bool((cover == 1).all())
    lo0, hi0 = tr.buckets[0][0]
    assert hi0 == tr.n_decay and hi0 - lo0 == tr.g['conv5/weights'].numel() + tr.g['conv6/weights'].numel()
    tr._started = set()
    tr.reduce_bucket(0)              # as the backward pass does after conv5 ...
    tr.reduce_bucket(1)
    tr.reduce_gradients()            # ... the rest, and wait
    want = float(sum(range(1, world + 1)))
    ok = all(bool((v.view(-1)[1:] == want).all()) for v in tr.g.values())
    ok = ok and float(tr.g['conv1_fullres/weights'].view(-1)[0]) == 10.0 * want
    flat = tr.grads.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok = ok and all(torch.equal(gathered[0], g) for g in gathered)   # replicas hold identical reduced gradients
    # BatchNorm moving statistics: every replica updated its copy from its own shard (0.9 * old + 0.1 * batch_i); the reference's
    # towers update ONE shared variable (main.py:555-560), so Trainer.apply combines them - 'mean' and the literal 'towers' form
    old = torch.linspace(0.0, 1.0, tr.moving.numel())
    batch = [old * 0 + 3.0 * (r + 1) + torch.arange(tr.moving.numel()) * 0.01 * (r + 1) for r in range(world)]
    for mode in ('mean', 'towers'):
        tr.bn_moving = mode
        tr._moving_prev.copy_(old)
        tr.moving.copy_(0.9 * old + 0.1 * batch[rank])
        tr.sync_moving_statistics()
        if mode == 'mean':
            want_m = 0.9 * old + 0.1 * sum(batch) / world
        else:
            want_m = old.clone()
            for r in range(world):                                   # tower order
                want_m = 0.9 * want_m + 0.1 * batch[r]
        ok = ok and bool(torch.allclose(tr.moving, want_m, rtol=1e-5, atol=1e-6))
        g2 = [torch.empty_like(tr.moving) for _ in range(world)]
        dist.all_gather(g2, tr.moving)
        ok = ok and all(torch.equal(g2[0], t) for t in g2)
    # replicas built from DIFFERENT initial values start from rank 0's (the towers share one variable set, main.py:555)
    gen2 = torch.Generator().manual_seed(100 + rank)
    p2 = jcm.init_part_detector(K, gen2, debug=True, device='cpu')
    sm2 = jcm.PairwiseParams.from_distribution(distr, names, K, 8, 12, device='cpu')
    tr2 = jcm.train.Trainer(p2, sm2, ctx, world_size=world)
    g3 = [torch.empty_like(tr2.flat) for _ in range(world)]
    dist.all_gather(g3, tr2.flat)
    ok = ok and all(torch.equal(g3[0], t) for t in g3)
    with open(os.path.join(out_dir, 'rank%d.json' % rank), 'w') as f:
        json.dump({'ok': bool(ok), 'n': tr.n}, f)
    dist.destroy_process_group()


def test_flat_gradient_all_reduce_world2_gloo(tmp_path, built_lib):
    world, port = 2, 29731
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [json.load(open(tmp_path / ('rank%d.json' % r))) for r in range(world)]
    assert all(r['ok'] for r in res)
    assert res[0]['n'] == res[1]['n']


def test_reference_arm_only_rank0_prints(tmp_path):
    """bench.py --impl reference under a 2-rank launch: rank 1 exits 0 without work or output."""
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1', MASTER_ADDR='127.0.0.1', MASTER_PORT='29733')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ''
