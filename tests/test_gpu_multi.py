"""-m gpu, needs >= 2 GPUs (skipped otherwise): the data-parallel step over NCCL.  Two replicas fed the SAME batch must end up
with exactly the parameters of a single-replica run (mean of identical gradients == the gradient, bit for bit), which checks
the bucketed, overlapped all-reduce (Trainer.reduce_bucket) against the plain path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (os.path.join(ROOT, 'joint-cnn-mrf_b200'), os.path.join(ROOT, 'oracle')):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    import jcm
    import jcm_oracle as orc
    K = 3
    names = orc.JOINT_NAMES[:K] + ['torso']

    def run(world_size):
        gen = torch.Generator().manual_seed(0)
        rng = np.random.default_rng(0)
        p = jcm.init_part_detector(K, gen, debug=True, device=dev)
        sm = jcm.PairwiseParams.from_distribution(orc.synthetic_pairwise(names, K, 8, 12, rng), names, K, 8, 12, device=dev)
        ctx = jcm.Context(n_joints=K, joint_names=names, flag_train=True, precision='bf16', debug=True, lmbd=0.01)
        tr = jcm.train.Trainer(p, sm, ctx, world_size=world_size, lr=1e-2, optimizer='momentum')
        x = torch.rand(2, 64, 96, 3, generator=gen).to(dev)
        y = torch.from_numpy(orc.synthetic_labels(2, 8, 12, K + 1, rng)).to(dev)
        for _ in range(3):
            tr.step(x, y)
        torch.cuda.synchronize()
        return tr.flat.clone()

    multi = run(world)
    single = run(1)
    ok = torch.equal(multi, single)
    gathered = [torch.empty_like(multi) for _ in range(world)]
    dist.all_gather(gathered, multi)
    ok = ok and all(torch.equal(gathered[0], g) for g in gathered)
    msg = 'ok' if ok else 'mismatch %g' % float((multi - single).abs().max())

    # DIFFERENT batches per replica (the real data-parallel case, main.py:511-517): after three steps every replica must hold the same
    # parameters, optimizer slots AND BatchNorm moving statistics (the towers of the reference update one shared variable,
    # main.py:555-560), and the moving statistics must be the mean-combined ones, not one replica's own
    def run_dp(bn_moving):
        gen = torch.Generator().manual_seed(0)
        rng = np.random.default_rng(0)
        p = jcm.init_part_detector(K, gen, debug=True, device=dev)
        sm = jcm.PairwiseParams.from_distribution(orc.synthetic_pairwise(names, K, 8, 12, rng), names, K, 8, 12, device=dev)
        ctx = jcm.Context(n_joints=K, joint_names=names, flag_train=True, precision='bf16', debug=True, lmbd=0.01)
        tr = jcm.train.Trainer(p, sm, ctx, world_size=world, lr=1e-2, optimizer='adam', bn_moving=bn_moving)
        g2 = torch.Generator().manual_seed(50 + rank)
        x = torch.rand(2, 64, 96, 3, generator=g2).to(dev)
        y = torch.from_numpy(orc.synthetic_labels(2, 8, 12, K + 1, np.random.default_rng(50 + rank))).to(dev)
        own = []
        for _ in range(3):
            tr.forward_backward(x, y)
            own.append(tr.moving.clone())           # this replica's own update, before the replicas are combined
            tr.apply()
        torch.cuda.synchronize()
        return tr, own

    for mode in ('mean', 'towers'):
        tr, own = run_dp(mode)
        for name, buf in (('flat', tr.flat), ('m', tr.m), ('v', tr.v), ('moving', tr.moving)):
            gathered = [torch.empty_like(buf) for _ in range(world)]
            dist.all_gather(gathered, buf)
            if not all(torch.equal(gathered[0], g) for g in gathered):
                ok, msg = False, msg + ' | %s differs across replicas (%s)' % (name, mode)
        if torch.equal(own[-1], tr.moving):
            ok, msg = False, msg + ' | moving statistics were not combined (%s)' % mode
    with open(os.path.join(out_dir, 'rank%d.txt' % rank), 'w') as f:
        f.write('ok' if ok else msg)
    dist.destroy_process_group()


def test_two_replicas_same_batch_equal_single_replica(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    mp.spawn(_worker, args=(2, 29741, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / ('rank%d.txt' % r)).read() for r in range(2)] == ['ok', 'ok']
