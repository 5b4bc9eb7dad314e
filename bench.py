#!/usr/bin/env python
"""Headline benchmark of the joint-cnn-mrf hot path on B200 (see BASELINE.json / SURVEY.md section 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload fwd16|train64] [--impl ours|reference]

One process per GPU (torchrun for N>1).  A step is one pass of the hot path over one synthetic batch:
  * fwd16   (BASELINE configs[1]): part detector + spatial model forward (+ both softmax-CE heads), batch 16 per GPU,
            K=7, 720x480, fp32-equivalent arithmetic (bf16x3 split products on the tensor cores), inference-mode BN.
  * train64 (BASELINE configs[2]/[3]): joint training step fwd+bwd + gradient all-reduce + clip + Adam, batch 64 per GPU,
            bf16 tensor-core operands with fp32 accumulation.
Rank 0 prints ONE JSON line (contract in the task statement): value = images/s with inputs resident in HBM,
e2e = the same through the public API with pinned-host inputs copied in and results copied out every step,
roofline = the conv implicit-GEMM kernel's achieved algorithmic TFLOP/s against the measured bf16 peak,
cpu_baseline = the CPU restatement of the reference graph (oracle) timed on this box's host cores.
`--impl reference` times that CPU restatement alone (TensorFlow 1.x cannot be installed: see DESIGN.md).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'joint-cnn-mrf_b200'))

import numpy as np
import torch

# workload -> (per-GPU batch, K, image H, W, training step?, precision)
WORKLOADS = {'fwd16': (16, 7, 480, 720, False, 'fp32'), 'train64': (64, 7, 480, 720, True, 'bf16'),
             'train_k14': (32, 14, 768, 1024, True, 'bf16')}   # BASELINE configs[4]: K=14, 96x128 maps, 32 images per GPU


def synthetic_pairwise(names, K, H, W, rng):
    """Pairwise-prior table for joint sets without FLIC statistics (K=14): non-negative, sum 1, smoothed histogram of 4000 displacements
    ~ N(0, (H/6)^2) centred at (H, W), as float64 [2H,2W] arrays keyed '<joint>_<cond>'."""
    c = np.array([1, 8, 28, 56, 70, 56, 28, 8, 1], dtype=np.float64) / 256
    out = {}
    for jn in names[:K]:
        for cn in names:
            if cn == jn:
                continue
            pd = np.zeros([2 * H, 2 * W])
            mu = rng.normal(0, H / 8, size=2)
            d = np.rint(rng.normal(mu, H / 6, size=(4000, 2))).astype(int)
            np.add.at(pd, (np.clip(H + d[:, 0], 0, 2 * H - 1), np.clip(W + d[:, 1], 0, 2 * W - 1)), 1)
            pd /= pd.sum()
            pd = np.apply_along_axis(lambda v: np.convolve(v, c, mode='same'), 0, pd)      # separable 9x9 binomial smoothing
            out[jn + '_' + cn] = np.apply_along_axis(lambda v: np.convolve(v, c, mode='same'), 1, pd)
    return out
PD_FWD_FLOP = 2 * 203.718e9      # per image, SURVEY Appendix A (K=7)
SM_FWD_FLOP = 2 * 1.469e9        # per image, K^2 (H+1)(W+1)HW MACs


def synthetic_labels(B, H, W, n_ch, rng):
    """3x3 binomial blob per channel at a random interior position (reference data.py:112-114,180-188)."""
    k = np.outer([1, 2, 1], [1, 2, 1]).astype(np.float32) / 16
    y = np.zeros([B, H, W, n_ch], dtype=np.float32)
    for b in range(B):
        for c in range(n_ch):
            r, q = int(rng.integers(1, H - 1)), int(rng.integers(1, W - 1))
            y[b, r - 1:r + 2, q - 1:q + 2, c] = k
    return y


def load_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return dict(bf16=float(p['bf16_tflops']), bf16_sustained=float(p.get('bf16_tflops_sustained', p['bf16_tflops'])),
                    hbm=float(p['hbm_gbs']), source='MEASURED_PEAKS.json')
    except Exception:
        return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle (CPU restatement of the reference TF graph), all host threads
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_step_fn(workload, sample_b):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import jcm_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    gen = torch.Generator().manual_seed(0)
    _, K, H, W, train, _ = WORKLOADS[workload]
    names = orc.JOINT_NAMES[:K] + ['torso'] if K <= 9 else ['j%02d' % i for i in range(K)] + ['torso']
    p = orc.init_part_detector(K, gen, dtype=torch.float32, requires_grad=train)
    if K <= 9:
        with np.load(os.path.join(ROOT, 'joint-cnn-mrf_b200', 'jcm', 'data', 'pairwise_distribution.npz')) as z:
            distr = {k: z[k] for k in z.files}
    else:
        distr = synthetic_pairwise(names, K, H // 8, W // 8, np.random.default_rng(0))
    sm = orc.init_spatial_model(distr, K, H // 8, W // 8, joint_names=names, dtype=torch.float32, requires_grad=train)
    x = torch.rand(sample_b, H, W, 3, generator=gen)
    y = torch.from_numpy(synthetic_labels(sample_b, H // 8, W // 8, K + 1, np.random.default_rng(0)))

    def step():
        if train:
            out = orc.tower_forward(x, y, p, sm, K, True, joint_names=names)
            params = [v for k, v in list(p.items()) + list(sm.items()) if v.requires_grad]
            torch.autograd.grad(out['loss'], params, allow_unused=True)
        else:
            with torch.no_grad():
                orc.tower_forward(x, y, p, sm, K, False, joint_names=names)
    return step


def time_cpu(step, n, warm):
    for _ in range(warm):
        step()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample_b = 2
    step = cpu_reference_step_fn(args.workload, sample_b)
    ts = time_cpu(step, args.steps, args.warmup)
    total = sum(ts)
    val = sample_b * len(ts) / total
    cores = os.cpu_count() or 1
    line = {
        'impl': 'reference', 'metric': metric_name(args.workload), 'value': val, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(ts), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': dict(workload_config(args.workload, args.gpus), per_gpu_batch=sample_b, global_batch=sample_b,
                       sample='CPU arm: each step is a bounded sample of %d images of the workload (its batch of %d would take minutes per step '
                              'on the host cores); images/s is per image, so the two arms compare per image' % (sample_b, WORKLOADS[args.workload][0])),
        'cpu_baseline': {'value': val, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d images per step (%s), torch-CPU fp32 restatement of the reference TF graph '
                                   '(TensorFlow 1.x not installable)' % (sample_b, args.workload)},
        'e2e': {'value': val, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit(line)


def metric_name(workload):
    """BASELINE.json's metric (images/sec fwd+bwd at 720x480, K=7); the other configurations are named as such."""
    return {'train64': 'images/sec fwd+bwd (720x480, K=7 joints)', 'fwd16': 'images/sec fwd (720x480, K=7 joints)',
            'train_k14': 'images/sec fwd+bwd (1024x768, K=14 joints)'}[workload]


def workload_config(workload, n_gpus):
    B, K, H, W, train, precision = WORKLOADS[workload]
    what = {'fwd16': 'BASELINE configs[1]: part-detector + spatial-model forward (+ softmax-CE heads), batch 16 per GPU, K=7, 720x480x3 '
                     'synthetic, fp32-equivalent (bf16x3 split) tensor-core convs, inference-mode BN',
            'train64': 'BASELINE configs[2]: joint training fwd+bwd + grad all-reduce + clip + Adam, batch 64 per GPU, K=7, 720x480x3 '
                       'synthetic, bf16 tensor-core operands / fp32 accumulation',
            'train_k14': 'BASELINE configs[4]: joint training fwd+bwd + grad all-reduce + clip + Adam, batch 32 per GPU, K=14 (196 pairwise '
                         'terms), 1024x768x3 synthetic, 96x128 heat maps, bf16 tensor-core operands / fp32 accumulation'}[workload]
    return {'workload': what, 'per_gpu_batch': B, 'global_batch': B * n_gpus, 'K': K, 'image': [H, W], 'heat_map': [H // 8, W // 8],
            'l2': 'inputs + activations per step (> 1 GB) exceed the 126 MB L2', 'parallelism': 'dp%d' % n_gpus}


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def measure(workload, steps, warmup, args, rank, world, local, dev, with_cpu_baseline, with_e2e=True):
    """One leg: W warm-up steps, K timed steps (device-resident inputs), then K end-to-end steps.  Returns the result dict (rank 0)
    or None (other ranks).  Everything allocated here is released before returning, so legs can follow each other."""
    import torch.distributed as dist
    import jcm
    from jcm import ops

    B, K, IH, IW, train, precision = WORKLOADS[workload]
    gen = torch.Generator().manual_seed(1234 + rank)
    wgen = torch.Generator().manual_seed(0)          # identical parameters on every replica
    p = jcm.init_part_detector(K, wgen, device=dev)
    if K <= 9:
        names = jcm.JOINT_NAMES[:K] + ['torso']
        distr = jcm.get_pairwise_distr()
    else:
        names = ['j%02d' % i for i in range(K)] + ['torso']
        distr = synthetic_pairwise(names, K, IH // 8, IW // 8, np.random.default_rng(0))
    sm = jcm.PairwiseParams.from_distribution(distr, names, K, IH // 8, IW // 8, device=dev)
    ctx = jcm.Context(n_joints=K, joint_names=names, flag_train=train, precision=precision,
                      bf16_activations=None if args.bf16_activations is None else bool(args.bf16_activations),
                      sm_tensor_core=False if args.sm_ffma else (True if args.sm_tc else None))

    x_host = torch.rand(B, IH, IW, 3, generator=gen).pin_memory()
    y_host = torch.from_numpy(synthetic_labels(B, IH // 8, IW // 8, K + 1, np.random.default_rng(rank))).pin_memory()
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

    trainer = None
    if train:
        from jcm import train as jtrain
        trainer = jtrain.Trainer(p, sm, ctx, world_size=world)

        def step(x, y):
            return trainer.step(x, y)['loss']
    else:
        def step(x, y):
            out = jcm.tower_forward(x, y, p, sm, ctx)
            return out['loss_pd'] + out['loss_sm']

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing
    for _ in range(warmup):
        step(x_dev, y_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = jcm.lib().jcm_launch_count()
    ops.PROFILE.clear()
    ops.PROFILE.enabled = rank == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        loss = step(x_dev, y_dev)
    e1.record()
    host_ms = 1e3 * (time.perf_counter() - t_host0) / steps      # CPU time per step of the issuing loop (no synchronisation inside it)
    barrier()
    ops.PROFILE.enabled = False
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = jcm.lib().jcm_launch_count() - launches0            # libjcm kernel launches inside the timed region (all K steps, this rank)
    clocks = sampler.stop() if rank == 0 else None
    conv_prof = ops.PROFILE.summary(steps, 'conv_igemm_kernel') if rank == 0 else None
    wgrad_prof = ops.PROFILE.summary(steps, 'conv_wgrad_kernel') if rank == 0 else None
    conv_big = ops.PROFILE.largest('conv_igemm_kernel') if rank == 0 else None
    conv_shapes = ops.PROFILE.by_shape('conv_igemm_kernel', steps) if rank == 0 else None
    wgrad_shapes = ops.PROFILE.by_shape('conv_wgrad_kernel', steps) if rank == 0 else None
    sm_prof = [ops.PROFILE.summary(steps, k) for k in ('spatial_model_fwd', 'spatial_model_bwd')] if rank == 0 else None
    ops.PROFILE.clear()
    # CPU time to issue ONE step into an empty stream (outside the timed region): in the loop above the launch queue fills up and the
    # CPU waits for the GPU, so host_ms_per_step there is a mix of both; this one is Python + ctypes + tensor-map encoding alone
    t1 = time.perf_counter()
    step(x_dev, y_dev)
    host_issue_ms = 1e3 * (time.perf_counter() - t1)
    barrier()

    # ---- replica consistency (N > 1, training): every replica must hold the same parameters, optimizer slots and BatchNorm
    # moving statistics after the timed steps (each saw DIFFERENT data): bit patterns summed as integers, compared across ranks
    replicas_identical = None
    if world > 1 and trainer is not None:
        sums = torch.stack([t.view(torch.int32).to(torch.int64).sum() for t in (trainer.flat, trainer.m, trainer.v, trainer.moving)])
        gathered = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(gathered, sums)
        replicas_identical = bool(all(torch.equal(gathered[0], g) for g in gathered))

    # ---- end to end through the public API: every step copies its inputs from pinned host memory (jcm.DeviceFeed: the copy of
    # step i+1 runs on a side stream while step i computes) and reads the loss back to the host
    ms_e2e = None
    h2d_gbs = None
    if with_e2e:
        res_host = torch.empty(1, dtype=torch.float32).pin_memory()
        feed = jcm.DeviceFeed(dev)
        # the host->device link of this box, measured alone (one batch, nothing else running): when a step's inputs take longer to
        # copy than the step takes to compute, e2e is bounded by this and not by the kernels
        feed.submit(x_host, y_host)
        feed.take()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        feed.submit(x_host, y_host)
        feed.take()
        c1.record()
        barrier()
        h2d_gbs = (x_host.numel() * 4 + y_host.numel() * 4) / (c0.elapsed_time(c1) * 1e-3) / 1e9
        for _ in range(min(warmup, 2)):
            feed.submit(x_host, y_host)
            step(*feed.take())
        barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        feed.submit(x_host, y_host)                       # step 0's inputs: inside the timed region, nothing to overlap with yet
        for i in range(steps):
            xd, yd = feed.take()
            if i + 1 < steps:
                feed.submit(x_host, y_host)               # next step's inputs, overlapped with this step's kernels
            loss = step(xd, yd)
            res_host.copy_(loss.reshape(1), non_blocking=True)
        e3.record()
        barrier()
        ms_e2e = max_over_ranks(e2.elapsed_time(e3))
    loss_value = float(loss.item()) if torch.is_tensor(loss) else float(loss)
    h2d = int(x_host.numel() * 4 + y_host.numel() * 4)
    sm_tc = bool(ctx.sm_tc)
    del trainer, p, sm, x_dev, y_dev, x_host, y_host, step, loss
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if rank != 0:
        return None

    peaks = load_peaks()
    imgs = B * world * steps
    value = imgs / (ms_total / 1e3)

    # CPU baseline on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if with_cpu_baseline:
        sample_b = 2
        st = cpu_reference_step_fn(workload, sample_b)
        ts = time_cpu(st, 3, 1)
        cpu = {'value': sample_b / min(ts), 'unit': 'images/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
               'sample': 'best of 3 steps of %d images (a bounded sample of the %s workload, not its batch of %d) after 1 warm-up; torch-CPU '
                         'fp32 restatement of the reference TF graph (TensorFlow 1.x not installable here)' % (sample_b, workload, B)}

    peak = peaks['bf16_sustained']
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture of this same command and binary (per launch, like
    # `achieved`): profiles/r02/ncu_full_<workload>_traffic.json is written by profiles/summarize_ncu.py from the capture
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02', 'ncu_full_%s_traffic.json' % workload)) as f:
            tj = json.load(f)
        tk = tj['kernels'].get('conv_igemm_pair_kernel') or tj['kernels']['conv_igemm_kernel']
        traffic = tk['dram_bytes_per_launch']
        traffic_src = 'profiles/r02/ncu_full_%s_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum, mean over %d launches of ' \
                      'conv_igemm_pair_kernel (%s)' % (workload, tk['launches'], tj.get('note', ''))
    except Exception:
        pass
    mult = 1 if precision == 'bf16' else 3
    roof = {'bound': 'tensor', 'kernel': 'conv_igemm_pair_kernel / conv_igemm_kernel (tcgen05 implicit GEMM, cta_group::2 for N = 256 tiles: forward + '
                                          'data-gradient convolutions, %d launches/step)' % conv_prof['launches_per_step'],
            'peak_nominal': 2250.0,
            'achieved': conv_prof['tflops'], 'peak': peak, 'unit': 'TFLOP/s', 'frac': conv_prof['tflops'] / peak,
            'traffic': traffic, 'traffic_source': traffic_src,
            'flops_per_launch': conv_prof['flops_per_step'] / max(conv_prof['launches_per_step'], 1),
            'ms_per_launch': conv_prof['ms_per_step'] / max(conv_prof['launches_per_step'], 1),
            'note': 'achieved = algorithmic conv FLOPs (2*MACs, SURVEY App. A) of the kernel\'s launches / their summed CUDA-event time, '
                    'measured live in this run; peak = bf16_tflops_sustained of %s (kernel timed inside a long power-capped step) - that '
                    'figure is what torch.matmul (cuBLAS) sustains on this pool, so frac > 1 means the kernel out-runs cuBLAS under the same '
                    'power cap, not that it exceeds the hardware: peak_nominal is the dense bf16 datasheet number; '
                    'tensor-core MMAs executed per algorithmic MAC: %d (fp32 config = bf16x3 split products, ceiling of frac 1/3)'
                    % (peaks['source'], mult),
            'share_of_step': conv_prof['ms_per_step'] / (ms_total / steps)}
    if conv_big:
        roof['largest_launch'] = {'layer': 'conv5 9x9 512->512 (forward / data gradient)', 'achieved': conv_big['tflops'],
                                  'frac': conv_big['tflops'] / peak, 'ms': conv_big['ms']}
    if conv_shapes:
        # per launch shape ("HxW Cin->Cout kernel", forward and data-gradient launches of one shape share a row): live CUDA-event time
        roof['shapes'] = [dict(d, frac=d['tflops'] / peak) for d in conv_shapes]
    if wgrad_prof and wgrad_prof['launches']:
        roof['conv_wgrad_kernel'] = {'achieved': wgrad_prof['tflops'], 'frac': wgrad_prof['tflops'] / peak,
                                     'launches_per_step': wgrad_prof['launches_per_step'],
                                     'share_of_step': wgrad_prof['ms_per_step'] / (ms_total / steps),
                                     'shapes': [dict(d, frac=d['tflops'] / peak) for d in wgrad_shapes]}
    if sm_prof and sm_prof[0]['launches']:
        # the spatial model (north_star's second kernel family): CUDA-event time of the whole fwd / bwd call (all of its kernels)
        roof['spatial_model'] = {
            'form': ('tensor cores: grouped Toeplitz GEMMs through conv_igemm_kernel, bf16 operands with the prior centred per pair (%s)'
                     % ('bf16 configuration' if precision == 'bf16' else 'opt-in for fp32 inference, --sm-tc')) if sm_tc
                    else 'fp32 FFMA2 kernels (sm_conv_kernel / sm_bwd_dp_kernel)',
            'fwd_ms': sm_prof[0]['ms_per_step'], 'bwd_ms': sm_prof[1]['ms_per_step'],
            'algorithmic_tflops_fwd': sm_prof[0]['tflops'], 'algorithmic_tflops_bwd': sm_prof[1]['tflops'] if sm_prof[1]['launches'] else None,
            'fp32_fma_peak_tflops': 73.0,
            'fp32_ffma2_attainable': 'register-operand FFMA2 streams reach 0.92-0.95 of that peak and every load or integer instruction takes an '
                                     'FMA issue cycle: 0.81 is the bound of this kernel\'s blocking (profiles/r02/ffma_probes/README.md)',
            'share_of_step': (sm_prof[0]['ms_per_step'] + sm_prof[1]['ms_per_step']) / (ms_total / steps)}
    line = {
        'metric': metric_name(workload), 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': steps, 'warmup': warmup,
        'ms_per_step': ms_total / steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'bf16' if precision == 'bf16' else 'f32',
        'dtype_detail': 'bf16 tensor-core operands and stored activations, fp32 accumulation, fp32 master weights / gradients / optimizer'
                        if precision == 'bf16' else 'fp32-equivalent: every product is 3 bf16 tensor-core MMAs (hi*hi + lo*hi + hi*lo), fp32 accumulation',
        'data': 'synthetic', 'config': workload_config(workload, world),
        'gpu_launches': int(launches),
        'gpu_launches_per_step': int(launches) // max(steps, 1),
        'host_ms_per_step': host_ms,
        'host_issue_ms': host_issue_ms,
        'host_note': 'host_issue_ms = CPU time to issue ONE step into an empty stream (Python + ctypes + tensor-map encoding; measured after '
                     'the timed region); host_ms_per_step = CPU time per step of the timed loop, where the CPU also waits for launch-queue '
                     'slots; the step is GPU-bound while host_issue_ms is below ms_per_step',
        'clocks': clocks,
        'roofline': roof,
        'loss': loss_value,
    }
    if ms_e2e is not None:
        line['e2e'] = {'value': imgs / (ms_e2e / 1e3), 'unit': 'images/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                       'h2d_link_gbs': h2d_gbs, 'h2d_ms_per_step_alone': h2d / (h2d_gbs * 1e9) * 1e3 if h2d_gbs else None,
                       'note': 'inputs copied from pinned host memory every step (overlapped with the previous step), loss read back; '
                               'h2d_link_gbs = this box\'s pinned host->device rate measured alone'}
    if replicas_identical is not None:
        line['replicas_identical'] = replicas_identical
        line['replicas_note'] = 'integer checksums of the parameter, Adam m / v and BatchNorm moving-statistic buffers compared across ranks after the timed steps'
    if cpu is not None:
        line['cpu_baseline'] = cpu
    return line


def run_ours(args):
    import torch.distributed as dist
    import jcm

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device - the jcm kernels have no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    jcm.lib()  # fail loudly if libjcm.so is missing

    line = measure(args.workload, args.steps, args.warmup, args, rank, world, local, dev,
                   with_cpu_baseline=(world == 1 and not args.no_cpu_baseline))
    # Extra legs: the other BASELINE configurations, so that the driver's record holds them too.  Short (5 steps after 3 warm-up), after
    # the headline measurement, never allowed to break it.  configs[1] (fwd16) is a single-GPU configuration; configs[4] (K=14) is
    # quoted on 8 GPUs and runs at every N.
    if args.extra_legs and args.workload == 'train64':
        extra = {}
        for wl in (['fwd16'] if world == 1 else []) + ['train_k14']:
            try:
                r = measure(wl, 5, 3, args, rank, world, local, dev, with_cpu_baseline=False, with_e2e=(world == 1))
                if r is not None:
                    keep = ('metric', 'value', 'unit', 'ms_per_step', 'dtype', 'config', 'e2e', 'gpu_launches_per_step', 'host_ms_per_step', 'host_issue_ms',
                            'loss', 'replicas_identical', 'steps', 'warmup', 'n_gpus')
                    extra[wl] = {k: r[k] for k in keep if k in r}
                    extra[wl]['roofline'] = {k: r['roofline'][k] for k in ('achieved', 'peak', 'frac', 'unit', 'share_of_step', 'spatial_model')
                                             if k in r['roofline']}
                    if wl == 'fwd16' and not args.sm_ffma and not args.sm_tc:
                        # the same leg with the spatial model in its centred tensor-core form (opt-in for fp32 inference)
                        import copy
                        a2 = copy.copy(args)
                        a2.sm_tc = True
                        r2 = measure(wl, 5, 3, a2, rank, world, local, dev, with_cpu_baseline=False, with_e2e=False)
                        if r2 is not None:
                            extra[wl]['sm_tensor_core_opt_in'] = {
                                'value': r2['value'], 'unit': r2['unit'], 'ms_per_step': r2['ms_per_step'],
                                'spatial_model': r2['roofline'].get('spatial_model'),
                                'note': 'Context(sm_tensor_core=True): bf16 Toeplitz GEMMs with the prior centred per pair; logits within 1e-4 of the '
                                        'oracle like the FFMA form (tests/test_gpu_parity.py::test_spatial_model_fp32_inference_on_tensor_cores_opt_in)'}
            except Exception as e:   # noqa: BLE001
                if rank == 0:
                    extra[wl] = dict(extra.get(wl, {}), error=repr(e)[:300])
        if rank == 0:
            line['extra_legs'] = extra
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


_RESULT_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version banner under NCCL_DEBUG=VERSION, measured on
    the GPU boxes: it ignores NCCL_DEBUG_FILE), so file descriptor 1 is pointed at stderr for everything else and the result line
    goes to a private duplicate of the original stdout."""
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        fd = os.dup(1)
        os.dup2(2, 1)
        _RESULT_OUT = os.fdopen(fd, 'w')
    return _RESULT_OUT


def emit(line):
    out = claim_stdout()
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default=None, choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra-legs', dest='extra_legs', action='store_false',
                    help='train64 only: skip the short fwd16 (configs[1], N=1) and train_k14 (configs[4]) legs reported under "extra_legs"')
    ap.add_argument('--sm-tc', action='store_true',
                    help='fp32 inference workloads: run the spatial model in its centred tensor-core form (opt-in; default FFMA)')
    ap.add_argument('--sm-ffma', action='store_true',
                    help='bf16 workloads: run the spatial model on the fp32 FFMA kernels (north_star form) instead of the tensor-core form')
    ap.add_argument('--bf16-activations', type=int, default=None,
                    help='store the post-ReLU activations in bf16 too (bf16 workloads); default: the library default')
    args = ap.parse_args()
    if args.workload is None:
        args.workload = default_workload()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


def default_workload():
    """train64 (the configuration the metric is quoted on) once the training step is built; fwd16 otherwise."""
    return 'train64' if os.path.exists(os.path.join(ROOT, 'joint-cnn-mrf_b200', 'jcm', 'train.py')) else 'fwd16'


if __name__ == '__main__':
    main()
