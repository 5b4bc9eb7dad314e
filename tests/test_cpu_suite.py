"""CPU-only checks (-m "not gpu"): golden vectors vs the oracle, the pairwise-prior table, the C ABI surface and the
host-side argument validation (no compute calls: there is no GPU here)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

import jcm_oracle as orc
import pairwise_prior as pp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')
NPZ = os.path.join(ROOT, 'joint-cnn-mrf_b200', 'jcm', 'data', 'pairwise_distribution.npz')
REF = '/root/reference'


# ------------------------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize('name', ['sm_small', 'sm_tiny_ragged'])
def test_oracle_reproduces_sm_golden(name):
    z = np.load(os.path.join(GOLD, name + '.npz'))
    names = [str(s) for s in z['names']]
    K = len(names) - 1
    sm = {k[3:]: torch.from_numpy(z[k]).double() for k in z.files if k.startswith('sm/')}
    cat = torch.from_numpy(z['heat_map']).double()
    for train in (0, 1):
        out = orc.spatial_model(cat, {k: v.clone() for k, v in sm.items()}, K, bool(train), joint_names=names)
        np.testing.assert_allclose(out.numpy(), z['out_train%d' % train], rtol=1e-12, atol=1e-12)
        assert np.array_equal(orc.get_joints_coords(orc.spatial_softmax(out)).numpy(), z['argmax_train%d' % train])


def test_oracle_reproduces_model_golden():
    sys.path.insert(0, GOLD)
    import make_golden as mg
    z = np.load(os.path.join(GOLD, 'model_debug_small.npz'))
    p, _ = mg.model_params(int(z['K']), int(z['seed']))
    assert abs(mg.params_checksum(p) - float(z['params_checksum'])) < 1e-6 * float(z['params_checksum'])
    p64 = {k: v.double() for k, v in p.items()}
    x = torch.from_numpy(z['x']).double()
    for train in (0, 1):
        logits = orc.model(x, {k: v.clone() for k, v in p64.items()}, int(z['K']), bool(train))
        np.testing.assert_allclose(logits.numpy(), z['logits_train%d' % train], rtol=1e-9, atol=1e-9)


# ------------------------------------------------------------------------------------------------ pairwise prior
def test_shipped_prior_table_structure():
    """90 keys '<joint>_<cond>' in joint_ids x joint_ids order, float64 (120,180), sum 1 (prepare_pairwise_distribution.py:46-56)."""
    with np.load(NPZ) as z:
        keys = list(z.files)
        want = [a + '_' + b for a in pp.JOINT_IDS for b in pp.JOINT_IDS if a != b]
        assert keys == want
        for k in keys:
            a = z[k]
            assert a.shape == (120, 180) and a.dtype == np.float64
            assert abs(a.sum() - 1.0) < 1e-9 and a.min() >= 0
        a = z['nose_torso']   # SURVEY Appendix C sanity values for the 'shipped' recipe
        assert np.unravel_index(a.argmax(), a.shape) == (49, 89)
        assert abs(a.max() - 0.013686) < 1e-5 and int((a != 0).sum()) == 800


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'data_FLIC.mat')), reason='reference mount not present')
def test_prior_regenerates_from_flic_and_matches_corrupt_pickle_structure():
    """The regenerated table equals the committed one, and its zero-run structure equals what survives in the
    (corrupt) pickle shipped with the reference - the only pin on the reference's actual initialisation."""
    keys = ['nose_torso', 'lwri_lelb', 'lsho_rsho', 'rhip_torso']
    regen = pp.regenerate(os.path.join(REF, 'data_FLIC.mat'), 'shipped', keys=set(keys))
    runs = pp.zero_runs_from_corrupt_pickle(os.path.join(REF, 'pairwise_distribution.pickle'))
    with np.load(NPZ) as z:
        for k in keys:
            np.testing.assert_allclose(regen[k], z[k], rtol=0, atol=1e-15)
            assert pp.zero_runs_of_array(regen[k]) == runs[k], k


def test_labels_follow_data_py():
    """data.py:106-114,180-189: a 3x3 binomial blob at int(row), int(col), clipped at the border."""
    pos = np.array([[[10.6, 20.2]] * 10, [[0.0, 0.0]] * 10, [[59.9, 89.9]] * 10])
    y = pp.target_heat_maps(pos)
    assert y.shape == (3, 60, 90, 10)
    assert abs(y[0, :, :, 0].sum() - 1) < 1e-6 and y[0, 10, 20, 0] == 0.25
    assert abs(y[1, :, :, 0].sum() - 9 / 16) < 1e-6      # corner blob is clipped
    assert y[2, 59, 89, 3] == 0.25


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, 'include', 'jcm.h')).read()
    declared = set(re.findall(r'\b(jcm_\w+)\s*\(', hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(built_lib)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    from jcm._lib import SIGNATURES
    assert declared == set(SIGNATURES), (declared ^ set(SIGNATURES))
    lib.jcm_version.restype = ctypes.c_int
    assert lib.jcm_version() >= 100
    lib.jcm_spatial_model_workspace.restype = ctypes.c_long
    assert lib.jcm_spatial_model_workspace(16, 60, 90, 7, 49) > 0


def test_sass_contains_blackwell_instructions(built_lib):
    """tcgen05.mma -> UTC*MMA, TMA -> UTMALDG, tcgen05.ld -> LDTM, packed fp32 FMA -> FFMA2 (B200_PROFILING.md)."""
    import subprocess
    sass = subprocess.run(['cuobjdump', '-sass', built_lib], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'LDTM', 'FFMA2', 'UTMASTG'):
        assert mnemonic in sass, mnemonic
    # the CTA-pair kernel of the N = 256 layers (conv_igemm_pair_kernel, cta_group::2) compiles to the 2-SM forms
    for mnemonic in ('UTCHMMA.2CTA', 'UTMALDG.4D.2CTA', 'UTCBAR.2CTA.MULTICAST'):
        assert mnemonic in sass, mnemonic


def test_register_budgets_of_the_bandwidth_bound_kernels(built_lib):
    """Occupancy guard: the HBM-bound glue kernels depend on many resident warps, and a small source change can make ptxas unroll one of
    them into a 255-register kernel (it happened to the diagonal reduce of the tensor-core spatial model: 2.3 ms instead of 0.8).
    No spills, and at most 64 registers per thread for the kernels launched with 256..1024 threads per block."""
    import re
    import subprocess
    out = subprocess.run(['cuobjdump', '--dump-resource-usage', built_lib], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r'Function ([^:\n]+):\s*\n\s*REG:(\d+) STACK:(\d+)', out):
        usage[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    assert len(usage) > 50, 'cuobjdump --dump-resource-usage gave no per-function lines'
    budget = {'smt_dp_reduce_kernel': 64, 'smt_pack_prior_kernel': 64, 'smt_dc_kernel': 64, 'smt_prep_kernel': 64, 'sm_finish_kernel': 64,
              'sm_bwd_dt_kernel': 64, 'sm_bwd_dh_kernel': 64, 'smt_dtsum_kernel': 64, 'smt_softplus_center_kernel': 64,
              'bn_stats_kernel': 64, 'clip_adam_kernel': 64, 'grad_prepare_kernel': 64, 'spatial_softmax_kernel': 64}
    seen = set()
    for name, (reg, stack) in usage.items():
        for k, lim in budget.items():
            if k in name:
                seen.add(k)
                assert reg <= lim and stack == 0, (name, reg, stack)
    assert seen == set(budget), set(budget) - seen


def test_argument_errors_do_not_need_a_gpu(built_lib):
    """Bad arguments are rejected on the host with a negative code and a message (ValueError in the Python layer)."""
    import jcm
    l = jcm.lib()
    rc = l.jcm_conv2d_fwd(None, None, None, None, None, None, 0, 1, 8, 8, 64, 64, 64, 3, 0, 1, None)
    assert rc == -1 and b'null pointer' in l.jcm_last_error()
    rc = l.jcm_spatial_softmax(None, 1, 10, 7, None, None)
    assert rc == -1
    with pytest.raises(ValueError):
        jcm.ops.spatial_softmax(torch.zeros(1, 4, 4, 2))          # CPU tensor: there is no CPU path
    with pytest.raises(ValueError):
        jcm.Context(n_joints=7, joint_names=['a', 'b'])
    with pytest.raises(ValueError):
        jcm.Context(precision='fp16')


def test_context_selects_the_spatial_model_form_by_precision(built_lib):
    """fp32 configuration: always the fp32 FFMA kernels (the 1e-3 / bit-exact arg-max path).  bf16 configuration: the tensor-core
    form unless switched off.  The workspace queries of both forms answer without a GPU."""
    import jcm
    assert not jcm.Context(n_joints=7).sm_tc
    assert jcm.Context(n_joints=7, precision='fp32', sm_tensor_core=True).sm_tc                      # opt-in, inference context
    assert not jcm.Context(n_joints=7, precision='fp32', sm_tensor_core=True, flag_train=True).sm_tc   # fp32 training: always FFMA
    assert jcm.Context(n_joints=7, precision='bf16').sm_tc
    assert not jcm.Context(n_joints=7, precision='bf16', sm_tensor_core=False).sm_tc
    l = jcm.lib()
    assert l.jcm_spatial_model_tc_workspace(64, 60, 90, 7, 49) > 0 and l.jcm_spatial_model_tc_bwd_workspace(64, 60, 90, 7, 49) > 0
    assert l.jcm_spatial_model_tc_workspace(2, 60, 300, 7, 49) < 0          # maps wider than 255 columns: not supported
    assert b'width' in l.jcm_last_error()
    rc = l.jcm_spatial_model_tc_fwd(None, None, None, None, None, None, None, None, None, 0, 2, 60, 90, 7, 49, None)
    assert rc == -1 and b'null pointer' in l.jcm_last_error()


def test_product_path_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, 'joint-cnn-mrf_b200', 'jcm')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            src = open(os.path.join(pkg, f)).read()
            assert 'jcm_oracle' not in src and 'import oracle' not in src and 'pairwise_prior' not in src.replace('oracle/pairwise_prior.py', ''), f


def test_checkpoint_roundtrip_under_tf_variable_names(tmp_path, built_lib):
    """save / restore of variables, BN moving statistics, optimizer slots and n_iters under the reference's
    TensorFlow names (main.py:604-617,663-666).  Host-side only: the buffers live on the CPU here."""
    import jcm
    K = 3
    names = jcm.JOINT_NAMES[:K] + ['torso']

    def make(seed):
        gen = torch.Generator().manual_seed(seed)
        p = jcm.init_part_detector(K, gen, debug=True, device='cpu')
        distr = {a + '_' + b: torch.rand(16, 24, generator=gen).numpy() for a in names[:K] for b in names if a != b}
        sm = jcm.PairwiseParams.from_distribution(distr, names, K, 8, 12, device='cpu')
        ctx = jcm.Context(n_joints=K, joint_names=names, flag_train=True, precision='bf16', debug=True)
        return jcm.train.Trainer(p, sm, ctx, lr=1e-3)

    a = make(1)
    gen = torch.Generator().manual_seed(9)
    a.m.copy_(torch.rand(a.n, generator=gen))
    a.v.copy_(torch.rand(a.n, generator=gen))
    a.p['conv3_halfres/BatchNorm/moving_mean'].copy_(torch.rand(64, generator=gen))
    a.t = 1234
    path = str(tmp_path / 'ckpt.npz')
    jcm.save_checkpoint(path, a)
    z = np.load(path)
    for name in ('conv1_fullres/weights', 'conv5/BatchNorm/moving_variance', 'bn_sm/BatchNorm/gamma', 'energy_lsho_torso', 'bias_lelb_lsho',
                 'conv6/biases/Adam', 'energy_lwri_lsho/Adam_1', 'n_iters'):
        assert name in z.files, name
    assert z['energy_lsho_torso'].shape == (1, 16, 24, 1) and z['conv1_fullres/weights'].shape == (5, 5, 3, 16)
    b = make(2)
    assert not torch.equal(a.flat, b.flat)
    jcm.load_checkpoint(path, b)
    from jcm.checkpoint import state_dict
    sa, sb = state_dict(a), state_dict(b)
    assert sorted(sa) == sorted(sb) and all(np.array_equal(sa[k], sb[k]) for k in sa) and b.t == 1234
    assert torch.equal(a.p['conv3_halfres/BatchNorm/moving_mean'], b.p['conv3_halfres/BatchNorm/moving_mean'])
    with pytest.raises(ValueError):
        jcm.load_checkpoint(path, jcm.train.Trainer(*_mk_parts(jcm, names, K), optimizer='momentum'))


def _mk_parts(jcm, names, K):
    gen = torch.Generator().manual_seed(3)
    p = jcm.init_part_detector(K, gen, debug=True, device='cpu')
    distr = {a + '_' + b: torch.rand(16, 24, generator=gen).numpy() for a in names[:K] for b in names if a != b}
    sm = jcm.PairwiseParams.from_distribution(distr, names, K, 8, 12, device='cpu')
    return p, sm, jcm.Context(n_joints=K, joint_names=names, flag_train=True, precision='bf16', debug=True)


def test_bench_stdout_carries_only_the_result_line():
    """bench.py contract: ONE JSON line on stdout.  NCCL writes its version banner to file descriptor 1 on the GPU boxes
    (NCCL_DEBUG=VERSION in their environment), so bench.py points fd 1 at stderr and keeps a private handle for the result."""
    import subprocess
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.claim_stdout(); os.write(1, b'NCCL version x\\n'); "
            "print('python-level noise'); bench.emit({'a': 1})" % ROOT)
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"a": 1}\n'
    assert 'NCCL version x' in r.stderr and 'python-level noise' in r.stderr


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's CPU stand-in: the oracle port on the host cores) prints ONE JSON line with the
    contract's keys; under torchrun every rank but 0 exits silently."""
    import json
    import subprocess
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'fwd16', '--steps', '1', '--warmup', '0']
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling', 'vs_baseline',
              'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in d, k
    assert d['impl'] == 'reference' and d['value'] > 0 and d['vs_baseline'] is None and 'workload' in d['config']
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK='1', WORLD_SIZE='2'))
    assert r1.returncode == 0 and r1.stdout.strip() == ''


def test_product_library_has_no_measurement_switches(built_lib):
    """The shipped libjcm.so reads no JCM_* environment variable and carries no skip-the-work / experimental kernel: those exist
    only in the -DJCM_EXPERIMENTS build (build.py --experiments -> libjcm_exp.so, loaded by tests/gpu_diag.py alone)."""
    data = open(built_lib, 'rb').read()
    for name in (b'JCM_CONV_DBG', b'JCM_MMA_SPLITN', b'JCM_CONV_NACC', b'JCM_CONV_CTA2', b'JCM_CONV_HALO', b'JCM_CONV_BGROUP',
                 b'JCM_CONV_PATCH_TW'):
        assert name not in data, name
    src = open(os.path.join(ROOT, 'joint-cnn-mrf_b200', 'jcm', '_lib.py')).read()
    assert 'environ' not in src                     # the package always loads jcm/libjcm.so


def test_packed_weight_cache_is_shared_and_epoch_validated(built_lib):
    """graph.Context.packed: one cache for every Context, invalidated by params_updated() (raw-pointer updates) and by torch's
    version counter - checked here on the bookkeeping alone (packing itself needs the GPU: tests/test_gpu_round2.py)."""
    import torch
    from jcm import graph
    calls = []
    orig = graph.ops.pack_weights
    graph.ops.pack_weights = lambda w, split, transpose=False: calls.append((w.data_ptr(), transpose)) or object()
    try:
        w = torch.zeros(3, 3, 16, 16)
        a, b = graph.Context(precision='bf16'), graph.Context(precision='bf16', flag_train=True)
        p1 = a.packed('l', w)
        assert b.packed('l', w) is p1 and len(calls) == 1          # the second context reuses the first one's planes
        graph.params_updated()
        p2 = b.packed('l', w)
        assert p2 is not p1 and a.packed('l', w) is p2 and len(calls) == 2
        w.add_(1.0)                                                 # torch-visible in-place update
        assert a.packed('l', w) is not p2 and len(calls) == 3
        assert a.packed('l', w, 'dgrad') is not a.packed('l', w) and len(calls) == 4
    finally:
        graph.ops.pack_weights = orig
