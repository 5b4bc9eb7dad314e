#!/bin/bash
# Profiling pass (run on the GPU box through gpurun, from the repo root):  bash profiles/run_profile.sh [workload]
# Produces in gpurun_out/ (small files only - a full report of a whole step is ~85 MB and stays on the box):
#   launches_<wl>.csv        every launch of the run with its device time (ncu, cold-cache, serialised: compare SHARES)
#   prof_<wl>_raw.csv        --set full counters of every conv / spatial-model launch of the first step (raw page as CSV)
# then here:  python profiles/launch_summary.py gpurun_out/launches_<wl>.csv 3   and   python profiles/ncu_summary.py ...
# Numbers printed by bench.py under ncu are never bench values.
WL=${1:-train64}
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
$NCU --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${WL}.log 2>&1
$NCU --set full --clock-control none --import-source on -k regex:"conv_igemm_kernel|conv_wgrad_kernel|sm_conv_kernel|sm_bwd_dp_kernel" \
    -c ${COUNT:-42} -f -o /tmp/prof_${WL} python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_${WL}.log 2>&1
$NCU -i /tmp/prof_${WL}.ncu-rep --page raw --csv > gpurun_out/prof_${WL}_raw.csv 2>/dev/null
$NCU -i /tmp/prof_${WL}.ncu-rep --page source --csv -k regex:sm_conv_kernel -c 1 > gpurun_out/prof_${WL}_smconv_source.csv 2>/dev/null
ls -la gpurun_out/ /tmp/prof_${WL}.ncu-rep
# the three spatial-model GEMM launches (tensor-core form) on their own: raw page -> gpurun_out/ncu_full_smtc_raw.csv
$NCU --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -c 3 -f -o /tmp/prof_smtc python tests/gpu_diag.py smtc64 \
    > gpurun_out/ncu_smtc.log 2>&1
$NCU -i /tmp/prof_smtc.ncu-rep --page raw --csv > gpurun_out/ncu_full_smtc_raw.csv 2>/dev/null
# the same at the K=14 / 96x128 / batch 32 shape (configs[4]): launch 14 = the forward GEMM after warm-up, then dL and dP of the first backward
$NCU --set full --clock-control none -k regex:conv_igemm_kernel -s 13 -c 3 -f -o /tmp/prof_smtc_k14 python tests/gpu_diag.py smtck14 \
    > gpurun_out/ncu_smtc_k14.log 2>&1
$NCU -i /tmp/prof_smtc_k14.ncu-rep --page raw --csv > gpurun_out/ncu_full_smtc_k14_raw.csv 2>/dev/null
