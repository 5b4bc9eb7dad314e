#!/bin/bash
# Profiling pass (run on the GPU box through gpurun, from the repo root):  bash profiles/run_profile.sh [workload]
# Produces in gpurun_out/: launches_<wl>.csv (every launch with its device time), conv_<wl>.ncu-rep / sm_<wl>.ncu-rep
# (--set full captures of the two hot kernels) and their raw-page CSVs.  Numbers printed under ncu are never bench values.
WL=${1:-fwd16}
mkdir -p gpurun_out
NCU=$(command -v ncu || echo /usr/local/cuda/bin/ncu)
# 1) launch list of two steady-state steps (skip the warm-up launches)
$NCU --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-300} -c ${COUNT:-200} --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_${WL}.log 2>&1
# 2) full capture of the dominant conv kernel (conv5 + conv6 of one step) and of the spatial-model kernel
$NCU --set full --clock-control none --import-source on -k regex:conv_igemm -s ${CONV_SKIP:-26} -c 2 -f -o gpurun_out/conv_${WL} \
    python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_conv_${WL}.log 2>&1
$NCU --set full --clock-control none --import-source on -k regex:sm_conv -s 1 -c 1 -f -o gpurun_out/sm_${WL} \
    python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sm_${WL}.log 2>&1
for r in conv_${WL} sm_${WL}; do
  $NCU -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/$r.raw.csv 2>/dev/null
done
ls -la gpurun_out/
