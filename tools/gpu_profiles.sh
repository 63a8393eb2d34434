#!/bin/bash
# ncu evidence: launch lists (the --metrics gpu__time_duration.sum pass of B200_PROFILING.md) of the default bench regions and
# `ncu --set full` captures of the hot kernels.  usage: gpu_profiles.sh OUTDIR [quick]
set -u
O=gpurun_out/${1:-prof}; mkdir -p $O
Q="--single-mode --no-workloads --no-cpu-baseline --no-e2e"
for wl in 2d_static_dense 3d_static_dense 1d_dynamic; do
  K=20; [ $wl = 1d_dynamic ] && K=64
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$wl.csv \
      python bench.py --workload $wl --steps $K --warmup 5 $Q > $O/launches_$wl.log 2>&1
done
cap() {  # name, kernel regex, skip, bench args...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/$name -f \
      python bench.py $Q "$@" > $O/$name.log 2>&1
}
cap prof_2d_roll20 k2d_rollout 4 --workload 2d_static_dense --steps 20 --warmup 5
cap prof_2d_roll20_131072 k2d_rollout 4 --workload 2d_static_dense --steps 20 --warmup 5 --envs 131072
cap prof_2d_step k2d_rollout 40 --workload 2d_static_dense --mode step --steps 64 --warmup 16
cap prof_3d_roll20 k3d_cache_rollout 4 --workload 3d_static_dense --steps 20 --warmup 40
cap prof_3d_step k3d_step_bytes 100 --workload 3d_static_dense --mode step --steps 64 --warmup 64
cap prof_1d_roll64 k1d_rollout 4 --workload 1d_dynamic --steps 64 --warmup 64
ls -la $O
