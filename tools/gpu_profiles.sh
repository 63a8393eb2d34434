#!/bin/bash
# One gpurun call (1 GPU): launch lists of the default benches and full ncu captures of every hot kernel.
set -u
O=gpurun_out/${1:-prof}; mkdir -p $O
Q="--no-cpu-baseline --no-e2e"
# launch lists (the --metrics gpu__time_duration.sum pass of B200_PROFILING.md)
for wl in 2d_static_dense 3d_static_dense 1d_dynamic; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$wl.csv \
      python bench.py --workload $wl --steps 64 --warmup 32 $Q > $O/launches_$wl.log 2>&1
done
cap() {  # name, kernel regex, skip, bench args...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/$name -f \
      python bench.py --single-mode --steps 64 --warmup 64 $Q "$@" > $O/$name.log 2>&1
}
cap prof_2d_roll16 k2d_rollout 10 --workload 2d_static_dense
cap prof_2d_step k2d_rollout 40 --workload 2d_static_dense --mode step
cap prof_3d_roll16 k3d_cache_rollout 10 --workload 3d_static_dense
cap prof_3d_step k3d_step_bytes 40 --workload 3d_static_dense --mode step
cap prof_1d_roll64 k1d_rollout 6 --workload 1d_dynamic
ls -la $O
