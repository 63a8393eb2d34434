#!/bin/bash
# ncu evidence: launch lists (the --metrics gpu__time_duration.sum pass of B200_PROFILING.md) of the default bench regions,
# `ncu --set full` captures of the hot kernels, and a DRAM-traffic pass over the shard sizes of the 1/2/4/8-GPU runs
# (tools/make_traffic.py turns it into profiles/traffic.json).  usage: gpu_profiles.sh OUTDIR
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
# the bit-record instantiation (DMP_OBS_BITS): one 20-step launch into device buffers, tools/time_kinds.py
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:Bits16 -s 6 -c 1 -o $O/prof_2d_bits_roll20 -f \
    python tools/time_kinds.py > $O/prof_2d_bits_roll20.log 2>&1
traffic() {  # workload, envs, steps, kernel regex, skip, extra args
  local wl=$1 envs=$2 K=$3 k=$4 skip=$5; shift 5
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:$k -s $skip -c 1 \
      --csv --log-file "$O/traffic_${wl}_${envs}_K${K}.csv" python bench.py $Q --workload $wl --envs $envs "$@" > /dev/null 2>&1
}
for e in 1048576 524288 262144 131072; do traffic 2d_static_dense $e 20 k2d_rollout 4 --steps 20 --warmup 5; done
traffic 2d_static_dense 1048576 1 k2d_rollout 40 --mode step --steps 64 --warmup 16
for e in 262144 131072 65536 32768; do traffic 3d_static_dense $e 20 k3d_cache_rollout 4 --steps 20 --warmup 40; done
traffic 3d_static_dense 262144 1 k3d_step_bytes 100 --mode step --steps 64 --warmup 64
traffic 1d_dynamic 65536 64 k1d_rollout 4 --steps 64 --warmup 64
traffic 2d_dynamic_dense 1048576 20 k2d_rollout 4 --steps 20 --warmup 40
ls -la $O
