#!/bin/bash
# compute-sanitizer over the kernels' parity tests (small sizes) and smoke():
#   memcheck  -- out-of-bounds / misaligned accesses in global and shared memory, incl. the guard-absorbed over-reads of the
#                3D kernels;
#   racecheck -- shared-memory hazards: the warp tiles are written by every lane, handed to the async proxy
#                (fence.proxy.async -> cp.async.bulk) and refilled in the next step; the 3D caches are cleared cooperatively.
# usage: gpu_sanitize.sh OUTDIR
set -u
O=gpurun_out/${1:-rs}; mkdir -p $O
K="kernels_agree or tall_columns or bulk_copy_out or rollout_shapes or stage or philox or shard"
CS="compute-sanitizer --error-exitcode 86 --print-limit 20"
timeout 1500 $CS --tool memcheck python -m pytest tests/test_parity_gpu.py tests/test_records_gpu.py -m gpu -x -q \
   -k "$K or record_rollout or host_stepper or bit_record or device_unpack" > $O/memcheck_parity.log 2>&1
echo "memcheck parity exit $?" | tee -a $O/memcheck_parity.log
timeout 600 $CS --tool memcheck python __graft_entry__.py smoke > $O/memcheck_smoke.log 2>&1
echo "memcheck smoke exit $?" | tee -a $O/memcheck_smoke.log
timeout 1500 $CS --tool racecheck --racecheck-report all python -m pytest tests/test_parity_gpu.py tests/test_records_gpu.py -m gpu -x -q \
   -k "bulk_copy_out or kernels_agree or rollout_shapes or tall_columns or record_rollout or bit_record" > $O/racecheck_parity.log 2>&1
echo "racecheck parity exit $?" | tee -a $O/racecheck_parity.log
timeout 900 $CS --tool racecheck --racecheck-report all python __graft_entry__.py smoke > $O/racecheck_smoke.log 2>&1
echo "racecheck smoke exit $?" | tee -a $O/racecheck_smoke.log
for f in memcheck_parity memcheck_smoke racecheck_parity racecheck_smoke; do echo "== $f"; tail -6 $O/$f.log; done
grep -c "Invalid\|out of bounds\|misaligned\|hazard" $O/*.log || true
