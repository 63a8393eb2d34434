#!/bin/bash
# compute-sanitizer memcheck over the kernels' parity tests (small sizes): out-of-bounds / misaligned accesses in global and
# shared memory, including the guard-absorbed over-reads of the 3D kernels.
set -u
O=gpurun_out/${1:-rs}; mkdir -p $O
CS="compute-sanitizer --tool memcheck --error-exitcode 86 --print-limit 20"
timeout 1500 $CS python -m pytest tests/test_parity_gpu.py -m gpu -x -q \
   -k "kernels_agree or tall_columns or bulk_copy_out or rollout_shapes or stage or philox or shard" > $O/memcheck_parity.log 2>&1
echo "memcheck parity exit $?" | tee -a $O/memcheck_parity.log
timeout 600 $CS python __graft_entry__.py smoke > $O/memcheck_smoke.log 2>&1
echo "memcheck smoke exit $?" | tee -a $O/memcheck_smoke.log
tail -8 $O/memcheck_parity.log; tail -6 $O/memcheck_smoke.log
grep -c "Invalid\|out of bounds\|misaligned" $O/memcheck_parity.log $O/memcheck_smoke.log || true
