#!/bin/bash
# Round-end check on a fresh box: GPU tests, smoke(), the default bench and the reference arm exactly as the driver runs
# them (--steps 20 --warmup 5), the policy-loop throughput.  usage: gpu_final.sh OUTDIR
set -u
O=gpurun_out/${1:-final}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -4 $O/smoke.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 700 $O/bench_reference.json
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.err; python tools/show_bench.py $O/bench_default.json
python tools/bench_policy_loop.py > $O/policy_loop.jsonl 2> $O/policy_loop.err; cat $O/policy_loop.jsonl | cut -c1-300
ls -la $O | head -30
