#!/bin/bash
# quick GPU check: tests + default bench + policy-loop throughput.  Outputs -> gpurun_out/$1
set -u
O=gpurun_out/${1:-rq}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.json
python bench.py --impl reference --steps 2000 --warmup 100 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 400 $O/bench_reference.json
python tools/bench_policy_loop.py > $O/policy_loop.jsonl 2> $O/policy_loop.err; cat $O/policy_loop.jsonl; tail -3 $O/policy_loop.err
