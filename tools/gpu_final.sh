#!/bin/bash
# Round-end check on a fresh box: GPU tests, smoke(), the default bench and the reference arm exactly as the driver runs
# them, the policy-loop throughput, and a refreshed launch list + ncu capture of the 1D kernel.
set -u
O=gpurun_out/${1:-final}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -4 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log; tail -4 $O/smoke.log
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 700 $O/bench_reference.json
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.err; python - $O/bench_default.json <<'PY'
import json,sys
d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print({k: d[k] for k in ("metric","value","unit","n_gpus","steps","warmup","ms_per_step","scaling","dtype","gpu_launches")})
print("roofline", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "i16", d["e2e_i16"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline_c"].get("value"), "clocks", d["clocks"])
PY
python tools/bench_policy_loop.py > $O/policy_loop.jsonl 2> $O/policy_loop.err; cat $O/policy_loop.jsonl | cut -c1-300
Q="--no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_1d_dynamic.csv \
    python bench.py --workload 1d_dynamic --steps 64 --warmup 32 $Q > $O/launches_1d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1d_rollout -s 6 -c 1 -o $O/prof_1d_roll64 -f \
    python bench.py --single-mode --steps 64 --warmup 64 $Q --workload 1d_dynamic > $O/prof_1d.log 2>&1
ls -la $O | head -30
