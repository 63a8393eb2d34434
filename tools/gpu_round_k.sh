#!/bin/bash
# GPU test-suite + the 1D benches
set -u
O=gpurun_out/${1:-rk}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -12 $O/pytest.log
B="python bench.py --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e"
for wl in 1d_dynamic 1d_static_step; do $B --workload $wl >> $O/b_1d.json 2>&1; done
$B --workload 1d_dynamic --rollout-k 16 --single-mode >> $O/b_1d.json 2>&1
$B --workload 1d_dynamic --envs 4194304 >> $O/b_1d.json 2>&1
for f in $O/b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in [x for x in open(sys.argv[1]).read().splitlines()]:
    if not l.startswith("{"):
        print("  |", l[:200]); continue
    d=json.loads(l); o=d.get("other_mode") or {}
    print("%s %.4e frac %.3f | other %s %.4e frac %.3f" % (d["config"]["workload"], d["value"], d["roofline"]["frac"], o.get("mode"), o.get("value",0), o.get("roofline_frac",0)))
PY
done
