#!/bin/bash
set -u
O=gpurun_out/${1:-rst}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
B="python bench.py --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e"
$B > $O/b_2d.json 2>&1
$B --workload 3d_static_dense > $O/b_3d.json 2>&1
$B --workload 3d_dynamic_dense > $O/b_3d_dyn.json 2>&1
$B --workload 2d_dynamic_dense > $O/b_2d_dyn.json 2>&1
$B --workload 1d_dynamic > $O/b_1d.json 2>&1
for f in $O/b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    l=[x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")][-1]; d=json.loads(l)
    o=d.get("other_mode") or {}
    print("%.3e frac %.3f | other %s %.3e frac %.3f" % (d["value"], d["roofline"]["frac"], o.get("mode"), o.get("value",0), o.get("roofline_frac",0)))
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1]).read()[-800:])
PY
done
ncu --set full --clock-control none --import-source on -k regex:k3d_step_span -s 40 -c 1 -o $O/prof_3d_step -f \
    python bench.py --workload 3d_static_dense --mode step --single-mode --steps 64 --warmup 64 --no-cpu-baseline --no-e2e > $O/ncu_3d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k2d_rollout -s 40 -c 1 -o $O/prof_2d_step -f \
    python bench.py --mode step --single-mode --steps 64 --warmup 64 --no-cpu-baseline --no-e2e > $O/ncu_2d.log 2>&1
