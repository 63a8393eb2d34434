#!/bin/bash
# A/B of an environment switch over the main workloads: gpu_ab.sh OUTDIR VAR=VALUE
set -u
O=gpurun_out/${1:-rab}; mkdir -p $O; SW=${2:-DMP_TILE_COPY=l}
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
B="python bench.py --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e"
for wl in 2d_static_dense 2d_dynamic_dense 3d_static_dense 3d_dynamic_dense; do
  $B --workload $wl > $O/b_${wl}_A.json 2>&1
  env $SW $B --workload $wl > $O/b_${wl}_B.json 2>&1
  $B --workload $wl > $O/b_${wl}_A2.json 2>&1
done
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
