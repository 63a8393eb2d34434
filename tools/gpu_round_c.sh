#!/bin/bash
# One gpurun call: GPU test-suite, then the 3D benches (byte-shadow kernels) with the u16 span kernel beside them.
set -u
O=gpurun_out/${1:-rc}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -15 $O/pytest.log
B="python bench.py --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e"
for wl in 3d_static_dense 3d_dynamic_dense; do
  $B --workload $wl >> $O/b_${wl}.json 2>&1
  DMP_3D_KERNEL=s $B --workload $wl --mode step --single-mode >> $O/b_${wl}_span.json 2>&1
  $B --workload $wl --mode step --single-mode >> $O/b_${wl}_bytes.json 2>&1
done
for f in $O/b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in [x for x in open(sys.argv[1]).read().splitlines()]:
    if not l.startswith("{"):
        print("  |", l[:200]); continue
    try:
        d=json.loads(l); o=d.get("other_mode") or {}
        print("%.4e frac %.3f | other %s %.4e frac %.3f" % (d["value"], d["roofline"]["frac"], o.get("mode"), o.get("value",0), o.get("roofline_frac",0)))
    except Exception as e:
        print("FAILED", e)
PY
done
