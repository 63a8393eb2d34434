#!/bin/bash
# One gpurun call: GPU test-suite, then launch-shape A/Bs of the two 3D hot kernels (waves per SM).
set -u
O=gpurun_out/${1:-rg}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -15 $O/pytest.log
B="python bench.py --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e --single-mode --workload 3d_static_dense"
for cfg in "2 b" "1 b" "2 a" "2 b"; do
  set -- $cfg
  echo "step wpb=$1 copy=$2" >> $O/b_step.json
  DMP_3D_STEP_WPB=$1 DMP_3D_STEP_COPY=$2 $B --mode step >> $O/b_step.json 2>&1
done
for w in 1 3 1; do
  echo "rollout wpb=$w" >> $O/b_roll.json
  DMP_3D_WPB=$w $B >> $O/b_roll.json 2>&1
done
for f in $O/b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in [x for x in open(sys.argv[1]).read().splitlines()]:
    if not l.startswith("{"):
        print("  |", l[:200]); continue
    try:
        d=json.loads(l); o=d.get("other_mode") or {}
        print("%.4e frac %.3f" % (d["value"], d["roofline"]["frac"]))
    except Exception as e:
        print("FAILED", e)
PY
done
