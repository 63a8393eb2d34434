#!/bin/bash
# timing experiments for the 3D kernels (no test run: DMP_3D_EXP bit 0 leaves the u16 maps stale on purpose)
set -u
O=gpurun_out/${1:-rf}; mkdir -p $O
B="python bench.py --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e --single-mode --workload 3d_static_dense"
for cfg in "0 b" "1 b" "2 b" "4 b" "3 b" "5 b" "2 a" "3 a" "5 a"; do
  set -- $cfg
  echo "step exp=$1 copy=$2" >> $O/b_step.json
  DMP_3D_EXP=$1 DMP_3D_STEP_COPY=$2 $B --mode step >> $O/b_step.json 2>&1
done
for x in 0 1; do
  echo "rollout exp=$x" >> $O/b_roll.json
  DMP_3D_EXP=$x $B >> $O/b_roll.json 2>&1
done
for f in $O/b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in [x for x in open(sys.argv[1]).read().splitlines()]:
    if not l.startswith("{"):
        print("  |", l[:200]); continue
    try:
        d=json.loads(l)
        print("%.4e frac %.3f" % (d["value"], d["roofline"]["frac"]))
    except Exception as e:
        print("FAILED", e)
PY
done
