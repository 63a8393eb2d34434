#!/bin/bash
# One gpurun call: GPU test-suite, A/B of the byte step kernel's copy path, ncu captures of the two 3D hot kernels.
set -u
O=gpurun_out/${1:-rd}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -15 $O/pytest.log
B="python bench.py --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e"
for c in a b a; do
  DMP_3D_STEP_COPY=$c $B --workload 3d_static_dense --mode step --single-mode >> $O/b_3dstep_$c.json 2>&1
done
$B --workload 3d_dynamic_dense --mode step --single-mode >> $O/b_3dstep_dyn.json 2>&1
$B --workload 3d_static_dense --single-mode >> $O/b_3droll.json 2>&1
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
Q="--no-cpu-baseline --no-e2e"
cap() {  # name, kernel regex, skip, bench args...
  local name=$1 k=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $O/$name -f \
      python bench.py --single-mode --steps 64 --warmup 64 $Q "$@" > $O/$name.log 2>&1
}
cap prof_3d_roll16 k3d_cache_rollout 10 --workload 3d_static_dense
cap prof_3d_step k3d_step_bytes 40 --workload 3d_static_dense --mode step
ls -la $O
