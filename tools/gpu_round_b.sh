#!/bin/bash
# One gpurun call: full GPU test-suite, then A/B of the tuning switches added this session
# (1D bulk copy-out: DMP_TILE_COPY=b|l; 3D single-step L2 knobs: DMP_3D_STEP_TUNE=0..5), then the default bench.
set -u
O=gpurun_out/${1:-rb}; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -5 $O/pytest.log
B="python bench.py --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e"
for m in l b l b; do
  DMP_TILE_COPY=$m $B --workload 1d_dynamic >> $O/b_1d_$m.json 2>&1
done
for m in l b; do
  DMP_TILE_COPY=$m $B --workload 1d_dynamic --envs 4194304 --single-mode >> $O/b_1d4m_$m.json 2>&1
done
for t in 0 1 2 3 4 5 0; do
  DMP_3D_STEP_TUNE=$t $B --workload 3d_static_dense --mode step --single-mode >> $O/b_3dstep_t$t.json 2>&1
done
for f in $O/b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in [x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")]:
    try:
        d=json.loads(l); o=d.get("other_mode") or {}
        print("%.4e frac %.3f | other %s %.4e frac %.3f" % (d["value"], d["roofline"]["frac"], o.get("mode"), o.get("value",0), o.get("roofline_frac",0)))
    except Exception as e:
        print("FAILED", e)
PY
done
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 1500 $O/bench_default.json
