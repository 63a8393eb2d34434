#!/bin/bash
# One gpurun call: GPU tests, A/B benches of the round's kernel changes, ncu captures.  Outputs -> gpurun_out/
set -u
O=gpurun_out/${1:-ra}; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log; tail -3 $O/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
B="python bench.py --steps 8192 --warmup 1024"
$B > $O/b_2d.json 2> $O/b_2d.err
DMP_PDL=0 $B --no-cpu-baseline --no-e2e > $O/b_2d_nopdl.json 2>&1
$B --workload 1d_dynamic --no-cpu-baseline --no-e2e > $O/b_1d.json 2>&1
$B --workload 1d_dynamic --rollout-k 16 --no-cpu-baseline --no-e2e > $O/b_1d_k16.json 2>&1
$B --workload 1d_static_step --no-cpu-baseline --no-e2e > $O/b_1d_static.json 2>&1
$B --workload 1d_dynamic --envs 4194304 --no-cpu-baseline --no-e2e > $O/b_1d_4m.json 2>&1
$B --workload 3d_static_dense --no-cpu-baseline --no-e2e > $O/b_3d.json 2>&1
DMP_PDL=0 $B --workload 3d_static_dense --no-cpu-baseline --no-e2e > $O/b_3d_nopdl.json 2>&1
DMP_3D_KERNEL=r $B --workload 3d_static_dense --mode step --single-mode --no-cpu-baseline --no-e2e > $O/b_3d_rows1.json 2>&1
$B --workload 3d_dynamic_dense --no-cpu-baseline --no-e2e > $O/b_3d_dyn.json 2>&1
for f in $O/b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    l=[x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")][-1]; d=json.loads(l)
    o=d.get("other_mode") or {}
    print("%.3e frac %.3f | other %s %.3e frac %.3f | e2e %.3e" % (d["value"], d["roofline"]["frac"], o.get("mode"), o.get("value",0), o.get("roofline_frac",0), d["e2e"]["value"]))
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1]).read()[-800:])
PY
done
# ncu: the new single-step 3D kernel and the 1D rollout kernel (full sets), plus launch lists
ncu --set full --clock-control none --import-source on -k regex:k3d_step_span -s 40 -c 1 -o $O/prof_3d_step_span -f \
    python bench.py --workload 3d_static_dense --mode step --single-mode --steps 64 --warmup 64 --no-cpu-baseline --no-e2e > $O/ncu_3d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k1d_rollout -s 20 -c 1 -o $O/prof_1d_roll -f \
    python bench.py --workload 1d_dynamic --single-mode --steps 64 --warmup 64 --no-cpu-baseline --no-e2e > $O/ncu_1d.log 2>&1
ls -la $O
