#!/bin/bash
# One gpurun call: bench line of every workload (kernel numbers + e2e), then launch lists and full ncu captures.
set -u
O=gpurun_out/${1:-all}; mkdir -p $O
for wl in 2d_static_dense 2d_dynamic_dense 3d_static_dense 3d_dynamic_dense 1d_dynamic 1d_static_step; do
  python bench.py --workload $wl --steps 8192 --warmup 1024 --no-cpu-baseline > $O/bench_$wl.json 2> $O/bench_$wl.err
done
python bench.py --workload 1d_dynamic --envs 4194304 --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e > $O/bench_1d_dynamic_4m.json 2>&1
for f in $O/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in [x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")]:
    d=json.loads(l); o=d.get("other_mode") or {}; r=d["roofline"]
    print("%.4e frac %.3f (layout %.3f) | other %s %.4e frac %.3f (layout %.3f) | e2e %.3e i16 %s" % (d["value"], r["frac"], r["frac_this_layout"], o.get("mode"), o.get("value",0), o.get("roofline_frac",0), o.get("roofline_frac_this_layout",0), d["e2e"]["value"], (d.get("e2e_i16") or {}).get("value")))
PY
done
bash tools/gpu_profiles.sh ${1:-all}/prof > $O/profiles.log 2>&1
ls $O $O/prof | head -60
