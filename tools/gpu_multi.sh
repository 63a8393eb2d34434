#!/bin/bash
# multi-GPU bench (one box): launched exactly like the driver does.  usage: gpu_multi.sh N outdir
set -u
N=${1:-2}; O=gpurun_out/${2:-multi$N}; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 8192 --warmup 1024 > $O/b_2d.json 2> $O/b_2d.err
run --steps 8192 --warmup 1024 --workload 3d_static_dense > $O/b_3d.json 2> $O/b_3d.err
run --steps 8192 --warmup 1024 --workload 2d_dynamic_dense > $O/b_2d_dyn.json 2> $O/b_2d_dyn.err
run --steps 8192 --warmup 1024 --workload 1d_dynamic > $O/b_1d.json 2> $O/b_1d.err
run --impl reference --steps 2000 --warmup 100 > $O/b_ref.json 2> $O/b_ref.err
for f in $O/b_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    l=[x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")][-1]; d=json.loads(l)
    o=d.get("other_mode") or {}
    print("n_gpus %s value %.3e frac %s | other %s %.3e frac %.3f | e2e %.3e" % (d.get("n_gpus"), d["value"], d.get("roofline",{}).get("frac"), o.get("mode"), o.get("value",0), o.get("roofline_frac",0), d["e2e"]["value"]))
except Exception as e:
    print("FAILED", e); print(open(sys.argv[1]).read()[-800:]); print(open(sys.argv[1][:-4]+"err").read()[-1500:])
PY
done
