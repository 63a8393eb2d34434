#!/bin/bash
# 2-GPU call: scaling sanity of the bench (torchrun, NCCL stats all-reduce), both arms, plus the ref3d action distribution.
set -u
O=gpurun_out/${1:-ri}; mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$T bench.py --gpus 2 --steps 8192 --warmup 1024 > $O/bench_2d_n2.json 2> $O/bench_2d_n2.err
$T bench.py --gpus 2 --steps 8192 --warmup 1024 --workload 3d_static_dense --no-e2e-i16 > $O/bench_3d_n2.json 2> $O/bench_3d_n2.err
$T bench.py --impl reference --gpus 2 --steps 2000 --warmup 100 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
python bench.py --workload 3d_static_dense --action-dist ref3d --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e > $O/bench_3d_ref3d.json 2> $O/bench_3d_ref3d.err
python bench.py --workload 3d_dynamic_dense --action-dist ref3d --steps 8192 --warmup 1024 --no-cpu-baseline --no-e2e > $O/bench_3dd_ref3d.json 2> $O/bench_3dd_ref3d.err
python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_k100.json 2> $O/bench_k100.err
for f in $O/bench_*.json; do echo "== $f"; tail -c 300 ${f%.json}.err; python - "$f" <<'PY'
import json,sys
for l in [x for x in open(sys.argv[1]).read().splitlines() if x.startswith("{")]:
    d=json.loads(l); o=d.get("other_mode") or {}; r=d.get("roofline") or {}
    print("n=%s steps=%s %.4e frac %s | other %s %.4e | e2e %.3e | launches %s | eps %s" % (d.get("n_gpus"), d.get("steps"), d["value"], r.get("frac"), o.get("mode"), o.get("value",0), d["e2e"]["value"], d.get("gpu_launches"), json.dumps(d.get("episode_stats"))[:160]))
PY
done
