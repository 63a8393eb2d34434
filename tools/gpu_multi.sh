#!/bin/bash
# multi-GPU bench (one box), launched exactly like the driver does.  usage: gpu_multi.sh "8 4 2" OUTDIR
set -u
NS=${1:-2}; O=gpurun_out/${2:-multi}; mkdir -p $O
for N in $NS; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
      bench.py --gpus $N --steps 20 --warmup 5 > $O/b_n$N.json 2> $O/b_n$N.err; echo "N=$N rc=$?"; tail -2 $O/b_n$N.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N \
      bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/ref_n$N.json 2> $O/ref_n$N.err
done
python tools/show_bench.py $O/b_n*.json $O/ref_n*.json
