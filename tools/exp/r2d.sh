#!/bin/bash
# batch D (2 GPUs): the multi-GPU bench on configs[4] at 2/8 scale: small smoke first, then full size with two row-cost weights
cd "$(dirname "$0")/../.."
O=gpurun_out/r2d; mkdir -p $O
run() { # name, env..., then bench args
  name=$1; shift
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 2 --steps ${STEPS:-10} --warmup 3 > $O/$name.json 2> $O/$name.err
  echo "$name rc $?" >> $O/status.log
}
STEPS=5 run small PNA_BENCH_C5_NODES_PER_GPU=100000 PNA_BENCH_C5_EDGES_PER_GPU=1000000
if ! grep -q '"metric"' $O/small.json; then echo "small failed" >> $O/status.log; tail -30 $O/small.err >> $O/status.log; exit 0; fi
run full_rc12 PNA_BENCH_ROW_COST=12
run full_rc32 PNA_BENCH_ROW_COST=32 PNA_BENCH_ALL_PLANES=0
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 3 > $O/ref.json 2> $O/ref.err
echo done >> $O/status.log
