#!/bin/bash
# batch O (2 GPUs): config 5 at 2/8 scale with the per-rank aggregation-alone timing
cd "$(dirname "$0")/../.."
O=gpurun_out/r2o; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench2.json 2> $O/bench2.err
echo "rc $?" >> $O/status.log
