#!/bin/bash
# batch N4: the 4-GPU bench line (configs[3], 60 000 superpixel graphs, graph-batch shard)
cd "$(dirname "$0")/../.."
O=gpurun_out/r2n4; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench4.json 2> $O/bench4.err
echo "rc $?" >> $O/status.log
