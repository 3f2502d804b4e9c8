#!/bin/bash
# batch N8: the 8-GPU bench line (configs[4], 10 M nodes / 100 M edges)
cd "$(dirname "$0")/../.."
O=gpurun_out/r2n8; mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench8.json 2> $O/bench8.err
echo "rc $?" >> $O/status.log
nvidia-smi topo -m > $O/topo.txt 2>&1
