#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2r; mkdir -p $O
for i in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --no-side-configs --no-cpu-baseline > $O/bench_$i.json 2> $O/bench_$i.err
done
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current,clocks.sm,clocks.mem --format=csv > $O/smi.txt 2>&1
echo done
