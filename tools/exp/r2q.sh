#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2q; mkdir -p $O
timeout 300 python tools/exp/e2e_trace.py > $O/e2e_plain.txt 2>&1
timeout 300 python tools/exp/e2e_trace.py --like-bench > $O/e2e_like_bench.txt 2>&1
echo done
