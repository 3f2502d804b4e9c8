#!/bin/bash
# batch J: bench_multi.py code paths on one GPU (world = 1), full test suite, full single-GPU bench line
cd "$(dirname "$0")/../.."
O=gpurun_out/r2j; mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
PNA_BENCH_FORCE_MULTI=1 PNA_BENCH_WORKLOAD=config4 PNA_BENCH_C4_GRAPHS=15000 timeout 600 python bench.py --steps 10 --warmup 3 > $O/multi_c4_w1.json 2> $O/multi_c4_w1.err; echo "c4 rc $?" >> $O/status.log
PNA_BENCH_FORCE_MULTI=1 timeout 600 python bench.py --steps 10 --warmup 3 > $O/multi_c5_w1.json 2> $O/multi_c5_w1.err; echo "c5 rc $?" >> $O/status.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc $?" >> $O/status.log
echo done
