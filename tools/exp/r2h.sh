#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2h; mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
for c in 2 5 4; do
  timeout 600 python tools/exp/agg_time.py --config $c --steps 20 --tag new2 >> $O/cfg.jsonl 2>> $O/err.log
done
for mb in 32 64 96; do
  PNA_EXP_L2_PERSIST_MB=$mb timeout 600 python tools/exp/agg_time.py --config 2 --steps 20 --tag persist$mb >> $O/cfg.jsonl 2>> $O/err.log
done
timeout 300 python tools/exp/layer_err.py > $O/layer_err.json 2>> $O/err.log
nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/probe tools/probes/probe_tmem_a.cu > $O/probe_tmem.log 2>&1 && timeout 60 /tmp/probe >> $O/probe_tmem.log 2>&1; echo "probe rc $?" >> $O/probe_tmem.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o /tmp/linear_ts tools/probes/linear_ts.cu > $O/probe_linear_ts.log 2>&1 && timeout 60 /tmp/linear_ts >> $O/probe_linear_ts.log 2>&1; echo "probe rc $?" >> $O/probe_linear_ts.log
echo done
