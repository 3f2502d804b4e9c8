#!/bin/bash
# batch G: new default (pipelined shared-memory reads, L1 gathers chosen per graph, batched tree merge): tests, configs, bench line
cd "$(dirname "$0")/../.."
O=gpurun_out/r2g; mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
for c in 2 5 4 3p 3 1; do
  timeout 600 python tools/exp/agg_time.py --config $c --steps 20 --tag new >> $O/cfg.jsonl 2>> $O/err.log
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -k regex:'k_rows|k_hub' --csv --log-file $O/ncu_cfg5.csv python tools/exp/agg_time.py --config 5 --once > /dev/null 2>> $O/err.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
echo done
