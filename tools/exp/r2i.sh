#!/bin/bash
# batch I: dynamic tail of the streamed kernel: tests + timing with and without
cd "$(dirname "$0")/../.."
O=gpurun_out/r2i; mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
for c in 2 5 2u; do
  timeout 600 python tools/exp/agg_time.py --config $c --steps 30 --tag dyn >> $O/cfg.jsonl 2>> $O/err.log
  PNA_B200_DYNAMIC_TAIL=0 timeout 600 python tools/exp/agg_time.py --config $c --steps 30 --tag static >> $O/cfg.jsonl 2>> $O/err.log
done
timeout 600 ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -k regex:'k_rows' --csv --log-file $O/ncu_cfg2.csv python tools/exp/agg_time.py --config 2 --once > /dev/null 2>> $O/err.log
echo done
