#!/bin/bash
# round-2 batch B: GPU tests, memory-bandwidth ceilings, per-config timing of the leaner stream kernel, full bench line
cd "$(dirname "$0")/../.."
O=gpurun_out/r2b; mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
timeout 120 python tools/membw.py > $O/membw.json 2>> $O/err.log
for c in 2 2u 5 4 1 3 3p; do
  timeout 600 python tools/exp/agg_time.py --config $c --steps 20 --tag lean >> $O/cfg.jsonl 2>> $O/err.log
done
PNA_B200_FOLD_FINALIZE=1 timeout 300 python tools/exp/agg_time.py --config 2 --tag lean_fold >> $O/cfg.jsonl 2>> $O/err.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_ref.json 2>> $O/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__inst_executed.sum --clock-control none -k regex:'k_rows|k_hub' --csv --log-file $O/ncu_cfg2.csv python tools/exp/agg_time.py --config 2 --once > /dev/null 2>> $O/err.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum --clock-control none -k regex:'k_rows|k_hub' --csv --log-file $O/ncu_cfg5.csv python tools/exp/agg_time.py --config 5 --once > /dev/null 2>> $O/err.log
echo done
