#!/bin/bash
# batch E: ring depth variants (NST x segment bytes) on configs 2 / 5, feature passes on top, parity tests of the new default
cd "$(dirname "$0")/../.."
O=gpurun_out/r2e; mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
for v in default st3 st4h2 h8 st4; do
  if [ $v = default ]; then unset PNA_B200_LIB; else export PNA_B200_LIB=$PWD/variants/$v/libpna_sm100.so; fi
  for c in 2 5; do
    timeout 600 python tools/exp/agg_time.py --config $c --steps 20 --tag $v >> $O/cfg.jsonl 2>> $O/err.log
  done
  PNA_B200_FEAT_SPLIT=stream2 timeout 600 python tools/exp/agg_time.py --config 2 --steps 20 --tag ${v}_fsplit >> $O/cfg.jsonl 2>> $O/err.log
done
unset PNA_B200_LIB
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed.sum,smsp__cycles_active.avg,sm__cycles_elapsed.max --clock-control none -k regex:'k_rows|k_hub' --csv --log-file $O/ncu_cfg5.csv python tools/exp/agg_time.py --config 5 --once > /dev/null 2>> $O/err.log
echo done
