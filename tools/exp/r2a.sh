#!/bin/bash
# round-2 experiment batch A: feature-split variants on config 2 + per-config launch lists
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
for rep in 1 2; do
for mode in default tiled2 stream2; do
  if [ $mode = default ]; then unset PNA_B200_FEAT_SPLIT; else export PNA_B200_FEAT_SPLIT=$mode; fi
  timeout 300 python tools/exp/agg_time.py --config 2 --tag $mode >> $O/cfg2_modes.jsonl 2>> $O/err.log
done
done
unset PNA_B200_FEAT_SPLIT
PNA_B200_FOLD_FINALIZE=1 timeout 300 python tools/exp/agg_time.py --config 2 --tag fold >> $O/cfg2_modes.jsonl 2>> $O/err.log
for mode in default tiled2 stream2; do
  PNA_B200_FEAT_SPLIT=$mode timeout 300 python tools/exp/agg_time.py --config 2u --tag $mode >> $O/cfg2u_modes.jsonl 2>> $O/err.log
done
# dram bytes per mode
for mode in default tiled2 stream2; do
  PNA_B200_FEAT_SPLIT=$mode timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
     --clock-control none -k regex:'k_rows|k_hub' --csv --log-file $O/ncu_cfg2_$mode.csv python tools/exp/agg_time.py --config 2 --once > /dev/null 2>> $O/err.log
done
# launch lists per config (kernel breakdown)
for c in 1 3 3p 4 5; do
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_rows|k_hub' --csv --log-file $O/ncu_cfg$c.csv \
     python tools/exp/agg_time.py --config $c --once > /dev/null 2>> $O/err.log
  timeout 600 python tools/exp/agg_time.py --config $c --steps 20 --tag base >> $O/cfg_all.jsonl 2>> $O/err.log
done
echo done
