#!/bin/bash
# batch F: is config 5 bound by hot source rows (L2 slice bandwidth)?  L1-allocating gathers, uniform-source control, write-only control
cd "$(dirname "$0")/../.."
O=gpurun_out/r2f; mkdir -p $O
for v in default l1; do
  if [ $v = default ]; then unset PNA_B200_LIB; else export PNA_B200_LIB=$PWD/variants/$v/libpna_sm100.so; fi
  for c in 5 5u 5w 2 4; do
    timeout 600 python tools/exp/agg_time.py --config $c --steps 20 --tag $v >> $O/cfg.jsonl 2>> $O/err.log
  done
done
echo done
