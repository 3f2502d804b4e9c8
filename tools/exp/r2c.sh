#!/bin/bash
# batch C: store-policy / occupancy / ring-size variants on configs 2, 4, 5 + ncu full profile of the stream kernel
# (reports are exported to CSV on the box: the .ncu-rep files exceed what gpurun copies back)
cd "$(dirname "$0")/../.."
O=gpurun_out/r2c; mkdir -p $O
for v in default plain minb8 half2k half8k; do
  if [ $v = default ]; then unset PNA_B200_LIB; else export PNA_B200_LIB=$PWD/variants/$v/libpna_sm100.so; fi
  for c in 2 5 4 3p; do
    timeout 600 python tools/exp/agg_time.py --config $c --steps 20 --tag $v >> $O/cfg.jsonl 2>> $O/err.log
  done
done
unset PNA_B200_LIB
prof() { # name config kernel-regex
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s 1 -c 1 -o /tmp/prof_$1 python tools/exp/agg_time.py --config $2 --once > /dev/null 2>> $O/err.log
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > $O/prof_$1_raw.csv 2>> $O/err.log
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > $O/prof_$1_source.csv 2>> $O/err.log
  ls -la /tmp/prof_$1.ncu-rep >> $O/err.log
}
prof cfg5 5 k_rows_stream
prof cfg2 2 k_rows_stream
prof cfg4 4 k_rows_tiled
du -sh $O >> $O/err.log
echo done
