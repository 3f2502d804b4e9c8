#!/bin/bash
# batch M (final evidence): ncu full captures of the final kernels (exported to CSV on the box), launch list of a bench run
cd "$(dirname "$0")/../.."
O=gpurun_out/r2m; mkdir -p $O
prof() { # name config kernel-regex skip
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c 1 -o /tmp/prof_$1 python tools/exp/agg_time.py --config $2 --once > /dev/null 2>> $O/err.log
  ncu -i /tmp/prof_$1.ncu-rep --page raw --csv > $O/prof_$1_raw.csv 2>> $O/err.log
  ncu -i /tmp/prof_$1.ncu-rep --page source --csv > $O/prof_$1_source.csv 2>> $O/err.log
}
prof cfg2 2 k_rows_stream 1
prof cfg5 5 k_rows_stream 1
prof cfg4 4 k_rows_tiled 1
prof cfg2_finalize 2 k_hub_finalize 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-side-configs > $O/bench_under_ncu.json 2>> $O/err.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'k_rows|k_hub' --csv --log-file $O/launches_cfg5.csv \
   python tools/exp/agg_time.py --config 5 --once > /dev/null 2>> $O/err.log
echo done
