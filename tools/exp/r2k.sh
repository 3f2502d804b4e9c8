#!/bin/bash
# batch K: oversubscription of the streamed kernel on config 2, PNAConv layer profile, config-5 path at world = 1
cd "$(dirname "$0")/../.."
O=gpurun_out/r2k; mkdir -p $O
for os in 1 2 3 4 6 8; do
  for c in 2 2u; do
    PNA_B200_OVERSUB=$os timeout 600 python tools/exp/agg_time.py --config $c --steps 30 --tag os$os >> $O/cfg.jsonl 2>> $O/err.log
  done
done
PNA_B200_OVERSUB=4 timeout 600 python tools/exp/agg_time.py --config 5 --steps 20 --tag os4 >> $O/cfg.jsonl 2>> $O/err.log
timeout 300 python tools/exp/layer_profile.py > $O/layer_profile.txt 2>> $O/err.log
PNA_BENCH_FORCE_MULTI=1 timeout 600 python bench.py --steps 10 --warmup 3 > $O/multi_c5_w1.json 2> $O/multi_c5_w1.err; echo "c5 rc $?" >> $O/status.log
echo done
