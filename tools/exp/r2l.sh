#!/bin/bash
# batch L: full GPU suite (incl. the tensor-core PNAConv path), layer profile, full bench line
cd "$(dirname "$0")/../.."
O=gpurun_out/r2l; mkdir -p $O
(timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
timeout 300 python tools/exp/layer_profile.py > $O/layer_profile.txt 2>> $O/err.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc $?" >> $O/status.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/status.log
echo done
