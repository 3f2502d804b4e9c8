#!/bin/bash
# batch P: final single-GPU verification: full GPU suite, smoke, full bench line, reference arm (short)
cd "$(dirname "$0")/../.."
O=gpurun_out/r2p; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/status.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc $?" >> $O/status.log
echo done
