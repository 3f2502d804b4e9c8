#!/bin/bash
# batch U (last GPU call of the round): the coefficient backward -- full GPU suite with it as the default, the backward tests
# again on the one-call path, the A/B timing, smoke
cd "$(dirname "$0")/../.."
O=gpurun_out/r2u; mkdir -p $O
(timeout 120 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
(PNA_B200_BWD=atomic timeout 60 python -m pytest tests -m gpu -x -q -k "backward or trains or compact or readouts" > $O/pytest_atomic.log 2>&1; echo "pytest rc $?" >> $O/pytest_atomic.log)
timeout 100 python tools/exp/bwd_ab.py --steps 3 > $O/bwd_ab.jsonl 2> $O/bwd_ab.err; echo "ab rc $?" >> $O/status.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/status.log
echo done
