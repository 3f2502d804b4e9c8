#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2s; mkdir -p $O
(timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/status.log
echo done
