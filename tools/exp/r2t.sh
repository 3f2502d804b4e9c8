#!/bin/bash
cd "$(dirname "$0")/../.."
O=gpurun_out/r2t; mkdir -p $O
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc $?" >> $O/status.log
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "layer_output_error or golden_dgl or conv_tensor_core" > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log)
echo done
