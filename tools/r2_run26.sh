#!/usr/bin/env bash
# Round-2 GPU run 26: the round's last check - full GPU suite and smoke() on the final tree
set -u
O=gpurun_out/r2_run26
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "rc=$?"
tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== done ==="
