#!/usr/bin/env bash
# Round-2 GPU run 20: config-5 share worst-bin sweep (new test), NUFFT hardware tests without their round-1 xfail marks
set -u
O=gpurun_out/r2_run20
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_zz_nufft.py -m gpu -q -rxXs -s -k "config5 or nufft" > $O/pytest.log 2>&1; echo "rc=$?"
grep -E "worst-bin excess|passed|failed|Error" $O/pytest.log | head -10
echo "=== done ==="
