#!/usr/bin/env bash
# Round-2 GPU run 1: tests with NUFFT as the default family, the worst-bin oracle sweep, a bench line, ncu captures
# of the NUFFT kernels.  Everything goes to gpurun_out/r2_run1/.
set -u
O=gpurun_out/r2_run1
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1

echo "=== 1. worst-bin sweep (full config 2) ==="
timeout 900 python tools/worst_bins.py > $O/worst_bins.log 2>&1; echo "rc=$?"
cp gpurun_out/worst_bins.json $O/ 2>/dev/null
tail -30 $O/worst_bins.log

echo "=== 2. GPU test suite ==="
timeout 1500 python -m pytest tests -m gpu -q -rxXs --deselect tests/test_gpu_fullsize.py::test_config2_worst_bins > $O/pytest_gpu.log 2>&1; echo "rc=$?"
tail -25 $O/pytest_gpu.log

echo "=== 3. bench (auto) ==="
timeout 600 python bench.py --steps 5 --warmup 3 --nufft-variants > $O/bench_auto.json 2> $O/bench_auto.err; echo "rc=$?"
tail -c 3000 $O/bench_auto.json

echo "=== 4. ragged probe (config-5 share x 1/4): nufft default vs direct ==="
timeout 300 python tools/probe_others.py 0.25 k1 2>&1 | tail -2
LKB_LS_RAGGED_NUFFT=0 timeout 300 python tools/probe_others.py 0.25 k1 2>&1 | tail -2

echo "=== 5. ncu: launch list + full captures of the NUFFT kernels ==="
B="python bench.py --steps 1 --warmup 3 --no-secondary --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_r02_bench_c2_nufft_v1.csv $B > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nufft_(spread|fft_pass|finish)_kernel" -s 6 -c 7 -o $O/r02_nufft_v1_global $B > $O/ncu_global.log 2>&1
LKB_NUFFT_FFT=smem timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nufft_fft_(cols|rows)_kernel" -c 2 -o $O/r02_nufft_v1_smem $B > $O/ncu_smem.log 2>&1
ls -la $O
echo "=== done ==="
