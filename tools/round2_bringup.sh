#!/usr/bin/env bash
# Round-2 bring-up of the NUFFT Lomb-Scargle path.  Run ON THE GPU BOX, e.g.
#   gpurun --timeout 900 -- 'bash tools/round2_bringup.sh > gpurun_out/round2_bringup.log 2>&1'
# Everything it writes goes under gpurun_out/.  Steps are ordered cheapest / most informative first; each one is
# independent, so a failure does not hide the later ones.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1

echo "=== 1. stage-by-stage check of the shared-grid path against the CPU harness ==="
timeout 300 python tools/nufft_gpu_check.py

echo "=== 2. the two xfail tests (shared grid incl. the chunk-pipelined host mode; ragged batches) ==="
timeout 600 python -m pytest tests/test_gpu_zz_nufft.py -m gpu -q -rxX -s 2>&1 | tail -15

echo "=== 3. headline bench: tensor path (today's default) vs NUFFT path and its switches ==="
for cfg in "auto" "nufft" "nufft LKB_NUFFT_TWIDDLE_CHAIN=1" "nufft LKB_NUFFT_TWIDDLE_CHAIN=1 LKB_NUFFT_GROUP_MB=96" \
           "nufft LKB_NUFFT_TWIDDLE_CHAIN=1 LKB_NUFFT_W=6" "nufft LKB_NUFFT_FFT=smem" \
           "nufft LKB_NUFFT_FFT=smem LKB_NUFFT_TWIDDLE_CHAIN=1" "nufft LKB_NUFFT_FFT=smem LKB_NUFFT_TWIDDLE_CHAIN=1 LKB_NUFFT_TILE=8192" \
           "nufft LKB_NUFFT_FFT=fused LKB_NUFFT_TWIDDLE_CHAIN=1"; do
  set -- $cfg
  algo=$1; shift
  tag=$(echo "$cfg" | tr ' =' '__')
  echo "--- bench --algo $algo  env: $*"
  env "$@" timeout 400 python bench.py --steps 5 --warmup 3 --algo "$algo" --no-secondary --no-cpu-baseline \
      > "gpurun_out/bench_${tag}.json" 2> "gpurun_out/bench_${tag}.err" || echo "bench failed (see gpurun_out/bench_${tag}.err)"
  python - "gpurun_out/bench_${tag}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.3e %s  ms/step %.2f  e2e %.3e  roofline %s %.1f %s (frac %.2f)  kernel_ms %.2f" % (
        d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["bound"], d["roofline"]["achieved"],
        d["roofline"]["unit"], d["roofline"]["frac"], d["roofline"]["kernel_ms"]))
except Exception as e:
    print("no bench line:", e)
PY
done

echo "=== 3b. ragged batches (config-5 share): direct kernel vs NUFFT path ==="
timeout 400 python tools/probe_others.py 0.125 k1
LKB_LS_RAGGED_NUFFT=1 timeout 400 python tools/probe_others.py 0.125 k1
LKB_LS_RAGGED_NUFFT=1 LKB_NUFFT_FFT=smem LKB_NUFFT_TWIDDLE_CHAIN=1 timeout 400 python tools/probe_others.py 0.125 k1

echo "=== 4. launch list and one full ncu capture of the NUFFT passes (only if step 3 produced numbers) ==="
if command -v ncu > /dev/null; then
  LKB_NUFFT_TWIDDLE_CHAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches_r02_bench_c2_nufft.csv python bench.py --steps 2 --warmup 3 --algo nufft \
      --no-secondary --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
  LKB_NUFFT_TWIDDLE_CHAIN=1 timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:nufft_fft_pass_kernel -c 2 -o gpurun_out/r02_nufft_fft_pass python bench.py --steps 1 --warmup 3 \
      --algo nufft --no-secondary --no-cpu-baseline > gpurun_out/ncu_pass.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:nufft_spread_kernel -c 1 \
      -o gpurun_out/r02_nufft_spread python bench.py --steps 1 --warmup 3 --algo nufft --no-secondary \
      --no-cpu-baseline > gpurun_out/ncu_spread.log 2>&1
fi
echo "=== done ==="
