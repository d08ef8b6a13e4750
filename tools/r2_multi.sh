#!/usr/bin/env bash
# Multi-GPU run: N = $1 ranks.  2-rank NCCL tests (N = 2), c2 weak scaling with pipelined all-gathers, c5 strong scaling.
set -u
N=${1:-2}
MODE=${2:-full}          # full: + monolithic all-gather variants, reference arm, all legs; trim: the two scaling lines only
O=gpurun_out/r2_multi_n$N
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv > $O/smi.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "2" ]; then
  echo "=== 2-rank NCCL tests ==="
  timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_zz_nccl_abi.py -m gpu -q -rxXs > $O/pytest_2rank.log 2>&1; echo "rc=$?"
  tail -8 $O/pytest_2rank.log
fi
echo "=== c2, N = $N: weak scaling, pipelined all-gathers ==="
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > $O/bench_c2_n$N.json 2> $O/bench_c2_n$N.err; echo "rc=$?"
tail -c 1800 $O/bench_c2_n$N.json; tail -3 $O/bench_c2_n$N.err
if [ "$MODE" = "pieces" ]; then
for c in 1 2; do
echo "=== c2, N = $N: --chunks $c ==="
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --chunks $c > $O/bench_c2_n${N}_chunks$c.json 2> $O/bench_c2_n${N}_chunks$c.err; echo "rc=$?"
python - $O/bench_c2_n${N}_chunks$c.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print("ms/step %.3f value %.4g e2e ms %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
except Exception as e:
    print("no line", e)
PY
done
fi
if [ "$MODE" = "full" ]; then
echo "=== c2, N = $N: one monolithic all-gather after the kernels (--chunks 1) ==="
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --chunks 1 > $O/bench_c2_n${N}_mono.json 2> $O/bench_c2_n${N}_mono.err; echo "rc=$?"
fi
python - $O/bench_c2_n$N.json $O/bench_c2_n${N}_mono.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms/step %.3f value %.4g e2e ms %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "no line", e)
PY
echo "=== c5, N = $N: strong scaling of 16384 ragged light curves ==="
timeout 1200 $TR bench.py --gpus $N --workload c5 --steps 3 --warmup 2 > $O/bench_c5_n$N.json 2> $O/bench_c5_n$N.err; echo "rc=$?"
tail -c 1500 $O/bench_c5_n$N.json; tail -3 $O/bench_c5_n$N.err
if [ "$MODE" = "full" ]; then
timeout 1200 $TR bench.py --gpus $N --workload c5 --steps 3 --warmup 2 --chunks 1 --no-cpu-baseline > $O/bench_c5_n${N}_mono.json 2> $O/bench_c5_n${N}_mono.err
fi
python - $O/bench_c5_n$N.json $O/bench_c5_n${N}_mono.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "ms/step %.3f value %.4g e2e ms %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "no line", e)
PY
if [ "$MODE" = "full" ]; then
echo "=== full default bench at N = $N (all legs) ==="
timeout 1500 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_full_n$N.json 2> $O/bench_full_n$N.err; echo "rc=$?"
tail -c 600 $O/bench_full_n$N.json; tail -3 $O/bench_full_n$N.err
echo "=== reference arm under torchrun (rank 0 works, the others exit 0) ==="
timeout 900 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref_n$N.json 2> $O/bench_ref_n$N.err; echo "rc=$?"
tail -c 700 $O/bench_ref_n$N.json
fi
echo "=== done ==="
