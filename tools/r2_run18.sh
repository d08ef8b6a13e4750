#!/usr/bin/env bash
# Round-2 GPU run 18 (2 GPUs): ramped chunk sizes of the host-buffer path, 2-rank tests, N = 2 lines after the last changes
set -u
O=gpurun_out/r2_run18
mkdir -p $O
export PYTHONUNBUFFERED=1
echo "=== 1. e2e with / without the chunk ramp (1 GPU) ==="
for v in "" "LKB_LS_NO_RAMP=1"; do
env $v timeout 600 python bench.py --no-secondary --no-cpu-baseline > $O/bench_e2e.json 2> $O/bench_e2e.err
python - $O/bench_e2e.json "$v" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("[%s] ms/step %.3f e2e ms %.3f (%.4g)" % (sys.argv[2], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("no bench line:", e)
PY
done
echo "=== 2. LS GPU tests (1 GPU view) + 2-rank NCCL tests ==="
timeout 1500 python -m pytest tests -m gpu -q -rxXs -k "ls or nufft or lomb or config2 or shared or ragged or periodogram or dist or nccl" > $O/pytest_ls.log 2>&1; echo "rc=$?"
tail -4 $O/pytest_ls.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "=== 3. N = 2: default line, c5 ==="
timeout 900 $TR bench.py --gpus 2 --no-secondary --no-cpu-baseline > $O/bench_c2_n2.json 2> $O/bench_c2_n2.err; echo "rc=$?"
timeout 1200 $TR bench.py --gpus 2 --workload c5 --steps 3 --warmup 2 --no-cpu-baseline > $O/bench_c5_n2.json 2> $O/bench_c5_n2.err; echo "rc=$?"
python - $O/bench_c2_n2.json $O/bench_c5_n2.json <<'PY'
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms/step %.3f value %.4g e2e ms %.3f" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
    except Exception as e:
        print(f, "no line", e)
PY
echo "=== done ==="
