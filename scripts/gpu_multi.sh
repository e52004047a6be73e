#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the tests that need >= 2 GPUs (one redux_ctx over all devices), then the
# bench line under torchrun exactly as the driver launches it.  Usage: scripts/gpu_multi.sh <tag> <N> [bench args]
tag=${1:-multi}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${tag}_gpus.txt 2>&1; nproc >> gpurun_out/${tag}_gpus.txt
python -m pytest tests -m gpu -x -q -k "multi_device or large_corpus or occupancy or two_ctas" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 3 --warmup 3 "$@" > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_bench_n$N.json").read().strip().splitlines() if l.startswith("{")][-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"), "ceiling", ((d.get("e2e") or {}).get("copy_ceiling") or {}).get("value"))
    print("strong", d.get("strong"))
    print("ctx", d.get("multi_device_ctx"))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench_n$N.err").read()[-3000:])
PY
python scripts/bench_small.py config5 > gpurun_out/${tag}_config5.log 2>&1; echo "config5 rc=$?"; tail -7 gpurun_out/${tag}_config5.log | cut -c1-330
