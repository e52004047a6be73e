#!/bin/bash
# The two commands the driver runs at round end (N = 1), as it runs them.  Usage: scripts/gpu_bench_default.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
( time python bench.py --impl reference ) > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc=$?"
( time python bench.py ) > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench rc=$?"
tail -4 gpurun_out/${tag}_bench_reference.err gpurun_out/${tag}_bench_default.err
python - <<PY
import json
for f in ("gpurun_out/${tag}_bench_reference.json", "gpurun_out/${tag}_bench_default.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value", d["value"], "e2e", d.get("e2e"), "cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generator" 2>&1 | tail -2
