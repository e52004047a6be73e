#!/bin/bash
# A/B of library builds on one GPU visit: for every variants/<name>.so (built here with different -D switches) copy it
# over redux_b200/libredux_b200.so, run a parity subset and the device-resident bench, print the kernel times.
# Usage: scripts/ab.sh <tag> <name> [<name> ...]
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  cp variants/$v.so redux_b200/libredux_b200.so
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or full_size or truncated or kat or million" > gpurun_out/${tag}_${v}_pytest.log 2>&1; echo "$v pytest rc=$? $(tail -1 gpurun_out/${tag}_${v}_pytest.log)"
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_${v}_bench.json 2> gpurun_out/${tag}_${v}_bench.err; echo "$v bench rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_${v}_bench.json").read().strip().splitlines()[-1])
    print("$v 8,14,16", d["roofline"]["kernel_ms"], "value", d["value"])
    for k, x in (d.get("classes") or {}).items(): print("$v", k, x["encode_ms"], x["decode_ms"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_${v}_bench.err").read()[-2000:])
PY
done
