#!/bin/bash
# One GPU visit: the whole parity suite, the device-resident bench line (all three parameter classes), the
# generic-path table.  Usage: scripts/gpu_round.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
    print("8,14,16", d["roofline"]["kernel_ms"], "value", d["value"])
    for k, v in (d.get("classes") or {}).items(): print(k, v["encode_ms"], v["decode_ms"])
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/${tag}_bench.err").read()[-3000:])
PY
python scripts/bench_generic.py > gpurun_out/${tag}_generic.log 2>&1; echo "generic rc=$?"; cat gpurun_out/${tag}_generic.log
