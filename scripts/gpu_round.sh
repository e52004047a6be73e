#!/bin/bash
# One GPU visit: parity tests, then bench lines per parameter triple (device-resident only), then the
# end-to-end probe with the host-side pipeline trace.  Usage: scripts/gpu_round.sh <tag>
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
for p in 8,14,16 8,22,24 8,30,32; do
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --params $p > gpurun_out/${tag}_bench_${p}.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${p}.log").read().strip().splitlines()[-1])
    print("$p", "enc", d["encode_MBps"], "dec", d["decode_MBps"], d["roofline"]["kernel_ms"])
except Exception as e:
    print("$p bench failed", e); print(open("gpurun_out/${tag}_bench_${p}.log").read()[-2000:])
PY
done
