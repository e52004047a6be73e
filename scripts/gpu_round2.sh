#!/bin/bash
# One GPU visit while tuning: host-emu-independent parity subset, bench lines (device-resident only) incl. the
# other parameter classes, then ONE ncu full capture of the (8,30,32) coders.  Usage: scripts/gpu_round2.sh <tag> [ncu]
tag=${1:-run}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ragged or full_size or truncated or kat or pretrained or million" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${tag}_pytest.log
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
if [ "$2" = "ncu" ]; then
  cmd="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-classes --params 8,30,32"
  ncu --set full --clock-control none --import-source on -k regex:lane -s 6 -c 2 -f -o gpurun_out/${tag}_wide_prof $cmd > gpurun_out/${tag}_wide_ncu.log 2>&1
  ls -la gpurun_out/${tag}_wide_prof.ncu-rep
fi
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-classes --blocks 16384 --params 8,30,34 > gpurun_out/${tag}_bench_huge.json 2>&1; echo "huge rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/${tag}_bench_huge.json') if l.startswith('{')][-1]); print('8,30,34 x16384', d['roofline']['kernel_ms'], d['value'])"
python scripts/bench_underfilled.py > gpurun_out/${tag}_underfilled.log 2>&1; echo "underfilled rc=$?"; cat gpurun_out/${tag}_underfilled.log
