#!/bin/bash
# Round-end measurements on one B200: smoke, parity tests, profile (launch list + full capture), default
# bench line, reference arm, the small-batch and generic-path tables.  Usage: scripts/final_round.sh <tag>
tag=${1:-r01_final2}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/${tag}_pytest.log
scripts/profile.sh ${tag} > gpurun_out/${tag}_profile.log 2>&1; echo "profile rc=$?"
python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${tag}_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>&1; echo "ref rc=$?"
for p in 8,22,24 8,30,32; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --params $p > gpurun_out/${tag}_bench_${p}.json 2>&1; done
python scripts/bench_generic.py > gpurun_out/${tag}_generic.log 2>&1; echo "generic rc=$?"; cat gpurun_out/${tag}_generic.log
python scripts/bench_small.py > gpurun_out/${tag}_small.log 2>&1; echo "small rc=$?"
