#!/bin/bash
# Round-end measurements on one B200: smoke, parity tests, profile (launch list + full captures of the narrow and the
# wide class), the driver's two bench commands, the generic-path / small-batch / under-filled / code_bits > 32 tables.
# Usage: scripts/final_round.sh <tag>
tag=${1:-r02_final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > gpurun_out/${tag}_gpu.txt 2>&1; nproc >> gpurun_out/${tag}_gpu.txt; lscpu | grep -E "Model name|Socket|NUMA" >> gpurun_out/${tag}_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${tag}_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/${tag}_pytest.log
scripts/profile.sh ${tag} --no-classes > gpurun_out/${tag}_profile.log 2>&1; echo "profile rc=$?"
cmd="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-classes --params 8,30,32"
ncu --set full --clock-control none --import-source on -k regex:lane -s 6 -c 2 -f -o gpurun_out/${tag}_wide_prof $cmd > gpurun_out/${tag}_wide_ncu.log 2>&1; echo "wide ncu rc=$?"
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "ref rc=$?"
python bench.py > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/${tag}_bench_default.json
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-classes --blocks 16384 --params 8,30,34 > gpurun_out/${tag}_bench_huge.json 2>&1; echo "huge rc=$?"
python scripts/bench_generic.py > gpurun_out/${tag}_generic.log 2>&1; echo "generic rc=$?"; cp gpurun_out/bench_generic.json gpurun_out/${tag}_generic.json
python scripts/bench_underfilled.py > gpurun_out/${tag}_underfilled.log 2>&1; echo "underfilled rc=$?"; cp gpurun_out/bench_underfilled.json gpurun_out/${tag}_underfilled.json
python scripts/bench_small.py > gpurun_out/${tag}_small.log 2>&1; echo "small rc=$?"; cp gpurun_out/bench_small.json gpurun_out/${tag}_small.json
