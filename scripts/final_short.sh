#!/bin/bash
# Short form of final_round.sh after a change to the narrow / wide lane kernels only: smoke, parity tests, launch
# list + full ncu captures, the driver's two bench commands.  Usage: scripts/final_short.sh <tag>
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
