#!/bin/bash
# GPU side of the profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then ONE
# full capture of the timed encode + decode launches.  Usage: scripts/profile.sh <tag> [bench args]
tag=${1:-prof}; shift
cmd="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline $*"
mkdir -p gpurun_out
$cmd > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lane -s 6 -c 2 -f -o gpurun_out/${tag}_prof $cmd > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/${tag}_*
