#!/bin/bash
# Cross-compiles the issue/latency microbenchmark for sm_100a (no GPU needed); the binary travels with gpurun.
cd "$(dirname "$0")" && /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o issue_peak.bin issue_peak.cu
