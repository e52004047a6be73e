#!/bin/bash
run() { echo "== $*"; env "$@" python scripts/e2e_probe.py 2>&1 | grep -E "encode_batch|decode_batch"; }
run A=1
run REDUX_PIPE_RAMP=0
run REDUX_PIPE_RAMP=0 REDUX_PIPE_CHUNKS=16
run REDUX_PIPE_CHUNKS=16
run REDUX_PIPE_CHUNKS=28
run CUDA_DEVICE_MAX_CONNECTIONS=8
REDUX_TRACE=1 python scripts/e2e_probe.py 2>&1 | grep trace | tail -52 | awk '{print $3, $5, $6, $7, $8, $9}'
