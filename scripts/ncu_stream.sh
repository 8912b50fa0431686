#!/bin/bash
# ncu full capture of the streaming kernel at given K (default 3); optional TWS_LIB via env
K=${1:-3}; TAG=${2:-stream}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:stream_step -s 3 -c 1 -f -o gpurun_out/prof_${TAG}_k$K python scripts/stream_check.py 4 $K --no-parity > gpurun_out/ncu_${TAG}_k$K.log 2>&1
tail -2 gpurun_out/ncu_${TAG}_k$K.log
