#!/bin/bash
# ncu --set full captures (one launch each): tile kernel K=2, K=3, streaming kernel K=4.
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -f -o gpurun_out/prof_tile_k2 python scripts/stream_check.py 3 2 --no-parity > gpurun_out/ncu_tile_k2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_step -s 3 -c 1 -f -o gpurun_out/prof_tile_k3 python scripts/stream_check.py 3 3 --no-parity > gpurun_out/ncu_tile_k3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stream_step -s 3 -c 1 -f -o gpurun_out/prof_stream_k4 python scripts/stream_check.py 4 4 --no-parity > gpurun_out/ncu_stream_k4.log 2>&1
tail -2 gpurun_out/ncu_*.log
