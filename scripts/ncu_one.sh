#!/bin/bash
# ncu --set full capture of one launch: ncu_one.sh <kernel-regex> <backend> <k> <tag>
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -f -o gpurun_out/prof_$4 python scripts/stream_check.py $2 $3 --no-parity > gpurun_out/ncu_$4.log 2>&1
tail -n 2 gpurun_out/ncu_$4.log
