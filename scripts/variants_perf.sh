#!/bin/bash
# perf only (no parity) of stream_check on the default lib and every build/variants/*.so: variants_perf.sh <backend> <ks>
mkdir -p gpurun_out; : > gpurun_out/variants.log
echo "== default" >> gpurun_out/variants.log
timeout 300 python scripts/stream_check.py $1 $2 --no-parity >> gpurun_out/variants.log 2>&1
for v in build/variants/*.so; do echo "== $v" >> gpurun_out/variants.log; TWS_LIB=$v timeout 300 python scripts/stream_check.py $1 $2 --no-parity >> gpurun_out/variants.log 2>&1; done
grep -E "==|perf|ERR|rror" gpurun_out/variants.log
