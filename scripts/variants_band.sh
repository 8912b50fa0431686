#!/bin/bash
# parity + perf of the band backend (k given, default 4) on the default lib and every build/variants/*.so
mkdir -p gpurun_out; : > gpurun_out/variants.log
echo "== default" >> gpurun_out/variants.log
timeout 300 python scripts/stream_check.py 5 ${1:-4} >> gpurun_out/variants.log 2>&1
for v in build/variants/*.so; do echo "== $v" >> gpurun_out/variants.log; TWS_LIB=$v timeout 300 python scripts/stream_check.py 5 ${1:-4} >> gpurun_out/variants.log 2>&1; done
grep -E "==|perf|PARITY|ERR|rror" gpurun_out/variants.log
