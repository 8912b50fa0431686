#!/bin/bash
# GPU visit: parity suite, packed-vs-scalar A/B, bench line with the host-step e2e leg, footprint probe.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
timeout 300 python scripts/perf_sweep.py 8192 > gpurun_out/sweep_packed.log 2>&1
TWS_LIB=build/variants/scalar.so timeout 300 python scripts/perf_sweep.py 8192 > gpurun_out/sweep_scalar.log 2>&1
timeout 600 python bench.py --steps 400 --warmup 40 > gpurun_out/bench.json 2> gpurun_out/bench.err
# footprint probe: which dimension slows the 32768^2 single-GPU run (168 vs 217 Gcell/s)?
for wh in "8192 32768" "32768 8192" "16384 16384" "32768 16384" "32768 32768"; do set -- $wh; timeout 300 python scripts/perf_sweep.py $1 $2 3:2,4:4 >> gpurun_out/footprint.log 2>&1; done
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep_packed.log gpurun_out/sweep_scalar.log gpurun_out/footprint.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
