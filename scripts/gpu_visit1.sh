#!/bin/bash
# GPU visit: parity tests, bench line + reference arm, variant sweep, ncu launch list + full capture.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py --steps 400 --warmup 40 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
python scripts/perf_sweep.py 8192 > gpurun_out/sweep_base.log 2>&1
for v in build/variants/*.so; do echo "== $v" >> gpurun_out/sweep_variants.log; TWS_LIB=$v python scripts/perf_sweep.py 8192 >> gpurun_out/sweep_variants.log 2>&1; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_step -s 2 -c 2 -f -o gpurun_out/prof_fused python bench.py --steps 8 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.csv
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/sweep_base.log gpurun_out/sweep_variants.log; tail -3 gpurun_out/bench.err
