#!/bin/bash
# Validation visit: parity tests, default bench line, stream/tb sweep, launch list.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
python bench.py --steps 400 --warmup 40 > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
: > gpurun_out/bench_sweep.jsonl
for cfg in "unfused 1" "fused 1" "tb 2" "tb 3" "stream 2" "stream 3" "stream 4"; do set -- $cfg; python bench.py --backend $1 --tb $2 --steps 240 --warmup 24 --no-cpu-baseline --no-e2e >> gpurun_out/bench_sweep.jsonl 2>> gpurun_out/bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.csv
nproc > gpurun_out/nproc.txt
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
