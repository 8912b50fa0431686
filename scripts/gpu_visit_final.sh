#!/bin/bash
# 1-GPU round: parity tests, default bench line (band k=4), reference arm, backend sweep, ncu launch list + full capture of the default kernel.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
: > gpurun_out/bench_sweep.jsonl
for cfg in "unfused 1" "fused 1" "tb 2" "tb 3" "stream 3" "stream 4" "band 1" "band 2" "band 3" "band 4"; do set -- $cfg; python bench.py --backend $1 --tb $2 --steps 240 --warmup 24 --no-cpu-baseline --no-e2e >> gpurun_out/bench_sweep.jsonl 2>> gpurun_out/bench.err; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 16 --warmup 8 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:band_step -s 2 -c 1 -f -o gpurun_out/prof_band_k4_final python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.csv
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-300; tail -3 gpurun_out/bench.err
