#!/bin/bash
# last 1-GPU check of the round: full parity suite, smoke(), default bench line.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/pytest_gpu.log; tail -1 gpurun_out/smoke.log; cut -c1-200 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
