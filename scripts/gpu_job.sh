#!/bin/bash
# gpu_job.sh <job> [args...] — the ONE entry point for everything this repo runs on a GPU box through
#   /usr/local/graft/bin/gpurun --timeout T -- 'bash scripts/gpu_job.sh <job> ...'
# Every job writes under gpurun_out/ (merged back by gpurun).  Jobs:
#   tests                      pytest -m gpu (full), smoke
#   band  [k-list]             parity (small grids vs oracle) + 8192^2 perf of the band kernel: default lib and every build/variants/*.so
#   bench [bench.py args]      python bench.py ... (JSON lines appended to gpurun_out/bench.jsonl)
#   benchN <N> [args]          torchrun --nproc-per-node N bench.py --gpus N ...
#   launches [bench.py args]   ncu launch list (gpu__time_duration.sum) of a short bench command
#   ncu <regex> <backend> <k> <tag>    ncu --set full of one launch (scripts/stream_check.py as the driver)
#   py <script.py> [args]      run a Python script from scripts/
#   refframe [frames]          the reference's frame timed from C++ through include/tws_terrain.hpp (scripts/refframe.cpp)
set -u
mkdir -p gpurun_out
job=${1:-tests}; shift || true
case "$job" in
  tests)
    timeout 1500 python -m pytest tests -x -q -m gpu "$@" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
    timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log ;;
  band)
    : > gpurun_out/band.log
    echo "== default" >> gpurun_out/band.log
    timeout 600 python scripts/stream_check.py 5 ${1:-4} >> gpurun_out/band.log 2>&1
    for v in build/variants/*.so; do
      [ -e "$v" ] || continue
      echo "== $v" >> gpurun_out/band.log
      TWS_LIB=$v timeout 600 python scripts/stream_check.py 5 ${1:-4} >> gpurun_out/band.log 2>&1
    done
    grep -E "==|perf|PARITY|ERR|rror" gpurun_out/band.log ;;
  bench)
    timeout 1500 python bench.py "$@" 2>>gpurun_out/bench.err | grep '^{' | tee -a gpurun_out/bench.jsonl ;;
  benchN)
    n=$1; shift
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n "$@" 2>>gpurun_out/bench.err | grep -E '^\{|STRIPS' | tee -a gpurun_out/bench.jsonl ;;
  launches)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py "$@" > gpurun_out/launches_bench.log 2>&1
    tail -n 3 gpurun_out/launches_bench.log ;;
  ncu)
    ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -f -o gpurun_out/prof_$4 python scripts/stream_check.py $2 $3 --no-parity > gpurun_out/ncu_$4.log 2>&1
    tail -n 2 gpurun_out/ncu_$4.log ;;
  refframe)
    mkdir -p build
    g++ -O2 -std=c++17 -Iinclude scripts/refframe.cpp -o build/refframe -Lterrainwatersim_b200 -ltws -Wl,-rpath,$PWD/terrainwatersim_b200 \
      && ./build/refframe ${1:-3000} | tee gpurun_out/refframe.log ;;
  py)
    s=$1; shift
    timeout 1500 python scripts/$s "$@" 2>&1 | tee gpurun_out/${s%.py}.log | tail -40 ;;
  *) echo "unknown job $job"; exit 2 ;;
esac
