#!/bin/bash
# 1-GPU: BASELINE config 4 grid (32768^2, 51.5 GB of state) on one B200 for every temporally blocked backend (T1 of the strong-scaling table).
mkdir -p gpurun_out
: > gpurun_out/bench_32768_1gpu.jsonl
for cfg in "band 4" "band 3" "stream 3" "tb 2"; do set -- $cfg; timeout 400 python bench.py --size 32768 --strong --backend $1 --tb $2 --steps 96 --warmup 12 --no-cpu-baseline --no-e2e >> gpurun_out/bench_32768_1gpu.jsonl 2>> gpurun_out/bench_32768.err; done
python - <<'PY'
import json
for l in open('gpurun_out/bench_32768_1gpu.jsonl'):
    j=json.loads(l); print(j['config']['backend'], j['config']['temporal_block'], j['config']['grid'], round(j['value'],1), 'ms/step', round(j['ms_per_step'],3))
PY
tail -3 gpurun_out/bench_32768.err
