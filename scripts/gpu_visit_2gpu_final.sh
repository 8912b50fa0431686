#!/bin/bash
# 2-GPU check of the round's final state: strip tests (multi-GPU + IPC), default bench at N=2 (with the e2e leg), strips verified.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_strips.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest_strips_2gpu.log
OUT=gpurun_out/scale_2gpu_final.jsonl; : > $OUT
tr() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus $n "$@" 2>>gpurun_out/scale2.err | grep -E '^\{|STRIPS' >> $OUT; }
tr 2 --no-cpu-baseline
tr 2 --size 1536 --steps 32 --warmup 8 --no-cpu-baseline --no-e2e --verify-strips
cat gpurun_out/pytest_strips_2gpu.log
python - <<'PY'
import json
for l in open('gpurun_out/scale_2gpu_final.jsonl'):
    if not l.startswith('{'): print(l.strip()); continue
    j=json.loads(l); print(j['n_gpus'], j['scaling'], j['config']['grid'], j['steps'], round(j['value'],1), 'per-gpu', round(j['per_gpu_value'],1), 'launches', j['gpu_launches'], 'e2e', (j.get('e2e') or {}).get('value'))
PY
tail -2 gpurun_out/scale2.err
