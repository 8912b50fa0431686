#!/bin/bash
# 2-GPU: band strips with the exchange fused into the step launch (one launch per block) — strip tests, N=1/2 weak, strong, verification.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_strips.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_strips_2gpu.log
OUT=gpurun_out/scale_2gpu.jsonl; : > $OUT
tr() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus $n "$@" 2>>gpurun_out/scale2.err | grep -E '^\{|STRIPS' >> $OUT; }
timeout 300 python bench.py --gpus 1 --steps 400 --warmup 40 --no-cpu-baseline --no-e2e 2>>gpurun_out/scale2.err | grep '^{' >> $OUT
tr 2 --steps 400 --warmup 40 --no-cpu-baseline --no-e2e
tr 2 --size 2048 --steps 40 --warmup 8 --no-cpu-baseline --no-e2e --verify-strips
tr 2 --size 640 --tb 3 --steps 30 --warmup 6 --no-cpu-baseline --no-e2e --verify-strips
tr 2 --size 32768 --strong --steps 96 --warmup 12 --no-cpu-baseline --no-e2e
tr 2 --steps 40 --warmup 8 --no-cpu-baseline
cat gpurun_out/pytest_strips_2gpu.log
python - <<'PY'
import json
for l in open('gpurun_out/scale_2gpu.jsonl'):
    if not l.startswith('{'): print(l.strip()); continue
    j=json.loads(l); print(j['n_gpus'], j['scaling'], j['config']['grid'], j['config']['backend'], j['config']['temporal_block'], round(j['value'],1), 'per-gpu', round(j['per_gpu_value'],1), 'ms/step', round(j['ms_per_step'],4), 'launches', j['gpu_launches'], 'e2e', (j.get('e2e') or {}).get('value'))
PY
tail -5 gpurun_out/scale2.err
