#!/bin/bash
# 2-GPU timing experiments: which part of the fused edge exchange costs time at 32768^2 (results are wrong by design when parts are disabled).
mkdir -p gpurun_out
OUT=gpurun_out/edge_debug_2gpu.jsonl; : > $OUT
tr() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus $n "$@" 2>>gpurun_out/scale2.err | grep -E '^\{' | sed -e "s/^{/{\"dbg\": \"$TWS_EDGE_DEBUG\", /" >> $OUT; }
for d in 0 7 5 3 6; do
  export TWS_EDGE_DEBUG=$d
  tr 2 --size 32768 --strong --steps 48 --warmup 8 --no-cpu-baseline --no-e2e
done
export TWS_EDGE_DEBUG=0
tr 2 --steps 40 --warmup 8 --no-cpu-baseline --no-e2e
tr 2 --steps 1200 --warmup 40 --no-cpu-baseline --no-e2e
python - <<'PY'
import json
for l in open('gpurun_out/edge_debug_2gpu.jsonl'):
    j=json.loads(l); print('dbg', j['dbg'], j['n_gpus'], j['scaling'], j['config']['grid'], j['steps'], round(j['value'],1), 'per-gpu', round(j['per_gpu_value'],1), 'ms/step', round(j['ms_per_step'],4))
PY
tail -3 gpurun_out/scale2.err
