#!/bin/bash
# 2-GPU: how many SMs the strips leave to the edge stream (0 = edge launches queue behind a full-width interior launch).
mkdir -p gpurun_out
OUT=gpurun_out/edge_sms_2gpu.jsonl; : > $OUT
tr() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+RANDOM%200)) bench.py --gpus $n "$@" 2>>gpurun_out/scale2.err | grep -E '^\{' | sed -e "s/^{/{\"edge_sms\": \"$TWS_EDGE_SMS\", /" >> $OUT; }
for e in 0 2 6; do
  export TWS_EDGE_SMS=$e
  tr 2 --steps 400 --warmup 40 --no-cpu-baseline --no-e2e
  tr 2 --size 32768 --strong --steps 96 --warmup 12 --no-cpu-baseline --no-e2e
done
export TWS_EDGE_SMS=2
tr 2 --backend band --tb 2 --steps 400 --warmup 40 --no-cpu-baseline --no-e2e
tr 2 --backend tb --tb 2 --steps 400 --warmup 40 --no-cpu-baseline --no-e2e
python - <<'PY'
import json
for l in open('gpurun_out/edge_sms_2gpu.jsonl'):
    j=json.loads(l); print('edge_sms', j['edge_sms'], j['n_gpus'], j['scaling'], j['config']['grid'], j['config']['backend'], j['config']['temporal_block'], round(j['value'],1), 'per-gpu', round(j['per_gpu_value'],1), 'ms/step', round(j['ms_per_step'],4))
PY
tail -3 gpurun_out/scale2.err
