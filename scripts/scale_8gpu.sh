#!/bin/bash
# 8-GPU visit: weak scaling at 8192^2 per GPU, strong scaling at 32768^2 (BASELINE config 4), config 5 ledger.
mkdir -p gpurun_out
OUT=gpurun_out/scale_8gpu.jsonl; : > $OUT
run() { # nproc, extra args...
  n=$1; shift
  if [ "$n" = "1" ]; then timeout 300 python bench.py --gpus 1 "$@" 2>>gpurun_out/scale.err | grep '^{' >> $OUT
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n "$@" 2>>gpurun_out/scale.err | grep '^{' >> $OUT; fi
}
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for n in 1 2 4 8; do run $n --steps 400 --warmup 40 --no-cpu-baseline --no-e2e; done
for n in 1 2 4 8; do run $n --size 32768 --strong --steps 100 --warmup 10 --no-cpu-baseline --no-e2e; done
run 8 --steps 200 --warmup 20 --no-cpu-baseline        # with the e2e leg at 8 GPUs
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 scripts/config5_ledger.py --size 65536 --steps 200 --fast-steps 2000 2>>gpurun_out/scale.err | grep CONFIG5 > gpurun_out/config5.json
python - <<'PY'
import json
for l in open('gpurun_out/scale_8gpu.jsonl'):
    j=json.loads(l); print(j['n_gpus'], j['scaling'], j['config']['grid'], round(j['value'],1), 'per-gpu', round(j['per_gpu_value'],1), 'ms/step', round(j['ms_per_step'],4), 'e2e', (j['e2e'] or {}).get('value'))
PY
cat gpurun_out/config5.json; tail -5 gpurun_out/scale.err
