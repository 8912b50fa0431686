#!/bin/bash
# 8-GPU visit (band kernel k=4, fused exchange): weak scaling 8192^2 per GPU (tiled scene), strong scaling 32768^2 (BASELINE config 4).
mkdir -p gpurun_out
OUT=gpurun_out/scale_8gpu.jsonl; : > $OUT
run() { n=$1; shift
  if [ "$n" = "1" ]; then timeout 300 python bench.py --gpus 1 "$@" 2>>gpurun_out/scale.err | grep -E '^\{|STRIPS' >> $OUT
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n+RANDOM%100)) bench.py --gpus $n "$@" 2>>gpurun_out/scale.err | grep -E '^\{|STRIPS' >> $OUT; fi
}
run 2 --size 1024 --steps 24 --warmup 8 --no-cpu-baseline --no-e2e --verify-strips
if ! grep -q STRIPS_VERIFIED $OUT; then echo "strip verification failed"; tail -20 gpurun_out/scale.err; exit 1; fi
run 8 --size 1024 --steps 24 --warmup 8 --no-cpu-baseline --no-e2e --verify-strips
for n in 1 2 4 8; do run $n --steps 400 --warmup 40 --no-cpu-baseline --no-e2e; done
for n in 2 4 8; do run $n --size 32768 --strong --steps 96 --warmup 12 --no-cpu-baseline --no-e2e; done
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/scale_8gpu.jsonl'):
    if not l.startswith('{'): print(l.strip()); continue
    j=json.loads(l); print(j['n_gpus'], j['scaling'], j['config']['grid'], j['config']['backend'], j['config']['temporal_block'], round(j['value'],1), 'per-gpu', round(j['per_gpu_value'],1), 'ms/step', round(j['ms_per_step'],4), 'launches', j['gpu_launches'])
PY
tail -4 gpurun_out/scale.err
