#!/bin/bash
# 1-GPU: band parity tests + bench after restoring the early fetch (edge pieces last on strips).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "band or strips or graph or host or frame" 2>&1 | tail -4 > gpurun_out/pytest_band.log
OUT=gpurun_out/bench_band_dyn4.jsonl; : > $OUT
python bench.py --size 32768 --strong --steps 48 --warmup 8 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | grep '^{' >> $OUT
for k in 4 3 2 1; do python bench.py --tb $k --steps 240 --warmup 24 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | grep '^{' >> $OUT; done
cat gpurun_out/pytest_band.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_band_dyn4.jsonl'):
    j=json.loads(l); print(j['config']['grid'], 'k', j['config']['temporal_block'], j['steps'], round(j['value'],1), 'ms/step', round(j['ms_per_step'],4))
PY
