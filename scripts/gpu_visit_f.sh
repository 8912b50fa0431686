#!/bin/bash
# 1-GPU: full parity suite with the dynamic band schedule (no extra band, whole-band segments), mip chain / GL tests, bench + balance.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
: > gpurun_out/bench_band_dyn2.jsonl
for cfg in "band 4" "band 3" "band 2"; do set -- $cfg; python bench.py --backend $1 --tb $2 --steps 240 --warmup 24 --no-cpu-baseline --no-e2e >> gpurun_out/bench_band_dyn2.jsonl 2>> gpurun_out/bench.err; done
python bench.py --size 32768 --strong --backend band --tb 4 --steps 96 --warmup 12 --no-cpu-baseline --no-e2e >> gpurun_out/bench_band_dyn2.jsonl 2>> gpurun_out/bench.err
ncu --metrics sm__cycles_active.avg,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_elapsed.max,smsp__inst_executed.sum,smsp__inst_executed.max,smsp__inst_executed.min,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:band_step -s 2 -c 1 --csv --log-file gpurun_out/band_dyn2_balance.csv python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cat gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_band_dyn2.jsonl'):
    j=json.loads(l); print(j['config']['backend'], j['config']['temporal_block'], j['config']['grid'], round(j['value'],1), 'ms/step', round(j['ms_per_step'],4))
PY
grep -v "^==" gpurun_out/band_dyn2_balance.csv | cut -d, -f13- | tail -10
tail -3 gpurun_out/bench.err
