#!/bin/bash
# 1-GPU: band kernel with the dynamic (guided) piece schedule — parity of the band backend, bench, SM balance from ncu.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "band or strips or graph or host or frame" 2>&1 | tail -6 > gpurun_out/pytest_band.log
: > gpurun_out/bench_band_dyn.jsonl
for cfg in "band 4" "band 3" "band 2" "band 1"; do set -- $cfg; python bench.py --backend $1 --tb $2 --steps 240 --warmup 24 --no-cpu-baseline --no-e2e >> gpurun_out/bench_band_dyn.jsonl 2>> gpurun_out/bench.err; done
python bench.py --size 1024 --backend band --tb 4 --steps 2000 --warmup 100 --no-cpu-baseline --no-e2e >> gpurun_out/bench_band_dyn.jsonl 2>> gpurun_out/bench.err
python bench.py --size 32768 --strong --backend band --tb 4 --steps 96 --warmup 12 --no-cpu-baseline --no-e2e >> gpurun_out/bench_band_dyn.jsonl 2>> gpurun_out/bench.err
ncu --metrics sm__cycles_active.avg,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_elapsed.max,smsp__inst_executed.sum,smsp__inst_executed.max,smsp__inst_executed.min,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:band_step -s 2 -c 1 --csv --log-file gpurun_out/band_dyn_balance.csv python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cat gpurun_out/pytest_band.log
python - <<'PY'
import json
for l in open('gpurun_out/bench_band_dyn.jsonl'):
    j=json.loads(l); print(j['config']['backend'], j['config']['temporal_block'], j['config']['grid'], round(j['value'],1), 'ms/step', round(j['ms_per_step'],4))
PY
grep -v "^==" gpurun_out/band_dyn_balance.csv | cut -d, -f13- | tail -12
tail -3 gpurun_out/bench.err
