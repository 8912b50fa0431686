#!/bin/bash
# 1-GPU: late fetch of the next piece — band parity tests, bench with and without the edge-first work list, SM balance.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "band or strips or graph or host or frame" 2>&1 | tail -4 > gpurun_out/pytest_band.log
OUT=gpurun_out/fake_edge2.jsonl; : > $OUT
for f in 0 1; do
  export TWS_BAND_FAKE_EDGE=$f
  python bench.py --size 32768 --strong --steps 48 --warmup 8 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | grep '^{' | sed -e "s/^{/{\"fake\": \"$f\", /" >> $OUT
  python bench.py --steps 240 --warmup 24 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | grep '^{' | sed -e "s/^{/{\"fake\": \"$f\", /" >> $OUT
done
export TWS_BAND_FAKE_EDGE=0
for k in 3 2 1; do python bench.py --tb $k --steps 240 --warmup 24 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | grep '^{' | sed -e "s/^{/{\"fake\": \"0\", /" >> $OUT; done
ncu --metrics sm__cycles_active.avg,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_elapsed.max,smsp__inst_executed.sum,smsp__inst_executed.max,smsp__inst_executed.min,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:band_step -s 2 -c 1 --csv --log-file gpurun_out/band_dyn3_balance.csv python bench.py --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > /dev/null 2>&1
cat gpurun_out/pytest_band.log
python - <<'PY'
import json
for l in open('gpurun_out/fake_edge2.jsonl'):
    j=json.loads(l); print('fake', j['fake'], j['config']['grid'], 'k', j['config']['temporal_block'], j['steps'], round(j['value'],1), 'ms/step', round(j['ms_per_step'],4))
PY
grep -v "^==" gpurun_out/band_dyn3_balance.csv | awk -F'","' '{print $(NF-2), $NF}' | tail -10
