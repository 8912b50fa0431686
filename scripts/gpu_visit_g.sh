#!/bin/bash
# 1-GPU timing experiment: does putting 2 x nstrips tiny pieces first in the work list slow a 32768^2 launch down?
mkdir -p gpurun_out
OUT=gpurun_out/fake_edge.jsonl; : > $OUT
for f in 0 1 3; do
  export TWS_BAND_FAKE_EDGE=$f
  python bench.py --size 32768 --strong --steps 48 --warmup 8 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | grep '^{' | sed -e "s/^{/{\"fake\": \"$f\", /" >> $OUT
  python bench.py --steps 240 --warmup 24 --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | grep '^{' | sed -e "s/^{/{\"fake\": \"$f\", /" >> $OUT
done
export TWS_BAND_FAKE_EDGE=1
ncu --metrics sm__cycles_active.avg,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_elapsed.max,smsp__inst_executed.sum,smsp__inst_executed.max,smsp__inst_executed.min,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:band_step -s 2 -c 1 --csv --log-file gpurun_out/fake_edge_ncu.csv python bench.py --size 32768 --strong --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > /dev/null 2>&1
export TWS_BAND_FAKE_EDGE=0
ncu --metrics sm__cycles_active.avg,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_elapsed.max,smsp__inst_executed.sum,smsp__inst_executed.max,smsp__inst_executed.min,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:band_step -s 2 -c 1 --csv --log-file gpurun_out/nofake_edge_ncu.csv python bench.py --size 32768 --strong --steps 16 --warmup 4 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/fake_edge.jsonl'):
    j=json.loads(l); print('fake', j['fake'], j['config']['grid'], j['steps'], round(j['value'],1), 'ms/step', round(j['ms_per_step'],4))
PY
for f in fake_edge_ncu nofake_edge_ncu; do echo $f; grep -v "^==" gpurun_out/$f.csv | awk -F'","' '{print $(NF-2), $NF}' | tail -11; done
