"""Summarise an .ncu-rep: key raw metrics + per-instruction stall hot spots (needs -lineinfo)."""
import csv, subprocess, sys
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max']
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
for i, k in enumerate(hdr):
    if k in KEYS: print(f"{k} = {r[i]} {units[i]}")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = []; blocks.append(cur); continue
    if cur is not None: cur.append(r)
b = blocks[0]; hdr = b[0]; data = b[1:]
ix = {k: i for i, k in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
stalls = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k: sum(int(r[ix[k]]) for r in data) for k in stalls}
print('total samples', tot, 'warp-inst', sum(int(r[ix['Instructions Executed']]) for r in data))
for k, v in sorted(agg.items(), key=lambda x: -x[1]):
    if v: print(f"  {k:28s} {v:7d} {100*v/tot:5.1f}%")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:ntop]:
    st = {k[6:]: int(r[ix[k]]) for k in stalls if int(r[ix[k]]) > 0}
    print(r[ix['# Samples']].rjust(6), r[ix['Instructions Executed']].rjust(9), r[1].strip()[:64].ljust(64), dict(sorted(st.items(), key=lambda x: -x[1])[:3]))
