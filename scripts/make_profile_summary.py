"""Turn the .ncu-rep / launch-list files a gpurun visit left in gpurun_out/ into the tracked
summaries under profiles/ (text + traffic.json that bench.py reads for roofline.traffic).
usage: python scripts/make_profile_summary.py r01 prof_reg_k1:fused_k1 prof_reg_k2c:fused_k2 ...
"""
import csv, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
tag_round = sys.argv[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
traffic_path = ROOT / "profiles" / "traffic.json"
traffic = json.loads(traffic_path.read_text()) if traffic_path.exists() else {}
for spec in sys.argv[2:]:
    name, tag = spec.split(":")
    rep = ROOT / "gpurun_out" / f"{name}.ncu-rep"
    out = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none summary of gpurun_out/{name}.ncu-rep (8192x8192 reference scene, one launch)"]
    for r in rows[2:3]:
        kv = dict(zip(hdr, r))
        lines.append("kernel = " + kv.get("Kernel Name", "?"))
        for i, k in enumerate(hdr):
            if k in KEYS or ("issue_stalled" in k and "per_issue_active" in k):
                lines.append(f"{k} = {r[i]} {units[i]}")
        def gb(x, u):
            v = float(x)
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
        rd = gb(kv["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
        wr = gb(kv["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")])
        traffic[tag] = int(rd + wr)
        lines.append(f"dram_bytes_per_launch = {int(rd + wr)}  ({(rd + wr) / (8192 * 8192):.2f} B per cell per launch)")
    (ROOT / "profiles" / f"{tag_round}_{tag}_ncu_full.txt").write_text("\n".join(lines) + "\n")
    print("\n".join(lines[-3:]))
traffic_path.write_text(json.dumps(traffic, indent=1, sort_keys=True) + "\n")
