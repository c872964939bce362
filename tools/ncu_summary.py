#!/usr/bin/env python3
"""Key metrics of every launch in an .ncu-rep (read here, on the CPU box):  python tools/ncu_summary.py gpurun_out/x.ncu-rep [elements_per_launch]
Prints time, instructions (per element if given), pipe utilisation, issue utilisation, DRAM traffic, stall reasons per issued instruction."""
import csv, subprocess, sys

rep = sys.argv[1]
elems = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
KEYS = [("gpu__time_duration.sum", "time"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("smsp__inst_executed.sum", "warp-inst"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu pipe %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
        ("sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "fmaheavy pipe %"),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
        ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe %"),
        ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput %"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts")]
for r in data:
    print("==", r[col["Kernel Name"]] if "Kernel Name" in col else "?")
    for k, name in KEYS:
        if k in col:
            print(f"   {name:22s} {r[col[k]]} {units[col[k]]}")
    if elems and "smsp__inst_executed.sum" in col:
        print(f"   thread-inst / element  {float(r[col['smsp__inst_executed.sum']]) * 32 / elems:.1f}")
    st = [(float(r[i]), h.split('issue_stalled_')[1].split('_per_')[0]) for h, i in col.items()
          if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "0")]
    print("   stalls (warps per issue):", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
