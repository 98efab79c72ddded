#!/usr/bin/env python3
"""Summarise an .ncu-rep (one kernel launch, `ncu --set full`) into a small text file for profiles/.
usage: summarize_ncu.py <report.ncu-rep> <units-per-launch> [label]"""
import collections
import csv
import io
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
label = sys.argv[3] if len(sys.argv) > 3 else "unit"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, val = rows[0], rows[1], rows[2]
M = {h: (val[i], unit[i]) for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_red.sum",
        "lts__t_sectors_srcunit_tex_op_red.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {rep}")
for k in want:
    if k in M:
        print(f"{k:75s} {M[k][0]} {M[k][1]}")
try:
    dr = float(M["dram__bytes_read.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[M["dram__bytes_read.sum"][1]]
    dw = float(M["dram__bytes_write.sum"][0]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[M["dram__bytes_write.sum"][1]]
    print(f"traffic (dram read+write) bytes per launch: {dr + dw:.4g}   per {label}: {(dr + dw) / units:.1f}")
    print(f"warp instructions per {label}: {float(M['smsp__inst_executed.sum'][0]) / units:.1f}")
except Exception as e:  # noqa
    print("traffic: n/a", e)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {c: i for i, c in enumerate(h)}
byop, wf, wfx, stall = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for r in rows[2:]:
    s = r[ix["Source"]].strip().split()
    if not s:
        continue
    op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
    byop[op] += int(r[ix["Instructions Executed"]] or 0)
    wf[op] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    wfx[op] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    for c in h:
        if c.startswith("stall_") and "Not Issued" not in c and r[ix[c]]:
            stall[c] += int(r[ix[c]])
print(f"SASS mix (warp instructions per {label}; shared wavefronts, excessive):")
for op, n in byop.most_common(14):
    print(f"  {op:8s} {n / units:8.1f}   {wf[op] / units:8.1f} {wfx[op] / units:8.1f}")
tot = sum(stall.values()) or 1
print("stall samples %:", {k[6:]: round(100 * v / tot, 1) for k, v in stall.most_common(8)})
