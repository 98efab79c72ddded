#!/usr/bin/env python3
"""Split the source page of an .ncu-rep of the Hex8 tile kernel into helper-warp and compute-warp code (the two USETMAXREG
instructions mark the regions) and print instructions / shared wavefronts / stall samples per region and the top stall sites.
usage: ncu_regions.py <report.ncu-rep> <elements-per-launch>"""
import collections
import csv
import io
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
ix = {c: i for i, c in enumerate(h)}
stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
region, regions = "prologue", collections.OrderedDict()
top = []
for k, r in enumerate(rows[2:]):
    s = r[ix["Source"]].strip()
    if "USETMAXREG" in s:
        region = "helper" if "DEALLOC" in s or ".DEC" in s.upper() else "compute"
        if region in regions:
            region += "2"
    d = regions.setdefault(region, collections.Counter())
    n = int(r[ix["Instructions Executed"]] or 0)
    d["inst"] += n
    d["wf"] += int(r[ix["L1 Wavefronts Shared"]] or 0)
    d["wfx"] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
    smp = int(r[ix["# Samples"]] or 0)
    d["samples"] += smp
    for c in stall_cols:
        if r[ix[c]]:
            d[c] += int(r[ix[c]])
    top.append((smp, k, region, s, n))
tot = sum(d["samples"] for d in regions.values()) or 1
for name, d in regions.items():
    st = {c[6:]: round(100 * d[c] / max(d["samples"], 1), 1) for c in stall_cols if d[c] * 20 > d["samples"]}
    print(f"{name:9s} inst/unit {d['inst'] / units:7.1f}  shared wf/unit {d['wf'] / units:6.1f} (excess {d['wfx'] / units:5.1f})  samples {100 * d['samples'] / tot:5.1f}%  {st}")
print("top stall sites:")
for smp, k, region, s, n in sorted(top, reverse=True)[:40]:
    print(f"  {100 * smp / tot:5.2f}%  #{k:5d} {region:8s} exec/unit {n / units:6.2f}  {s[:90]}")
