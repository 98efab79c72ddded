#!/usr/bin/env python3
"""Per-instruction shared-memory wavefronts of one kernel from an `ncu --set full --import-source on` report.

    ncu -i gpurun_out/prof_tile64_v6.ncu-rep --page source --csv > /tmp/src_sass.csv
    python scripts/ncu_shared_wavefronts.py /tmp/src_sass.csv 2000376 [min wavefronts per element to list]

Prints the total wavefronts per element (second argument: elements per launch) and every LDS / STS whose share exceeds the threshold,
with its executions per element and the ideal (conflict-free) count - the numbers behind "Flush bank-conflict model" in
profiles/r01/README.md."""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    elements = float(sys.argv[2])
    threshold = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    header = rows[1]
    ix = {c: i for i, c in enumerate(header)}
    total = excess = 0.0
    lines = []
    for n, r in enumerate(rows[2:]):
        w = int(r[ix["L1 Wavefronts Shared"]] or 0)
        if w == 0:
            continue
        ideal = int(r[ix["L1 Wavefronts Shared Ideal"]] or 0)
        ex = int(r[ix["Instructions Executed"]] or 0)
        total += w
        excess += max(w - ideal, 0)
        lines.append((n, r[ix["Source"]].strip()[:56], ex / elements, w / elements, ideal / elements))
    print(f"shared-memory wavefronts per element: {total / elements:.2f}  (excess over ideal: {excess / elements:.2f})")
    for n, src, ex, w, ideal in lines:
        if w >= threshold:
            print(f"{n:5d} {src:<56s} exec/el={ex:6.3f} wavefronts/el={w:6.2f} ideal={ideal:6.2f} per instr={w / max(ex, 1e-12):5.2f}")


if __name__ == "__main__":
    main()
