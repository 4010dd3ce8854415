#!/usr/bin/env python
"""Warp-instructions and stall samples per CUDA source line of a kernel: rep regex [top]  (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
hdr = rows[his[0]]
end = his[1] if len(his) > 1 else len(rows)
ie = hdr.index("Instructions Executed"); ia = hdr.index("Warp Stall Sampling (All Samples)")
lines = []
for r in rows[his[0] + 1:end]:
    if len(r) == len(hdr) and r[0].isdigit():
        lines.append((int(r[ie]), int(r[ia]), int(r[0]), r[1].strip()))
tot = sum(l[0] for l in lines); tots = sum(l[1] for l in lines)
print("source lines %d  warp-instructions %d  stall samples %d" % (len(lines), tot, tots))
for e, a, no, src in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% instr %5.1f%% samples  line %4d  %s" % (100.0 * e / tot, 100.0 * a / tots, no, src[:110]))
