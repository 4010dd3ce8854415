#!/usr/bin/env python
"""Instruction budget of a kernel by execution-count bucket (phases of a persistent kernel): rep regex lo:hi:name ..."""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[his[0]]; end = his[1] if len(his) > 1 else len(rows)
seen, data = set(), []
for r in rows[his[0] + 1:end]:
    if len(r) == len(hdr) and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); data.append(r)
ie = hdr.index("Instructions Executed"); ia = hdr.index("Warp Stall Sampling (All Samples)")
tot = sum(int(r[ie]) for r in data); tots = sum(int(r[ia]) for r in data)
counts = sorted(set(int(r[ie]) for r in data), reverse=True)
# cluster execution counts within 8 %
clusters = []
for r in sorted(data, key=lambda r: -int(r[ie])):
    n = int(r[ie])
    if clusters and n > 0 and abs(clusters[-1][0] - n) <= 0.08 * clusters[-1][0]:
        c = clusters[-1]; c[1] += 1; c[2] += n; c[3] += int(r[ia])
    else:
        clusters.append([n, 1, n, int(r[ia])])
print("total warp-instructions %d, stall samples %d" % (tot, tots))
for n, lines, instr, samples in sorted(clusters, key=lambda c: -c[2])[:10]:
    print("exec/line ~%9d  sass lines %5d  warp-instr %11d (%4.1f%%)  stall samples %6d (%4.1f%%)" % (n, lines, instr, 100.0 * instr / tot, samples, 100.0 * samples / max(1, tots)))
