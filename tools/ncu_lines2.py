#!/usr/bin/env python
"""Per source line of a kernel (all source files): warp-instructions per read and stall samples by reason.
    ncu_lines2.py rep regex reads [top] [byline]"""
import csv, io, os, subprocess, sys
rep, pat, reads = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
lines, fname, hdr = [], "?", None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed")
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; si = [hdr.index(h) for h in stalls]
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        lines.append((fname, int(r[0]), int(r[ie]), [int(r[i] or 0) for i in si], r[1].strip()))
tot = sum(sum(l[3]) for l in lines)
print("warp-instructions per read %.2f, stall samples %d" % (sum(l[2] for l in lines) / reads, tot))
tot_by = [sum(l[3][k] for l in lines) for k in range(len(stalls))]
print("samples by reason: " + "  ".join("%s %.1f%%" % (stalls[k][6:], 100.0 * tot_by[k] / tot) for k in range(len(stalls)) if tot_by[k] * 200 > tot))
if len(sys.argv) > 5 and sys.argv[5] == "byline":
    lines.sort(key=lambda l: (l[0], l[1]))
else:
    lines.sort(key=lambda l: -sum(l[3]))
for fn, no, e, st, src in lines[:top]:
    s = sum(st)
    if not s and not e: continue
    why = sorted(range(len(stalls)), key=lambda k: -st[k])[:3]
    print("%-14s %4d %6.2f i/rd %5.2f%% [%s]  %s" % (fn[4:18], no, e / reads, 100.0 * s / tot,
          ", ".join("%s %.1f" % (stalls[k][6:12], 100.0 * st[k] / tot) for k in why if st[k]), src[:80]))
