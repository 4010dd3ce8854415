#!/usr/bin/env python
"""Hot instructions of one kernel from an .ncu-rep (source page): python tools/ncu_hot.py rep [regex] [topN] [context addr]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else "count_"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hdr = rows[his[0]]
end = his[1] if len(his) > 1 else len(rows)
data, seen = [], set()
for r in rows[his[0] + 1:end]:
    if len(r) == len(hdr) and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); data.append(r)
ia = hdr.index("Warp Stall Sampling (All Samples)"); ie = hdr.index("Instructions Executed")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ia]) for r in data)
print("instructions %d  samples %d  warp-instructions executed %d" % (len(data), tot, sum(int(r[ie]) for r in data)))
agg = {}
for r in data:
    for i in cols:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i] or 0)
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
for r in sorted(data, key=lambda r: -int(r[ia]))[:top]:
    reasons = {hdr[i][6:]: int(r[i] or 0) for i in cols if int(r[i] or 0) > 0}
    print("%6s %10s  %-58s %s" % (r[ia], r[ie], r[1].strip()[:58], dict(sorted(reasons.items(), key=lambda kv: -kv[1])[:3])))
if len(sys.argv) > 4:
    key = sys.argv[4]
    idx = [k for k, r in enumerate(data) if key in r[1]][0]
    n = int(sys.argv[5]) if len(sys.argv) > 5 else 60
    for r in data[max(0, idx - n):idx + n]:
        print("%6s %10s  %s" % (r[ia], r[ie], r[1].strip()[:100]))
