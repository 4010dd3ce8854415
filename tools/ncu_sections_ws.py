#!/usr/bin/env python
"""Warp-instructions per read and stall samples by section of count_planes_ws_kernel (line ranges of mdg_planes_ws.cuh
found by their marker comments): rep reads"""
import csv, io, os, re, subprocess, sys
rep, reads = sys.argv[1], float(sys.argv[2])
src = open(sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "mapdamage_b200", "csrc", "mdg_planes_ws.cuh")).read().split("\n")
marks = [("helpers (mbarrier, plane bytes)", r"^__device__ __forceinline__ void mbar_init"), ("set-up", r"^count_planes_ws_kernel"),
         ("stage: addresses", r"auto stage_window"), ("stage: planes, masks, store, events", r"auto emit = "),
         ("stage: loads + loop", r"if constexpr \(kNW > 0\)"), ("parse_read", r"auto parse_read"), ("prefetch / issue", r"auto prefetch_headers"),
         ("parse loop + lists", r"// ---- parse: one read per thread"), ("work lists, mode", r"// ---- reads this kernel does not count"),
         ("stage loop (waits, items, barriers)", r"// ---- stage: one thread per"), ("consumer set-up / spill", r"// =+ consumers"),
         ("consumer flush", r"auto flush = "), ("consumer loop (wait, bounds)", r"int since_flush = 0"), ("consumer count", r"// ---- count: this thread"),
         ("final drain", r"// ---- everybody")]
starts = []
for name, pat in marks:
    for i, l in enumerate(src):
        if re.search(pat, l):
            starts.append((i + 1, name)); break
starts.sort()
def section(fn, no):
    if "planes_ws" not in fn: return "inlined: " + fn
    cur = "top"
    for s, name in starts:
        if no >= s: cur = name
    return cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:count_planes_ws"],
                     capture_output=True, text=True).stdout
agg, fname, hdr = {}, "?", None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; ie = hdr.index("Instructions Executed")
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]; si = [hdr.index(h) for h in stalls]
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        a = agg.setdefault(section(fname, int(r[0])), [0] + [0] * len(si))
        a[0] += int(r[ie])
        for k, i in enumerate(si): a[k + 1] += int(r[i] or 0)
tot = sum(sum(v[1:]) for v in agg.values()); ti = sum(v[0] for v in agg.values())
print("%-40s %7s %6s | top stall reasons (%% of all samples)" % ("section", "i/read", "samp%"))
for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1][1:])):
    s = sum(v[1:])
    why = sorted(range(len(stalls)), key=lambda k: -v[k + 1])[:4]
    print("%-40s %7.2f %6.1f | %s" % (name[:40], v[0] / reads, 100.0 * s / tot, ", ".join("%s %.1f" % (stalls[k][6:], 100.0 * v[k + 1] / tot) for k in why if v[k + 1])))
print("%-40s %7.2f" % ("total", ti / reads))
