#!/usr/bin/env python
"""Warp-instructions and stall samples of count_planes_kernel by section of mdg_planes.cuh: rep"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:count_planes"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
hdr = rows[his[0]]; end = his[1] if len(his) > 1 else len(rows)
ie = hdr.index("Instructions Executed"); ia = hdr.index("Warp Stall Sampling (All Samples)")
L = [(int(r[0]), int(r[ie]), int(r[ia])) for r in rows[his[0] + 1:end] if len(r) == len(hdr) and r[0].isdigit()]
tot = sum(x[1] for x in L); ts = sum(x[2] for x in L)
lines = open('mapdamage_b200/csrc/mdg_planes.cuh').read().split('\n')
def find(pat):
    for i, l in enumerate(lines):
        if pat in l: return i + 1
marks = [("transpose", find("__device__ __forceinline__ uint32_t nibbles_to_planes")), ("setup", find("template <int kThreads>")), ("spill", find("auto spill = [&]")),
         ("flush", find("const int LA = L + A;")), ("stage_window", find("// ---- stage: the plane words of one window")), ("parse_read", find("__shared__ uint32_t indel_here;")),
         ("prefetch", find("// L2 prefetch of a tile two ahead")), ("tile loop/parse", find("constexpr int PREP = 2;")), ("lists/mode", find("// ---- reads this kernel does not count go")),
         ("stage loop", find("// ---- pull the tile after next towards L2")), ("count", find("// ---- count: this thread's window word"))]
marks = sorted([m for m in marks if m[1]], key=lambda m: m[1])
reads = float(sys.argv[2]) if len(sys.argv) > 2 else 4166667.0
print("warp-instructions %d (%.1f per read), stall samples %d" % (tot, tot / reads, ts))
for (n, a), (n2, b) in zip(marks, marks[1:] + [("eof", 10 ** 6)]):
    ins = sum(x[1] for x in L if a <= x[0] < b); sm = sum(x[2] for x in L if a <= x[0] < b)
    print("%-16s lines %4d-%4d  %5.1f%% instr (%5.1f/read) %5.1f%% samples" % (n, a, b, 100 * ins / tot, ins / reads, 100 * sm / ts))
