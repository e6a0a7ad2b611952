#!/usr/bin/env python
"""Per-instruction listing of a kernel from an ncu report: offset, warp instructions executed, average active threads,
stall samples, SASS.  Consecutive instructions with the same execution count are also summarised as regions.
  python tools/ncu_sass.py rep.ncu-rep <kernel substring> [--full]"""
import csv, io, subprocess, sys

rep, pat = sys.argv[1], sys.argv[2]
full = "--full" in sys.argv
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = None; hdr = None; data = {}
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]; data[cur] = []; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur and hdr and len(r) == len(hdr):
        data[cur].append(dict(zip(hdr, r)))
for k, ins in data.items():
    if pat not in k:
        continue
    base = int(ins[0]["Address"], 16)
    tot = sum(int(i["Instructions Executed"]) for i in ins)
    print("##", k[:80], "total warp instr", tot)
    regions = []
    for i in ins:
        off = int(i["Address"], 16) - base
        n = int(i["Instructions Executed"]); th = float(i["Avg. Threads Executed"] or 0); sm = int(i["# Samples"] or 0)
        if full:
            print("%5x %10d %5.1f %6d  %s" % (off, n, th, sm, i["Source"].strip()))
        if regions and regions[-1][2] == n:
            regions[-1][1] = off; regions[-1][3] += 1; regions[-1][4] += sm; regions[-1][5] += th
        else:
            regions.append([off, off, n, 1, sm, th])
    if not full:
        for a, b, n, c, sm, th in regions:
            if n * c > tot * 0.002:
                print("%5x-%5x  exec %10d x %3d instrs = %5.2f%%  thr %5.1f  samples %6d" % (a, b, n, c, 100.0 * n * c / tot, th / c, sm))
