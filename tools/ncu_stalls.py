#!/usr/bin/env python
"""Stall-reason samples of a kernel from an ncu report, split into address regions.
  python tools/ncu_stalls.py rep.ncu-rep <kernel substring> [split offsets in hex ...]"""
import csv, io, subprocess, sys, collections
rep, pat = sys.argv[1], sys.argv[2]
splits = [int(x, 16) for x in sys.argv[3:]]
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
    reg = collections.defaultdict(collections.Counter)
    for i in ins:
        off = int(i["Address"], 16) - base
        ri = sum(1 for s in splits if off >= s)
        for c in hdr:
            if c.startswith("stall_") and "Not Issued" not in c:
                reg[ri][c] += int(i[c] or 0)
        reg[ri]["instr"] += int(i["Instructions Executed"])
    for ri in sorted(reg):
        tot = sum(v for c, v in reg[ri].items() if c.startswith("stall_"))
        print("region", ri, "warp instr", reg[ri]["instr"], "samples", tot)
        for c, v in reg[ri].most_common():
            if c.startswith("stall_") and v > 0.01 * tot:
                print("   %-24s %6.1f%%" % (c, 100.0 * v / tot))
