#!/usr/bin/env python
"""Per-source-line warp instruction counts of a kernel: joins the per-SASS-instruction page of an ncu report with the
line table of the cubin the report was taken from (nvdisasm -g on the cubin inside libzfb.so).
  python tools/ncu_lines.py rep.ncu-rep <kernel substring> [top N]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "vizaly-foresight_b200", "libzfb.so")], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur = hdr = None
data = {}
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = r[1]; data[cur] = []; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if cur and hdr and len(r) == len(hdr):
        data[cur].append(dict(zip(hdr, r)))
src_cache = {}
def src_line(f, n):
    if f not in src_cache:
        try: src_cache[f] = open(f).read().split("\n")
        except OSError: src_cache[f] = []
    L = src_cache[f]
    return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""
for k, ins in data.items():
    if pat not in k:
        continue
    # the kernel's section in the disassembly: mangled names carry the same identifier
    ident = re.search(r"(\w+_kernel)", k).group(1)
    sec = [m.start() for m in re.finditer(r"^\.text\._ZN3zfb\w*?\d+%s\w*:" % ident, dis, re.M)]
    best = None
    for s in sec:
        e = dis.find("//--------------------- .text.", s)
        body = dis[s:e if e > 0 else len(dis)]
        n = len(re.findall(r"/\*[0-9a-f]{4,6}\*/\s+\S", body))
        if n == len(ins):
            best = body
    if best is None:
        print("## %s: no section with %d instructions in the cubin (different build?)" % (k[:60], len(ins)))
        continue
    line_of = {}
    f, n = "?", 0
    for l in best.split("\n"):
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            f, n = m.group(1), int(m.group(2)); continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+\S", l)
        if m:
            line_of[int(m.group(1), 16)] = (f, n)
    base = int(ins[0]["Address"], 16)
    agg = collections.defaultdict(lambda: [0, 0.0, 0])
    tot = 0
    for i in ins:
        off = int(i["Address"], 16) - base
        c = int(i["Instructions Executed"]); th = float(i["Avg. Threads Executed"] or 0); sm = int(i["# Samples"] or 0)
        a = agg[line_of.get(off, ("?", 0))]
        a[0] += c; a[1] += c * th; a[2] += sm
        tot += c
    print("## %s\ntotal warp instr %d" % (k[:70], tot))
    for (f, n), (c, cth, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%5.1f%% n=%11d thr/inst=%5.1f samp=%6d %s:%d  %s" % (100.0 * c / tot, c, cth / max(c, 1), sm, os.path.basename(f), n, src_line(f, n)))
    print()
