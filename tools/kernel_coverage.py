"""Small driver for compute-sanitizer: touches every kernel family once on small shapes (run on the GPU box):
   python tools/kernel_coverage.py   (touches every kernel family once; compute-sanitizer is closed on this pool)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge

pkg = ge.load_package()
c = pkg.Codec(0)
rng = np.random.default_rng(0)


def rt(shape, dtype, rate, wra=1):
    f = (rng.standard_normal(shape) * 10).astype(dtype)
    mb = pkg.maxbits(dtype, len(shape), rate, wra)
    d = torch.from_numpy(f).cuda()
    s = c.encode(d, mb)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    o = c.decode(s, shape, tdt, mb, orig=d)
    p = c.partials()
    o2 = c.decode(s, shape, tdt, mb)
    torch.cuda.synchronize()
    return p.n


for shape in [(8, 8, 256), (12, 8, 516), (7, 9, 12), (5, 6, 7), (4, 5, 260)]:
    for rate in (4, 8, 16):
        rt(shape, np.float32, rate)
rt((8, 8, 256), np.float32, 3, wra=0)
for shape in [(8, 8, 256), (7, 9, 12), (5, 6, 7)]:
    for rate in (1, 8, 32):
        rt(shape, np.float64, rate)
for n in (10, 4096, 4099):
    for rate in (8, 32):
        rt((n,), np.float32, rate)
    rt((n,), np.float32, 8, wra=0)
    rt((n,), np.float64, 8)
rt((17, 19), np.float32, 8)
rt((64, 64), np.float64, 8)
a = torch.randn(10001, device="cuda")
b = a + 0.01 * torch.randn_like(a)
c.metrics(a[:10000], b[:10000])
cnt, oob = c.histogram(0, 0.0, 0.05, a[:10000], b[:10000])
h = (rng.standard_normal((8, 16, 64)) * 3).astype(np.float32)
mb = pkg.maxbits(np.float32, 3, 8)
s = c.compress_host(h, mb)
g = c.decompress_host(s, h.shape, np.float32, mb, orig=h)
p = c.metrics_host(h, g)
torch.cuda.synchronize()
print("sanitize driver done", p.n, int(cnt.sum().item()))
