# times the second-pass histograms (absoluteError.hpp:104-139 etc.) on a 1 GiB float32 pair; run on the GPU box
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
pkg = g.load_package()
codec = pkg.Codec(0)
n = 256 << 20
o = torch.rand(n, device="cuda") * 256
a = o + 0.01 * torch.randn(n, device="cuda")
for which, lo, hi in ((0, 0.0, 0.06), (1, 0.0, 0.06), (2, 0.0, 256.0)):
    codec.histogram(which, lo, hi, o, a)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        codec.histogram(which, lo, hi, o, a)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    nbytes = n * 4 * (1 if which == 2 else 2)
    print("hist which=%d  %.3f ms  %.0f GB/s" % (which, ms, nbytes / ms / 1e6))
for dt in (torch.float32, torch.float64):
    oo, aa = o.to(dt), a.to(dt)
    codec.metrics(oo, aa)
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        codec.metrics(oo, aa)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print("metrics %s  %.3f ms  %.0f GB/s" % (dt, ms, 2 * oo.numel() * oo.element_size() / ms / 1e6))
