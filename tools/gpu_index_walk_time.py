# times the sequential index rebuild of a bare variable-rate stream (zfb_decode_var_dev without an index)
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as g
import bench
pkg = g.load_package()
codec = pkg.Codec(0)
for shape in ((64, 256, 256), (128, 512, 512)):
    f = bench.make_field(torch, shape, "float32", "lognormal", 0, torch.device("cuda", 0))
    m = pkg.mode_accuracy(0.01)
    s, nbytes = codec.encode_var(f, m)
    bare = s.clone()
    torch.cuda.synchronize()
    t = time.time()
    out = codec.decode_var(bare, shape, torch.float32, m)
    torch.cuda.synchronize()
    dt = time.time() - t
    nb = pkg.block_count(shape)
    ref = codec.decode_var(s, shape, torch.float32, m)
    print("shape", shape, "blocks", nb, "bare decode %.3f s  (%.2f us per block)" % (dt, dt / nb * 1e6), "equal", bool(torch.equal(out, ref)))
# the same with a coarse index (every 128th entry): parallel expansion
for shape in ((128, 512, 512), (256, 1024, 1024)):
    f = bench.make_field(torch, shape, "float32", "lognormal", 0, torch.device("cuda", 0))
    m = pkg.mode_accuracy(0.01)
    full = torch.zeros(pkg.index_count(shape) + 1, dtype=torch.int64, device="cuda")
    s, nbytes = codec.encode_var(f, m, index=full)
    coarse = torch.cat([full[:-1][::128], full[-1:]]).contiguous()
    torch.cuda.synchronize()
    t = time.time()
    got = codec.index_expand(s, shape, torch.float32, m, coarse, 128)
    torch.cuda.synchronize()
    dt = time.time() - t
    print("shape", shape, "blocks", pkg.block_count(shape), "coarse entries", coarse.numel(), "expand %.1f ms" % (dt * 1e3), "equal", bool(torch.equal(got, full)))
