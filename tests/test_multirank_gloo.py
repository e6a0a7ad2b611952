"""world_size-2 CPU test (gloo) of the N>1 host logic: slab sharding, the partials all-reduce used by
bench.py (vizaly-foresight_b200/dist.py), CBench's brick partitioner, and the slab-concatenation identity.
The per-rank partials come from the oracle here (no GPU); the combine code is the product's."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, smooth_field


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, shape, mb, q):
    import sys

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    import oracle as O
    from conftest import smooth_field as sf

    pkg = ge.load_package()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = sf(shape, seed=77, amp=40.0)
    thick = shape[0] // world
    mine = np.ascontiguousarray(full[rank * thick:(rank + 1) * thick])
    stream = O.compress(mine, mb)
    dec = O.decompress(stream, mine.shape, np.float32, mb)
    m = O.metrics(mine, dec)
    p = pkg.Partials(m["sum_sq"], m["abs_sum"], m["rel_sum"], m["absolute_error"], m["relative_error"], m["max"],
                     -m["min"], mine.size)
    t = torch.frombuffer(bytearray(bytes(p)), dtype=torch.int64).clone()
    import importlib.util
    spec = importlib.util.spec_from_file_location("zfb_dist", os.path.join(ROOT, "vizaly-foresight_b200", "dist.py"))
    zd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(zd)
    zd.allreduce_partials(t, dist)
    g = pkg.Partials.from_buffer_copy(t.numpy().tobytes())
    streams = [None] * world
    dist.all_gather_object(streams, stream.tobytes())
    if rank == 0:
        q.put((pkg.metric_values(g), b"".join(streams)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_slab_sharding_and_partials_allreduce(oracle, pkg):
    import torch.multiprocessing as mp

    shape, mb, world = (16, 24, 40), 512, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, shape, mb, q)) for r in range(world)]
    for p in procs:
        p.start()
    vals, stream_cat = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = smooth_field(shape, seed=77, amp=40.0)
    whole = oracle.compress(full, mb)
    # slab streams concatenate to the single-rank stream (fixed rate, slab thickness % 4 == 0)
    assert stream_cat == whole.tobytes()
    dec = oracle.decompress(whole, shape, np.float32, mb)
    ref = oracle.metrics(full, dec)
    assert vals["n"] == full.size
    for k in ("absolute_error", "relative_error", "minmax", "min", "max"):
        assert vals[k] == ref[k], k
    for k in ("mse", "psnr"):
        assert abs(vals[k] - ref[k]) <= 1e-12 * abs(ref[k]), k


def test_brick_partition_covers_the_field(pkg):
    """CBench-parity sharding (getPartition): bricks tile the field without overlap for 1..16 ranks."""
    ext = (64, 32, 48)
    for world in (1, 2, 4, 8, 16):
        seen = np.zeros(ext, dtype=np.int32)
        for r in range(world):
            cnt, off = pkg.partition(r, world, ext)
            seen[off[0]:off[0] + cnt[0], off[1]:off[1] + cnt[1], off[2]:off[2] + cnt[2]] += 1
        assert (seen == 1).all()
