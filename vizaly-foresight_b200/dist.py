"""Multi-GPU combine of the metric partials: the one exchange step of the hot path.

The reference does this with blocking MPI_Allreduce calls on MPI_COMM_WORLD (absoluteError.hpp:80-92,
relativeError.hpp:93-105, meansquareError.hpp:70-71, psnrError.hpp:75-83, minmaxMetric.hpp:77-81).  The product path
is the C ABI's zfb_allreduce_partials (csrc/zfb_comm.cuh): ONE ncclAllGather of the 64-byte zfb_partials and a
rank-ordered combine on the device.  This module is the same exchange written with torch.distributed -- one
all_gather of the 8 x int64 view (raw bits) and the same rank-ordered combine -- for back ends without NCCL:
tests/test_multirank_gloo.py runs it on CPU tensors with gloo, world size 2.
"""


def allreduce_partials(p_i64, dist):
    """In-place combine of a zfb_partials laid out as 8 x int64 (raw bits): all_gather + combine in rank order, so the
    result is bitwise the same on every rank and equal to zfb_partials_combine over the ranks."""
    import torch

    world = dist.get_world_size()
    if world == 1:
        return p_i64
    parts = [torch.empty_like(p_i64) for _ in range(world)]
    dist.all_gather(parts, p_i64)
    allp = torch.stack(parts)                      # world x 8 (int64 bits)
    f = allp.view(torch.float64)
    acc = f[0].clone()
    n = allp[0, 7].clone()
    for r in range(1, world):                      # fixed order: rank 0, 1, 2, ...
        acc[:3] += f[r, :3]
        acc[3:7] = torch.maximum(acc[3:7], f[r, 3:7])
        n += allp[r, 7]
    p_i64.view(torch.float64)[:7] = acc[:7]
    p_i64[7] = n
    return p_i64


def allreduce_histogram(counts_i64, dist):
    """SUM all-reduce of 1024 bin counters (absoluteError.hpp:129, relativeError.hpp:141, minmaxMetric.hpp:122)."""
    dist.all_reduce(counts_i64, op=dist.ReduceOp.SUM)
    return counts_i64
