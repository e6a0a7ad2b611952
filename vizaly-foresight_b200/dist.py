"""Multi-GPU combine of the metric partials: the one exchange step of the hot path.

The reference does this with blocking MPI_Allreduce calls on MPI_COMM_WORLD (absoluteError.hpp:80-92,
relativeError.hpp:93-105, meansquareError.hpp:70-71, psnrError.hpp:75-83, minmaxMetric.hpp:77-81).  Here
the same quantities ride two tiny all-reduces over NCCL / NVLink (SUM of f64[3] + n, MAX of f64[4]; the
minimum is stored negated so MAX covers it).  The tensor is the context's device-resident zfb_partials
viewed as 8 x int64 (Codec.partials_tensor()); on CPU tensors with the gloo backend the identical code
path is what tests/test_multirank_gloo.py exercises.
"""


def allreduce_partials(p_i64, dist, scratch=None):
    """In-place all-reduce of a zfb_partials laid out as 8 x int64 (raw bits)."""
    import torch

    pf = p_i64.view(torch.float64)
    if scratch is None:
        scratch = (torch.empty(4, dtype=torch.float64, device=p_i64.device),
                   torch.empty(4, dtype=torch.float64, device=p_i64.device))
    s, m = scratch
    s[:3] = pf[:3]
    s[3] = p_i64[7].to(torch.float64)  # n < 2^53: exact in float64
    m[:] = pf[3:7]
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    pf[:3] = s[:3]
    pf[3:7] = m
    p_i64[7] = s[3].round().to(torch.int64)
    return p_i64


def allreduce_histogram(counts_i64, dist):
    """SUM all-reduce of 1024 bin counters (absoluteError.hpp:129, relativeError.hpp:141, minmaxMetric.hpp:122)."""
    dist.all_reduce(counts_i64, op=dist.ReduceOp.SUM)
    return counts_i64
