#!/usr/bin/env python
"""Where is the ceiling of the end-to-end number at N GPUs?  One process per GPU (torchrun), each moving pinned host
buffers to and from its GPU at the same time as all the others: H2D alone, D2H alone, both directions at once, and the
host-side copy pageable -> pinned that the plugin's pageable path adds.  Rank 0 prints one JSON line with per-rank and
aggregate GB/s.   torchrun --nproc-per-node N tools/pcie_probe.py [--mib 1024]"""
import argparse, json, os, time
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = a.mib << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_in.fill_(1)
h_out = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out.fill_(2)
pageable = torch.ones(n, dtype=torch.uint8)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

def timed(fn):
    best = 1e9
    for _ in range(a.reps):
        barrier()
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t1 = time.perf_counter()
        best = min(best, t1 - t0)
    return n / best / 1e9

def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
def host_copy():
    h_in.copy_(pageable)

res = {"h2d": timed(h2d), "d2h": timed(d2h), "both_each_direction": timed(both), "host_copy_pageable_to_pinned_1thread": timed(host_copy)}
aff = sorted(os.sched_getaffinity(0))
vals = torch.tensor([res["h2d"], res["d2h"], res["both_each_direction"], res["host_copy_pageable_to_pinned_1thread"]], dtype=torch.float64, device="cuda")
if world > 1:
    allv = [torch.empty_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    per = [[round(float(x), 2) for x in v.tolist()] for v in allv]
    agg = [round(sum(p[i] for p in per), 1) for i in range(4)]
    print(json.dumps({"probe": "pcie", "n_gpus": world, "mib_per_transfer": a.mib, "columns": ["h2d", "d2h", "both (each direction)", "host copy, 1 thread"],
                      "per_rank_GBps": per, "aggregate_GBps": agg, "host_cpus_visible": len(aff)}))
if world > 1:
    dist.destroy_process_group()
