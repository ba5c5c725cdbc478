#!/usr/bin/env python3
"""Device->host copy rate per rank, one rank at a time vs all ranks at once (torchrun, GPU box):
what bounds the host-buffer step (bb_env_step_host: one pinned D2H of 41 B/env + a stream
synchronise per step per rank) when 8 ranks run it together.

    python -m torch.distributed.run --nproc-per-node 8 tools/d2h_scaling.py > gpurun_out/d2h_scaling.json
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bbgpu.dist import pin_to_gpu_numa

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
bind = pin_to_gpu_numa(local, world) if "--no-pin" not in sys.argv else {"pinned": False}
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def rate(nbytes, reps, sync_each):
    d = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    h = torch.zeros(nbytes, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
        if sync_each:
            torch.cuda.current_stream().synchronize()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


sizes = {"packed_step_11.8MB": (45 * 262144, 200, True), "dense_step_320MB": (1221 * 262144, 10, True)}
res = {}
for name, (nb, reps, se) in sizes.items():
    alone = 0.0
    for r in range(world):                # one rank at a time
        barrier()
        if r == rank:
            alone = rate(nb, reps, se)
        barrier()
    barrier()
    together = rate(nb, reps, se)         # all ranks at once
    barrier()
    res[name] = (alone, together)
rows = [None] * world
payload = {"rank": rank, "binding": bind, **{k: {"alone_GBs": v[0], "all_ranks_GBs": v[1]} for k, v in res.items()}}
if world > 1:
    dist.all_gather_object(rows, payload)
else:
    rows = [payload]
if rank == 0:
    out = {"world": world, "host_cpus": os.cpu_count(), "ranks": rows}
    for k in sizes:
        a = [r[k]["alone_GBs"] for r in rows]
        t = [r[k]["all_ranks_GBs"] for r in rows]
        out[k] = {"alone_mean_GBs": sum(a) / world, "together_mean_GBs": sum(t) / world, "together_total_GBs": sum(t),
                  "efficiency": sum(t) / sum(a)}
    print(json.dumps(out, indent=1))
if world > 1:
    dist.destroy_process_group()
