#!/usr/bin/env python3
"""K1 launches of independent env batches alternated over S streams (GPU box).

A single-step launch over 262,144 envs ends with a drain: a few long warps (trio searches) keep
their SMs while the rest idle.  Consecutive bench launches work on independent env batches, so
launch k+1 may start while launch k drains when they sit on different streams.  Prints the
per-launch time for S = 1, 2, 3, 4 streams, raw ctypes calls (no torch stream lookups)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bbgpu import capi

n = int(os.environ.get("N", 262144))
M, K = 8, int(os.environ.get("K", 800))
L = capi.lib()
envs = [capi.EnvHandle(n, 42, b * n) for b in range(M)]
outs = [dict(a=torch.zeros(n, dtype=torch.int32, device="cuda"), r=torch.zeros(n, device="cuda"),
             t=torch.zeros(n, dtype=torch.uint8, device="cuda"), m=torch.zeros((3, n), dtype=torch.int64, device="cuda")) for _ in range(M)]
stats = torch.zeros(64, dtype=torch.int64, device="cuda")
for e in envs:
    e.step_random(64)
torch.cuda.synchronize()


def run(S, K):
    streams = [torch.cuda.Stream() for _ in range(S)]
    sp = [s.cuda_stream for s in streams]
    args = [(envs[b].h, 1, o["a"].data_ptr(), o["r"].data_ptr(), o["t"].data_ptr(), o["m"].data_ptr(), stats.data_ptr())
            for b, o in enumerate(outs)]
    f = L.bb_env_step_random
    for k in range(40):
        f(*args[k % M], sp[k % S])
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
    main = torch.cuda.current_stream()
    e0.record(main)
    for s in streams:
        s.wait_event(e0)
    t0 = time.perf_counter()
    for k in range(K):
        f(*args[k % M], sp[k % S])
    t_issue = time.perf_counter() - t0
    for s, e in zip(streams, ends):
        e.record(s)
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e) for e in ends)
    return ms * 1e3 / K, t_issue * 1e6 / K


for S in (1, 2, 3, 4, 8):
    if M % S:
        continue
    best = min(run(S, K) for _ in range(3))
    print("streams=%d  %.2f us per launch (%.2f G env-steps/s), host issue %.2f us per launch" % (S, best[0], n / best[0] / 1e3, best[1]), flush=True)

# D2H bandwidth of a dense-observation sized block (1,216 B/env) from device to pinned host memory
nb = 1216 * n
d = torch.zeros(nb, dtype=torch.uint8, device="cuda")
h = torch.zeros(nb, dtype=torch.uint8).pin_memory()
for _ in range(3):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print("D2H %d MB pinned: %.2f ms = %.1f GB/s -> %.1f M env-steps/s ceiling for dense obs" % (nb >> 20, dt * 1e3, nb / dt / 1e9, n / dt / 1e6))
import numpy as np
# host-side expansion cost (single thread numpy) for reference
b = np.random.randint(0, 2**63, n, dtype=np.int64).view(np.uint64)
t0 = time.perf_counter()
x = np.unpackbits(b.view(np.uint8).reshape(-1, 8), axis=1, bitorder="little").reshape(-1, 8, 8).astype(np.float32)
print("numpy expand_board: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
print("host cores", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
os.system("nvidia-smi topo -m 2>/dev/null | head -20; lscpu | grep -E 'Model name|Socket|NUMA|^CPU\\(s\\)' ")
