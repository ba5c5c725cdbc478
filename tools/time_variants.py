#!/usr/bin/env python3
"""Time K1 (random-policy step, 262,144 envs, 8 rotating batches) for the library named by BBGPU_LIB."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bbgpu import capi
n, M, K = 262144, 8, 400
envs = [capi.EnvHandle(n, 42, b * n) for b in range(M)]
outs = [dict(a=torch.zeros(n, dtype=torch.int32, device="cuda"), r=torch.zeros(n, device="cuda"),
             t=torch.zeros(n, dtype=torch.uint8, device="cuda"), m=torch.zeros((3, n), dtype=torch.int64, device="cuda")) for _ in range(M)]
stats = torch.zeros(16, dtype=torch.int64, device="cuda")
for e in envs:
    e.step_random(64)
def run(k):
    o = outs[k % M]
    envs[k % M].step_random(1, o["a"], o["r"], o["t"], o["m"], stats)
for k in range(40):
    run(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(K):
    run(k)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / K
e0.record()
envs[0].step_random(256)
e1.record()
torch.cuda.synchronize()
print("%-28s launch %.1f us  %.2f G env-steps/s | fused256: %.2f G/s" % (os.path.basename(os.environ.get("BBGPU_LIB", "default")), us, n / us / 1e3, n * 256 / (e0.elapsed_time(e1) * 1e-3) / 1e9))
