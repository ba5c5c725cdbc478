#!/usr/bin/env python3
"""Launch each kernel a few times at bench sizes (for ncu captures of K1-K4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bbgpu import capi
dev = "cuda"
n = 262144
envs = [capi.EnvHandle(n, 42, b * n) for b in range(4)]
a = torch.zeros(n, dtype=torch.int32, device=dev); r = torch.zeros(n, device=dev)
t = torch.zeros(n, dtype=torch.uint8, device=dev); m = torch.zeros((3, n), dtype=torch.int64, device=dev)
st = torch.zeros(64, dtype=torch.int64, device=dev)
for e in envs:
    e.step_random(64)
for k in range(8):
    envs[k % 4].step_random(1, a, r, t, m, st)                # K1 random policy (bench kernel)
envs[0].observe(None, None, m)
for k in range(2):
    envs[0].step(a, r, t, m, None, None, None)                 # K1 given actions
T, N = 128, 262144
rw, v = torch.randn(T, N, device=dev), torch.randn(T, N, device=dev)
d = (torch.rand(T, N, device=dev) < 0.07).float(); lv = torch.randn(N, device=dev)
adv, ret = torch.empty_like(rw), torch.empty_like(rw); mom = torch.zeros(2, dtype=torch.float64, device=dev)
for _ in range(2):
    capi.gae(rw, v, d, lv, 0.99, 0.95, adv, ret, mom)          # K4
nn = 524288
logits = torch.randn(nn, 192, device=dev)
mask = torch.randint(-2 ** 62, 2 ** 62, (3, nn), dtype=torch.int64, device=dev) | 1
act = torch.empty(nn, dtype=torch.int32, device=dev); lp, en = torch.empty(nn, device=dev), torch.empty(nn, device=dev)
for _ in range(2):
    capi.masked_sample(logits, mask, nn, 1, 1, 0, act, lp, en) # K3
board = torch.randint(-2 ** 62, 2 ** 62, (nn,), dtype=torch.int64, device=dev)
pieces = torch.randint(0, 37, (nn,), dtype=torch.int32, device=dev) * 0x010101
obs = torch.empty((nn, 4, 8, 8), device=dev)
dense = torch.empty((nn, 192), dtype=torch.uint8, device=dev)
for _ in range(2):
    capi.unpack_obs(board, pieces, mask, nn, obs=obs, mask_dense=dense, n=nn)   # K2
torch.cuda.synchronize()
print("ok")
