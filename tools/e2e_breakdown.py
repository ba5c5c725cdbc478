#!/usr/bin/env python3
"""Where does one host-buffer env step spend its time? (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bbgpu import capi
n = 262144
h = capi.EnvHandle(n, 42)
pin = lambda dt, *s: torch.zeros(s or (n,), dtype=dt).pin_memory()
a = pin(torch.int32)
blk = capi.pinned_result_block(n)
r, t, board, pieces, mask, eps, epl, info = (blk[k] for k in ("rewards", "term", "board", "pieces", "mask", "ep_score", "ep_len", "info"))
def timeit(fn, k=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e6
ctr = [0]
def sample():
    ctr[0] += 1; h.sample_valid_actions(ctr[0], None, a)
print("sample_valid_actions (kernel + D2H 1 MB + sync): %.0f us" % timeit(sample))
def full():
    sample(); h.step_host(a, r, t, board, pieces, mask, eps, epl, info)
def lean():
    sample(); h.step_host(a, r, t, board, pieces, mask, None, None, None)
def minimal():
    sample(); h.step_host(a, r, t, None, None, None, None, None, None)
print("sample + step_host all outputs (53 B/env, one block): %.0f us" % timeit(full))
r2, t2 = pin(torch.float32), pin(torch.uint8)
def scattered():
    sample(); h.step_host(a, r2, t2, board, pieces, mask, eps, epl, info)
print("sample + step_host all outputs, separate arrays:       %.0f us" % timeit(scattered))
from bbgpu.vec_env import VectorizedBlockBlastEnv
venv = VectorizedBlockBlastEnv(n, seed=3, output="numpy", reuse_buffers=True)
venv.reset()
def api():
    venv.step(venv.sample_valid_actions())
print("VectorizedBlockBlastEnv numpy API sample + step:       %.0f us" % timeit(api))
print("sample + step_host obs only (41 B/env):    %.0f us" % timeit(lean))
print("sample + step_host rewards/term only:      %.0f us" % timeit(minimal))
d = torch.zeros(13893632 // 8, dtype=torch.int64, device="cuda"); hp = torch.zeros(13893632 // 8, dtype=torch.int64).pin_memory()
print("one 13.9 MB D2H copy + sync: %.0f us" % timeit(lambda: (hp.copy_(d, non_blocking=True), torch.cuda.current_stream().synchronize())))
d1 = torch.zeros(1048576 // 8, dtype=torch.int64, device="cuda"); hp1 = torch.zeros(1048576 // 8, dtype=torch.int64).pin_memory()
print("one 1 MB D2H copy + sync: %.0f us" % timeit(lambda: (hp1.copy_(d1, non_blocking=True), torch.cuda.current_stream().synchronize())))
print("one 1 MB H2D copy + sync: %.0f us" % timeit(lambda: (d1.copy_(hp1, non_blocking=True), torch.cuda.current_stream().synchronize())))
