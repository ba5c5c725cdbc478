#!/usr/bin/env python3
"""Time K3 (masked sample, 524,288 rows, f32 and bf16 logits, with and without entropy) for every library
given on the command line (BBGPU_LIB variants); one subprocess per library."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] != "--one":
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, "--one"], env=dict(os.environ, BBGPU_LIB=os.path.abspath(lib)))
    sys.exit(0)
sys.path.insert(0, ROOT)
import torch
from bbgpu import capi
n = 524288
logits = torch.randn(n, 192, device="cuda")
lb = logits.bfloat16()
mask = torch.randint(-2 ** 62, 2 ** 62, (3, n), dtype=torch.int64, device="cuda") | 1
act = torch.empty(n, dtype=torch.int32, device="cuda")
lp, en = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
import time
t_end = time.time() + 1.0
while time.time() < t_end:                       # bring the clocks up before timing anything
    capi.masked_sample(logits, mask, n, 1, 1, 0, act, lp, en)
    torch.cuda.synchronize()
def t(fn):
    best = 1e9
    for rep in range(3):
        for _ in range(10): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 200 * 1e3)
    return best
r = [t(lambda: capi.masked_sample(logits, mask, n, 1, 1, 0, act, lp, en)), t(lambda: capi.masked_sample(lb, mask, n, 1, 1, 0, act, lp, en)),
     t(lambda: capi.masked_sample(lb, mask, n, 1, 1, 0, act, lp, None)), t(lambda: capi.masked_sample(lb, mask, n, 1, 1, 2, act, lp, en))]
print("%-20s f32+ent %.1f us (%.3f) | bf16+ent %.1f us (%.3f) | bf16 sample only %.1f us (%.3f) | bf16 evaluate+ent %.1f us" % (
    os.path.basename(os.environ.get("BBGPU_LIB", "default")), r[0], 804 * n / r[0] / 1e3 / 6537.3, r[1], 420 * n / r[1] / 1e3 / 6537.3,
    r[2], 416 * n / r[2] / 1e3 / 6537.3, r[3]), flush=True)
