#!/usr/bin/env python3
"""Time K1 (random-policy step, 262,144 envs, 8 rotating batches, launches alternated over two
streams) for every library given on the command line; one subprocess per library."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] != "--one":
    for lib in sys.argv[1:]:
        env = dict(os.environ, BBGPU_LIB=os.path.abspath(lib))
        subprocess.run([sys.executable, __file__, "--one"], env=env)
    sys.exit(0)
sys.path.insert(0, ROOT)
import torch
from bbgpu import capi
n, M, K = 262144, 8, 800
L = capi.lib()
envs = [capi.EnvHandle(n, 42, b * n) for b in range(M)]
outs = [dict(a=torch.zeros(n, dtype=torch.int32, device="cuda"), r=torch.zeros(n, device="cuda"),
             t=torch.zeros(n, dtype=torch.uint8, device="cuda"), m=torch.zeros((3, n), dtype=torch.int64, device="cuda")) for _ in range(M)]
stats = torch.zeros(64, dtype=torch.int64, device="cuda")
for e in envs:
    e.step_random(64)
torch.cuda.synchronize()
args = [(envs[b].h, 1, o["a"].data_ptr(), o["r"].data_ptr(), o["t"].data_ptr(), o["m"].data_ptr(), stats.data_ptr()) for b, o in enumerate(outs)]
f = L.bb_env_step_random
res = []
for S in (1, 2):
    streams = [torch.cuda.Stream() for _ in range(S)]
    sp = [s.cuda_stream for s in streams]
    best = 1e9
    for rep in range(4):
        for k in range(40):
            f(*args[k % M], sp[k % S])
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(S)]
        e0.record(torch.cuda.current_stream())
        for s in streams:
            s.wait_event(e0)
        for k in range(K):
            f(*args[k % M], sp[k % S])
        for s, e in zip(streams, ends):
            e.record(s)
        torch.cuda.synchronize()
        best = min(best, max(e0.elapsed_time(e) for e in ends) * 1e3 / K)
    res.append(best)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); envs[0].step_random(256); e1.record(); torch.cuda.synchronize()
print("%-22s 1 stream %.2f us | 2 streams %.2f us (%.2f G/s, frac %.4f) | fused256 %.2f G/s" % (
    os.path.basename(os.environ.get("BBGPU_LIB", "default")), res[0], res[1], n / res[1] / 1e3,
    129 * n / res[1] / 1e3 / 6537.3, n * 256 / (e0.elapsed_time(e1) * 1e-3) / 1e9), flush=True)
