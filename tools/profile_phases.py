#!/usr/bin/env python3
"""Debug tool: build libbbgpu with -DBB_PROFILE (clock64 phase timers in K1) into a scratch
copy and print where a warp's cycles go.  Run on a GPU box."""
import ctypes as C, os, shutil, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bbgpu import build as B
tmp = tempfile.mkdtemp()
objs = []
for src in B.SOURCES:
    o = os.path.join(tmp, src.replace(".cu", ".o"))
    subprocess.check_call([B._nvcc()] + B._host_compiler_flags() + [f for f in B.NVCC_FLAGS if f not in ("-v", "-warn-spills")][:-4] +
                          ["-Xcompiler", "-fPIC", "-DBB_PROFILE", "-c", os.path.join(B.CSRC, src), "-o", o], stderr=subprocess.DEVNULL)
    objs.append(o)
lib = os.path.join(tmp, "libbbgpu_prof.so")
subprocess.check_call([B._nvcc()] + B._host_compiler_flags() + ["-shared", "-o", lib] + objs + ["-lcudart"])
from bbgpu import capi
capi.LIB_PATH = lib
n = 262144
h = capi.EnvHandle(n, 42)
h.step_random(64)
stats = torch.zeros(64, dtype=torch.int64, device="cuda")
K = 1
for _ in range(K):
    h.step_random(1, None, None, None, None, stats)
torch.cuda.synchronize()
s = stats.cpu().tolist()
warps = n // 32 * K
print("per warp-step: total cycles %.0f | classify phase %.0f | team loop %.0f | team rounds %.2f | deal iterations %.2f"
      % (s[4] / warps, s[5] / warps, s[6] / warps, s[7] / warps, s[8] / warps))
print("max warp cycles %d, max rounds %d" % (s[9], s[10]))
print("per-warp cycle histogram (bins of 8192 cycles = 4.2 us):", s[16:48])
print("team rounds per warp histogram:", s[48:56])
for v in sorted(s[56:64], reverse=True):
    print("  heavy warp: %d cycles, unit loops %d cycles, %d unit iterations, %d rounds" % (v >> 32, ((v >> 16) & 0xFFFF) << 6, (v >> 8) & 0xFF, v & 0xFF))
R = max(s[7], 1)
print("per round: H items %.2f | unit-loop iterations %.1f (max %d) | open cycles %.0f | unit-loop cycles %.0f (%.0f per unit)"
      % (s[15] / R, s[11] / R, s[12], s[13] / R, s[14] / R, s[14] / max(s[11], 1)))
