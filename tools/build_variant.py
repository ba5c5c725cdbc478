#!/usr/bin/env python3
"""Build a tuning variant of libbbgpu.so: tools/build_variant.py OUT.so -DX=Y ..."""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bbgpu import build as B
out, defs = sys.argv[1], sys.argv[2:]
tmp = tempfile.mkdtemp()
objs = []
for src in B.SOURCES:
    o = os.path.join(tmp, src.replace(".cu", ".o"))
    r = subprocess.run([B._nvcc()] + B._host_compiler_flags() + B.NVCC_FLAGS + defs + ["-c", os.path.join(B.CSRC, src), "-o", o],
                       capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr)
    for line in r.stderr.splitlines():
        if "bb_step_kernel" in line and "Compiling" in line:
            tag = line.split("'")[1][:24]
        if "Used" in line and src == "bb_env_kernels.cu":
            print(out, defs, line.strip())
    objs.append(o)
subprocess.check_call([B._nvcc()] + B._host_compiler_flags() + ["-shared", "-o", out] + objs + ["-lcudart"])
