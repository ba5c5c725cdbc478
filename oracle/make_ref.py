#!/usr/bin/env python3
"""Recipe for oracle/_ref: a runnable copy of the UNMODIFIED reference for the CPU arm of bench.py.

    python oracle/make_ref.py            (build container; /root/reference must be mounted)

The reference is pure Python + numpy + torch, so "building" it is copying its ``src/`` tree
(the files on the hot path: game/, environment/, agents/, models/, utils/) to ``oracle/_ref/src``
and placing the 15-line ``gymnasium`` import stub (tests/golden/_gym_stub — the reference only
subclasses ``gym.Env`` and declares spaces; gymnasium is absent from the image) beside it.
``oracle/_ref/`` is git-ignored (no reference source ever enters the history) but NOT
gpurun-ignored, so it travels to the GPU box like the built ``.so`` files; bench.py's
``--impl reference`` arm and ``cpu_baseline`` leg then time the real
``VectorizedBlockBlastEnv(64, seed=42)`` loop (scripts/benchmark.py:101-144) and the real
``PPOAgent.select_actions / update`` (scripts/train.py:169-209) on the box's host cores
(``kind: "reference"``), and fall back to the oracle port (``kind: "port"``) only when this
directory is absent.  Test infrastructure: nothing in the product package imports it.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BB_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def main():
    src = os.path.join(REF, "src")
    if not os.path.isdir(src):
        print("make_ref: %s not found; oracle/_ref left as it is" % src)
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    shutil.copytree(src, os.path.join(DST, "src"), ignore=ignore)
    shutil.copytree(os.path.join(os.path.dirname(HERE), "tests", "golden", "_gym_stub"), os.path.join(DST, "_gym_stub"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    with open(os.path.join(DST, "PROVENANCE.txt"), "w") as f:
        f.write("copied unmodified from %s by oracle/make_ref.py; git-ignored; CPU arm of bench.py only\n" % src)
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    print("make_ref: %d files -> %s" % (n, DST))
    return 0


if __name__ == "__main__":
    sys.exit(main())
