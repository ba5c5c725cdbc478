#!/usr/bin/env python3
"""Seconds per PPO iteration (collect + update) of a given schedule, eager vs CUDA graphs (GPU box).
    python tools/train_iter_time.py [num_envs] [T] [minibatch] [epochs] [precision]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bbgpu import dist as bbdist
from bbgpu.ppo import PPOAgent, PPOConfig
from bbgpu.rollout import RolloutBuffer
from bbgpu.train import RolloutRunner
from bbgpu.vec_env import VectorizedBlockBlastEnv

n, T, mb, ep = (int(x) for x in (sys.argv[1:5] + ["64", "128", "2048", "10"][len(sys.argv) - 1:])[:4])
prec = sys.argv[5] if len(sys.argv) > 5 else "bf16"
rank, world, local = bbdist.init()          # under torchrun: data-parallel ranks, gradient all-reduce per step
torch.cuda.set_device(local)
for graph in ((True,) if world > 1 else (False, True)):
    agent = PPOAgent(PPOConfig(batch_size=mb, num_epochs=ep, precision=prec), seed=1, global_env_offset=rank * n)
    if os.environ.get("BB_NO_EARLY_ALLREDUCE"):
        agent.bucket.split = None                      # A/B: one all-reduce after backward instead of two slices
    agent.train()
    venv = VectorizedBlockBlastEnv(n, seed=1, output="packed", global_env_offset=rank * n)
    buf = RolloutBuffer(T, n)
    run = RolloutRunner(venv, agent, buf, use_graph=graph)
    for it in range(6):
        if it == 3:
            torch.cuda.synchronize()
            tc = tu = 0.0
        t0 = time.perf_counter()
        lv = run.run()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        m = agent.update(buf, lv, use_graph=graph)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if it >= 3:
            tc += t1 - t0
            tu += t2 - t1
    if rank == 0:
      print("world=%d early_allreduce=%s NCCL_ALGO=%s NCCL_PROTO=%s " % (world, not os.environ.get("BB_NO_EARLY_ALLREDUCE"),
            os.environ.get("NCCL_ALGO"), os.environ.get("NCCL_PROTO")), end="")
      print("envs=%d T=%d mb=%d epochs=%d %s graph=%s: collect %.1f ms, update %.1f ms, %.0f samples/s  (entropy %.3f)"
          % (n, T, mb, ep, prec, graph, tc / 3 * 1e3, tu / 3 * 1e3, n * T * 3 / (tc + tu), m["entropy"]), flush=True)
    venv.close()
