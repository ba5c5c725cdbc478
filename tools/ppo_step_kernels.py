#!/usr/bin/env python3
"""GPU time per kernel of ONE PPO minibatch step (gather -> CNN fwd/bwd -> loss tail -> clip -> Adam)
at a given minibatch size, from torch.profiler (GPU box).
    python tools/ppo_step_kernels.py [minibatch=2048] [envs=64] [T=128]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from bbgpu.ppo import PPOAgent, PPOConfig
from bbgpu.rollout import RolloutBuffer
from bbgpu.train import RolloutRunner
from bbgpu.vec_env import VectorizedBlockBlastEnv

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 128
agent = PPOAgent(PPOConfig(batch_size=mb, num_epochs=1, precision="bf16"), seed=1)
agent.train()
venv = VectorizedBlockBlastEnv(n, seed=1, output="packed")
buf = RolloutBuffer(T, n)
run = RolloutRunner(venv, agent, buf, use_graph=False)
lv = run.run()
agent.update(buf, lv)                                    # warm-up (cuDNN autotune, optimizer state)
mean, std = buf.advantage_mean_std()
ms = torch.stack([mean, std]).float()
idx = torch.randperm(n * T, device="cuda")[:mb]
sums = torch.zeros(6, dtype=torch.float64, device="cuda")
for _ in range(3):
    g = buf.gather(idx, ms)
    agent._fused_step(g["obs"], g["mask"], g["actions"], g["logp"], g["adv"], g["ret"], sums)
torch.cuda.synchronize()
reps = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(reps):
        g = buf.gather(idx, ms)
        agent._fused_step(g["obs"], g["mask"], g["actions"], g["logp"], g["adv"], g["ret"], sums)
    torch.cuda.synchronize()
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev)
print("minibatch %d: %.3f ms of GPU kernel time per step, %d kernel launches per step" % (mb, tot / reps / 1e3, sum(e.count for e in ev) // reps))
for e in sorted(ev, key=lambda e: -e.device_time_total)[:28]:
    print("%6.1f%% %8.1f us x%3d  %s" % (100 * e.device_time_total / tot, e.device_time_total / reps, e.count // reps, e.key[:110]))
