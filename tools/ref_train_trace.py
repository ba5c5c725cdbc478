#!/usr/bin/env python3
"""Run the UNMODIFIED reference's training loop body (scripts/train.py:169-209) on the CPU for a few
updates and print the per-update metrics — the learning-dynamics yardstick for bbgpu.train
(build container only: reads /root/reference or oracle/_ref).

    python tools/ref_train_trace.py [updates] [seed]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_runner  # noqa: E402

ref_runner._paths()
import numpy as np  # noqa: E402
import torch  # noqa: E402
from agents.ppo import PPOAgent, PPOConfig, RolloutBuffer  # noqa: E402
from environment.wrappers import VectorizedBlockBlastEnv  # noqa: E402

updates = int(sys.argv[1]) if len(sys.argv) > 1 else 5
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 42
np.random.seed(seed)
torch.manual_seed(seed)
n_envs, T = 64, 128
vec = VectorizedBlockBlastEnv(num_envs=n_envs, seed=seed)
agent = PPOAgent(config=PPOConfig(batch_size=2048, num_epochs=10), device=torch.device("cpu"))
agent.train()
buf = RolloutBuffer(buffer_size=T, num_envs=n_envs)
obs, _ = vec.reset()
t0 = time.time()
for u in range(updates):
    buf.reset()
    scores, lens = [], []
    for step in range(T):
        actions, log_probs, values = agent.select_actions(obs)
        next_obs, rewards, terminated, truncated, infos = vec.step(actions)
        dones = np.logical_or(terminated, truncated)
        buf.add(board=obs["board"], pieces=obs["pieces"], action_mask=obs["action_mask"], action=actions,
                log_prob=log_probs, reward=rewards, done=dones.astype(np.float32), value=values)
        for term, info in zip(terminated, infos):
            if term:
                scores.append(info.get("final_score", info.get("score", 0)))
                lens.append(info.get("moves", 0))
        obs = next_obs
    m = agent.update(buf, agent.get_values(obs))
    row = {"update": u + 1, "step": (u + 1) * n_envs * T, "wall_s": time.time() - t0, "avg_score": float(np.mean(scores)),
           "avg_length": float(np.mean(lens)), "episodes": len(scores), **{k: float(v) for k, v in m.items()}}
    print(json.dumps(row), flush=True)
