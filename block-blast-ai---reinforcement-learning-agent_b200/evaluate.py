"""Batched evaluation of a trained agent — the reference's ``scripts/evaluate.py:23-90``
(one env, ``deterministic=True`` argmax rollouts) run for many episodes at once on the GPU.

Each of ``num_episodes`` envs plays exactly one episode (``BB_ENV_NO_AUTO_RESET``: a finished env
stays GAME_OVER and its further steps are no-ops), so the returned statistics are per-episode
like the reference's: scores, lengths, lines cleared, max combo.

    python -m bbgpu.evaluate --checkpoint checkpoints/best.pt --episodes 1000
"""
import argparse
import json

import numpy as np
import torch

from . import capi
from .ppo import PPOAgent


@torch.no_grad()
def evaluate(agent, num_episodes=100, seed=0, deterministic=True, max_steps=10000, reward_config=None,
             return_actions=False):
    """``return_actions``: also return ``actions`` int32 [steps, n] (the action every env was given at every
    step; entries after an env's game over are ignored by the env) and ``rewards_per_episode`` — enough to
    replay every episode through another implementation of the rules (tests do, through the oracle)."""
    dev = agent.device
    n = int(num_episodes)
    h = capi.EnvHandle(n, seed, 0, reward_config, capi.ENV_NO_AUTO_RESET)
    board = torch.zeros(n, dtype=torch.int64, device=dev)
    pieces = torch.zeros(n, dtype=torch.int32, device=dev)
    mask = torch.zeros((3, n), dtype=torch.int64, device=dev)
    rewards = torch.zeros(n, dtype=torch.float32, device=dev)
    term = torch.zeros(n, dtype=torch.uint8, device=dev)
    total_reward = torch.zeros(n, dtype=torch.float64, device=dev)
    done = torch.zeros(n, dtype=torch.bool, device=dev)
    was_training = agent.training
    agent.eval()                                       # scripts/evaluate.py evaluates in eval mode
    steps = 0
    log = []
    while steps < max_steps and not bool(done.all()):
        h.observe(board, pieces, mask)
        act, _, _ = agent.act({"board": board, "pieces": pieces, "mask": mask}, deterministic=deterministic)
        if return_actions:
            log.append(act.clone())
        h.step(act, rewards, term, None, None, None, None)
        total_reward += torch.where(done, torch.zeros_like(rewards), rewards).double()
        done |= term.bool()
        steps += 1
    st = h.get_state()
    h.close()
    if was_training:
        agent.train()
    scores, lengths = st["score"].astype(np.int64), st["moves"].astype(np.int64)
    extra = {}
    if return_actions:
        extra = {"actions": torch.stack(log).cpu().numpy() if log else np.zeros((0, n), np.int32),
                 "rewards_per_episode": total_reward.cpu().numpy(), "lines": st["lines_total"].astype(np.int64),
                 "max_combos": st["max_streak"].astype(np.int64)}
    return {**extra, "num_episodes": n, "mean_score": float(scores.mean()), "std_score": float(scores.std()),
            "max_score": int(scores.max()), "min_score": int(scores.min()), "mean_length": float(lengths.mean()),
            "max_length": int(lengths.max()), "mean_lines": float(st["lines_total"].mean()),
            "max_combo": int(st["max_streak"].max()), "mean_reward": float(total_reward.mean().item()),
            "finished": int(done.sum().item()), "scores": scores, "lengths": lengths}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", required=True)
    ap.add_argument("--episodes", type=int, default=100)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--stochastic", action="store_true")
    a = ap.parse_args()
    agent = PPOAgent()
    agent.load(a.checkpoint)
    res = evaluate(agent, a.episodes, a.seed, deterministic=not a.stochastic)
    print(json.dumps({k: v for k, v in res.items() if not isinstance(v, np.ndarray)}, indent=1))


if __name__ == "__main__":
    main()
