"""Masked-PPO training loop on the device-resident path — the reference's
``scripts/train.py:train`` (lines 60-316) with the same config keys, loop shape, metric names
and checkpoint files, run by one process per GPU.

    python -m bbgpu.train --config cfg.yaml [--resume ckpt.pt] [--seed 42]
    torchrun --nproc-per-node 8 -m bbgpu.train --config cfg.yaml        (config 4 / 5)

Config quirks of the reference are honoured: ``num_epochs`` is read from the
``ppo`` section (both shipped YAMLs put it under ``training``, so the default 10 applies),
``batch_size`` from ``training``, the ``network`` / ``environment`` sections are ignored.
Keys of ours live under ``b200:`` (precision, reseed_on_reset, log_every_update).
"""
import argparse
import json
import os
import time

import numpy as np
import torch

from . import dist
from .logger import Logger, TensorBoardLogger, tensorboard_tags
from .ppo import PPOAgent, PPOConfig
from .rollout import RolloutBuffer
from .vec_env import VectorizedBlockBlastEnv


def load_config(path):
    import yaml
    with open(path) as f:
        return yaml.safe_load(f)


def set_seed(seed):
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def collect_rollout(vec_env, agent, buffer, obs, ep):
    """scripts/train.py:173-203: one rollout of buffer.buffer_size steps, all on the device.
    ``ep`` accumulates episode statistics as device scalars [count, sum score, max score, sum len]."""
    buffer.reset()
    for _ in range(buffer.buffer_size):
        actions, log_probs, values = agent.act(obs)
        buffer.add_obs(obs, actions, log_probs, values)
        obs, rewards, terminated, truncated, infos = vec_env.step(actions)
        buffer.add_outcome(rewards, terminated.float())
        t = terminated
        sc = torch.where(t, infos["ep_score"], torch.zeros_like(infos["ep_score"]))
        ep[0] += t.sum()
        ep[1] += sc.sum()
        ep[2] = torch.maximum(ep[2], sc.max())
        ep[3] += torch.where(t, infos["ep_len"], torch.zeros_like(infos["ep_len"])).sum()
    return obs


def train(config, resume_path=None, seed=42, progress_callback=None, max_updates=None):
    rank, world, local = dist.init()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    set_seed(seed + rank)
    ppo_c, tr_c = config.get("ppo", {}), config.get("training", {})
    rew_c, log_c, ours = config.get("rewards", {}), config.get("logging", {}), config.get("b200", {})
    paths = config.get("paths", {})
    ckpt_dir, log_dir = paths.get("checkpoint_dir", "checkpoints"), paths.get("log_dir", "logs")
    if rank == 0:
        os.makedirs(ckpt_dir, exist_ok=True)
        os.makedirs(log_dir, exist_ok=True)

    num_envs = tr_c.get("num_envs", 64)                       # whole job; sharded over ranks
    offset, n_local = dist.shard(num_envs)
    vec_env = VectorizedBlockBlastEnv(n_local, seed=seed, reward_config=rew_c or None, output="packed",
                                      global_env_offset=offset, reseed_on_reset=bool(ours.get("reseed_on_reset", False)))
    agent = PPOAgent(PPOConfig(
        learning_rate=ppo_c.get("learning_rate", 3e-4), gamma=ppo_c.get("gamma", 0.99),
        gae_lambda=ppo_c.get("gae_lambda", 0.95), clip_epsilon=ppo_c.get("clip_epsilon", 0.2),
        entropy_coef=ppo_c.get("entropy_coef", 0.01), value_coef=ppo_c.get("value_coef", 0.5),
        max_grad_norm=ppo_c.get("max_grad_norm", 0.5), num_epochs=ppo_c.get("num_epochs", 10),
        batch_size=tr_c.get("batch_size", 2048), precision=ours.get("precision", "fp32")), device)
    agent.train()                                             # train mode during rollout too (scripts/train.py:122)
    start_step = 0
    if resume_path and os.path.exists(resume_path):
        agent.load(resume_path)
        try:
            start_step = int(os.path.splitext(os.path.basename(resume_path))[0].split("_")[-1])
        except ValueError:
            pass
    rollout_steps = tr_c.get("rollout_steps", 128)
    buffer = RolloutBuffer(rollout_steps, n_local, device=device)
    total_timesteps = tr_c.get("total_timesteps", 50_000_000)
    log_interval, save_interval = log_c.get("log_interval", 100), log_c.get("save_interval", 1000)

    obs, _ = vec_env.reset()
    global_step, num_updates, best_score = start_step, 0, 0.0
    history = []
    # same files / keys / TensorBoard tags as scripts/train.py:136-137, 231-258 (rank 0 only)
    logger = Logger(log_dir, name="ppo_b200") if rank == 0 else None
    tb_logger = TensorBoardLogger(log_dir, name="ppo_b200") if rank == 0 and log_c.get("tensorboard", True) else None
    t0 = time.time()
    try:
        while global_step < total_timesteps:
            ep = [torch.zeros((), dtype=torch.int64, device=device) for _ in range(4)]
            obs = collect_rollout(vec_env, agent, buffer, obs, ep)
            global_step += num_envs * rollout_steps
            last_values = agent.values(obs)
            metrics = agent.update(buffer, last_values)
            num_updates += 1
            cnt, ssum, smax, lsum = dist.all_reduce_scalars([ep[0].item(), ep[1].item(), 0, ep[3].item()], device=device)
            smax = dist.all_reduce_scalars([ep[2].item()], op="max", device=device)[0]
            elapsed = time.time() - t0
            avg_score = ssum / cnt if cnt else 0.0
            row = {"step": global_step, "fps": (global_step - start_step) / max(elapsed, 1e-9), "avg_score": avg_score,
                   "max_score": smax, "best_score": max(best_score, avg_score), "avg_length": lsum / cnt if cnt else 0.0,
                   "episodes": cnt, "wall_s": elapsed, **metrics}
            history.append(row)
            if rank == 0:
                if avg_score > best_score:
                    best_score = avg_score
                    agent.save(os.path.join(ckpt_dir, "best.pt"))
                if num_updates % log_interval == 0 or num_updates <= 10 or ours.get("log_every_update"):
                    logger.log(row, global_step)
                    if tb_logger is not None:
                        tb_logger.log_metrics(tensorboard_tags(row), global_step)
                    print("update %d step=%d fps=%.0f avg_score=%.1f max=%d len=%.1f entropy=%.3f kl=%.4f clip=%.3f"
                          % (num_updates, global_step, row["fps"], avg_score, smax, row["avg_length"], metrics["entropy"],
                             metrics["approx_kl"], metrics["clip_fraction"]), flush=True)
                if num_updates % save_interval == 0:
                    agent.save(os.path.join(ckpt_dir, "checkpoint_%d.pt" % global_step))
                    agent.save(os.path.join(ckpt_dir, "latest.pt"))
            if progress_callback is not None and not progress_callback(
                    {"total_steps": global_step, "mean_score": avg_score, "best_score": best_score,
                     "episodes": cnt, "fps": row["fps"]}):
                break
            if max_updates and num_updates >= max_updates:
                break
    finally:
        if rank == 0:
            agent.save(os.path.join(ckpt_dir, "final.pt"))
            if logger.metrics_history:
                logger.save_summary()
            if tb_logger is not None:
                tb_logger.close()
        vec_env.close()
    return history


def main():
    ap = argparse.ArgumentParser(description="B200 masked-PPO training for Block Blast")
    ap.add_argument("--config", required=True)
    ap.add_argument("--resume", default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--max-updates", type=int, default=None)
    a = ap.parse_args()
    train(load_config(a.config), a.resume, a.seed, max_updates=a.max_updates)


if __name__ == "__main__":
    main()
