"""Timed loops over the REAL reference (oracle/_ref, see make_ref.py) for bench.py's CPU arm.
Test / benchmark infrastructure only — never imported by the product package.

env_loop:  VectorizedBlockBlastEnv(64, seed=42); reset(); actions = sample_valid_actions(); step(actions)
           — the method of the reference's scripts/benchmark.py:101-144.
ppo_iteration: the body of scripts/train.py:169-209 with PPOConfig(batch_size=2048, num_epochs=10)
           on the CPU: select_actions -> vec_env.step -> buffer.add per step, get_values, agent.update.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available():
    return os.path.isdir(os.path.join(REF, "src", "game"))


def _paths():
    for p in (os.path.join(REF, "_gym_stub"), os.path.join(REF, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)


def env_loop(args):
    """One process: (env-steps done, seconds).  args = (proc index, n_envs, seconds, min_steps)."""
    k, n_envs, seconds, min_steps = args
    _paths()
    import numpy as np
    from environment.wrappers import VectorizedBlockBlastEnv
    np.random.seed(k)
    vec = VectorizedBlockBlastEnv(num_envs=n_envs, seed=42 + 1000 * k)
    vec.reset()
    for _ in range(3):
        vec.step(vec.sample_valid_actions())
    t0 = time.perf_counter()
    steps = 0
    while steps < min_steps or time.perf_counter() - t0 < seconds:
        vec.step(vec.sample_valid_actions())
        steps += 1
    return steps * n_envs, time.perf_counter() - t0


def env_loop_all_cores(n_procs, seconds, n_envs=64, min_steps=1):
    """(env-steps/s over all processes, total env-steps)."""
    import multiprocessing as mp
    if n_procs == 1:
        res = [env_loop((0, n_envs, seconds, min_steps))]
    else:
        with mp.get_context("fork").Pool(n_procs) as pool:
            res = pool.map(env_loop, [(k, n_envs, seconds, min_steps) for k in range(n_procs)])
    return sum(r[0] for r in res) / max(r[1] for r in res), sum(r[0] for r in res)


def ppo_iteration(epochs_timed=1, threads=None):
    """The reference's PPO iteration on the CPU with the real classes: the full 128-step collect
    (select_actions -> vec_env.step -> buffer.add), get_values, and agent.update with ``epochs_timed`` of
    the schedule's 10 epochs (each 4 minibatches of 2,048); the full schedule's time is
    collect + 10 x (update time per epoch).  Returns (samples/s of the full schedule, dict of parts)."""
    _paths()
    import numpy as np
    import torch
    if threads:
        torch.set_num_threads(int(threads))
    from agents.ppo import PPOAgent, PPOConfig, RolloutBuffer
    from environment.wrappers import VectorizedBlockBlastEnv
    np.random.seed(0)
    torch.manual_seed(0)
    n_envs, T = 64, 128
    vec = VectorizedBlockBlastEnv(num_envs=n_envs, seed=42)
    agent = PPOAgent(config=PPOConfig(batch_size=2048, num_epochs=epochs_timed), device=torch.device("cpu"))
    agent.train()
    buf = RolloutBuffer(buffer_size=T, num_envs=n_envs)
    obs, _ = vec.reset()
    t0 = time.perf_counter()
    for step in range(T):
        actions, log_probs, values = agent.select_actions(obs)
        next_obs, rewards, terminated, truncated, infos = vec.step(actions)
        dones = np.logical_or(terminated, truncated)
        buf.add(board=obs["board"], pieces=obs["pieces"], action_mask=obs["action_mask"], action=actions,
                log_prob=log_probs, reward=rewards, done=dones.astype(np.float32), value=values)
        obs = next_obs
    last_values = agent.get_values(obs)
    t_collect = time.perf_counter() - t0
    t0 = time.perf_counter()
    metrics = agent.update(buf, last_values)
    t_update = time.perf_counter() - t0
    per_epoch = t_update / epochs_timed
    iteration = t_collect + 10 * per_epoch
    return n_envs * T / iteration, dict(collect_s=t_collect, epoch_s=per_epoch, iteration_s=iteration,
                                       epochs_timed=epochs_timed, torch_threads=torch.get_num_threads(),
                                       entropy=float(metrics.get("entropy", 0.0)),
                                       clip_fraction=float(metrics.get("clip_fraction", 0.0)))
