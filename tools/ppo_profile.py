#!/usr/bin/env python3
"""Where does a PPO minibatch step spend its time? (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bbgpu.ppo import PPOAgent, PPOConfig
from bbgpu.rollout import RolloutBuffer
from bbgpu.train import collect_rollout
from bbgpu.vec_env import VectorizedBlockBlastEnv
bench = "--bench" in sys.argv
torch.backends.cudnn.benchmark = bench
dev = torch.device("cuda")
n, T, mb = 65536, 8, 32768
agent = PPOAgent(PPOConfig(batch_size=mb, num_epochs=1, precision="bf16"), dev); agent.train()
venv = VectorizedBlockBlastEnv(n, seed=1, output="packed")
buf = RolloutBuffer(T, n, device=dev)
obs, _ = venv.reset()
ep = [torch.zeros((), dtype=torch.int64, device=dev) for _ in range(4)]
for it in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    obs = collect_rollout(venv, agent, buf, obs, ep)
    last = agent.values(obs)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    m = agent.update(buf, last)
    torch.cuda.synchronize(); t2 = time.perf_counter()
print("cudnn.benchmark=%s collect %.3f s  update %.3f s (%d minibatches of %d -> %.1f ms each)" % (bench, t1 - t0, t2 - t1, n * T // mb, mb, (t2 - t1) / (n * T // mb) * 1e3))
# breakdown of one minibatch
gen = buf.iter_minibatches(mb)
torch.cuda.synchronize(); t0 = time.perf_counter()
o, mk, a, lp, adv, ret = next(gen)
torch.cuda.synchronize(); t1 = time.perf_counter()
o = o.contiguous(memory_format=torch.channels_last)
with torch.autocast("cuda", dtype=torch.bfloat16):
    _, nlp, ent, v = agent.network.evaluate_actions(o, mk, a)
torch.cuda.synchronize(); t2 = time.perf_counter()
loss = -(nlp * adv).mean() + 0.5 * ((v.float() - ret) ** 2).mean() - 0.01 * ent.mean()
loss.backward()
torch.cuda.synchronize(); t3 = time.perf_counter()
torch.nn.utils.clip_grad_norm_(agent.network.parameters(), 0.5); agent.optimizer.step()
torch.cuda.synchronize(); t4 = time.perf_counter()
print("minibatch: gather+unpack %.1f ms | forward %.1f ms | backward %.1f ms | clip+adam %.1f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, nlp, ent, v = agent.network.evaluate_actions(o, mk, a)
    loss = -(nlp * adv).mean() + 0.5 * ((v.float() - ret) ** 2).mean() - 0.01 * ent.mean()
    loss.backward()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
