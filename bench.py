#!/usr/bin/env python3
"""bench.py — env-steps/s of the fused Block Blast step kernel (BASELINE.json config 3:
random valid-action policy, 262,144 envs per GPU), with roofline, CPU baseline and the
end-to-end number through the drop-in API.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = ONE launch of the fused K1 kernel over ONE batch of `--envs` environments
(policy action pick, placement, line clears, scoring, Philox trio regeneration with the
solvability search, game over, reward, auto-reset, next mask), reading and writing the full
packed protocol (129 algorithmic bytes per env-step).  To keep every launch HBM-cold the
bench rotates over `--batches` independent env batches whose combined footprint exceeds the
126 MB L2 (stated in config.l2).  Weak scaling: every rank owns its own batches; there is no
data-path collective (envs are independent), only the timing barrier.

The reference arm (--impl reference) times the CPU restatement of the reference's
VectorizedBlockBlastEnv(64) + sample_valid_actions loop (oracle/bb_oracle.py, same cell-grid
algorithm and serial per-env Python loop as the reference; the Python reference itself cannot
travel to the GPU box) on all host cores (one process per core).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NCU_DRAM_BYTES_PER_LAUNCH = 12.63e6 + 0.05e6   # measured by ncu for one 262,144-env launch (see traffic_source)
ALGO_BYTES_PER_ENV_STEP = 129          # SURVEY.md §8d / DESIGN.md: 48 R + 48 W + 4 + 4 + 1 + 24
STATE_BYTES, OUT_BYTES = 48, 33
METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arms
def _py_port_worker(args):
    """One process: the Python oracle port of VectorizedBlockBlastEnv(64) + sample_valid_actions
    (method of the reference's scripts/benchmark.py:101-144)."""
    seed, n_envs, seconds, min_steps = args
    import numpy as np
    from oracle import bb_oracle as O
    rng = np.random.RandomState(seed)
    envs = [O.Env(seed=seed * 1000 + i, rng_factory=O.numpy_rng_factory) for i in range(n_envs)]
    vec = O.VecEnv(envs)
    vec.reset()
    for _ in range(2):
        vec.step(vec.sample_valid_actions(rng))
    t0 = time.perf_counter()
    steps = 0
    while steps < min_steps or time.perf_counter() - t0 < seconds:
        vec.step(vec.sample_valid_actions(rng))
        steps += 1
    return steps * n_envs, time.perf_counter() - t0


def cpu_python_port(n_procs, seconds, n_envs=64, min_steps=1):
    import multiprocessing as mp
    if n_procs == 1:
        res = [_py_port_worker((0, n_envs, seconds, min_steps))]
    else:
        with mp.get_context("fork").Pool(n_procs) as pool:
            res = pool.map(_py_port_worker, [(k, n_envs, seconds, min_steps) for k in range(n_procs)])
    return sum(r[0] for r in res) / max(r[1] for r in res), sum(r[0] for r in res)


def cpu_c_port(n_threads, seconds, n_envs=4096):
    """The plain-C oracle port, random-valid rollout, OpenMP over envs."""
    import numpy as np
    from bbgpu import philox
    from oracle import bb_oracle_c as OC
    chunk = 64
    streams = philox.candidate_trios(42, np.arange(n_envs), 4096)
    env = OC.CVecEnv(streams, n_threads=n_threads)
    rs = np.random.RandomState(0)
    words = rs.randint(0, 2 ** 32, size=(chunk, n_envs), dtype=np.uint64).astype(np.uint32)
    env.random_rollout(8, words)
    t0 = time.perf_counter()
    done = 0
    while time.perf_counter() - t0 < seconds:
        d, _, _ = env.random_rollout(chunk, words)
        done += d
        if env.stats()[:, 7].max() > 3500:      # candidate streams nearly used up: restart them
            env = OC.CVecEnv(streams, n_threads=n_threads)
    return done / (time.perf_counter() - t0), done


def cpu_ppo_port(seconds_hint=20.0):
    """PPO samples/s of the reference's schedule on the host CPU (config 1: 64 envs x 128 steps,
    10 epochs x 4 minibatches of 2048), from TIMED COMPONENTS: the Python oracle env step, one
    64-sample CNN forward and one 2048-sample forward+backward+Adam step of the same
    5.29 M-parameter network in torch on the CPU.  A full iteration takes ~2 minutes on 8 cores
    (BASELINE.md), so the components are timed and the schedule's total is computed."""
    import numpy as np
    import torch
    from bbgpu.network import BlockBlastNetwork
    from oracle import bb_oracle as O
    torch.manual_seed(0)
    net = BlockBlastNetwork()
    net.train()
    opt = torch.optim.Adam(net.parameters(), lr=3e-4, eps=1e-5)
    envs = [O.Env(seed=100 + i, rng_factory=O.numpy_rng_factory) for i in range(64)]
    vec = O.VecEnv(envs)
    obs, _ = vec.reset()
    rng = np.random.RandomState(0)
    t0 = time.perf_counter()
    n_env = 0
    while time.perf_counter() - t0 < seconds_hint * 0.25:
        obs, *_ = vec.step(vec.sample_valid_actions(rng))
        n_env += 1
    t_env = (time.perf_counter() - t0) / n_env
    b, p = torch.from_numpy(obs["board"]), torch.from_numpy(obs["pieces"])
    with torch.no_grad():
        net.forward(b, p)
        t0 = time.perf_counter()
        for _ in range(3):
            net.forward(b, p)
        t_fwd = (time.perf_counter() - t0) / 3
    x = torch.rand(2048, 4, 8, 8).round()
    m = torch.ones(2048, 192)
    a = torch.zeros(2048, dtype=torch.long)
    t_upd = []
    for _ in range(2):
        t0 = time.perf_counter()
        _, lp, ent, v = net.evaluate_actions(x, m, a)
        loss = -(lp.mean()) + 0.5 * (v ** 2).mean() - 0.01 * ent.mean()
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 0.5)
        opt.step()
        t_upd.append(time.perf_counter() - t0)
    t_mb = min(t_upd)
    iteration = 128 * (t_env + t_fwd) + 40 * t_mb
    return 8192 / iteration, dict(env_vec_step_s=t_env, fwd64_s=t_fwd, minibatch2048_step_s=t_mb,
                                  iteration_s=iteration, torch_threads=torch.get_num_threads())


def kernel_rooflines(dev, peak):
    """CUDA-event timings of the companion kernels at config-3/4 sizes, inputs larger than L2;
    achieved = algorithmic bytes / time."""
    import torch
    from bbgpu import capi
    out = {}

    def timeit(fn, iters=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e-3 / iters

    # context for the write-dominated K2: what a pure-write stream (torch fill_) reaches on this GPU
    big = torch.empty(256 * 1024 * 1024, dtype=torch.float32, device=dev)
    t = timeit(lambda: big.fill_(1.0), iters=10)
    out["write_only_fill_GBs"] = big.numel() * 4 / t / 1e9
    del big
    # K4 GAE: T=128 x N=262,144, 20 B/sample (+4N last values) = 671 MB per launch
    T, N = 128, 262144
    r, v = torch.randn(T, N, device=dev), torch.randn(T, N, device=dev)
    d = (torch.rand(T, N, device=dev) < 0.07).float()
    lv = torch.randn(N, device=dev)
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    mom = torch.zeros(2, dtype=torch.float64, device=dev)
    t = timeit(lambda: capi.gae(r, v, d, lv, 0.99, 0.95, adv, ret, mom))
    b = 20 * T * N + 4 * N
    out["K4_gae"] = {"samples": T * N, "us": t * 1e6, "algorithmic_bytes": b, "achieved_GBs": b / t / 1e9,
                     "frac_of_hbm_peak": b / t / 1e9 / peak, "samples_per_sec": T * N / t}
    del r, v, d, adv, ret
    # K3 masked sample: n=524,288 rows of f32 logits: 804 B/row = 421 MB per launch
    n = 524288
    logits = torch.randn(n, 192, device=dev)
    mask = torch.randint(-2 ** 62, 2 ** 62, (3, n), dtype=torch.int64, device=dev) | 1
    act = torch.empty(n, dtype=torch.int32, device=dev)
    lp, en = torch.empty(n, device=dev), torch.empty(n, device=dev)
    t = timeit(lambda: capi.masked_sample(logits, mask, n, 1, 1, 0, act, lp, en))
    b = 804 * n
    out["K3_masked_sample_f32"] = {"rows": n, "us": t * 1e6, "algorithmic_bytes": b, "achieved_GBs": b / t / 1e9,
                                   "frac_of_hbm_peak": b / t / 1e9 / peak}
    lb = logits.bfloat16()
    t = timeit(lambda: capi.masked_sample(lb, mask, n, 1, 1, 0, act, lp, en))
    b = 420 * n
    out["K3_masked_sample_bf16"] = {"rows": n, "us": t * 1e6, "algorithmic_bytes": b, "achieved_GBs": b / t / 1e9,
                                    "frac_of_hbm_peak": b / t / 1e9 / peak}
    # K3 backward (PPO update path): 768 B logits read + 768 B grad written + 36 B mask/action/grads per row
    gl = torch.empty_like(logits)
    g1, g2 = torch.randn(n, device=dev), torch.randn(n, device=dev)
    t = timeit(lambda: capi.masked_head_backward(logits, mask, n, act, g1, g2, gl))
    b = (768 * 2 + 36) * n
    out["K3_backward_f32"] = {"rows": n, "us": t * 1e6, "algorithmic_bytes": b, "achieved_GBs": b / t / 1e9,
                              "frac_of_hbm_peak": b / t / 1e9 / peak}
    del logits, lb, gl
    # K2 obs unpack: 36 B in, 1,024 B f32 planes out (+ 192 B u8 mask) per env
    board = torch.randint(-2 ** 62, 2 ** 62, (n,), dtype=torch.int64, device=dev)
    pieces = torch.randint(0, 37, (n,), dtype=torch.int32, device=dev) * 0x010101
    obs = torch.empty((n, 4, 8, 8), device=dev)
    t = timeit(lambda: capi.unpack_obs(board, pieces, mask, n, obs=obs, n=n))
    b = (12 + 1024) * n
    out["K2_unpack_obs_f32"] = {"envs": n, "us": t * 1e6, "algorithmic_bytes": b, "achieved_GBs": b / t / 1e9,
                                "frac_of_hbm_peak": b / t / 1e9 / peak}
    obs16 = torch.empty((n, 4, 8, 8), dtype=torch.bfloat16, device=dev)
    t = timeit(lambda: capi.unpack_obs(board, pieces, mask, n, obs=obs16, n=n))
    b = (12 + 512) * n
    out["K2_unpack_obs_bf16"] = {"envs": n, "us": t * 1e6, "algorithmic_bytes": b, "achieved_GBs": b / t / 1e9,
                                 "frac_of_hbm_peak": b / t / 1e9 / peak}
    for k in ("K2_unpack_obs_f32", "K2_unpack_obs_bf16"):
        out[k]["frac_of_write_only_fill"] = out[k]["achieved_GBs"] / out["write_only_fill_GBs"]
    del board, pieces, obs, obs16
    # BatchNorm + ReLU + residual add of the CNN (bf16 NHWC [rows, 128]): forward reads x twice and the
    # skip once, writes y; backward reads (x, y, dy) twice, writes dx and dskip
    rows, ch = 32768 * 64, 128
    mk = lambda: torch.randn(rows, ch, device=dev).to(torch.bfloat16)
    x, skip, y, dy, dx, dsk = mk(), mk(), mk(), mk(), mk(), mk()
    gamma, beta = torch.rand(ch, device=dev) + 0.5, torch.randn(ch, device=dev)
    rm, rv = torch.zeros(ch, device=dev), torch.ones(ch, device=dev)
    sm, sr = torch.empty(ch, device=dev), torch.empty(ch, device=dev)
    dg, db = torch.empty(ch, device=dev), torch.empty(ch, device=dev)
    ws = torch.empty(capi.bn_workspace_size(ch), device=dev)
    t = timeit(lambda: capi.bn_relu_forward(x, skip, gamma, beta, None, rm, rv, 0.1, 1e-5, True, y, sm, sr, ws, rows, ch))
    b = 4 * rows * ch * 2
    out["BN_relu_residual_forward_bf16"] = {"rows": rows, "channels": ch, "us": t * 1e6, "algorithmic_bytes": b,
                                            "achieved_GBs": b / t / 1e9, "frac_of_hbm_peak": b / t / 1e9 / peak}
    t = timeit(lambda: capi.bn_relu_backward(x, y, dy, gamma, sm, sr, dx, dsk, dg, db, ws, rows, ch))
    b = 8 * rows * ch * 2
    out["BN_relu_residual_backward_bf16"] = {"rows": rows, "channels": ch, "us": t * 1e6, "algorithmic_bytes": b,
                                             "achieved_GBs": b / t / 1e9, "frac_of_hbm_peak": b / t / 1e9 / peak}
    return out


def gpu_ppo_leg(rank, world, dev, n_envs, T, minibatch, epochs, precision, chunk):
    """Masked-PPO collect + GAE + update on the device-resident path (BASELINE config 4 shape:
    131,072 envs per GPU); returns samples/s and the phase times.  Two warm-up iterations (cuDNN
    autotuning of the conv shapes happens in the first, allocator growth in the second)."""
    import torch
    import torch.distributed as dist
    from bbgpu.ppo import PPOAgent, PPOConfig
    from bbgpu.rollout import RolloutBuffer
    from bbgpu.train import collect_rollout
    from bbgpu.vec_env import VectorizedBlockBlastEnv
    agent = PPOAgent(PPOConfig(batch_size=minibatch, num_epochs=epochs, precision=precision), dev)
    agent.train()
    venv = VectorizedBlockBlastEnv(n_envs, seed=42, output="packed", global_env_offset=(64 + rank) * n_envs)
    buf = RolloutBuffer(T, n_envs, device=dev)
    obs, _ = venv.reset()
    orig_act = agent.act
    agent.act = lambda o, deterministic=False: orig_act(o, deterministic, chunk)
    times = {}
    for it in range(3):
        ep = [torch.zeros((), dtype=torch.int64, device=dev) for _ in range(4)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        obs = collect_rollout(venv, agent, buf, obs, ep)
        last = agent.values(obs, chunk)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        metrics = agent.update(buf, last)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t2 = time.perf_counter()
        times = dict(collect_s=t1 - t0, update_s=t2 - t1, total_s=t2 - t0)
    tt = torch.tensor([times["total_s"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    venv.close()
    return dict(samples_per_sec=world * n_envs * T / float(tt.item()), envs_per_gpu=n_envs, rollout_steps=T,
                minibatch=minibatch, epochs=epochs, precision=precision, act_chunk=chunk, **times,
                entropy=metrics["entropy"], approx_kl=metrics["approx_kl"],
                network="BlockBlastNetwork 5,290,113 params: convs/linears cuDNN/cuBLAS via PyTorch; BatchNorm+ReLU(+residual) "
                        "= bb_bn_relu_* kernels, loss tail = bb_ppo_loss (bf16 path)",
                grad_allreduce="1 flat NCCL all-reduce of 21.2 MB per optimiser step" if world > 1 else "n/a (1 GPU)")


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_envs = 64
    per_step_s = 0.06                      # ~64 envs x ~1 ms per Python env-step
    budget = min(150.0, max(5.0, (args.steps + args.warmup) * per_step_s))
    # warm-up + timed steps folded into a time-bounded run per process
    t0 = time.perf_counter()
    value, total = cpu_python_port(cores, budget, n_envs=n_envs, min_steps=max(1, min(args.steps, 50)))
    wall = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_envs * cores / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": "random valid-action policy, VectorizedBlockBlastEnv(64) per process, "
                                   "%d processes (one per host core)" % cores,
                       "step": "one 64-env vec step per process (bounded sample of config 3)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d env-steps in %.1f s: oracle/bb_oracle.py (Python restatement of the "
                                       "reference's cell-grid engine, serial 64-env loop) x %d processes"
                                       % (total, wall, cores)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=262144, help="envs per batch (= per launch) per GPU")
    ap.add_argument("--batches", type=int, default=8, help="independent env batches rotated per GPU (L2-cold launches)")
    ap.add_argument("--preroll", type=int, default=64, help="untimed random steps to reach the steady-state mix")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 100)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ppo", action="store_true")
    ap.add_argument("--ppo-envs", type=int, default=131072, help="envs per GPU for the PPO leg (config 4: 1,048,576 / 8)")
    ap.add_argument("--ppo-steps", type=int, default=8)
    ap.add_argument("--ppo-minibatch", type=int, default=32768)
    ap.add_argument("--ppo-epochs", type=int, default=2)
    ap.add_argument("--ppo-precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    # stdout must carry exactly ONE JSON line: libraries (NCCL prints its version banner there)
    # get stderr; the line is written to the saved descriptor at the end
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from bbgpu import capi
    from bbgpu.vec_env import VectorizedBlockBlastEnv

    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    capi.lib()

    n, M, K, W = args.envs, args.batches, args.steps, max(3, args.warmup)
    seed = 42
    # independent batches; global env ids are disjoint across batches and ranks
    envs, outs = [], []
    for b in range(M):
        envs.append(capi.EnvHandle(n, seed, global_env_offset=(rank * M + b) * n))
        outs.append(dict(actions=torch.zeros(n, dtype=torch.int32, device=dev),
                         rewards=torch.zeros(n, dtype=torch.float32, device=dev),
                         term=torch.zeros(n, dtype=torch.uint8, device=dev),
                         mask=torch.zeros((3, n), dtype=torch.int64, device=dev)))
    stats = torch.zeros(4, dtype=torch.int64, device=dev)
    for e in envs:
        e.step_random(args.preroll)          # desynchronise episodes (all envs start in lockstep)
    torch.cuda.synchronize()

    def launch(k):
        b = k % M
        o = outs[b]
        envs[b].step_random(1, o["actions"], o["rewards"], o["term"], o["mask"], stats)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(W):
        launch(k)
    barrier()
    stats.zero_()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    ev0.record()
    for k in range(K):
        launch(W + k)
    ev1.record()
    barrier()
    wall = time.perf_counter() - t_wall0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    s = stats.cpu().tolist()
    assert s[0] == n * K, "kernel did not process the expected number of env-steps"
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    value = world * n * K / (ms_max * 1e-3)

    # L2-warm variant (single batch, state+outputs 21 MB stay in L2): reported beside, not as value
    for k in range(20):
        envs[0].step_random(1, outs[0]["actions"], outs[0]["rewards"], outs[0]["term"], outs[0]["mask"], stats)
    barrier()
    ev0.record()
    kw = min(K, 1000)
    for k in range(kw):
        envs[0].step_random(1, outs[0]["actions"], outs[0]["rewards"], outs[0]["term"], outs[0]["mask"], stats)
    ev1.record()
    barrier()
    ms_warm = ev0.elapsed_time(ev1)

    # fused rollout: ONE launch runs 256 steps with the env state in registers
    ev0.record()
    envs[0].step_random(256, None, None, None, None, stats)
    ev1.record()
    barrier()
    ms_fused = ev0.elapsed_time(ev1)

    # rollout kernel: one launch = 16 steps, EVERY step's actions/rewards/terminated/masks written ([16, N] arrays)
    S = 16
    RA = torch.zeros((S, n), dtype=torch.int32, device=dev); RR = torch.zeros((S, n), dtype=torch.float32, device=dev)
    RT = torch.zeros((S, n), dtype=torch.uint8, device=dev); RM = torch.zeros((S, 3, n), dtype=torch.int64, device=dev)
    for b in range(2):
        envs[b].rollout_random(S, RA, RR, RT, RM, stats)
    barrier()
    ev0.record()
    for k in range(32):
        envs[k % M].rollout_random(S, RA, RR, RT, RM, stats)
    ev1.record()
    barrier()
    ms_roll = ev0.elapsed_time(ev1)
    del RA, RR, RT, RM

    # ------------------------------------------------------------------ e2e through the drop-in API
    Ke = args.e2e_steps or min(K, 100)
    venv = VectorizedBlockBlastEnv(n, seed=seed, output="numpy", global_env_offset=(world * M + rank) * n,
                                   reuse_buffers=True)
    venv.reset()
    for _ in range(3):
        venv.step(venv.sample_valid_actions())
    barrier()
    t0 = time.perf_counter()
    n_term = 0
    for _ in range(Ke):
        a = venv.sample_valid_actions()                    # kernel + D2H 4 B/env (numpy actions, as the reference)
        obs, rew, term, trunc, infos = venv.step(a)        # H2D 4 B/env, K1, D2H packed obs 41 B/env
        n_term += int(np.count_nonzero(term))              # the caller reads the result (np.sum over bools is 10x slower)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * n * Ke / float(te.item())
    # same loop, but the caller touches the dense reference-layout observation every step
    # (board (N,8,8) f32, pieces (N,3,8,8) f32, action_mask (N,192) int8 expanded on the host)
    Kd = min(Ke, 5)
    t0 = time.perf_counter()
    for _ in range(Kd):
        obs, rew, term, trunc, infos = venv.step(venv.sample_valid_actions())
        _ = obs["board"], obs["pieces"], obs["action_mask"]
    e2e_dense = n * Kd / (time.perf_counter() - t0)
    h2d = 4 * n
    d2h = 4 * n + (4 + 1 + 8 + 4 + 24 + 4 + 4) * n

    ppo = None
    for e in envs:
        e.close()
    del outs
    torch.cuda.empty_cache()
    kernels = kernel_rooflines(dev, load_peaks()[0]) if rank == 0 else None
    torch.cuda.empty_cache()
    if not args.no_ppo:
        ppo = gpu_ppo_leg(rank, world, dev, args.ppo_envs, args.ppo_steps, args.ppo_minibatch, args.ppo_epochs,
                          args.ppo_precision, None)

    if rank == 0:
        peak, peak_src = load_peaks()
        per_launch_s = ms_max * 1e-3 / K
        achieved = ALGO_BYTES_PER_ENV_STEP * n / per_launch_s / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": "random valid-action policy, %d envs per GPU per launch (BASELINE config 3)" % n,
                       "envs_per_gpu_per_launch": n, "batches_rotated": M, "seed": seed,
                       "l2": "rotating %d independent env batches per GPU: %.0f MB of state+outputs > 126 MB L2, "
                             "every launch reads its state from HBM" % (M, M * n * (STATE_BYTES + OUT_BYTES) / 1e6),
                       "protocol": "packed: state 48 B R+W, action 4 B, reward 4 B, terminated 1 B, mask 24 B",
                       "parallelism": "env shards per GPU, no data-path collective"},
            "gpu_launches": K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH,
                         "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch "
                                           "(profiles/r1_k1_final_ncu_summary.txt); reads are the 12.6 MB of state, the "
                                           "21 MB of outputs + state write-back were still in the 126 MB L2 when the capture ended",
                         "bound_note": "integer-issue bound, not HBM bound: ALU pipe 63% busy while an SM is active, SMs active 78% "
                                       "of the launch, 28.9 M warp instructions, 21.9 of 32 lanes active per instruction (ncu)",
                         "kernel": "bb_step_kernel<true>",
                         "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * n, "peak_source": peak_src,
                         "launch_us": per_launch_s * 1e6},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "dense_obs_materialised_env_steps_per_sec_rank0": e2e_dense, "api": "VectorizedBlockBlastEnv(output='numpy', reuse_buffers=True): sample_valid_actions() + "
                                        "step(actions); numpy results are zero-copy views of double-buffered pinned memory, "
                                        "packed obs expanded lazily on host"},
            "clocks": clocks,
            "extra": {"l2_warm_single_batch_env_steps_per_sec": n * kw / (ms_warm * 1e-3),
                      "fused_256_step_launch_env_steps_per_sec": n * 256 / (ms_fused * 1e-3),
                      "rollout16_all_outputs_env_steps_per_sec": n * 16 * 32 / (ms_roll * 1e-3),
                      "episodes": s[1], "mean_episode_len": (s[3] / s[1]) if s[1] else None,
                      "mean_final_score": (s[2] / s[1]) if s[1] else None, "wall_s_timed_region": wall},
        }
        line["kernels"] = kernels
        if ppo is not None:
            line["ppo"] = ppo
        if not args.no_cpu:
            cores = os.cpu_count() or 1
            v_py, tot_py = cpu_python_port(1, args.cpu_seconds)
            v_c, tot_c = cpu_c_port(cores, min(args.cpu_seconds, 10.0))
            line["cpu_baseline"] = {"value": v_py, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "%d env-steps: oracle/bb_oracle.py VecEnv(64) + sample_valid_actions, "
                                              "serial Python like the reference (wrappers.py:93-108)" % tot_py}
            line["cpu_baseline_c_port"] = {"value": v_c, "unit": UNIT, "cores": cores, "kind": "port",
                                           "sample": "%d env-steps: oracle/bb_oracle.c random-valid rollout, 4096 envs, "
                                                     "OpenMP over all host cores" % tot_c}
            if ppo is not None:
                v_ppo, parts = cpu_ppo_port(args.cpu_seconds)
                line["ppo"]["cpu_baseline"] = {"value": v_ppo, "unit": "samples/s", "cores": parts["torch_threads"],
                                               "kind": "port", "sample": "reference schedule (64 envs x 128 steps, 10 epochs x 4 x 2048) "
                                               "computed from timed components", **parts}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
