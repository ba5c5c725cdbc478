"""GPU parity tests for K1 (fused env step) through the C ABI, against the oracle and the
golden traces recorded from the reference.  Bit-exact: boards, 3x64-bit masks, float32 reward
bit patterns, done flags, scores, episode stats, candidate-trio consumption."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def _dev_buffers(torch, n):
    dev = "cuda"
    return dict(actions=torch.zeros(n, dtype=torch.int32, device=dev), rewards=torch.zeros(n, dtype=torch.float32, device=dev),
                term=torch.zeros(n, dtype=torch.uint8, device=dev), mask=torch.zeros((3, n), dtype=torch.int64, device=dev),
                ep_score=torch.full((n,), -1, dtype=torch.int32, device=dev), ep_len=torch.full((n,), -1, dtype=torch.int32, device=dev),
                info=torch.zeros(n, dtype=torch.int32, device=dev), board=torch.zeros(n, dtype=torch.int64, device=dev),
                pieces=torch.zeros(n, dtype=torch.int32, device=dev))


def _gpu_step(torch, h, B, actions):
    B["actions"].copy_(torch.from_numpy(np.ascontiguousarray(actions, np.int32)))
    B["ep_score"].fill_(-1)
    B["ep_len"].fill_(-1)
    h.step(B["actions"], B["rewards"], B["term"], B["mask"], B["ep_score"], B["ep_len"], B["info"])
    h.observe(B["board"], B["pieces"], None)
    torch.cuda.synchronize()
    return dict(rewards=B["rewards"].cpu().numpy(), terminated=B["term"].cpu().numpy(),
                mask=B["mask"].cpu().numpy().view(np.uint64).T.copy(), ep_score=B["ep_score"].cpu().numpy(),
                ep_len=B["ep_len"].cpu().numpy(), info=B["info"].cpu().numpy().view(np.uint32),
                board=B["board"].cpu().numpy().view(np.uint64), pieces=B["pieces"].cpu().numpy().view(np.uint32))


def _pieces4(p):
    return np.stack([p & 0xFF, (p >> 8) & 0xFF, (p >> 16) & 0xFF, (p >> 24) & 0xFF], axis=1).astype(np.uint8)


@pytest.mark.parametrize("name", ["vec_trace.npz", "vec_trace_cfg.npz"])
def test_gpu_replays_reference_golden_trace(torch, name):
    """config 2 of BASELINE.json: fixed piece/action sequence, bit-exact vs the reference."""
    from bbgpu import capi
    tr = np.load(os.path.join(G, name))
    cfg = json.loads(str(tr["reward_cfg_json"]))
    n = tr["actions"].shape[1]
    h = capi.EnvHandle(n, int(tr["seed"]), 0, cfg)
    B = _dev_buffers(torch, n)
    h.observe(B["board"], B["pieces"], B["mask"])
    torch.cuda.synchronize()
    assert np.array_equal(B["board"].cpu().numpy().view(np.uint64), tr["board0"])
    assert np.array_equal(_pieces4(B["pieces"].cpu().numpy().view(np.uint32)), tr["pieces0"])
    assert np.array_equal(B["mask"].cpu().numpy().view(np.uint64).T, tr["mask0"])
    for t in range(tr["actions"].shape[0]):
        o = _gpu_step(torch, h, B, tr["actions"][t])
        assert np.array_equal(o["board"], tr["board"][t]), t
        assert np.array_equal(_pieces4(o["pieces"]), tr["pieces"][t]), t
        assert np.array_equal(o["mask"], tr["mask"][t]), t
        assert np.array_equal(o["rewards"].view(np.uint32), tr["rewards"][t].view(np.uint32)), t
        assert np.array_equal(o["terminated"].astype(bool), tr["terminated"][t]), t
        assert np.array_equal((o["info"] & 1).astype(bool), tr["invalid"][t]), t
        tt = tr["terminated"][t]
        assert np.array_equal(o["ep_score"][tt], tr["ep_score"][t][tt]), t
        assert np.array_equal(o["ep_len"][tt], tr["ep_len"][t][tt]), t
        if t % 250 == 0 or t == tr["actions"].shape[0] - 1:
            st = h.get_state()
            assert np.array_equal(st["score"], tr["score"][t]) and np.array_equal(st["streak"], tr["streak"][t])
            assert np.array_equal(st["moves"], tr["moves"][t]) and np.array_equal(st["lines_total"], tr["lines_total"][t])
            assert np.array_equal(st["max_streak"], tr["max_streak"][t]) and np.array_equal(st["blocks_total"], tr["blocks_total"][t])
            assert np.array_equal(st["draw_ctr"], tr["draws"][t])
            assert np.array_equal(st["aux"] & 0xFF, tr["holes"][t])
    h.close()


@pytest.mark.parametrize("reseed", [False, True])
def test_gpu_vs_c_oracle_many_envs(torch, reseed):
    from bbgpu import capi, philox
    from oracle import bb_oracle_c as OC
    n, T, seed, off = 4096, 120, 1234, 1000
    streams = philox.candidate_trios(seed, off + np.arange(n), 512)
    ora = OC.CVecEnv(streams, reseed=reseed, n_threads=8)
    h = capi.EnvHandle(n, seed, off, None, capi.ENV_RESEED_ON_RESET if reseed else 0)
    B = _dev_buffers(torch, n)
    rs = np.random.RandomState(0)
    _, _, m = ora.export()
    n_term = 0
    for t in range(T):
        # uniformly random valid action per env from the oracle's mask, ~2% garbage actions
        bits = np.unpackbits(m.view(np.uint8).reshape(n, 24), axis=1, bitorder="little")
        r = rs.rand(n, 192) * bits
        acts = r.argmax(axis=1).astype(np.int32)
        bad = rs.rand(n) < 0.02
        acts[bad] = rs.randint(-100, 300, bad.sum())
        oo = ora.step(acts)
        go = _gpu_step(torch, h, B, acts)
        assert np.array_equal(oo["board"], go["board"]), t
        assert np.array_equal(oo["pieces"], _pieces4(go["pieces"])), t
        assert np.array_equal(oo["mask"], go["mask"]), t
        assert np.array_equal(oo["rewards"].view(np.uint32), go["rewards"].view(np.uint32)), t
        assert np.array_equal(oo["terminated"], go["terminated"]), t
        assert np.array_equal(oo["invalid"], (go["info"] & 1).astype(np.uint8)), t
        tt = oo["terminated"].astype(bool)
        n_term += tt.sum()
        assert np.array_equal(oo["ep_score"][tt], go["ep_score"][tt]) and np.array_equal(oo["ep_len"][tt], go["ep_len"][tt])
        m = oo["mask"]
    st, os_ = h.get_state(), ora.stats()
    assert np.array_equal(st["score"], os_[:, 0]) and np.array_equal(st["moves"], os_[:, 2])
    if not reseed:
        assert np.array_equal(st["draw_ctr"], os_[:, 7])
    assert n_term > 1000 and not ora.exhausted()
    h.close()


def test_fused_random_policy_parity_and_multistep_equivalence(torch):
    from bbgpu import capi, philox
    from oracle import bb_oracle_c as OC
    n, T, seed = 1024, 96, 77
    streams = philox.candidate_trios(seed, np.arange(n), 512)
    ora = OC.CVecEnv(streams, n_threads=8)
    words = philox.policy_words(seed, np.arange(n), np.arange(T))
    h1 = capi.EnvHandle(n, seed)
    h2 = capi.EnvHandle(n, seed)
    B = _dev_buffers(torch, n)
    stats1 = torch.zeros(4, dtype=torch.int64, device="cuda")
    stats2 = torch.zeros(4, dtype=torch.int64, device="cuda")
    eps = score = length = 0
    for t in range(T):
        _, _, m = ora.export()
        bits = np.unpackbits(m.view(np.uint8).reshape(n, 24), axis=1, bitorder="little").astype(np.int64)
        nv = bits.sum(1)
        k = philox.mulhi32(words[t], nv).astype(np.int64)
        csum = np.cumsum(bits, axis=1)
        acts = (csum <= k[:, None]).sum(axis=1).astype(np.int32)   # index of the (k+1)-th set bit
        h1.step_random(1, B["actions"], B["rewards"], B["term"], B["mask"], stats1)
        torch.cuda.synchronize()
        assert np.array_equal(B["actions"].cpu().numpy(), acts), t
        oo = ora.step(acts)
        assert np.array_equal(oo["rewards"].view(np.uint32), B["rewards"].cpu().numpy().view(np.uint32)), t
        assert np.array_equal(oo["terminated"], B["term"].cpu().numpy()), t
        assert np.array_equal(oo["mask"], B["mask"].cpu().numpy().view(np.uint64).T), t
        tt = oo["terminated"].astype(bool)
        eps += tt.sum(); score += oo["ep_score"][tt].sum(); length += oo["ep_len"][tt].sum()
    h2.step_random(T, None, None, None, None, stats2)      # one launch, state in registers
    torch.cuda.synchronize()
    s1, s2 = h1.get_state(), h2.get_state()
    assert s1.tobytes() == s2.tobytes()
    assert stats1.cpu().tolist() == stats2.cpu().tolist() == [n * T, int(eps), int(score), int(length)]
    h1.close(); h2.close()


def test_shards_reproduce_the_single_device_run(torch):
    """trajectories depend on (seed, global env id, actions) only."""
    from bbgpu import capi
    n, seed, T = 2048, 5, 200
    whole = capi.EnvHandle(n, seed, 0)
    lo = capi.EnvHandle(n // 2, seed, 0)
    hi = capi.EnvHandle(n // 2, seed, n // 2)
    for h in (whole, lo, hi):
        h.step_random(T)
    torch.cuda.synchronize()
    w = whole.get_state()
    # policy_ctr/draw_ctr are per env, so the records must match field for field
    assert w[: n // 2].tobytes() == lo.get_state().tobytes()
    assert w[n // 2:].tobytes() == hi.get_state().tobytes()
    for h in (whole, lo, hi):
        h.close()


def test_state_roundtrip_and_host_step_equals_device_step(torch):
    from bbgpu import capi
    n, seed = 1000, 3          # deliberately not a multiple of the block size
    a = capi.EnvHandle(n, seed)
    a.step_random(37)
    rec = a.get_state()
    b = capi.EnvHandle(n, 999)
    b.set_state(rec)
    assert b.get_state().tobytes() == rec.tobytes()
    b.close()
    b = capi.EnvHandle(n, seed)     # same seed: same Philox streams from the restored counters
    b.set_state(rec)
    B = _dev_buffers(torch, n)
    a.observe(None, None, B["mask"])
    torch.cuda.synchronize()
    m = B["mask"].cpu().numpy().view(np.uint64).T
    bits = np.unpackbits(np.ascontiguousarray(m).view(np.uint8).reshape(n, 24), axis=1, bitorder="little")
    acts = (np.random.RandomState(0).rand(n, 192) * bits).argmax(1).astype(np.int32)
    go = _gpu_step(torch, a, B, acts)
    pin = lambda dt, shape=(n,): torch.zeros(shape, dtype=dt).pin_memory()
    ha, hr, ht = pin(torch.int32), pin(torch.float32), pin(torch.uint8)
    hb, hp, hm, hs, hl = pin(torch.int64), pin(torch.int32), pin(torch.int64, (3, n)), pin(torch.int32), pin(torch.int32)
    ha.numpy()[:] = acts
    hi_ = pin(torch.int32)
    b.step_host(ha, hr, ht, hb, hp, hm, hs, hl, hi_)
    assert np.array_equal(hi_.numpy().view(np.uint32), go["info"])
    assert np.array_equal(hr.numpy().view(np.uint32), go["rewards"].view(np.uint32))
    assert np.array_equal(ht.numpy(), go["terminated"])
    assert np.array_equal(hb.numpy().view(np.uint64), go["board"])
    assert np.array_equal(hp.numpy().view(np.uint32), go["pieces"])
    assert np.array_equal(hm.numpy().view(np.uint64).T, go["mask"])
    assert a.get_state().tobytes() == b.get_state().tobytes()
    a.close(); b.close()


def test_no_auto_reset_single_env_semantics(torch):
    """BlockBlastEnv (no wrapper): after game over every action is rejected with -10."""
    from bbgpu import capi
    n = 256
    h = capi.EnvHandle(n, 11, 0, None, capi.ENV_NO_AUTO_RESET)
    B = _dev_buffers(torch, n)
    done = np.zeros(n, bool)
    for t in range(120):
        h.step_random(1, B["actions"], B["rewards"], B["term"], B["mask"], None)
        torch.cuda.synchronize()
        r, tm = B["rewards"].cpu().numpy(), B["term"].cpu().numpy().astype(bool)
        assert (r[done] == -10.0).all() and not tm[done].any()
        assert (B["mask"].cpu().numpy()[:, done] == 0).all()
        done |= tm
    assert done.sum() > n // 2
    st = h.get_state()
    assert ((st["aux"] >> 16) & 1).astype(bool).tolist() == done.tolist()
    h.close()


def test_full_size_properties_262144_envs(torch):
    """BASELINE config 3 size: invariants that need no oracle."""
    from bbgpu import capi
    import hostrules as H
    n, T = 262144, 60
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    h = capi.EnvHandle(n, 42)
    for _ in range(T):
        h.step_random(1, None, None, None, None, stats)
    torch.cuda.synchronize()
    s = stats.cpu().tolist()
    assert s[0] == n * T and s[1] > 0
    mean_len = s[3] / s[1]
    assert 8 < mean_len < 25, mean_len          # random play: ~14.7 moves per episode
    st = h.get_state()
    b = np.ascontiguousarray(st["board"])
    # no full row / column survives a step
    rows = b.view(np.uint8).reshape(n, 8)
    assert not (rows == 0xFF).any()
    cols = np.bitwise_and.reduce(rows, axis=1)
    assert not cols.any()
    # stored holes / center counters equal a recomputation
    L = H.lib()
    idx = np.random.RandomState(0).choice(n, 2000, replace=False)
    for i in idx:
        assert L.bbh_holes(int(b[i])) == int(st["aux"][i] & 0xFF)
        assert L.bbh_center(int(b[i])) == int((st["aux"][i] >> 8) & 0xFF)
    # every env has a legal move (auto-reset) and unused pieces only
    mask = torch.zeros((3, n), dtype=torch.int64, device="cuda")
    h.observe(None, None, mask)
    assert bool((mask != 0).any(dim=0).all())
    # determinism: same seed, same trajectory
    h2 = capi.EnvHandle(n, 42)
    h2.step_random(T)
    torch.cuda.synchronize()
    assert h2.get_state().tobytes() == st.tobytes()
    h.close(); h2.close()


def test_rollout_kernel_equals_step_by_step(torch):
    """bb_env_rollout_random: one launch, every step's outputs materialised."""
    from bbgpu import capi
    n, S, seed = 5000, 24, 13
    a, b = capi.EnvHandle(n, seed), capi.EnvHandle(n, seed)
    A = torch.zeros((S, n), dtype=torch.int32, device="cuda")
    R = torch.zeros((S, n), dtype=torch.float32, device="cuda")
    Tm = torch.zeros((S, n), dtype=torch.uint8, device="cuda")
    M = torch.zeros((S, 3, n), dtype=torch.int64, device="cuda")
    st1 = torch.zeros(4, dtype=torch.int64, device="cuda")
    st2 = torch.zeros(4, dtype=torch.int64, device="cuda")
    a.rollout_random(S, A, R, Tm, M, st1)
    B = _dev_buffers(torch, n)
    for t in range(S):
        b.step_random(1, B["actions"], B["rewards"], B["term"], B["mask"], st2)
        assert torch.equal(A[t], B["actions"]) and torch.equal(Tm[t], B["term"]) and torch.equal(M[t], B["mask"]), t
        assert torch.equal(R[t].view(torch.int32), B["rewards"].view(torch.int32)), t
    torch.cuda.synchronize()
    assert a.get_state().tobytes() == b.get_state().tobytes() and st1.tolist() == st2.tolist()
    a.close(); b.close()


def test_config4_size_and_tiny_sizes(torch):
    """1,048,576 envs (BASELINE config 4's total) on one GPU, and the 1-env / odd-size corners."""
    from bbgpu import capi
    n = 1_048_576
    stats = torch.zeros(4, dtype=torch.int64, device="cuda")
    h = capi.EnvHandle(n, 7)
    h.step_random(20, None, None, None, None, stats)
    torch.cuda.synchronize()
    s = stats.cpu().tolist()
    assert s[0] == n * 20 and s[1] > n // 2          # ~1 episode per 14.7 steps per env
    st = h.get_state()
    assert int(st["moves"].max()) <= 20 and int(st["policy_ctr"].min()) == 20 == int(st["policy_ctr"].max())
    h.close()
    for m in (1, 31, 33, 129):
        a = capi.EnvHandle(m, 3)
        B = _dev_buffers(torch, m)
        a.step_random(50, B["actions"], B["rewards"], B["term"], B["mask"], None)
        torch.cuda.synchronize()
        big = capi.EnvHandle(200, 3)
        big.step_random(50)
        torch.cuda.synchronize()
        assert a.get_state().tobytes() == big.get_state()[:m].tobytes()      # same global ids -> same trajectories
        a.close(); big.close()
    # empty GAE / sample calls are no-ops
    e = torch.zeros((0, 8), device="cuda")
    capi.gae(e, e, e, torch.zeros(8, device="cuda"), 0.99, 0.95, e, e, None)
    torch.cuda.synchronize()


def test_hundred_candidate_exhaustion_on_gpu(torch):
    """engine.py:171-172 through the speculative warp deal: envs that reject all 100 candidates
    keep the last one; compared with the oracle (setup shared with tests/test_host_rules.py)."""
    from bbgpu import capi
    from oracle import bb_oracle_c as OC
    from test_host_rules import _exhaustion_setup
    n, seed = 40000, 31337
    board, target, pieces, streams = _exhaustion_setup(n, seed)
    ora = OC.CVecEnv(streams, n_threads=8)
    for k in range(n):
        ora.set_board(k, board, [0, 5, 9, 0b110])
    h = capi.EnvHandle(n, seed)
    st = h.get_state()
    st["board"] = board
    st["pieces"] = pieces
    st["draw_ctr"] = 1
    h.set_state(st)
    B = _dev_buffers(torch, n)
    acts = np.full(n, 1, np.int32)
    oo = ora.step(acts)
    go = _gpu_step(torch, h, B, acts)
    draws = (go["info"] >> 11) & 0x7F
    assert (draws == 100).sum() >= 20 and int(draws.max()) == 100
    assert np.array_equal(oo["board"], go["board"]) and np.array_equal(oo["pieces"], _pieces4(go["pieces"]))
    assert np.array_equal(oo["mask"], go["mask"]) and np.array_equal(oo["terminated"], go["terminated"])
    assert np.array_equal(oo["rewards"].view(np.uint32), go["rewards"].view(np.uint32))
    fresh = oo["terminated"] == 0
    assert np.array_equal(ora.stats()[fresh, 7], h.get_state()["draw_ctr"][fresh])
    # and one more ordinary step from there stays in lock step
    m = oo["mask"]
    bits = np.unpackbits(m.view(np.uint8).reshape(n, 24), axis=1, bitorder="little")
    acts2 = (np.random.RandomState(0).rand(n, 192) * bits).argmax(1).astype(np.int32)
    oo2 = ora.step(acts2)
    go2 = _gpu_step(torch, h, B, acts2)
    assert np.array_equal(oo2["board"], go2["board"]) and np.array_equal(oo2["mask"], go2["mask"])
    assert np.array_equal(oo2["rewards"].view(np.uint32), go2["rewards"].view(np.uint32))
    h.close()


# ------------------------------------------------------------------ injected candidate trios (bb_env_set_trios)
def _pcg64_streams(seed, n, L):
    """What the reference's per-env numpy Generator returns for successive rng.choice(37, size=3)
    calls (engine.py:109, pieces.py:354), env i seeded with seed + i (wrappers.py:35-38)."""
    streams = np.zeros((n, L, 3), np.uint8)
    for i in range(n):
        rng = np.random.default_rng(seed + i)
        for d in range(L):
            streams[i, d] = rng.choice(37, size=3, replace=True)
    return streams


def test_gpu_replays_the_references_own_seeded_pcg64_games(torch):
    """vec_trace_seeded.npz = VectorizedBlockBlastEnv(4, seed=42) of the unmodified reference: numpy
    PCG64 trios, re-seeded on every reset.  The GPU plays it from the injected trio table."""
    from bbgpu import capi
    tr = np.load(os.path.join(G, "vec_trace_seeded.npz"))
    n, seed = tr["actions"].shape[1], int(tr["seed"])
    h = capi.EnvHandle(n, 0, 0, None, capi.ENV_RESEED_ON_RESET)
    h.set_trios(_pcg64_streams(seed, n, 64))
    h.reset()
    B = _dev_buffers(torch, n)
    h.observe(B["board"], B["pieces"], B["mask"])
    torch.cuda.synchronize()
    assert np.array_equal(B["board"].cpu().numpy().view(np.uint64), tr["board0"])
    assert np.array_equal(_pieces4(B["pieces"].cpu().numpy().view(np.uint32)), tr["pieces0"])
    assert np.array_equal(B["mask"].cpu().numpy().view(np.uint64).T, tr["mask0"])
    n_term = 0
    for t in range(tr["actions"].shape[0]):
        o = _gpu_step(torch, h, B, tr["actions"][t])
        assert np.array_equal(o["board"], tr["board"][t]), t
        assert np.array_equal(_pieces4(o["pieces"]), tr["pieces"][t]), t
        assert np.array_equal(o["mask"], tr["mask"][t]), t
        assert np.array_equal(o["rewards"].view(np.uint32), tr["rewards"][t].view(np.uint32)), t
        assert np.array_equal(o["terminated"].astype(bool), tr["terminated"][t]), t
        assert np.array_equal((o["info"] & 1).astype(bool), tr["invalid"][t]), t
        tt = tr["terminated"][t]
        n_term += int(tt.sum())
        assert np.array_equal(o["ep_score"][tt], tr["ep_score"][t][tt]), t
        assert np.array_equal(o["ep_len"][tt], tr["ep_len"][t][tt]), t
    st = h.get_state()
    assert np.array_equal(st["score"], tr["score"][-1]) and np.array_equal(st["moves"], tr["moves"][-1])
    assert np.array_equal(st["lines_total"], tr["lines_total"][-1]) and np.array_equal(st["max_streak"], tr["max_streak"][-1])
    assert n_term > 20
    h.close()


def test_gpu_replays_reference_play_random_game_kats(torch):
    """engine_kats.json: GameEngine(seed) first trios and play_random_game(seed) results of the
    reference (engine.py:538-576) for 40 seeds.  The engine's PCG64 stream feeds both the trio draws
    and the move choice, so the oracle plays each game once to log (candidate trios, actions); the
    GPU then plays all 40 games side by side from the injected trios and must end with the
    REFERENCE's recorded score / moves / lines / max combo / blocks."""
    from bbgpu import capi
    from oracle import bb_oracle as O
    kats = json.load(open(os.path.join(G, "engine_kats.json")))
    seeds = sorted(int(s) for s in kats["random_game"])
    logs = []
    for s in seeds:
        draw = O.numpy_rng_factory(s)
        trios = []

        def rec(draw=draw, trios=trios):
            t = [int(x) for x in draw()]
            trios.append(t)
            return t
        g = O.Game(rec)
        acts = []
        while not g.over:
            mv = [(i, r, c) for i in range(3) if not g.used[i]
                  for r in range(O.N) for c in range(O.N) if O.fits(g.grid, g.trio[i], r, c)]
            if not mv:
                break
            i, r, c = mv[draw.rng.choice(len(mv))]
            acts.append(i * 64 + r * 8 + c)
            g.move(i, r, c)
        logs.append((trios, acts))
    n = len(seeds)
    L = max(len(t) for t, _ in logs) + 1
    T = max(len(a) for _, a in logs)
    table = np.zeros((n, L, 3), np.uint8)
    actions = np.zeros((T + 1, n), np.int32)
    for k, (trios, acts) in enumerate(logs):
        table[k, :len(trios)] = trios
        actions[:len(acts), k] = acts
    h = capi.EnvHandle(n, 0, 0, None, capi.ENV_NO_AUTO_RESET)
    h.set_trios(table)
    h.reset()
    B = _dev_buffers(torch, n)
    h.observe(None, B["pieces"], None)
    torch.cuda.synchronize()
    first = _pieces4(B["pieces"].cpu().numpy().view(np.uint32))[:, :3]
    for k, s in enumerate(seeds):
        if str(s) in kats["first_trio"]:
            assert first[k].tolist() == kats["first_trio"][str(s)], s
    done = np.zeros(n, bool)
    for t in range(T + 1):
        o = _gpu_step(torch, h, B, actions[t])
        live = np.array([t < len(a) for _, a in logs])
        assert not (o["info"][live & ~done] & 1).any(), t          # every logged move is legal on the GPU too
        assert ((o["info"][~live] & 1) == 1).all()                  # finished games reject everything (-10)
        done |= o["terminated"].astype(bool)
    st = h.get_state()
    for k, s in enumerate(seeds):
        got = [int(st["score"][k]), int(st["moves"][k]), int(st["lines_total"][k]), int(st["max_streak"][k]),
               int(st["blocks_total"][k])]
        assert got == kats["random_game"][str(s)], (s, got)
        assert int(st["draw_ctr"][k]) == len(logs[k][0])
    assert done.all()
    h.close()


def test_injected_trios_equal_the_philox_streams_they_copy(torch):
    """An injected table filled with the env's own Philox candidates reproduces the Philox run
    (accept / reject decisions, draw counters, wrap-around of a short table is never reached)."""
    from bbgpu import capi, philox
    n, seed, off, T = 3000, 9, 500, 150
    a = capi.EnvHandle(n, seed, off)
    b = capi.EnvHandle(n, 12345, off)
    b.set_trios(philox.candidate_trios(seed, off + np.arange(n), 256))
    b.reset()
    a.reset()
    # reset() on `a` consumed one more Philox candidate than the constructor's deal: align b
    sa = a.get_state()
    sb = b.get_state()
    sb["draw_ctr"] = sa["draw_ctr"]
    sb["pieces"] = sa["pieces"]
    b.set_state(sb)
    Ba, Bb = _dev_buffers(torch, n), _dev_buffers(torch, n)
    rs = np.random.RandomState(1)
    a.observe(None, None, Ba["mask"])
    torch.cuda.synchronize()
    m = Ba["mask"].cpu().numpy().view(np.uint64).T.copy()
    for t in range(T):
        bits = np.unpackbits(m.view(np.uint8).reshape(n, 24), axis=1, bitorder="little")
        acts = (rs.rand(n, 192) * bits).argmax(axis=1).astype(np.int32)
        ga, gb = _gpu_step(torch, a, Ba, acts), _gpu_step(torch, b, Bb, acts)
        for k in ("board", "pieces", "mask", "terminated", "info", "ep_score", "ep_len"):
            assert np.array_equal(ga[k], gb[k]), (t, k)
        assert np.array_equal(ga["rewards"].view(np.uint32), gb["rewards"].view(np.uint32)), t
        m = ga["mask"]
    assert a.get_state().tobytes() == b.get_state().tobytes()
    b.set_trios(None)                     # back to Philox: the handle keeps working
    b.step_random(3)
    torch.cuda.synchronize()
    a.close(); b.close()


def test_step_outputs_board_pieces_stats_and_mask_in(torch):
    """ABI v2 additions: bb_env_step writes the packed next observation and episode statistics
    itself; bb_env_step_random fed with the previous mask (mask_in) equals the recomputing form."""
    from bbgpu import capi
    n, seed, T = 5000, 21, 60
    a, b = capi.EnvHandle(n, seed), capi.EnvHandle(n, seed)
    Ba, Bb = _dev_buffers(torch, n), _dev_buffers(torch, n)
    sa = torch.zeros(8, dtype=torch.int64, device="cuda")
    sb = torch.zeros(8, dtype=torch.int64, device="cuda")
    a.observe(None, None, Ba["mask"])
    for t in range(T):
        a.step_random(1, Ba["actions"], Ba["rewards"], Ba["term"], Ba["mask"], sa, mask_in=Ba["mask"])
        b.step_random(1, Bb["actions"], Bb["rewards"], Bb["term"], Bb["mask"], sb)
        torch.cuda.synchronize()
        for k in ("actions", "rewards", "term", "mask"):
            assert torch.equal(Ba[k], Bb[k]), (t, k)
    assert a.get_state().tobytes() == b.get_state().tobytes()
    assert sa.tolist() == sb.tolist() and sa[1].item() > 100 and sa[4].item() >= sa[2].item() / sa[1].item()
    # the same actions through bb_env_step: board / pieces / stats written by the step kernel
    c, d = capi.EnvHandle(n, seed), capi.EnvHandle(n, seed)
    sc = torch.zeros(8, dtype=torch.int64, device="cuda")
    Bc = _dev_buffers(torch, n)
    eps = score = length = 0
    best = 0
    for t in range(T):
        d.step_random(1, Bb["actions"], Bb["rewards"], Bb["term"], Bb["mask"], None)
        c.step(Bb["actions"], Bc["rewards"], Bc["term"], Bc["mask"], Bc["ep_score"], Bc["ep_len"], Bc["info"],
               board_out=Bc["board"], pieces_out=Bc["pieces"], stats=sc)
        torch.cuda.synchronize()
        st = c.get_state()
        assert np.array_equal(Bc["board"].cpu().numpy().view(np.uint64), st["board"]), t
        assert np.array_equal(Bc["pieces"].cpu().numpy().view(np.uint32), st["pieces"]), t
        assert torch.equal(Bc["mask"], Bb["mask"]) and torch.equal(Bc["rewards"], Bb["rewards"])
        tt = Bc["term"].cpu().numpy().astype(bool)
        es = Bc["ep_score"].cpu().numpy()[tt]
        eps += int(tt.sum()); score += int(es.sum()); length += int(Bc["ep_len"].cpu().numpy()[tt].sum())
        best = max(best, int(es.max()) if len(es) else 0)
    assert sc.tolist()[:5] == [n * T, eps, score, length, best]
    assert sc.tolist()[:5] == sa.tolist()[:5]
    for h in (a, b, c, d):
        h.close()


def test_dense_host_step_equals_packed_host_step(torch):
    """bb_env_step_host_dense (reference-layout obs in one pinned block) against the packed host
    step expanded on the host, plus the on-demand info tail (bb_env_fetch_step_info)."""
    from bbgpu import capi
    from bbgpu.vec_env import expand_board, expand_pieces, expand_mask
    n, seed, T = 777, 4, 40
    a, b = capi.EnvHandle(n, seed), capi.EnvHandle(n, seed)
    a.step_random(11); b.step_random(11)
    blk = capi.pinned_result_block(n)
    dense = capi.pinned_dense_block(n)
    ha = torch.zeros(n, dtype=torch.int32).pin_memory()
    tail = [torch.zeros(n, dtype=torch.int32).pin_memory() for _ in range(3)]
    b.observe_host_dense(dense["_block"])
    st = b.get_state()
    assert np.array_equal(dense["board"].numpy(), expand_board(st["board"]))
    assert np.array_equal(dense["pieces"].numpy(), expand_pieces(st["pieces"]))
    for t in range(T):
        a.sample_valid_actions(t + 1, None, ha)
        if t % 7 == 3:
            ha.numpy()[::5] = 191                                   # some invalid actions too
        a.step_host(ha, blk["rewards"], blk["term"], blk["board"], blk["pieces"], blk["mask"], blk["ep_score"],
                    blk["ep_len"], blk["info"])
        b.step_host_dense(ha, dense["_block"])
        assert np.array_equal(dense["rewards"].numpy().view(np.uint32), blk["rewards"].numpy().view(np.uint32)), t
        assert np.array_equal(dense["term"].numpy(), blk["term"].numpy()), t
        assert np.array_equal(dense["board"].numpy(), expand_board(blk["board"].numpy().view(np.uint64))), t
        assert np.array_equal(dense["pieces"].numpy(), expand_pieces(blk["pieces"].numpy().view(np.uint32))), t
        assert np.array_equal(dense["action_mask"].numpy(), expand_mask(blk["mask"].numpy().view(np.uint64))), t
        b.fetch_step_info(*tail)
        tt = blk["term"].numpy().astype(bool)
        assert np.array_equal(tail[2].numpy(), blk["info"].numpy()), t
        assert np.array_equal(tail[0].numpy()[tt], blk["ep_score"].numpy()[tt]), t
        assert np.array_equal(tail[1].numpy()[tt], blk["ep_len"].numpy()[tt]), t
    assert a.get_state().tobytes() == b.get_state().tobytes()
    # prefix-only transfer (no info tail requested) leaves the same prefix
    blk2 = capi.pinned_result_block(n)
    a2 = capi.EnvHandle(n, seed)
    a2.set_state(a.get_state())
    a.sample_valid_actions(99, None, ha)
    a.step_host(ha, blk["rewards"], blk["term"], blk["board"], blk["pieces"], blk["mask"], blk["ep_score"], blk["ep_len"], blk["info"])
    a2.step_host(ha, blk2["rewards"], blk2["term"], blk2["board"], blk2["pieces"], blk2["mask"], None, None, None)
    for k in ("rewards", "term", "board", "pieces", "mask"):
        assert torch.equal(blk[k], blk2[k]), k
    assert a.get_state().tobytes() == a2.get_state().tobytes()
    for h in (a, b, a2):
        h.close()
