"""Pins the oracles (oracle/bb_oracle.py, oracle/bb_oracle.c) against vectors recorded from
the UNMODIFIED reference by tests/golden/make_golden.py.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import bb_oracle as O
from oracle import bb_oracle_c as OC

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


# ------------------------------------------------------------------ engine KATs (PCG64 path)
def test_first_trios_and_random_games_python_oracle():
    with open(os.path.join(G, "engine_kats.json")) as f:
        kats = json.load(f)
    for s, trio in kats["first_trio"].items():
        g = O.Game(O.numpy_rng_factory(int(s)))
        assert g.trio == trio
    for s, want in kats["random_game"].items():
        st = O.play_random_game(int(s))
        assert [st["score"], st["moves"], st["lines"], st["max_combo"], st["blocks"]] == want, s


def test_survey_kats():
    # GameEngine(seed=42) trio = [DIAG2_TL_BR, L_3, Z_H]
    g = O.Game(O.numpy_rng_factory(42))
    assert [O.PIECE_NAMES[p] for p in g.trio] == ["DIAG2_TL_BR", "L_3", "Z_H"]
    assert O.play_random_game(0) == dict(score=41, moves=11, lines=0, max_combo=0, blocks=41)
    assert O.play_random_game(1)["score"] == 459


# ------------------------------------------------------------------ vec traces
def _check_step(tr, t, board, pieces, mask, rewards, term, invalid, stats, ep_score, ep_len, draws=True):
    assert np.array_equal(board, tr["board"][t]), t
    assert np.array_equal(pieces, tr["pieces"][t]), t
    assert np.array_equal(mask, tr["mask"][t]), t
    assert rewards.dtype == np.float32
    assert np.array_equal(rewards.view(np.uint32), tr["rewards"][t].view(np.uint32)), t
    assert np.array_equal(term.astype(bool), tr["terminated"][t]), t
    assert np.array_equal(invalid.astype(bool), tr["invalid"][t]), t
    assert np.array_equal(stats[:, 0], tr["score"][t]), t
    assert np.array_equal(stats[:, 1], tr["streak"][t]), t
    assert np.array_equal(stats[:, 2], tr["moves"][t]), t
    assert np.array_equal(stats[:, 3], tr["lines_total"][t]), t
    assert np.array_equal(stats[:, 4], tr["max_streak"][t]), t
    assert np.array_equal(stats[:, 5], tr["blocks_total"][t]), t
    assert np.array_equal(stats[:, 6], tr["holes"][t]), t
    if draws:
        assert np.array_equal(stats[:, 7], tr["draws"][t]), t
    tt = tr["terminated"][t]
    assert np.array_equal(ep_score[tt], tr["ep_score"][t][tt]), t
    assert np.array_equal(ep_len[tt], tr["ep_len"][t][tt]), t


@pytest.mark.parametrize("name", ["vec_trace.npz", "vec_trace_cfg.npz"])
def test_c_oracle_replays_reference_trace(name):
    tr = load(name)
    cfg = json.loads(str(tr["reward_cfg_json"]))
    env = OC.CVecEnv(tr["streams"], reward_cfg=cfg)
    b, p, m = env.export()
    assert np.array_equal(b, tr["board0"]) and np.array_equal(p, tr["pieces0"]) and np.array_equal(m, tr["mask0"])
    assert np.array_equal(env.stats()[:, 7], tr["draws0"])
    T = tr["actions"].shape[0]
    for t in range(T):
        out = env.step(tr["actions"][t])
        _check_step(tr, t, out["board"], out["pieces"], out["mask"], out["rewards"], out["terminated"],
                    out["invalid"], env.stats(), out["ep_score"], out["ep_len"])
    assert not env.exhausted()
    assert not tr["truncated"].any()


def _py_vec(tr, n_steps, envs, do_reset=True):
    vec = O.VecEnv(envs)
    if do_reset:
        vec.reset()
    for t in range(n_steps):
        obs, rew, term, trunc, infos = vec.step(tr["actions"][t])
        board = np.array([O.grid_to_u64(e.game.grid) for e in vec.envs], dtype=np.uint64)
        pieces = np.array([e.game.trio + [sum(int(u) << k for k, u in enumerate(e.game.used))]
                           for e in vec.envs], dtype=np.uint8)
        mask = np.array([[sum(1 << k for k in range(64) if obs["action_mask"][i][p * 64 + k]) for p in range(3)]
                         for i in range(vec.num_envs)], dtype=np.uint64)
        stats = np.array([[e.game.score, e.game.streak, e.game.moves, e.game.lines_total, e.game.max_streak,
                           e.game.blocks_total, O.holes(e.game.grid), e.game.draws] for e in vec.envs])
        invalid = np.array([i["invalid_action"] for i in infos])
        ep_score = np.array([i.get("final_score", -1) for i in infos])
        ep_len = np.array([i["moves"] for i in infos])
        yield t, board, pieces, mask, rew, term, invalid, stats, ep_score, ep_len, obs


def test_python_oracle_replays_reference_trace():
    tr = load("vec_trace.npz")
    n = tr["streams"].shape[0]

    def mk(i):
        it = iter(tr["streams"][i])
        # the recorded stream starts at the deal of VectorizedBlockBlastEnv.reset(); the
        # oracle Env deals once in its constructor, so no further reset here.
        return O.Env(draw=lambda: next(it))
    envs = [mk(i) for i in range(n)]
    for t, board, pieces, mask, rew, term, invalid, stats, eps, epl, obs in _py_vec(tr, 500, envs, do_reset=False):
        _check_step(tr, t, board, pieces, mask, rew, term, invalid, stats, eps, epl)
        assert np.array_equal(obs["board"].reshape(n, -1).sum(1).astype(np.int32), tr["obs_board_sum"][t])
        assert np.array_equal(obs["pieces"].reshape(n, 3, -1).sum(2).astype(np.int32), tr["obs_pieces_sum"][t])


def test_python_oracle_seeded_reset_path():
    """VectorizedBlockBlastEnv(seed=42): every auto-reset re-seeds env i with 42+i
    (wrappers.py:102 -> block_blast_env.py:215 -> engine.py:137-138)."""
    tr = load("vec_trace_seeded.npz")
    n = tr["actions"].shape[1]
    seed = int(tr["seed"])
    envs = [O.Env(seed=seed + i, rng_factory=O.numpy_rng_factory) for i in range(n)]
    for t, board, pieces, mask, rew, term, invalid, stats, eps, epl, _ in _py_vec(tr, tr["actions"].shape[0], envs):
        _check_step(tr, t, board, pieces, mask, rew, term, invalid, stats, eps, epl, draws=False)


def test_c_oracle_seeded_reset_path():
    tr = load("vec_trace_seeded.npz")
    n = tr["actions"].shape[1]
    seed = int(tr["seed"])
    L = 64
    streams = np.zeros((n, L, 3), np.uint8)
    for i in range(n):
        rng = np.random.default_rng(seed + i)
        for d in range(L):
            streams[i, d] = rng.choice(37, size=3, replace=True)
    env = OC.CVecEnv(streams, reseed=True)
    for t in range(tr["actions"].shape[0]):
        out = env.step(tr["actions"][t])
        _check_step(tr, t, out["board"], out["pieces"], out["mask"], out["rewards"], out["terminated"],
                    out["invalid"], env.stats(), out["ep_score"], out["ep_len"], draws=False)
    assert not env.exhausted()


# ------------------------------------------------------------------ GAE / policy maths
def test_gae_oracles_match_reference():
    g = load("gae_golden.npz")
    for tag in "abc":
        args = (g[f"{tag}_rewards"], g[f"{tag}_values"], g[f"{tag}_dones"], g[f"{tag}_last"],
                float(g["gamma"]), float(g["lam"]))
        adv, ret = O.gae(*args)
        assert np.array_equal(adv, g[f"{tag}_adv"]) and np.array_equal(ret, g[f"{tag}_ret"])
        adv_c, ret_c = OC.gae(*args)
        assert np.array_equal(adv_c, g[f"{tag}_adv"]) and np.array_equal(ret_c, g[f"{tag}_ret"])
        np.testing.assert_allclose(O.normalize_advantages(adv), g[f"{tag}_norm"], rtol=1e-6, atol=1e-7)


def test_policy_terms_match_reference():
    g = load("policy_golden.npz")
    probs, logp, ent = O.masked_policy_terms(g["logits"], g["mask"], g["actions"])
    np.testing.assert_allclose(probs, g["probs"], rtol=2e-6, atol=1e-8)
    np.testing.assert_allclose(logp, g["log_prob"], rtol=2e-6, atol=2e-7)
    np.testing.assert_allclose(ent, g["entropy"], rtol=2e-6, atol=2e-7)
    assert np.array_equal(np.argmax(probs, axis=-1), g["argmax"])
    # masked-out actions carry exactly zero probability
    assert (probs[g["mask"] == 0] == 0).all()
