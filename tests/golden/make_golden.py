#!/usr/bin/env python3
"""Record golden vectors by RUNNING THE UNMODIFIED REFERENCE (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Needs /root/reference (read-only mount) — it does not exist on the GPU box, so the outputs
are committed under tests/golden/ and this script is the provenance record.  ``gymnasium`` is
absent from the image: tests/golden/_gym_stub provides the handful of names the reference's
env module touches.  Nothing from the reference is copied; only its
outputs are stored.

Fixtures written:
  engine_kats.json     GameEngine(seed) first trios; play_random_game(seed) statistics
                       (numpy PCG64 path: pins the oracle's rules incl. the regeneration DFS)
  vec_trace.npz        VectorizedBlockBlastEnv(8 envs, injected candidate-trio streams):
                       2,500 vec steps = 20,000 env-steps with injected invalid actions;
                       every observable per step
  vec_trace_cfg.npz    same, 4 envs x 600 steps, non-default reward_config
  vec_trace_seeded.npz VectorizedBlockBlastEnv(4, seed=42): the re-seed-on-reset path
  gae_golden.npz       RolloutBuffer.compute_returns_and_advantages + get_samples normalisation
  policy_golden.npz    BlockBlastNetwork masking / Categorical log-prob / masked entropy
"""
import json
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("BB_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_gym_stub"))
sys.path.insert(0, os.path.join(REF, "src"))
sys.path.insert(0, ROOT)

from game.engine import GameEngine, play_random_game  # noqa: E402  (reference)
from game.pieces import PIECE_LIST  # noqa: E402
from environment.wrappers import VectorizedBlockBlastEnv  # noqa: E402

import bbgpu  # noqa: E402  (product package: only its host Philox replica is used here)
from bbgpu import philox  # noqa: E402


class FakeRng:
    """Stands in for engine.rng: each choice(37, size=3) returns the next candidate trio."""

    def __init__(self, stream):
        self.stream = stream
        self.cursor = 0

    def choice(self, n, size=None, replace=True):
        assert (n, size, replace) == (37, 3, True), (n, size, replace)
        row = self.stream[self.cursor]
        self.cursor += 1
        return np.array(row, dtype=np.int64)


def grid_u64(grid):
    v = 0
    for r in range(8):
        for c in range(8):
            if grid[r, c]:
                v |= 1 << (r * 8 + c)
    return v


def mask_u64x3(mask192):
    m = np.asarray(mask192).reshape(3, 64)
    return [sum(1 << k for k in range(64) if m[p, k]) for p in range(3)]


def record_vec_trace(path, n_envs, n_steps, seed, reward_config=None, real_seed=None,
                     invalid_every=97):
    rs = np.random.RandomState(1234 + n_envs)
    if real_seed is None:
        streams = philox.candidate_trios(seed, np.arange(n_envs), 4096)
        venv = VectorizedBlockBlastEnv(num_envs=n_envs, seed=None, reward_config=reward_config)
        rngs = []
        for i, e in enumerate(venv.envs):
            e.engine.rng = FakeRng(streams[i])
            rngs.append(e.engine.rng)
    else:
        streams = np.zeros((n_envs, 1, 3), np.uint8)
        venv = VectorizedBlockBlastEnv(num_envs=n_envs, seed=real_seed, reward_config=reward_config)
        rngs = None
    obs, infos = venv.reset()

    def snapshot():
        board = np.array([grid_u64(e.engine.board.grid) for e in venv.envs], dtype=np.uint64)
        pieces = np.array([[PIECE_LIST.index(p) for p in e.engine.current_pieces]
                           + [sum(int(u) << k for k, u in enumerate(e.engine.pieces_used))]
                           for e in venv.envs], dtype=np.uint8)
        draws = np.array([r.cursor for r in rngs], dtype=np.int64) if rngs else np.zeros(n_envs, np.int64)
        return board, pieces, draws

    keys = ("actions", "rewards", "terminated", "truncated", "invalid", "board", "pieces", "mask",
            "draws", "score", "streak", "moves", "lines_total", "max_streak", "blocks_total",
            "holes", "ep_score", "ep_len", "obs_board_sum", "obs_pieces_sum")
    rec = {k: [] for k in keys}
    b0, p0, d0 = snapshot()
    init = dict(board0=b0, pieces0=p0, draws0=d0,
                mask0=np.array([mask_u64x3(obs["action_mask"][i]) for i in range(n_envs)], dtype=np.uint64))
    for t in range(n_steps):
        acts = np.zeros(n_envs, dtype=np.int32)
        for i in range(n_envs):
            va = np.where(obs["action_mask"][i])[0]
            if invalid_every and (t * n_envs + i) % invalid_every == invalid_every - 1:
                bad = np.where(obs["action_mask"][i] == 0)[0]
                choice = rs.randint(0, 4)
                if choice == 0 or len(bad) == 0:
                    acts[i] = rs.choice([-1, -64, 192, 200, 1000])
                else:
                    acts[i] = bad[rs.randint(len(bad))]
            else:
                acts[i] = va[rs.randint(len(va))]
        obs, rew, term, trunc, infos = venv.step(acts)
        board, pieces, draws = snapshot()
        rec["actions"].append(acts)
        rec["rewards"].append(rew.copy())
        rec["terminated"].append(term.copy())
        rec["truncated"].append(trunc.copy())
        rec["invalid"].append(np.array([bool(i["invalid_action"]) for i in infos]))
        rec["board"].append(board)
        rec["pieces"].append(pieces)
        rec["mask"].append(np.array([mask_u64x3(obs["action_mask"][i]) for i in range(n_envs)], dtype=np.uint64))
        rec["draws"].append(draws)
        eng = [e.engine for e in venv.envs]
        rec["score"].append(np.array([g.score for g in eng], dtype=np.int64))
        rec["streak"].append(np.array([g.combo_count for g in eng], dtype=np.int32))
        rec["moves"].append(np.array([g.moves_made for g in eng], dtype=np.int32))
        rec["lines_total"].append(np.array([g.total_lines_cleared for g in eng], dtype=np.int32))
        rec["max_streak"].append(np.array([g.max_combo for g in eng], dtype=np.int32))
        rec["blocks_total"].append(np.array([g.total_blocks_placed for g in eng], dtype=np.int32))
        rec["holes"].append(np.array([g.board.count_holes() for g in eng], dtype=np.int32))
        rec["ep_score"].append(np.array([i.get("final_score", -1) for i in infos], dtype=np.int64))
        rec["ep_len"].append(np.array([i["moves"] if t_ else -1 for i, t_ in zip(infos, term)], dtype=np.int32))
        # the dense observation is a pure function of (board, pieces): keep checksums only
        rec["obs_board_sum"].append(obs["board"].reshape(n_envs, -1).sum(1).astype(np.int32))
        rec["obs_pieces_sum"].append(obs["pieces"].reshape(n_envs, 3, -1).sum(2).astype(np.int32))
    out = {k: np.stack(v) for k, v in rec.items()}
    out.update(init)
    out["streams"] = streams[:, : int(out["draws"].max()) + 2] if rngs else streams
    out["seed"] = np.int64(seed if real_seed is None else real_seed)
    out["reward_cfg_json"] = np.array(json.dumps(reward_config or {}))
    assert out["rewards"].dtype == np.float32
    np.savez_compressed(path, **out)
    n_term = int(out["terminated"].sum())
    print("%s: %d env-steps, %d episodes, %d invalid, max draws %d, max score %d"
          % (os.path.basename(path), n_envs * n_steps, n_term, int(out["invalid"].sum()),
             int(out["draws"].max()), int(out["score"].max())))


def record_engine_kats(path):
    out = {"first_trio": {}, "random_game": {}}
    for s in (42, 43, 44, 0, 1, 7):
        g = GameEngine(seed=s)
        out["first_trio"][str(s)] = [PIECE_LIST.index(p) for p in g.current_pieces]
    for s in range(40):
        st = play_random_game(seed=s)
        out["random_game"][str(s)] = [int(st["score"]), int(st["moves_made"]),
                                      int(st["total_lines_cleared"]), int(st["max_combo"]),
                                      int(st["total_blocks_placed"])]
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print("engine_kats.json: seed 42 trio", out["first_trio"]["42"], "game0", out["random_game"]["0"])


def record_gae(path):
    import torch  # noqa: F401
    from agents.ppo import RolloutBuffer
    rs = np.random.RandomState(7)
    out = {}
    for tag, (T, N) in {"a": (16, 8), "b": (128, 5), "c": (1, 3)}.items():
        buf = RolloutBuffer(T, N)
        buf.rewards[:] = rs.randn(T, N).astype(np.float32) * 2
        buf.values[:] = rs.randn(T, N).astype(np.float32)
        buf.dones[:] = (rs.rand(T, N) < 0.15).astype(np.float32)
        last = rs.randn(N).astype(np.float32)
        buf.compute_returns_and_advantages(last, 0.99, 0.95)
        adv = buf.advantages.reshape(-1)
        norm = (adv - adv.mean()) / (adv.std() + 1e-8)   # ppo.py:196, same expression
        out.update({f"{tag}_rewards": buf.rewards.copy(), f"{tag}_values": buf.values.copy(),
                    f"{tag}_dones": buf.dones.copy(), f"{tag}_last": last,
                    f"{tag}_adv": buf.advantages.copy(), f"{tag}_ret": buf.returns.copy(),
                    f"{tag}_norm": norm.astype(np.float32)})
    out["gamma"] = np.float64(0.99)
    out["lam"] = np.float64(0.95)
    np.savez_compressed(path, **out)
    print("gae_golden.npz written")


def record_policy(path):
    import torch
    from models.network import BlockBlastNetwork
    torch.manual_seed(0)
    np.random.seed(11)        # the reference's sample_valid_actions draws from numpy's GLOBAL generator (block_blast_env.py:323)
    net = BlockBlastNetwork()
    net.eval()
    rs = np.random.RandomState(3)
    B = 64
    venv = VectorizedBlockBlastEnv(num_envs=B, seed=5)
    obs, _ = venv.reset()
    for _ in range(7):
        obs, *_ = venv.step(venv.sample_valid_actions())
    board = torch.from_numpy(obs["board"])
    pieces = torch.from_numpy(obs["pieces"])
    mask = torch.from_numpy(obs["action_mask"]).float()
    with torch.no_grad():
        raw, value = net.forward(board, pieces, None)
        raw = raw * 3.0   # widen the logit range a little; masking maths is what is pinned
        acts = torch.tensor([int(rs.choice(np.where(obs["action_mask"][i])[0])) for i in range(B)])

        class Fixed(BlockBlastNetwork):
            pass
        # evaluate the reference's own masking/log-prob/entropy code on these raw logits
        masked = raw + torch.where(mask.bool(), torch.zeros_like(raw), torch.full_like(raw, float("-inf")))
        orig_forward = net.forward
        net.forward = lambda b, p, m=None: (masked if m is not None else raw, value)
        a2, logp, ent, _ = net.get_action_and_value(board, pieces, mask, action=acts)
        det, logp_det, _, _ = net.get_action_and_value(board, pieces, mask, deterministic=True)
        net.forward = orig_forward
        probs = torch.softmax(masked, dim=-1)
    np.savez_compressed(path, logits=raw.numpy(), mask=obs["action_mask"].astype(np.uint8),
                        actions=acts.numpy(), log_prob=logp.numpy(), entropy=ent.numpy(),
                        argmax=det.numpy(), log_prob_argmax=logp_det.numpy(), probs=probs.numpy())
    print("policy_golden.npz written: mean entropy %.4f" % float(ent.mean()))


def main():
    os.makedirs(HERE, exist_ok=True)
    if "--only-policy" in sys.argv:
        record_policy(os.path.join(HERE, "policy_golden.npz"))
        return
    record_engine_kats(os.path.join(HERE, "engine_kats.json"))
    record_vec_trace(os.path.join(HERE, "vec_trace.npz"), 8, 2500, seed=42)
    record_vec_trace(os.path.join(HERE, "vec_trace_cfg.npz"), 4, 600, seed=7,
                     reward_config=dict(line_clear_base=0.7, block_placed=0.013, game_over_penalty=-2.5,
                                        hole_penalty=-0.11, center_bonus=0.3, combo_multiplier_bonus=0.37,
                                        survival_bonus=0.0021))
    record_vec_trace(os.path.join(HERE, "vec_trace_seeded.npz"), 4, 400, seed=0, real_seed=42)
    record_gae(os.path.join(HERE, "gae_golden.npz"))
    record_policy(os.path.join(HERE, "policy_golden.npz"))


if __name__ == "__main__":
    main()
