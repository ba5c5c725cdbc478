"""CPU ORACLE (test infrastructure, NOT product code) — plain-Python restatement of the
reference's Block Blast rules, reward shaping, vec-env auto-reset, GAE and masked
log-prob/entropy maths.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / reference arm
may import this module, and only as the checker / the CPU baseline.  The product
(``bbgpu``) never imports it and has no CPU fallback.

It deliberately uses the reference's *representation and cost structure* (a cell grid and
per-cell Python loops, one env object per env stepped serially) and not the product's
bitboards, so that it is an independent statement of the rules.

Pinned against the reference by ``tests/golden/make_golden.py`` (runs the unmodified
reference from /root/reference in the build container and records traces) and
``tests/test_oracle_golden.py`` (replays them through this file).  Citations are
``file:line`` into the reference checkout.

The piece shapes come from ``oracle/piece_table.json`` which tools/gen_piece_tables.py
derives from ``src/game/pieces.py`` (PIECE_LIST order, pieces.py:244-318).
"""
import json
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

with open(os.path.join(_HERE, "piece_table.json")) as _f:
    PIECE_TABLE = json.load(_f)
#: per piece index: tuple of (dr, dc) cells, top-left normalised (pieces.py:60-68)
PIECE_CELLS = [tuple((int(r), int(c)) for r, c in row["cells"]) for row in PIECE_TABLE]
PIECE_H = [row["h"] for row in PIECE_TABLE]
PIECE_W = [row["w"] for row in PIECE_TABLE]
PIECE_NAMES = [row["name"] for row in PIECE_TABLE]
NUM_PIECES = len(PIECE_CELLS)
N = 8  # board size (board.py:25)

#: reward_config defaults, block_blast_env.py:63-71 (order = the C-ABI's reward_cfg[7])
REWARD_KEYS = ("line_clear_base", "block_placed", "game_over_penalty", "hole_penalty",
               "center_bonus", "combo_multiplier_bonus", "survival_bonus")
REWARD_DEFAULTS = dict(line_clear_base=1.0, block_placed=0.01, game_over_penalty=-1.0,
                       hole_penalty=-0.05, center_bonus=0.02, combo_multiplier_bonus=0.5,
                       survival_bonus=0.001)


# --------------------------------------------------------------------------- board rules
def new_grid():
    return [[0] * N for _ in range(N)]


def fits(grid, piece, row, col):
    """board.py:71-93 — every block in bounds and on an empty cell."""
    for dr, dc in PIECE_CELLS[piece]:
        r, c = row + dr, col + dc
        if r < 0 or r >= N or c < 0 or c >= N:
            return False
        if grid[r][c] != 0:
            return False
    return True


def put(grid, piece, row, col):
    """board.py:95-115 (caller has validated)."""
    for dr, dc in PIECE_CELLS[piece]:
        grid[row + dr][col + dc] = 1


def has_anchor(grid, piece):
    """board.py:133-142 — scan anchors r <= 8-h, c <= 8-w."""
    for r in range(N - PIECE_H[piece] + 1):
        for c in range(N - PIECE_W[piece] + 1):
            if fits(grid, piece, r, c):
                return True
    return False


def sweep_lines(grid):
    """board.py:144-193 / engine.py:226-238 — detect ALL full rows and columns first,
    then zero them; returns (n_rows, n_cols)."""
    rows = [r for r in range(N) if all(grid[r][c] == 1 for c in range(N))]
    cols = [c for c in range(N) if all(grid[r][c] == 1 for r in range(N))]
    for r in rows:
        for c in range(N):
            grid[r][c] = 0
    for c in cols:
        for r in range(N):
            grid[r][c] = 0
    return len(rows), len(cols)


def holes(grid):
    """board.py:195-216 — empty cells whose 4 neighbours are filled or off-board."""
    k = 0
    for r in range(N):
        for c in range(N):
            if grid[r][c] != 0:
                continue
            blocked = 0
            for dr, dc in ((-1, 0), (1, 0), (0, -1), (0, 1)):
                rr, cc = r + dr, c + dc
                if rr < 0 or rr >= N or cc < 0 or cc >= N or grid[rr][cc] == 1:
                    blocked += 1
            if blocked == 4:
                k += 1
    return k


def center_openness(grid):
    """board.py:236-243 — 1 - filled(rows 2..5, cols 2..5)/16, a Python float."""
    s = 0
    for r in range(2, 6):
        for c in range(2, 6):
            s += grid[r][c]
    return 1.0 - (s / 16.0)


def grid_to_u64(grid):
    v = 0
    for r in range(N):
        for c in range(N):
            if grid[r][c]:
                v |= 1 << (r * 8 + c)
    return v


def u64_to_grid(v):
    return [[(v >> (r * 8 + c)) & 1 for c in range(N)] for r in range(N)]


# --------------------------------------------------------------------------- engine
class Game:
    """engine.py:81-454 restated.  ``draw()`` returns one candidate trio (3 piece indices)
    and stands for one ``rng.choice(37, size=3, replace=True)`` call (pieces.py:350-355)."""

    def __init__(self, draw):
        self.draw = draw
        self.draws = 0  # candidates consumed so far (parity counter, not in the reference)
        self._zero()
        self.deal()

    def _zero(self):
        # engine.py:107-122 / :140-149
        self.grid = new_grid()
        self.trio = [0, 0, 0]
        self.used = [False, False, False]
        self.score = 0
        self.streak = 0          # reference name: combo_count
        self.moves = 0
        self.lines_total = 0
        self.over = False
        self.max_streak = 0      # reference name: max_combo
        self.blocks_total = 0

    def reset(self, draw=None):
        """engine.py:127-153; a non-None ``draw`` models the re-seed at :137-138."""
        if draw is not None:
            self.draw = draw
        self._zero()
        self.deal()

    # -- trio regeneration: engine.py:155-238
    def deal(self):
        for _ in range(100):
            self.trio = [int(x) for x in self.draw()]
            self.draws += 1
            self.used = [False, False, False]
            if self._solvable([row[:] for row in self.grid], [False, False, False]):
                return
        # 100 rejections: keep the last candidate (engine.py:171-172)

    def _solvable(self, grid, used):
        """engine.py:181-224 — depth-3 search over piece order and anchors with line
        clears applied after each simulated placement."""
        if all(used):
            return True
        for i in range(3):
            if used[i]:
                continue
            p = self.trio[i]
            for r in range(N - PIECE_H[p] + 1):
                for c in range(N - PIECE_W[p] + 1):
                    if fits(grid, p, r, c):
                        g2 = [row[:] for row in grid]
                        put(g2, p, r, c)
                        sweep_lines(g2)
                        u2 = used[:]
                        u2[i] = True
                        if self._solvable(g2, u2):
                            return True
        return False

    # -- queries: engine.py:326-388
    def legal(self, i, r, c):
        if i < 0 or i >= 3 or self.used[i] or self.over:
            return False
        return fits(self.grid, self.trio[i], r, c)

    def mask3(self):
        """engine.py:364-380 — bool (3,8,8); used pieces give an all-False plane."""
        m = np.zeros((3, N, N), dtype=bool)
        for i in range(3):
            if self.used[i]:
                continue
            for r in range(N):
                for c in range(N):
                    if fits(self.grid, self.trio[i], r, c):
                        m[i, r, c] = True
        return m

    def any_move(self):
        return any((not self.used[i]) and has_anchor(self.grid, self.trio[i]) for i in range(3))

    # -- one move: engine.py:390-454.  Returns None when rejected, else a dict.
    def move(self, i, r, c):
        if not self.legal(i, r, c):
            return None
        p = self.trio[i]
        n = len(PIECE_CELLS[p])
        put(self.grid, p, r, c)
        self.used[i] = True
        self.moves += 1
        self.blocks_total += n
        nr, nc = sweep_lines(self.grid)
        lines = nr + nc
        if lines > 0:
            self.streak += 1
            self.max_streak = max(self.max_streak, self.streak)
            self.lines_total += lines
        else:
            self.streak = 0
        # engine.py:240-312 with the already-updated streak (:261) and lines*8 blocks (:427)
        gain = n
        if lines > 0:
            gain += (lines * N * 10) * min(lines, 4) * min(self.streak + 1, 8)
        self.score += gain
        if all(self.used):
            self.deal()
        if not self.any_move():
            self.over = True
        return dict(blocks=n, rows=nr, cols=nc, lines=lines,
                    combo_mult=(min(lines, 4) if lines > 0 else 1),  # engine.py:451
                    gain=gain, game_over=self.over)

    def filled(self):
        return sum(sum(row) for row in self.grid)


# --------------------------------------------------------------------------- gym env
class Env:
    """block_blast_env.py:20-323 restated (reset/step/reward/obs/mask/info)."""

    def __init__(self, draw=None, reward_config=None, seed=None, rng_factory=None):
        """Either pass ``draw`` (candidate-trio source that survives resets — the
        reference's ``seed=None`` behaviour with an injected rng) or ``seed`` +
        ``rng_factory(seed) -> draw`` (the reference's re-seed-on-every-reset behaviour,
        block_blast_env.py:212-215 -> engine.py:137-138)."""
        self.cfg = dict(REWARD_DEFAULTS)
        if reward_config:
            self.cfg.update(reward_config)
        self.seed = seed
        self.rng_factory = rng_factory
        if draw is None:
            draw = rng_factory(seed)
        self.game = Game(draw)
        self.prev_holes = 0
        self.prev_center = 1.0

    def reset(self, seed=None):
        if seed is not None:
            self.seed = seed
        self.game.reset(self.rng_factory(self.seed) if (self.seed is not None and self.rng_factory) else None)
        self.prev_holes = 0
        self.prev_center = 1.0
        return self.obs(), self.info()

    @staticmethod
    def decode(a):
        """block_blast_env.py:104-118 (Python floor division, so a<0 gives piece -1)."""
        return a // 64, (a % 64) // 8, a % 8

    def reward(self, res):
        """block_blast_env.py:148-193, same float64 operation order."""
        cfg = self.cfg
        r = 0.0
        r += res["blocks"] * cfg["block_placed"]
        r += cfg["survival_bonus"]
        if res["lines"] > 0:
            lr = res["lines"] * cfg["line_clear_base"]
            lr *= res["combo_mult"]
            r += lr
            if res["combo_mult"] > 1:
                r += (res["combo_mult"] - 1) * cfg["combo_multiplier_bonus"]
        if res["game_over"]:
            r += cfg["game_over_penalty"]
        h = holes(self.game.grid)
        d = h - self.prev_holes
        if d > 0:
            r += d * cfg["hole_penalty"]
        self.prev_holes = h
        o = center_openness(self.game.grid)
        if o >= self.prev_center:
            r += cfg["center_bonus"] * 0.1
        self.prev_center = o
        return r

    def step(self, a):
        """block_blast_env.py:224-264."""
        i, r, c = self.decode(int(a))
        if not self.game.legal(i, r, c):
            info = self.info()
            info["invalid_action"] = True
            return self.obs(), -10.0, False, False, info
        res = self.game.move(i, r, c)
        rew = self.reward(res)
        return self.obs(), rew, res["game_over"], False, self.info(res)

    def obs(self):
        """engine.py:478-507 + block_blast_env.py:134-146."""
        g = self.game
        board = np.array(g.grid, dtype=np.float32)
        pieces = np.zeros((3, N, N), dtype=np.float32)
        for i in range(3):
            if not g.used[i]:
                for dr, dc in PIECE_CELLS[g.trio[i]]:
                    pieces[i, dr, dc] = 1.0
        return dict(board=board, pieces=pieces,
                    action_mask=g.mask3().reshape(-1).astype(np.int8))

    def info(self, res=None):
        """block_blast_env.py:266-288."""
        g = self.game
        d = dict(score=g.score, moves=g.moves, lines_cleared=g.lines_total,
                 max_combo=g.max_streak, blocks_placed=g.blocks_total,
                 board_fill=g.filled() / 64, holes=holes(g.grid), invalid_action=False)
        if res:
            d["last_move"] = dict(blocks_placed=res["blocks"], lines_cleared=res["lines"],
                                  combo_multiplier=res["combo_mult"], score_gained=res["gain"])
        return d

    def valid_actions(self):
        return np.where(self.game.mask3().reshape(-1))[0].tolist()


class VecEnv:
    """wrappers.py:14-141 restated: serial loop, auto-reset that returns the reset obs
    but the terminal step's reward/terminated, ``final_score`` stashed in info."""

    def __init__(self, envs):
        self.envs = list(envs)
        self.num_envs = len(self.envs)

    def reset(self):
        out = [e.reset() for e in self.envs]
        return self._stack([o for o, _ in out]), [i for _, i in out]

    def step(self, actions):
        obs, infos = [], []
        rewards = np.zeros(self.num_envs, dtype=np.float32)
        term = np.zeros(self.num_envs, dtype=bool)
        trunc = np.zeros(self.num_envs, dtype=bool)
        for k, (e, a) in enumerate(zip(self.envs, actions)):
            o, r, t, tr, info = e.step(int(a))
            if t or tr:
                info["terminal_observation"] = o
                info["final_score"] = info["score"]
                o, _ = e.reset()
            obs.append(o)
            rewards[k] = r            # float64 -> float32 cast, wrappers.py:105
            term[k] = t
            trunc[k] = tr
            infos.append(info)
        return self._stack(obs), rewards, term, trunc, infos

    @staticmethod
    def _stack(obs):
        return {k: np.stack([o[k] for o in obs]) for k in ("board", "pieces", "action_mask")}

    def sample_valid_actions(self, rng=np.random):
        """wrappers.py:133-136 / block_blast_env.py:318-323."""
        out = []
        for e in self.envs:
            va = e.valid_actions()
            out.append(rng.choice(va) if va else 0)
        return np.array(out)


def numpy_rng_factory(seed):
    """``np.random.default_rng(seed)`` as a candidate-trio source (engine.py:109,
    pieces.py:354) — used to pin the oracle against the reference's seeded KATs."""
    rng = np.random.default_rng(seed)

    def draw():
        return rng.choice(NUM_PIECES, size=3, replace=True)
    draw.rng = rng
    return draw


def play_random_game(seed):
    """engine.py:538-576 — move picked with the *engine's* rng among get_valid_moves()
    (piece-major, then row, then col order, engine.py:348-362)."""
    draw = numpy_rng_factory(seed)
    g = Game(draw)
    while not g.over:
        mv = [(i, r, c) for i in range(3) if not g.used[i]
              for r in range(N) for c in range(N) if fits(g.grid, g.trio[i], r, c)]
        if not mv:
            break
        g.move(*mv[draw.rng.choice(len(mv))])
    return dict(score=g.score, moves=g.moves, lines=g.lines_total, max_combo=g.max_streak,
                blocks=g.blocks_total)


# --------------------------------------------------------------------------- PPO-side maths
def gae(rewards, values, dones, last_values, gamma, lam):
    """ppo.py:141-169 — float32 arrays (T,N), python-float gamma/lambda.
    Returns (advantages, returns)."""
    rewards = np.asarray(rewards, np.float32)
    values = np.asarray(values, np.float32)
    dones = np.asarray(dones, np.float32)
    T = rewards.shape[0]
    adv = np.zeros_like(rewards)
    last = 0
    for t in reversed(range(T)):
        nnt = 1.0 - dones[t]
        nv = last_values if t == T - 1 else values[t + 1]
        delta = rewards[t] + gamma * nv * nnt - values[t]
        last = delta + gamma * lam * nnt * last
        adv[t] = last
    return adv, adv + values


def normalize_advantages(adv):
    """ppo.py:196 — whole-buffer mean / population std + 1e-8, in float32 numpy."""
    a = np.asarray(adv, np.float32).reshape(-1)
    return (a - a.mean()) / (a.std() + 1e-8)


def masked_policy_terms(logits, mask, action):
    """network.py:172-262 for *given* actions, float32 numpy.
    Returns (probs, log_prob[action], entropy)."""
    logits = np.asarray(logits, np.float32)
    m = np.asarray(mask).astype(bool)
    z = np.where(m, logits, -np.inf).astype(np.float32)            # :175-180
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z, dtype=np.float32)
    probs = e / e.sum(axis=-1, keepdims=True, dtype=np.float32)      # F.softmax :210
    # torch.distributions.Categorical(probs): probs /= sum; logits = log(clamp(probs, eps, 1-eps))
    eps = np.finfo(np.float32).eps
    pn = probs / probs.sum(axis=-1, keepdims=True, dtype=np.float32)
    logp_all = np.log(np.clip(pn, eps, 1.0 - eps))
    idx = np.asarray(action).astype(np.int64)
    logp = np.take_along_axis(logp_all, idx[..., None], axis=-1)[..., 0]
    # _masked_entropy :232-262
    mp = probs * m.astype(np.float32)
    s = np.maximum(mp.sum(axis=-1, keepdims=True, dtype=np.float32), np.float32(1e-10))
    q = mp / s
    ent = -(q * np.log(np.maximum(q, np.float32(1e-10))) * m.astype(np.float32)).sum(axis=-1, dtype=np.float32)
    return probs, logp.astype(np.float32), ent.astype(np.float32)
