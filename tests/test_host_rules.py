"""CPU-only: the product's rules header (csrc/bb_rules.cuh — the per-env body of the sm_100a
step kernel) compiled for the host and fuzzed against the oracle.  Catches rule bugs without
a GPU; the GPU tests repeat the comparison through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import hostrules as H
from bbgpu import philox
from oracle import bb_oracle as O
from oracle import bb_oracle_c as OC


def rand_board(rs, fill):
    bits = rs.rand(64) < fill
    return int(sum(1 << k for k in range(64) if bits[k]))


def test_reference_kats_on_bitboards():
    L = H.lib()
    # tests/test_board.py:229-240: 64 anchors for SINGLE, 40 for I_H on an empty board
    assert bin(L.bbh_valid(0, 0)).count("1") == 64
    assert bin(L.bbh_valid(0, 13)).count("1") == 40
    # :262-376 full board clears 8 rows + 8 cols
    ln = C.c_int(0)
    assert L.bbh_clear(2 ** 64 - 1, C.byref(ln)) == 0 and ln.value == 16
    # row 3 and column 5 (15 blocks) -> (1,1)
    b = (0xFF << 24) | sum(1 << (r * 8 + 5) for r in range(8))
    assert L.bbh_clear(b, C.byref(ln)) == 0 and ln.value == 2
    # :387-400 plus pattern has 2 holes... restated: a cell enclosed by 4 filled neighbours
    g = O.new_grid()
    for r, c in ((0, 1), (1, 0), (1, 2), (2, 1)):
        g[r][c] = 1
    assert L.bbh_holes(O.grid_to_u64(g)) == O.holes(g) == 2   # (1,1) and the corner (0,0)
    # :402-416 center openness 1.0 / 0.0
    assert L.bbh_center(0) == 0 and L.bbh_center(0x00003C3C3C3C0000) == 16


def test_primitives_match_cell_oracle():
    L = H.lib()
    rs = np.random.RandomState(0)
    for fill in (0.05, 0.3, 0.5, 0.7, 0.95):
        for _ in range(120):
            b = rand_board(rs, fill)
            g = O.u64_to_grid(b)
            assert L.bbh_holes(b) == O.holes(g)
            assert L.bbh_center(b) == sum(g[r][c] for r in range(2, 6) for c in range(2, 6))
            for p in range(37):
                want = sum(1 << (r * 8 + c) for r in range(8) for c in range(8) if O.fits(g, p, r, c))
                assert L.bbh_valid(b, p) == want
            g2 = [row[:] for row in g]
            for r in rs.choice(8, rs.randint(0, 3), replace=False):
                g2[r] = [1] * 8
            for c in rs.choice(8, rs.randint(0, 3), replace=False):
                for r in range(8):
                    g2[r][c] = 1
            b2 = O.grid_to_u64(g2)
            ln = C.c_int(0)
            got = L.bbh_clear(b2, C.byref(ln))
            nr, nc = O.sweep_lines(g2)
            assert got == O.grid_to_u64(g2) and ln.value == nr + nc


def test_every_trio_is_solvable_on_an_empty_board():
    """bb_reset_state accepts the first candidate without searching; engine.py:155-172 would
    too, because every trio fits an empty board."""
    L = H.lib()
    for p0 in range(37):
        for p1 in range(p0, 37):
            for p2 in range(p1, 37):
                assert L.bbh_solvable(0, p0, p1, p2) == 1
    assert OC.trio_solvable(0, 36, 36, 36)[0] and OC.trio_solvable(0, 15, 16, 36)[0]


def test_pruned_solver_equals_reference_dfs_on_scattered_boards():
    L = H.lib()
    rs = np.random.RandomState(1)
    n_rej = 0
    for fill in (0.15, 0.3, 0.45, 0.55, 0.65, 0.75, 0.85):
        for _ in range(2500):
            ln = C.c_int(0)
            b = L.bbh_clear(rand_board(rs, fill), C.byref(ln))
            p = [int(x) for x in rs.randint(0, 37, 3)]
            want, _ = OC.trio_solvable(b, *p)
            assert bool(L.bbh_solvable(b, *p)) == want, (hex(b), p)
            n_rej += not want
    assert n_rej > 3000


def test_pruned_solver_on_structured_boards():
    """Boards built to need line clears: nearly-full rows/columns plus noise."""
    L = H.lib()
    rs = np.random.RandomState(2)
    checked = 0
    for _ in range(6000):
        g = [[1 if rs.rand() < 0.35 else 0 for _ in range(8)] for _ in range(8)]
        for _ in range(rs.randint(1, 4)):
            k = rs.randint(8)
            gap = set(rs.choice(8, rs.randint(1, 4), replace=False).tolist())
            if rs.rand() < 0.5:
                for c in range(8):
                    g[k][c] = 0 if c in gap else 1
            else:
                for r in range(8):
                    g[r][k] = 0 if r in gap else 1
        ln = C.c_int(0)
        b = L.bbh_clear(O.grid_to_u64(g), C.byref(ln))
        p = [int(x) for x in rs.choice([5, 6, 13, 14, 15, 16, 17, 34, 35, 36, 0, 1, 2, 9, 26, 30], 3)]
        want, _ = OC.trio_solvable(b, *p)
        assert bool(L.bbh_solvable(b, *p)) == want, (hex(b), p)
        checked += 1
    assert checked == 6000


@pytest.mark.parametrize("flags,reseed", [(0, False), (1, True)])
def test_host_env_step_is_bit_exact_with_c_oracle(flags, reseed):
    n, T, seed = 192, 400, 42
    streams = philox.candidate_trios(seed, np.arange(n), 1024)
    ora = OC.CVecEnv(streams, reseed=reseed)
    host = H.HostEnv(n, seed, flags=flags)
    b, p, m = ora.export()
    assert np.array_equal(b, host.state["board"]) and np.array_equal(p, host.pieces4()) and np.array_equal(m, host.masks())
    rs = np.random.RandomState(5)
    for t in range(T):
        m = host.masks()
        acts = np.zeros(n, np.int32)
        for i in range(n):
            va = [pp * 64 + k for pp in range(3) for k in range(64) if (int(m[i, pp]) >> k) & 1]
            acts[i] = va[rs.randint(len(va))] if rs.rand() > 0.03 else rs.randint(-70, 260)
        oo, ho = ora.step(acts), host.step(acts)
        assert np.array_equal(oo["board"], host.state["board"]), t
        assert np.array_equal(oo["pieces"], host.pieces4()), t
        assert np.array_equal(oo["mask"], ho["mask"]), t
        assert np.array_equal(oo["rewards"].view(np.uint32), ho["rewards"].view(np.uint32)), t
        assert np.array_equal(oo["terminated"], ho["terminated"]), t
        assert np.array_equal(oo["invalid"], ho["info"] & 1), t
        tt = oo["terminated"].astype(bool)
        assert np.array_equal(oo["ep_score"][tt], ho["ep_score"][tt]) and np.array_equal(oo["ep_len"][tt], ho["ep_len"][tt])
        st = ora.stats()
        for k, f in enumerate(("score", "streak", "moves", "lines_total", "max_streak", "blocks_total")):
            assert np.array_equal(st[:, k], host.state[f]), (t, f)
        if not reseed:
            assert np.array_equal(st[:, 7], host.state["draw_ctr"]), t
    assert not ora.exhausted()


def test_host_philox_equals_numpy_replica():
    L = H.lib()
    for env in (0, 1, 77, 2 ** 33 + 5):
        for d in (0, 1, 1000):
            w = L.bbh_draw_trio(12345678901234567, env, d)
            t = philox.candidate_trios(12345678901234567, [env], 1, first_draw=d)[0, 0]
            assert [w & 0xFF, (w >> 8) & 0xFF, (w >> 16) & 0xFF] == t.tolist()


def test_fused_random_policy_matches_oracle_with_host_policy_words():
    """bb_env_step_random's action rule: k-th valid action in np.where(mask) order,
    k = mulhi(word, n_valid) (block_blast_env.py:313-323 picks uniformly among the same set)."""
    n, T, seed = 64, 300, 9
    streams = philox.candidate_trios(seed, np.arange(n), 1024)
    ora = OC.CVecEnv(streams)
    host = H.HostEnv(n, seed)
    words = philox.policy_words(seed, np.arange(n), np.arange(T))
    for t in range(T):
        _, _, m = ora.export()
        acts = np.zeros(n, np.int32)
        for i in range(n):
            va = [pp * 64 + k for pp in range(3) for k in range(64) if (int(m[i, pp]) >> k) & 1]
            acts[i] = va[int(philox.mulhi32(words[t, i], len(va)))]
        ho = host.step(None)
        assert np.array_equal(ho["actions"], acts), t
        oo = ora.step(acts)
        assert np.array_equal(oo["board"], host.state["board"]), t
        assert np.array_equal(oo["rewards"].view(np.uint32), ho["rewards"].view(np.uint32)), t


def _exhaustion_setup(n, seed):
    """States whose third placement leaves a board on which only SINGLE fits: every row and column
    has exactly one empty cell, no two of them adjacent (not even diagonally).  A candidate trio is
    solvable only if it contains a SINGLE (8 % of draws), so a few envs in 10^5 reject all 100
    candidates and keep the last one (engine.py:171-172)."""
    empties = [(r, (3 * r) % 8) for r in range(8)]
    target = (2 ** 64 - 1)
    for r, c in empties:
        target &= ~(1 << (r * 8 + c))
    board = target & ~(1 << 1)                                   # cell (0,1) also empty: the SINGLE goes there
    pieces = 0 | (5 << 8) | (9 << 16) | (0b110 << 24)            # trio (SINGLE, TRIO_H, TRIO_L1), pieces 1 and 2 used
    from bbgpu import philox
    L = 112
    streams = np.zeros((n, L, 3), np.uint8)
    streams[:, 1:] = philox.candidate_trios(seed, np.arange(n), L - 1, first_draw=1)   # cursor 1 <-> draw_ctr 1
    return board, target, pieces, streams


def test_hundred_candidate_exhaustion_host_vs_oracle():
    n, seed = 40000, 31337
    board, target, pieces, streams = _exhaustion_setup(n, seed)
    ora = OC.CVecEnv(streams, n_threads=8)
    host = H.HostEnv(n, seed)
    for k in range(n):
        ora.set_board(k, board, [0, 5, 9, 0b110])
    host.state["board"] = board
    host.state["pieces"] = pieces
    host.state["draw_ctr"] = 1
    acts = np.full(n, 1, np.int32)                               # piece 0 (SINGLE) at row 0, col 1
    oo, ho = ora.step(acts), host.step(acts)
    draws = (ho["info"] >> 11) & 0x7F
    assert (draws == 100).sum() >= 1, "no env exhausted its 100 candidates: enlarge n"
    assert np.array_equal(oo["board"], host.state["board"]) and np.array_equal(oo["pieces"], host.pieces4())
    assert np.array_equal(oo["mask"], ho["mask"]) and np.array_equal(oo["terminated"], ho["terminated"])
    assert np.array_equal(oo["rewards"].view(np.uint32), ho["rewards"].view(np.uint32))
    # draws consumed: oracle counts its constructor draw too
    fresh = oo["terminated"] == 0
    assert np.array_equal(ora.stats()[fresh, 7], host.state["draw_ctr"][fresh])
    # both outcomes of an exhausted deal occur: game over (nothing of the kept trio fits) and
    # play continues (the kept, unsolvable trio still contains a piece that fits: engine.py:440-441)
    ex = draws == 100
    assert ex.sum() >= 20 and oo["terminated"][ex].any() and not oo["terminated"][ex].all()
