"""ctypes loader for the plain-C oracle (oracle/bb_oracle.c).  Test infrastructure only —
see the header of bb_oracle.c for who may use it."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libbboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.bbo_env_size.restype = C.c_int64
        L.bbo_trio_solvable.restype = C.c_int
        L.bbo_trio_solvable.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        L.bbo_random_rollout.restype = C.c_int64
        L.bbo_env_dfs_nodes.restype = C.c_int64
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class CVecEnv:
    """n oracle envs fed by per-env candidate-trio streams (uint8 [n, L, 3])."""

    def __init__(self, streams, reward_cfg=None, reseed=False, n_threads=1):
        L = lib()
        streams = np.ascontiguousarray(streams, dtype=np.uint8)
        assert streams.ndim == 3 and streams.shape[2] == 3
        self.streams = streams  # keep alive: the C side holds pointers into it
        self.n = streams.shape[0]
        self.n_threads = n_threads
        self.sz = L.bbo_env_size()
        self.buf = np.zeros(self.n * self.sz, dtype=np.uint8)
        from .bb_oracle import REWARD_DEFAULTS, REWARD_KEYS
        cfg = dict(REWARD_DEFAULTS)
        if reward_cfg:
            cfg.update(reward_cfg)
        self.cfg = np.array([cfg[k] for k in REWARD_KEYS], dtype=np.float64)
        for k in range(self.n):
            L.bbo_env_init(C.c_void_p(self.buf.ctypes.data + k * self.sz),
                           C.c_void_p(streams[k].ctypes.data), C.c_int64(streams.shape[1]),
                           _p(self.cfg), C.c_int(1 if reseed else 0))

    def _env(self, k):
        return C.c_void_p(self.buf.ctypes.data + k * self.sz)

    def export(self):
        """(board u64[n], pieces u8[n,4], mask u64[n,3]) of the current states."""
        L = lib()
        board = np.zeros(self.n, np.uint64)
        pieces = np.zeros((self.n, 4), np.uint8)
        mask = np.zeros((self.n, 3), np.uint64)
        for k in range(self.n):
            L.bbo_env_export(self._env(k), C.c_void_p(board.ctypes.data + 8 * k),
                             C.c_void_p(pieces.ctypes.data + 4 * k),
                             C.c_void_p(mask.ctypes.data + 24 * k))
        return board, pieces, mask

    def stats(self):
        """int64 [n, 8]: score, streak, moves, lines_total, max_streak, blocks_total, holes, draws."""
        L = lib()
        out = np.zeros((self.n, 8), np.int64)
        for k in range(self.n):
            L.bbo_env_stats(self._env(k), C.c_void_p(out.ctypes.data + 64 * k))
        return out

    def step(self, actions):
        """Vectorised step with auto-reset. Returns dict of arrays (post-step/post-reset obs)."""
        L = lib()
        a = np.ascontiguousarray(actions, dtype=np.int32)
        n = self.n
        out = dict(rewards=np.zeros(n, np.float32), terminated=np.zeros(n, np.uint8),
                   invalid=np.zeros(n, np.uint8), board=np.zeros(n, np.uint64),
                   pieces=np.zeros((n, 4), np.uint8), mask=np.zeros((n, 3), np.uint64),
                   ep_score=np.full(n, -1, np.int32), ep_len=np.full(n, -1, np.int32))
        L.bbo_vec_step(C.c_void_p(self.buf.ctypes.data), C.c_int64(n), _p(a), _p(out["rewards"]),
                       _p(out["terminated"]), _p(out["invalid"]), _p(out["board"]),
                       _p(out["pieces"]), _p(out["mask"]), _p(out["ep_score"]), _p(out["ep_len"]),
                       C.c_int(self.n_threads))
        return out

    def set_board(self, k, board, pieces4):
        """Overwrite env k's grid, trio and used bits (counters and stream cursor untouched)."""
        pc = np.ascontiguousarray(pieces4, dtype=np.uint8)
        lib().bbo_env_set_board(self._env(k), C.c_uint64(int(board)), _p(pc))

    def exhausted(self):
        L = lib()
        return any(L.bbo_env_exhausted(self._env(k)) for k in range(self.n))

    def dfs_nodes(self):
        L = lib()
        return sum(L.bbo_env_dfs_nodes(self._env(k)) for k in range(self.n))

    def random_rollout(self, n_steps, words):
        L = lib()
        words = np.ascontiguousarray(words, dtype=np.uint32)
        assert words.size >= n_steps * self.n
        ep = C.c_int64(0)
        ss = C.c_int64(0)
        done = L.bbo_random_rollout(C.c_void_p(self.buf.ctypes.data), C.c_int64(self.n),
                                    C.c_int64(n_steps), _p(words), C.byref(ep), C.byref(ss),
                                    C.c_int(self.n_threads))
        return done, ep.value, ss.value


def trio_solvable(board, p0, p1, p2):
    nodes = C.c_int64(0)
    ok = lib().bbo_trio_solvable(C.c_uint64(int(board)), int(p0), int(p1), int(p2), C.byref(nodes))
    return bool(ok), nodes.value


def gae(rewards, values, dones, last_values, gamma, lam):
    L = lib()
    r = np.ascontiguousarray(rewards, np.float32)
    v = np.ascontiguousarray(values, np.float32)
    d = np.ascontiguousarray(dones, np.float32)
    lv = np.ascontiguousarray(last_values, np.float32)
    T, N = r.shape
    adv = np.zeros_like(r)
    ret = np.zeros_like(r)
    L.bbo_gae(_p(r), _p(v), _p(d), _p(lv), C.c_double(gamma), C.c_double(lam), _p(adv), _p(ret),
              C.c_int64(T), C.c_int64(N))
    return adv, ret
