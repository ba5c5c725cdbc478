"""ctypes loader for tests/hostbuild/libbbrules_host.so — the product's rules header
(csrc/bb_rules.cuh) compiled for the host so it can be fuzzed against the oracle without a
GPU.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostbuild")
_LIB = None

STATE_DTYPE = np.dtype([("board", "<u8"), ("pieces", "<u4"), ("aux", "<u4"), ("score", "<i4"),
                        ("streak", "<i4"), ("moves", "<i4"), ("lines_total", "<i4"),
                        ("max_streak", "<i4"), ("blocks_total", "<i4"), ("draw_ctr", "<u4"),
                        ("policy_ctr", "<u4")])


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _DIR], stderr=subprocess.DEVNULL)
        L = C.CDLL(os.path.join(_DIR, "libbbrules_host.so"))
        L.bbh_state_size.restype = C.c_int64
        L.bbh_valid.restype = C.c_uint64
        L.bbh_valid.argtypes = [C.c_uint64, C.c_int]
        L.bbh_clear.restype = C.c_uint64
        L.bbh_clear.argtypes = [C.c_uint64, C.POINTER(C.c_int)]
        L.bbh_holes.argtypes = [C.c_uint64]
        L.bbh_center.argtypes = [C.c_uint64]
        L.bbh_solvable.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.bbh_solvable_fast.argtypes = [C.c_uint64, C.c_int, C.c_int, C.c_int]
        L.bbh_draw_trio.restype = C.c_uint32
        L.bbh_draw_trio.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        assert L.bbh_state_size() == STATE_DTYPE.itemsize == 48
        _LIB = L
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class HostEnv:
    """n envs stepped by the host build of bb_env_apply."""

    def __init__(self, n, seed, env_offset=0, flags=0, cfg=None):
        from oracle.bb_oracle import REWARD_DEFAULTS, REWARD_KEYS
        c = dict(REWARD_DEFAULTS)
        if cfg:
            c.update(cfg)
        self.cfg = np.array([c[k] for k in REWARD_KEYS], np.float64)
        self.n, self.seed, self.off, self.flags = n, seed, env_offset, flags
        self.state = np.zeros(n, STATE_DTYPE)
        lib().bbh_reset(_p(self.state), C.c_int64(n), C.c_uint64(seed), C.c_int64(env_offset), C.c_uint32(flags))

    def masks(self):
        m = np.zeros((self.n, 3), np.uint64)
        lib().bbh_masks(_p(self.state), C.c_int64(self.n), _p(m))
        return m

    def step(self, actions=None):
        n = self.n
        a = None if actions is None else np.ascontiguousarray(actions, np.int32)
        out = dict(rewards=np.zeros(n, np.float32), terminated=np.zeros(n, np.uint8), info=np.zeros(n, np.uint32),
                   gain=np.zeros(n, np.int32), ep_score=np.full(n, -1, np.int32), ep_len=np.full(n, -1, np.int32),
                   mask=np.zeros((n, 3), np.uint64), actions=np.zeros(n, np.int32))
        lib().bbh_step(_p(self.state), C.c_int64(n), _p(a), _p(self.cfg), C.c_uint64(self.seed), C.c_int64(self.off),
                       C.c_uint32(self.flags), _p(out["rewards"]), _p(out["terminated"]), _p(out["info"]), _p(out["gain"]),
                       _p(out["ep_score"]), _p(out["ep_len"]), _p(out["mask"]), _p(out["actions"]))
        return out

    def pieces4(self):
        p = self.state["pieces"]
        return np.stack([p & 0xFF, (p >> 8) & 0xFF, (p >> 16) & 0xFF, (p >> 24) & 0xFF], axis=1).astype(np.uint8)


def work():
    out = (C.c_longlong * 7)()
    lib().bbh_work(out)
    return dict(zip(("valid_calls", "fast_accept", "fast_reject", "pack_iters", "clear_iters", "slow", "branches"), list(out)))


def work_reset():
    lib().bbh_work_reset()
