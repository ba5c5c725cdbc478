"""bbgpu — B200-native batched Block Blast simulator and masked-PPO rollout path.

Drop-in for the reference's hot path only: ``VectorizedBlockBlastEnv`` /
``BlockBlastEnv`` reset/step/action-mask, ``RolloutBuffer``, ``PPOAgent`` collect/update.
The compute is hand-written sm_100a CUDA behind a C ABI (``include/bbgpu.h``,
``csrc/`` -> ``libbbgpu.so``); this Python layer only mirrors the reference's class API on
top of it.  There is no CPU fallback: every op raises if the CUDA library is missing.

Heavy modules (torch, the CUDA library) load lazily on first attribute access.
"""
__version__ = "0.1.0"

_LAZY = {
    "VectorizedBlockBlastEnv": "vec_env",
    "BlockBlastEnv": "vec_env",
    "BlockBlastEnvFlat": "vec_env",
    "GameState": "vec_env",
    "Logger": "logger",
    "TensorBoardLogger": "logger",
    "MetricsTracker": "logger",
    "RolloutBuffer": "rollout",
    "PPOAgent": "ppo",
    "PPOConfig": "ppo",
    "BlockBlastNetwork": "network",
}


def __getattr__(name):
    import importlib
    if name in _LAZY:
        mod = importlib.import_module("." + _LAZY[name], __name__)
        return getattr(mod, name)
    if name in ("philox", "capi", "vec_env", "rollout", "ppo", "network", "build", "dist", "train", "evaluate", "logger"):
        return importlib.import_module("." + name, __name__)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
