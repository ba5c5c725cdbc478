"""Import stub standing in for `gymnasium` (absent from this image, no network).

Test infrastructure only: lets tests/golden/make_golden.py import the *reference*
environment classes, which only subclass ``gym.Env`` and declare spaces.  Never put this
directory on a product path.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def reset(self, seed=None, options=None):
        return None

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        return getattr(self.env, name)


def register(*args, **kwargs):
    return None
