"""Importable alias for the package directory
``block-blast-ai---reinforcement-learning-agent_b200/`` (its name is not a Python identifier).

``import bbgpu`` executes that directory's ``__init__.py`` with this module's ``__path__``
pointing at it, so ``bbgpu.philox``, ``bbgpu.capi`` ... are the files in the package directory.
"""
import os as _os

_PKG_DIR = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                         "block-blast-ai---reinforcement-learning-agent_b200")
__path__ = [_PKG_DIR]
_init = _os.path.join(_PKG_DIR, "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"), globals())
