"""Importable alias of the product package.

The package directory required by the repo layout is ``rlaif-v_b200/`` (hyphenated, hence not a
Python identifier).  This shim makes it importable as ``rlaifv_b200``: submodules resolve from
``rlaif-v_b200/`` and its ``__init__`` is executed in this module's namespace.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rlaif-v_b200")
__path__ = [_real]
_init = _os.path.join(_real, "__init__.py")
with open(_init) as _f:
    exec(compile(_f.read(), _init, "exec"))
del _f, _init
