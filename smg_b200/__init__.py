"""Importable alias of the product package.

The package directory is `smg-multimodal-grasping_b200/` (the name the repo
layout prescribes); a hyphen cannot appear in an `import` statement, so this
alias points its `__path__` at that directory and runs its `__init__`.
Use `import smg_b200` everywhere.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "smg-multimodal-grasping_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _f
