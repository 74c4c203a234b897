"""Importable alias of the ``fpl-plus_b200`` package directory (a hyphen is not a valid
module name; the sources live in ../fpl-plus_b200)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "fpl-plus_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
