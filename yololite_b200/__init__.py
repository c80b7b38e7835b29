"""Importable alias of the package directory ``yololite-official-repo_b200/`` (a hyphen is not a valid
module name).  ``import yololite_b200`` resolves every submodule from that directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "yololite-official-repo_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
