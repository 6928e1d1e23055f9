"""Import alias: the package sources live in ``driving-dirty_b200/`` (the repository's layout
contract); a hyphen is not importable, so this stub points ``driving_dirty_b200`` at that
directory and runs its ``__init__``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "driving-dirty_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f, _real
