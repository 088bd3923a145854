"""Import alias: `fwgym_b200` -> ../fixed-wing-gym_b200/ (the directory name required by the layout contract is not
a valid Python identifier)."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "fixed-wing-gym_b200")
__path__.insert(0, _real)
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
