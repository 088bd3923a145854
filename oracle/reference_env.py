"""CPU ORACLE (test infrastructure): run the UNMODIFIED reference env `gym_fixed_wing/fixed_wing.py` in this container.

The reference file imports `gym`, `matplotlib` and `pyfly` (fixed_wing.py:1-7), none of which is installed.  This
module registers in-memory stand-ins for exactly the names the file touches, with `pyfly.pyfly.PyFly` bound to the
restated simulator (oracle/pyfly_restated.py), and then imports the reference file from where it lies under
/root/reference.  Nothing is copied.  It is used (a) to generate the golden fixtures under tests/golden/ (script:
oracle/make_golden.py) and (b) by the CPU tests that check oracle/env_restated.py against the reference's own code.
/root/reference does not exist on the GPU box: callers must handle `reference_available() == False`.
"""
import importlib.util
import os
import sys
import types

import numpy as np

from . import pyfly_restated

REFERENCE_ROOT = os.environ.get("FWGYM_REFERENCE_ROOT", "/root/reference")
_REF_FILE = os.path.join(REFERENCE_ROOT, "gym_fixed_wing", "fixed_wing.py")
_module = None


def reference_available():
    return os.path.isfile(_REF_FILE)


def reference_config_path(name="fixed_wing_config.json"):
    return os.path.join(REFERENCE_ROOT, "gym_fixed_wing", name)


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.low = np.asarray(low, dtype=np.float64)
        self.high = np.asarray(high, dtype=np.float64)
        if shape is not None and self.low.shape != tuple(shape):
            self.low = np.full(shape, low, dtype=np.float64)
            self.high = np.full(shape, high, dtype=np.float64)
        self.shape = self.low.shape
        self.dtype = np.dtype(dtype)


class _Dict:
    def __init__(self, spaces):
        self.spaces = spaces


def _np_random(seed=None):
    # gym's real implementation hashes the seed; parity runs inject their own stream instead.
    return np.random.RandomState(seed), seed


def _install_shims():
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    if "gym" not in sys.modules:
        gym = mod("gym")
        gym.Env = type("Env", (), {})
        gym.GoalEnv = type("GoalEnv", (), {})
        gym.spaces = mod("gym.spaces")
        gym.spaces.Box = _Box
        gym.spaces.Dict = _Dict
        gym.utils = mod("gym.utils")
        gym.utils.seeding = mod("gym.utils.seeding")
        gym.utils.seeding.np_random = _np_random
    if "matplotlib" not in sys.modules:
        mpl = mod("matplotlib")
        mpl.pyplot = mod("matplotlib.pyplot")
        mpl.gridspec = mod("matplotlib.gridspec")
    pyfly = mod("pyfly")
    pyfly.pyfly = mod("pyfly.pyfly")
    pyfly.pyfly.PyFly = pyfly_restated.PyFly
    pyfly.pid_controller = mod("pyfly.pid_controller")
    pyfly.pid_controller.PIDController = pyfly_restated.PIDController


def load_reference_module():
    """Import /root/reference/gym_fixed_wing/fixed_wing.py (unmodified) over the stand-ins."""
    global _module
    if _module is not None:
        return _module
    if not reference_available():
        raise FileNotFoundError(_REF_FILE)
    _install_shims()
    spec = importlib.util.spec_from_file_location("_reference_fixed_wing", _REF_FILE)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    _module = m
    return m


def make_reference_env(config_path=None, config_kw=None, sim_config_kw=None):
    m = load_reference_module()
    if config_path is None:
        config_path = reference_config_path()
    return m.FixedWingAircraft(config_path, config_kw=config_kw, sim_config_kw=sim_config_kw)
