"""fwgym_b200 — B200-native batched fixed-wing simulator (hot path of eivindeb/fixed-wing-gym).

The directory is named `fixed-wing-gym_b200/` (not an importable identifier); import it as `fwgym_b200` through the
alias package at the repo root.
"""
from .config import CompiledConfig, ConfigError  # noqa: F401
from .vec_env import FixedWingVecEnv, HostStepper, Box  # noqa: F401
from .env import FixedWingAircraft  # noqa: F401
