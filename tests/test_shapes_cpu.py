"""The shape header the env kernels are specialised on (csrc/env_shapes_gen.h) must be the one the generator produces
from the shipped configuration files - a stale header would silently route a shipped configuration to the generic
kernels (or, worse, describe a structure no configuration has)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generated_shapes_are_current():
    spec = importlib.util.spec_from_file_location("gen_env_shapes", os.path.join(ROOT, "scripts", "gen_env_shapes.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    with open(g.OUT) as f:
        assert f.read() == g.render(), "run `python scripts/gen_env_shapes.py` and rebuild"
