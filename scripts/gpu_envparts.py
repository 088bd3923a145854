"""Diagnosis: env-kernel / dynamics time per step for config variants (noise, turbulence, obs length)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as ge
ge.build()
from fwgym_b200 import FixedWingVecEnv
from fwgym_b200.config import DEFAULT_ENV_CONFIG
N = 65536
variants = {
    "turb+noise": ({"observation": {"noise": {"mean": 0, "var": 0.1}}}, {"turbulence": True, "turbulence_intensity": "moderate"}),
    "turb only": (None, {"turbulence": True, "turbulence_intensity": "moderate"}),
    "noise only": ({"observation": {"noise": {"mean": 0, "var": 0.1}}}, {"turbulence": False}),
    "neither": (None, {"turbulence": False}),
}
for name, (ck, sk) in variants.items():
    vec = FixedWingVecEnv(DEFAULT_ENV_CONFIG, N, config_kw=ck, sim_config_kw=sk, seed=3)
    vec.reset()
    acts = torch.rand((40, N, 3), device="cuda") * 2 - 1
    for i in range(5):
        vec.step_tensors(acts[i])
    vec.set_profiling(True)
    for i in range(5, 35):
        vec.step_tensors(acts[i])
    d, e, n = vec.profile()
    print("%-12s dyn %.4f ms env %.4f ms" % (name, d / n, e / n), flush=True)
    vec.close()
