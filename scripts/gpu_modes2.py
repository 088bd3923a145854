import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from fwgym_b200 import FixedWingVecEnv
from fwgym_b200.config import DEFAULT_ENV_CONFIG
n = 65536
P = os.path.dirname(DEFAULT_ENV_CONFIG)
for name, cfg in (("shipped", DEFAULT_ENV_CONFIG), ("rand", os.path.join(P, "fixed_wing_config_randomised.json"))):
    vec = FixedWingVecEnv(cfg, n, config_kw=bench.CONFIG_KW, sim_config_kw=bench.SIM_KW, seed=1)
    vec.reset()
    acts = torch.rand((40, n, 3), device="cuda") * 2 - 1
    for i in range(5):
        vec.step_tensors(acts[i])
    vec.reset_counters(); vec.set_profiling(True)
    for i in range(20):
        vec.step_tensors(acts[5 + i])
    d, e, k = vec.profile()
    c = vec.counters()
    print(name, vec.kernel_variant(), "dyn %.1f us env %.1f us" % (d / k * 1e3, e / k * 1e3), "lane_eff %.3f passes %d failures %d resets %d" % (c["warp_steps"] / (32.0 * c["warp_max_attempts"]), c["warp_max_attempts"] / 20, c["failures"], c["resets"]))
    vec.close()
