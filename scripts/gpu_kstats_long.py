"""Stationary behaviour of the bench workload: per block of 50 env steps, the dynamics-kernel time and the dopri5
stragglers (max attempts of any aircraft in a step).  DESIGN.md 4.4."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fwgym_b200 import FixedWingVecEnv
from fwgym_b200.config import DEFAULT_ENV_CONFIG
n = 65536
vec = FixedWingVecEnv(DEFAULT_ENV_CONFIG, n, config_kw=bench.CONFIG_KW, sim_config_kw=bench.SIM_KW, seed=1)
vec.reset()
torch.manual_seed(0)
acts = torch.rand((16, n, 3), device="cuda") * 2 - 1
T = int(sys.argv[1]) if len(sys.argv) > 1 else 600
dyn, kmax = [], []
for t in range(T):
    vec.set_profiling(True)
    vec.step_tensors(acts[t % 16])
    d, e, _ = vec.profile()
    vec.set_profiling(False)
    dyn.append(d * 1e3)
    kmax.append(int(vec.last_attempts().max()))
    if t % 50 == 49:
        dd, kk = torch.tensor(dyn[-50:]), torch.tensor(kmax[-50:])
        print("steps %3d-%3d dyn mean %.1f median %.1f max %.1f us | max k: median %d max %d, steps with k>=12: %d, k>=20: %d"
              % (t - 49, t, dd.mean(), dd.median(), dd.max(), int(kk.median()), int(kk.max()), int((kk >= 12).sum()), int((kk >= 20).sum())))
