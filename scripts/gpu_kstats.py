"""dopri5 attempt counts per aircraft along the bench workload: distribution, tail, and how well last step's count
predicts this step's (the priority start of the attempt kernel uses it).  DESIGN.md 4.4."""
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
acts = torch.rand((80, n, 3), device="cuda") * 2 - 1
prev = None
for t in range(80):
    vec.set_profiling(True)
    vec.step_tensors(acts[t])
    d, e, _ = vec.profile()
    vec.set_profiling(False)
    k = vec.last_attempts().clone().float()
    if t % 5 == 4 or t < 3:
        line = "step %2d dyn %.1f us | k mean %.2f p99 %.0f p99.9 %.0f max %.0f | k>=8: %d k>=12: %d" % (
            t, d * 1e3, k.mean(), k.quantile(0.99), k.quantile(0.999), k.max(), int((k >= 8).sum()), int((k >= 12).sum()))
        if prev is not None:
            big = k >= 8
            line += " | of k>=8 now, prev k>=6: %.2f; corr %.2f" % (float((prev[big] >= 6).float().mean()) if big.any() else 0.0,
                                                                      float(torch.corrcoef(torch.stack([k, prev]))[0, 1]))
        print(line)
    prev = k
