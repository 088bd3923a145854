"""Is the previous step's dopri5 attempt count a predictor of this step's (bench workload, stationary mix)?"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fwgym_b200 import FixedWingVecEnv
from fwgym_b200.config import DEFAULT_ENV_CONFIG
n = 65536
vec = FixedWingVecEnv(DEFAULT_ENV_CONFIG, n, config_kw=bench.CONFIG_KW, sim_config_kw=bench.SIM_KW, seed=20261017)
vec.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
ks = []
for t in range(240):
    a = torch.rand((n, 3), generator=g, device="cuda") * 2 - 1
    _, _, done, _ = vec.step_tensors(a)
    if t >= 200:
        ks.append(vec.last_attempts().cpu().numpy().copy())
ks = np.array(ks).astype(np.float64)
c = [np.corrcoef(ks[i], ks[i + 1])[0, 1] for i in range(len(ks) - 1)]
print("corr(k_t, k_t+1): mean %.3f" % np.mean(c))
a, b = ks[:-1].ravel(), ks[1:].ravel()
for thr in (5, 6, 7):
    print("P(k_t+1 >= %d) = %.4f ; P(k_t+1 >= %d | k_t >= %d) = %.4f ; P(k_t >= %d | k_t+1 >= %d) = %.4f"
          % (thr, (b >= thr).mean(), thr, thr, (b[a >= thr] >= thr).mean(), thr, thr, (a[b >= thr] >= thr).mean()))
print("mean k_t+1 by k_t:", {int(k): round(float(b[a == k].mean()), 2) for k in range(1, 10) if (a == k).sum() > 100})
# smooth actions (a policy) instead of iid random ones: same question
vec.reset()
act = torch.zeros((n, 3), device="cuda")
ks = []
for t in range(240):
    act = 0.95 * act + 0.05 * (torch.rand((n, 3), generator=g, device="cuda") * 2 - 1) * 3
    vec.step_tensors(act.clamp(-1, 1))
    if t >= 200:
        ks.append(vec.last_attempts().cpu().numpy().copy())
ks = np.array(ks).astype(np.float64)
print("smooth actions: corr %.3f, mean k %.2f" % (np.mean([np.corrcoef(ks[i], ks[i + 1])[0, 1] for i in range(len(ks) - 1)]), ks.mean()))
