"""Which aircraft become dopri5 stragglers (k >= 12 attempts in one env step)?  Stationary bench workload; state at the
START of the step against the attempt count of the step.  DESIGN.md 4.4."""
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
for t in range(200):
    vec.step_tensors(acts[t % 16])
names = ["Va", "alpha", "beta", "omega_p", "omega_q", "omega_r", "roll", "pitch"]
S, K = [], []
for t in range(200, 320):
    st = vec.get_named_state(names)
    S.append(torch.stack([st[nm] for nm in names]).clone())      # [8, n] at the start of the step
    vec.step_tensors(acts[t % 16])
    K.append(vec.last_attempts().clone())
S = torch.stack(S, 1).reshape(len(names), -1)
K = torch.stack(K).reshape(-1).float()
big = K >= 12
print("aircraft-steps %d, stragglers (k>=12) %d, k>=20 %d, max %d" % (K.numel(), int(big.sum()), int((K >= 20).sum()), int(K.max())))
for i, nm in enumerate(names):
    x = S[i]
    print("%-8s all: mean %.3f p1 %.3f p99 %.3f | stragglers: mean %.3f min %.3f max %.3f" % (
        nm, x.mean(), x.float().quantile(0.01), x.float().quantile(0.99), x[big].mean(), x[big].min(), x[big].max()))
Va, al = S[0], S[1].abs()
om = S[3:6].abs().max(0).values
for thr in (3, 5, 8, 10):
    sel = Va < thr
    print("Va < %2d: selects %.3f%% of aircraft-steps, recall of k>=12 %.2f, of k>=20 %.2f" % (
        thr, 100 * sel.float().mean(), float(sel[big].float().mean()), float(sel[K >= 20].float().mean())))
for thr in (0.5, 1.0, 1.5):
    sel = al > thr
    print("|alpha| > %.1f: selects %.3f%%, recall k>=12 %.2f, k>=20 %.2f" % (thr, 100 * sel.float().mean(), float(sel[big].float().mean()), float(sel[K >= 20].float().mean())))
for thr in (3.0, 5.0, 8.0):
    sel = om > thr
    print("max|omega| > %.0f: selects %.3f%%, recall k>=12 %.2f, k>=20 %.2f" % (thr, 100 * sel.float().mean(), float(sel[big].float().mean()), float(sel[K >= 20].float().mean())))
