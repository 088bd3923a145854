import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fwgym_b200 import FixedWingVecEnv
from fwgym_b200.config import DEFAULT_ENV_CONFIG
cfg = os.path.join(os.path.dirname(DEFAULT_ENV_CONFIG), "fixed_wing_config_dev.json")
vec = FixedWingVecEnv(cfg, 64, sim_config_kw={"turbulence": False}, seed=3)
vec.reset()
st, rows = vec.get_state(), vec.state_rows()
for k in ("omega_p", "omega_q", "omega_r"):
    st[rows.index(k), :8] = 1e153
vec.set_state(st)
acts = torch.zeros((64, 3), dtype=torch.float64, device=vec.device)
for t in range(3):
    obs, rew, done, term = vec.step_tensors(acts)
    s2 = vec.get_state()
    print(t, "done", done[:10].tolist(), "term", term[:10].tolist(), "k", vec.last_attempts()[:10].tolist())
    print("   omega_p", s2[rows.index("omega_p"), :3].tolist(), "u", s2[rows.index("velocity_u"), :3].tolist(), "status", s2[rows.index("sim_status"), :3].tolist())
print(vec.kernel_variant(), vec.counters())
