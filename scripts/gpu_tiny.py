import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fwgym_b200 import FixedWingVecEnv
from fwgym_b200.config import DEFAULT_ENV_CONFIG
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
v = FixedWingVecEnv(DEFAULT_ENV_CONFIG, n, sim_config_kw={"turbulence": False}, seed=1)
v.reset()
a = torch.rand((n, 3), device="cuda") * 2 - 1
for _ in range(3):
    v.step_tensors(a)
torch.cuda.synchronize()
print("ok", v.counters())
