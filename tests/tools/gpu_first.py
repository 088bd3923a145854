import os
import sys, time, numpy as np, torch
_R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import fwgym_b200
from fwgym_b200 import FixedWingVecEnv
from oracle import harness
import parity_utils as pu
cfg = harness.config_path()
N=8; T=30
vec = FixedWingVecEnv(cfg, N, sim_config_kw={"turbulence": False}, seed=11)
orc = pu.make_oracles(N, cfg, None, {"turbulence": False}, 11)
rng = np.random.RandomState(1)
acts = rng.uniform(-1.2, 1.2, (T, N, 3))
out = pu.run_parity(vec, orc, acts)
print("obs", np.max(out["obs"]), "rew", np.max(out["rew"]), "state", np.max(out["state"]), {k:v for k,v in out.items() if not isinstance(v,list)})
print("state err per step", np.array(out["state"])[:10])
print(vec.counters())
# throughput smoke
N=65536
vec2 = FixedWingVecEnv(cfg, N, sim_config_kw={"turbulence": False}, seed=1)
vec2.reset()
a = torch.rand((N,3), device="cuda")*2-1
for _ in range(5): vec2.step_tensors(a)
torch.cuda.synchronize(); t0=time.time()
for _ in range(20): vec2.step_tensors(a)
torch.cuda.synchronize(); dt=time.time()-t0
print("env-steps/s", N*20/dt, vec2.counters())
import ctypes
fl=ctypes.c_double(); ms=ctypes.c_double()
from fwgym_b200 import _capi
_capi.lib().fw_dfma_peak(0, ctypes.byref(fl), ctypes.byref(ms)); print("DFMA peak TFLOP/s", fl.value/1e12, ms.value)
