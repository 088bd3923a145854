"""Upper bound of longest-first scheduling on the real kernel: run an env step, record every aircraft's attempt count k,
restore the state, and run the SAME step again with the adoption order sorted by k (descending / ascending / random).
Only the attempt kernel's duration changes.  DESIGN.md 4.4."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fwgym_b200 import FixedWingVecEnv, _capi
from fwgym_b200.config import DEFAULT_ENV_CONFIG
n = 65536
vec = FixedWingVecEnv(DEFAULT_ENV_CONFIG, n, config_kw=bench.CONFIG_KW, sim_config_kw=bench.SIM_KW, seed=1)
vec.reset()
acts = torch.rand((24, n, 3), device="cuda") * 2 - 1
for i in range(8):
    vec.step_tensors(acts[i])
lib = vec._lib
res = {"natural": [], "k descending (oracle LPT)": [], "k ascending": [], "random": []}
for t in range(8, 20):
    snap = vec.get_state().clone()
    vec.step_tensors(acts[t])
    k = vec.last_attempts().clone()
    orders = {"natural": None, "k descending (oracle LPT)": torch.argsort(-k, stable=True).int(),
              "k ascending": torch.argsort(k, stable=True).int(), "random": torch.randperm(n, device="cuda").int()}
    for name, o in orders.items():
        vec.set_state(snap)
        _capi.check(lib.fw_debug_set_order(vec._h, ctypes.c_void_p(o.data_ptr() if o is not None else 0)))
        vec.set_profiling(True)
        vec.step_tensors(acts[t])
        d, e, _ = vec.profile()
        vec.set_profiling(False)
        res[name].append(d * 1e3)
    _capi.check(lib.fw_debug_set_order(vec._h, None))
    vec.set_state(snap)
    vec.step_tensors(acts[t])
for name, v in res.items():
    print("%-28s dynamics kernels %.1f us (min %.1f max %.1f)" % (name, sum(v) / len(v), min(v), max(v)))
