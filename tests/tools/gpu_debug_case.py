import sys, os
_R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import numpy as np, torch
import __graft_entry__ as ge; ge.build()
from fwgym_b200 import FixedWingVecEnv
from oracle import harness
from oracle.cases import CASES
from oracle.make_golden import SEED
import parity_utils as pu
name = sys.argv[1]
c = CASES[name]
g = np.load(os.path.join("tests/golden", "case_%s.npz" % name))
vec = FixedWingVecEnv(harness.config_path(c["config"]), c["n"], config_kw=c["config_kw"], sim_config_kw=c["sim_kw"], seed=SEED, keep_terminal_obs=True)
vec.enable_f64_outputs(True)
vec.reset()
print("reset obs err", pu.rel_err(vec._obs64.cpu().numpy(), g["obs"][0], 1e-3).max())
for t, a in enumerate(g["actions"]):
    _, _, done, term = vec.step_tensors(torch.as_tensor(a, dtype=torch.float64, device=vec.device))
    d = done.cpu().numpy().astype(bool)
    k = vec.last_attempts().cpu().numpy()
    st = pu.rel_err(pu.gpu_state(vec), g["state"][t], 1e-3).max(axis=1)
    print(t, "done", d.astype(int), "gold", g["done"][t].astype(int), "term", term.cpu().numpy(), "k", k, "gk", g["k"][t], "state err", np.array2string(st, precision=1))
    if not np.array_equal(d, g["done"][t]):
        break
