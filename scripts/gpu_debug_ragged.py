"""Which env / step / quantity exceeds the parity tolerance at a ragged env count (debug helper)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import parity_utils as pu
from oracle import harness
from oracle.cases import CASES
from fwgym_b200 import FixedWingVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 197
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 9
c = CASES["turb_noise"]
vec = FixedWingVecEnv(harness.config_path(c["config"]), n, config_kw=c["config_kw"], sim_config_kw=c["sim_kw"], seed=seed,
                      keep_terminal_obs=True)
orc = pu.make_oracles(n, harness.config_path(c["config"]), c["config_kw"], c["sim_kw"], seed)
acts = np.random.RandomState(n).uniform(-1, 1, (6, n, 3))
vec.enable_f64_outputs(True)
vec.reset()
obs_o = np.stack([np.asarray(o.reset(), dtype=np.float64).ravel() for o in orc])
e = pu.rel_err(vec._obs64.cpu().numpy(), obs_o, 1e-3)
print("reset obs err max %.3e at env %d" % (e.max(), np.unravel_index(e.argmax(), e.shape)[0]))
for t, a in enumerate(acts):
    _, _, done_g, term_g = vec.step_tensors(torch.as_tensor(a, dtype=torch.float64, device=vec.device))
    res = [o.step(a[i]) for i, o in enumerate(orc)]
    obs_o = np.stack([np.asarray(r[0], dtype=np.float64).ravel() for r in res])
    rew_o = np.array([r[1] for r in res])
    k_o = np.array([o.attempts_last() for o in orc]); k_g = vec.last_attempts().cpu().numpy()
    so = np.stack([o.ode_state() for o in orc]); sg = pu.gpu_state(vec)
    eo = pu.rel_err(vec._obs64.cpu().numpy(), obs_o, 1e-3); es = pu.rel_err(sg, so, 1e-3); er = pu.rel_err(vec._rew64.cpu().numpy(), rew_o, 1e-3)
    bad = np.where((eo.max(1) > 1e-9) | (es.max(1) > 1e-9))[0]
    print("step %d: obs %.3e (env %d) state %.3e (env %d, row %d) rew %.3e k_mismatch %s done g/o %d/%d bad envs %s"
          % (t, eo.max(), eo.max(1).argmax(), es.max(), es.max(1).argmax(), es[es.max(1).argmax()].argmax(), er.max(),
             np.where(k_g != k_o)[0].tolist(), int(done_g.sum()), sum(bool(r[2]) for r in res), bad.tolist()[:12]))
    for i in bad[:3]:
        print("   env %d: k g/o %d/%d state g %s\n            state o %s" % (i, k_g[i], k_o[i], np.array2string(sg[i][:13], precision=6), np.array2string(so[i][:13], precision=6)))
