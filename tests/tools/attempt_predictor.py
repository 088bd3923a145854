"""How well is the dopri5 attempt count of an env step predictable after its FIRST attempt?  Drives the CPU oracle with a
manual RK45 loop; writes /tmp/kpred_pairs.npy for scripts/sched_sim_sorted.py.  Test tooling; DESIGN.md 4.4."""
import os, sys, numpy as np, collections, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import scipy.integrate
from scipy.integrate import RK45
from oracle import harness
import bench
rec = []
orig = scipy.integrate.solve_ivp
def patched(fun, t_span, y0, **kw):
    # shadow run with the same fun is NOT possible (fun has side effects on PyFly state) -> do the real run manually
    s = RK45(fun, t_span[0], y0, t_span[1])
    first = None
    while s.status == "running":
        s.step()
        if first is None:
            first = (s.t, s.h_abs, (s.nfev - 2) // 6)
    k = (s.nfev - 2) // 6
    rec.append((first[0], first[1], first[2], k))
    class R: pass
    r = R(); r.y = np.array([s.y]).T; r.t = np.array([t_span[0], s.t]); r.nfev = s.nfev; r.status = 0; r.success = True
    return r
import oracle.pyfly_restated as pr
pr.scipy.integrate.solve_ivp = patched
for wid in range(6):
    env = harness.make_env("restated", harness.config_path(), bench.CONFIG_KW, bench.SIM_KW)
    run = harness.OracleRunner(env, seed=1234, env_id=wid)
    run.reset()
    rng = np.random.RandomState(wid)
    for _ in range(250):
        run.step(rng.uniform(-1, 1, 3))
r = np.array(rec); dt = 0.01
t1, h2, k1, k = r[:, 0], r[:, 1], r[:, 2].astype(int), r[:, 3].astype(int)
rem = np.where(t1 >= dt, 0, np.ceil((dt - t1) / h2 - 1e-9)).astype(int)
pred = k1 + rem
print("mean k", k.mean(), "corr(pred,k)", np.corrcoef(pred, k)[0, 1])
print("exact %.3f  within1 %.3f  under-predicted %.3f" % ((pred == k).mean(), (abs(pred - k) <= 1).mean(), (pred < k).mean()))
c = collections.Counter(zip(pred.tolist(), k.tolist()))
for (p_, k_), n in sorted(c.items()): print("pred", p_, "actual", k_, n)
np.save("/tmp/kpred_pairs.npy", np.stack([pred, k], 1))
