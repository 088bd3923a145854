"""Oracle-side statistics of the dopri5 attempt count per env step (k) on the bench workload: histogram, relation to the
first accepted step, autocorrelation.  Test tooling (drives the CPU oracle); used for DESIGN.md 4.4."""
import os, sys, numpy as np, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import scipy.integrate
from oracle import harness
import bench
rec = []
orig = scipy.integrate.solve_ivp
def patched(fun, t_span, y0, **kw):
    sol = orig(fun, t_span, y0, **kw)
    # sol.t = [0, t1, t2, ... dt]; first accepted step size t1 (<= h0 after rejections)
    rec.append((sol.t[1] - sol.t[0], (sol.nfev - 2) // 6, len(sol.t) - 1))
    return sol
scipy.integrate.solve_ivp = patched
import oracle.pyfly_restated as pr
pr.scipy.integrate.solve_ivp = patched
for wid in range(6):
    env = harness.make_env("restated", harness.config_path(), bench.CONFIG_KW, bench.SIM_KW)
    run = harness.OracleRunner(env, seed=1234, env_id=wid)
    run.reset()
    rng = np.random.RandomState(wid)
    for _ in range(250):
        run.step(rng.uniform(-1, 1, 3))
r = np.array(rec)
h1, k, acc = r[:, 0], r[:, 1].astype(int), r[:, 2].astype(int)
print("n", len(r), "mean k", k.mean(), "hist", collections.Counter(k.tolist()))
print("rejections", collections.Counter((k - acc).tolist()))
for kk in sorted(set(k.tolist())):
    m = k == kk
    print(kk, m.sum(), "first accepted step: min %.2e med %.2e max %.2e" % (h1[m].min(), np.median(h1[m]), h1[m].max()))
# autocorrelation of k between consecutive steps of the same env (records are sequential per env, 250 per env + reset steps)
kk = k
a, b = kk[:-1], kk[1:]
print("corr(k_t, k_t+1) =", np.corrcoef(a, b)[0, 1])
for v in sorted(set(a.tolist())):
    m = a == v
    print("k_t=%d -> mean k_t+1 %.2f  P(k_t+1>=5)=%.2f n=%d" % (v, b[m].mean(), (b[m] >= 5).mean(), m.sum()))
