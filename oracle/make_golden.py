"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference env
(/root/reference/gym_fixed_wing/fixed_wing.py over the stand-ins of oracle/reference_env.py, physics from the restated
PyFly) on Philox streams.  Build-container only.   python -m oracle.make_golden
"""
import os

import numpy as np

from . import harness
from .cases import CASES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
SEED = 77


def case_actions(name, c):
    rng = np.random.RandomState(abs(hash(name)) % (2 ** 31) if False else sum(map(ord, name)))
    return rng.uniform(-c["amp"], c["amp"], (c["steps"], c["n"], 3))


def run_case(name, c, kind="reference"):
    runners = [harness.OracleRunner(harness.make_env(kind, harness.config_path(c["config"]), c["config_kw"], c["sim_kw"]),
                                    SEED, i) for i in range(c["n"])]
    acts = case_actions(name, c)
    obs = [np.stack([np.asarray(r.reset(), dtype=np.float64).ravel() for r in runners])]
    rew, done, k, state, term_obs = [], [], [], [], []
    for a in acts:
        res = [r.step(a[i]) for i, r in enumerate(runners)]
        obs.append(np.stack([np.asarray(x[0], dtype=np.float64).ravel() for x in res]))
        rew.append([float(x[1]) for x in res])
        done.append([bool(x[2]) for x in res])
        k.append([r.attempts_last() for r in runners])
        state.append(np.stack([r.ode_state() for r in runners]))
        term_obs.append(np.stack([np.asarray(x[3].get("terminal_observation", np.full_like(obs[0][0], np.nan)),
                                             dtype=np.float64).ravel() for x in res]))
    return dict(actions=acts, obs=np.array(obs), rew=np.array(rew), done=np.array(done), k=np.array(k),
                state=np.array(state), term_obs=np.array(term_obs))


def main():
    for name, c in CASES.items():
        out = run_case(name, c)
        np.savez_compressed(os.path.join(GOLDEN, "case_%s.npz" % name), **out)
        print("%-14s steps %3d envs %d dones %3d mean k %.2f" % (name, c["steps"], c["n"], out["done"].sum(), out["k"].mean()))


if __name__ == "__main__":
    main()
