"""Golden fixtures for the episode metrics: FixedWingAircraft.get_metric of the UNMODIFIED reference
file (fixed_wing.py:1095-1162), evaluated at every episode end of the parity cases that finish episodes.
Build-container only.   python -m oracle.make_golden_metrics  ->  tests/golden/metrics_<case>.npz

Row layout = the device's (csrc/layout.h EP_*): return (sum of the rewards step() returned, Monitor's "r"), length, control_variation, success_all, settling_time_all,
success_time_frac_all, then per target: avg_error, total_error, end_error, rise_time, overshoot, success,
settling_time, success_time_frac.  NaN where the reference yields nan or has no entry for that state.
"""
import os
import warnings

import numpy as np

from . import harness
from .cases import CASES
from .make_golden import GOLDEN, SEED, case_actions

METRIC_CASES = ["failure", "success_done", "norm_step2", "dev_history"]
PER_TARGET = ["avg_error", "total_error", "end_error", "rise_time", "overshoot", "success", "settling_time",
              "success_time_frac"]


def episode_row(env, rise_kw, ep_return):
    tn = list(env.target.keys())
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = {name: env.get_metric(name, **(rise_kw if name == "rise_time" else {})) for name in PER_TARGET}
        cv = env.get_metric("control_variation")["all"]
    goal = env.goal_enabled
    row = [float(ep_return), float(env.steps_count), float(cv),
           float(m["success"]["all"]) if goal else np.nan,
           float(m["settling_time"]["all"]) if goal else np.nan,
           float(m["success_time_frac"]["all"]) if goal else np.nan]
    for t in tn:
        for name in PER_TARGET:
            v = m[name].get(t, np.nan)
            row.append(float(v))
    return row


def run_case(name, c):
    rows = []
    runners = []
    for i in range(c["n"]):
        r = harness.OracleRunner(harness.make_env("reference", harness.config_path(c["config"]), c["config_kw"],
                                                  c["sim_kw"]), SEED, i)
        rise_kw = {}
        for mm in r.env.cfg.get("metrics", []):
            if mm["name"] == "rise_time":
                rise_kw = {"low": mm.get("low", 0.1), "high": mm.get("high", 0.9)}
        state = {"t": 0}
        r.on_done = (lambda env, info, i=i, state=state, rise_kw=rise_kw, r=r:
                     rows.append([state["t"], i] + episode_row(env, rise_kw, r.ep_return)))
        r._state = state
        runners.append(r)
    acts = case_actions(name, c)
    for r in runners:
        r.reset()
    for t, a in enumerate(acts):
        for i, r in enumerate(runners):
            r._state["t"] = t
            r.step(a[i])
    return np.array(rows, dtype=np.float64)


def main():
    for name in METRIC_CASES:
        rows = run_case(name, CASES[name])
        np.savez_compressed(os.path.join(GOLDEN, "metrics_%s.npz" % name), rows=rows)
        print("%-14s episodes %d" % (name, len(rows)))


if __name__ == "__main__":
    main()
