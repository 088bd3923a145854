#!/usr/bin/env python
"""The reference's examples/evaluate_controller.py on the GPU: every scenario of a test set flown at once, one env each
(fwgym_b200.evaluate).

    python examples/evaluate_controller.py tests/golden/test_set_wind_none.npz --PID
    python examples/evaluate_controller.py tests/golden/test_set_wind_none.npz --model-path tests/golden/mlp_controller.npz
    python examples/evaluate_controller.py path/to/test_set.npy --PID --turbulence-intensity moderate

`path_to_file`: the reference's pickled test set (.npy, list of {"state", "target"}) or the plain-array fixture
(.npz).  `--model-path`: an .npz with the arrays of a stable-baselines PPO2 MlpPolicy and its VecNormalize statistics
(oracle/make_golden_policy.py writes one from examples/models/mlp_controller).  Prints the reference's result table
(nan-mean per metric and state, evaluate_controller.py:32-41) and optionally saves the result dictionary in the
reference's layout (res[metric][state] = [value per scenario], res["rewards"])."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("path_to_file", help="test set (.npy as the reference ships them, or the .npz fixture)")
    ap.add_argument("--PID", dest="use_pid", action="store_true", help="use the PID controller")
    ap.add_argument("--model-path", help=".npz with the MlpPolicy arrays (tests/golden/mlp_controller.npz)")
    ap.add_argument("--env-config-path", help="env configuration JSON (default: the examples configuration)")
    ap.add_argument("--turbulence-intensity", default="none", choices=["none", "light", "moderate", "severe"])
    ap.add_argument("--save", help="write the result dictionary to this .npy")
    a = ap.parse_args()
    if not a.use_pid and not a.model_path:
        ap.error("give --PID or --model-path")
    import __graft_entry__ as ge
    ge.build()
    from fwgym_b200 import evaluate
    from fwgym_b200.config import DEFAULT_ENV_CONFIG
    cfg = a.env_config_path or os.path.join(os.path.dirname(DEFAULT_ENV_CONFIG), "fixed_wing_config_examples.json")
    scenarios = evaluate.load_test_set(a.path_to_file)
    controller = "pid" if a.use_pid else dict(np.load(a.model_path))
    res, vec = evaluate.evaluate_on_set(scenarios, cfg, controller=controller,
                                        turbulence_intensity=a.turbulence_intensity)
    print("%d scenarios, mean episode length %.1f steps" % (len(scenarios), float(np.mean(res["lengths"]))))
    for key, val in sorted(evaluate.summarise(res).items()):
        print("\t%-28s %.4f" % (key, val))
    if a.save:
        out = {m: res[m] for m in evaluate.DEFAULT_METRICS}
        out["rewards"] = [np.asarray(r) for r in res["rewards"]]
        np.save(a.save, out, allow_pickle=True)
    vec.close()
