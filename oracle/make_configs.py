"""Regenerate the env-config data files under fixed-wing-gym_b200/params/ from the reference's JSON configs and the
scenario fixture under tests/golden/ from its test set.  Run in the build container only (/root/reference does not
exist on the GPU box).  These are DATA (the reference's config-JSON surface, README.md:21-160), re-serialised
compactly with sorted keys; no reference source code is copied.

    python -m oracle.make_configs
"""
import json
import os

import numpy as np

REF = os.environ.get("FWGYM_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = os.path.join(HERE, "..", "fixed-wing-gym_b200", "params")
GOLDEN = os.path.join(HERE, "..", "tests", "golden")

SOURCES = {
    "fixed_wing_config.json": "gym_fixed_wing/fixed_wing_config.json",
    "fixed_wing_config_dev.json": "gym_fixed_wing/fixed_wing_config_dev.json",
    "fixed_wing_config_examples.json": "gym_fixed_wing/examples/fixed_wing_config.json",
    "fixed_wing_config_cnn.json": "gym_fixed_wing/examples/models/cnn_controller/fixed_wing_config.json",
}


# simulator-parameter randomisation (fixed_wing.py:523-570): no shipped configuration carries a "model" block and
# config_kw cannot add one (set_config_attrs only descends into existing keys), so the parity case has its own file:
# the default configuration + a model block exercising every branch (per-parameter var / clip, a negative original
# with a relative clip, an inertia entry the dynamics never read, a zero original that is skipped) + a simulator
# attribute range.
RANDOMISED_SIMULATOR = {
    "model": {"var_type": "relative", "var": 0.1, "clip": 0.25, "distribution": "gaussian",
              "parameters": [{"name": "mass"}, {"name": "C_L_alpha", "var": 0.2}, {"name": "C_m_q"},
                             {"name": "Jx"}, {"name": "C_D_q"}, {"name": "M", "var": 0.05, "clip": 0.1},
                             {"name": "e"}, {"name": "C_l_p", "clip": 0.05}, {"name": "k_motor"}, {"name": "b"}]},
    "rho": {"low": 1.1, "high": 1.3},
}


def write_randomised():
    with open(os.path.join(PARAMS, "fixed_wing_config.json")) as f:
        cfg = json.load(f)
    cfg["simulator"].update(RANDOMISED_SIMULATOR)
    with open(os.path.join(PARAMS, "fixed_wing_config_randomised.json"), "w") as f:
        json.dump(cfg, f, sort_keys=True, separators=(",", ":"))
        f.write("\n")


# Reward / target "zoo" (fixed_wing.py:495-507, :684-746, :948, :977-985): no shipped configuration uses the reward
# classes action.value / state.value / success / step / goal, the quadratic function class, or linear / sinusoidal
# targets, and config_kw cannot ADD keys, so the parity cases get their own file: the default configuration with one
# reward factor per branch (three terms), and the slope / amplitude / period keys present on every target state so that
# a case can switch a state's class with config_kw.
ZOO_FACTORS = [
    {"class": "state", "type": "error", "name": "roll", "function_class": "linear", "scaling": 3.2, "max": 0.3,
     "shaping": True, "sign": -1},
    {"class": "state", "type": "error", "name": "pitch", "function_class": "quadratic", "scaling": 2.0,
     "shaping": True, "sign": -1},
    {"class": "state", "type": "error", "name": "Va", "function_class": "exponential", "scaling": 400.0,
     "shaping": True, "sign": -1},
    {"class": "action", "type": "delta", "name": "action", "function_class": "linear", "scaling": 60, "window_size": 5,
     "shaping": False, "sign": -1},
    {"class": "action", "type": "bound", "name": "action_bound", "function_class": "linear", "scaling": 1,
     "shaping": False, "sign": -1},
    {"class": "action", "type": "value", "name": "action_value", "function_class": "linear", "scaling": 30,
     "max": 0.08, "shaping": False, "sign": -1},
    {"class": "state", "type": "value", "name": "omega_q", "function_class": "quadratic", "scaling": 10.0,
     "shaping": False, "sign": -1},
    {"class": "state", "type": "value", "name": "Va", "function_class": "exponential", "scaling": 9000.0,
     "shaping": False, "sign": -1},
    {"class": "success", "name": "success_time", "value": "timesteps", "function_class": "linear", "scaling": 2000,
     "shaping": False, "sign": 1},
    {"class": "success", "name": "success_bonus", "value": 5.0, "function_class": "quadratic", "scaling": 50,
     "shaping": False, "sign": 1},
    {"class": "step", "name": "step", "value": 1, "function_class": "linear", "scaling": 100, "shaping": False,
     "sign": -1},
    {"class": "goal", "type": "per_state", "name": "goal_state", "value": 0.3, "function_class": "linear",
     "scaling": 1, "shaping": False, "sign": 1},
    {"class": "goal", "type": "all", "name": "goal_all", "value": 1.0, "function_class": "linear", "scaling": 2,
     "shaping": True, "sign": 1},
    {"class": "state", "type": "int_error", "name": "roll", "function_class": "linear", "scaling": 300.0, "max": 0.2,
     "shaping": False, "sign": -1},
]
ZOO_TARGET_KEYS = {"slope_low": 1.0, "slope_high": 5.0, "amplitude_low": 2.0, "amplitude_high": 8.0,
                   "period_low": 40, "period_high": 90}


def write_zoo():
    with open(os.path.join(PARAMS, "fixed_wing_config.json")) as f:
        cfg = json.load(f)
    cfg["integration_window"] = 10
    cfg["reward"]["factors"] = ZOO_FACTORS
    cfg["reward"]["terms"] = [{"function_class": "linear", "weight": 1}, {"function_class": "exponential", "weight": 0.5},
                              {"function_class": "quadratic", "weight": 0.25}]
    for st in cfg["target"]["states"]:
        st.update(ZOO_TARGET_KEYS)
    with open(os.path.join(PARAMS, "fixed_wing_config_zoo.json"), "w") as f:
        json.dump(cfg, f, sort_keys=True, separators=(",", ":"))
        f.write("\n")


def main():
    for out, src in SOURCES.items():
        with open(os.path.join(REF, src)) as f:
            cfg = json.load(f)
        cfg.pop("render", None)   # visualisation is out of scope
        with open(os.path.join(PARAMS, out), "w") as f:
            json.dump(cfg, f, sort_keys=True, separators=(",", ":"))
            f.write("\n")
    write_randomised()
    write_zoo()
    ex = os.path.join(REF, "gym_fixed_wing", "examples")
    scen = np.load(os.path.join(ex, "test_sets", "test_set_wind_none_step20-20-3.npy"), allow_pickle=True)
    skeys = sorted(scen[0]["state"].keys())
    tkeys = sorted(scen[0]["target"].keys())
    np.savez_compressed(os.path.join(GOLDEN, "test_set_wind_none.npz"),
                        state_keys=np.array(skeys), target_keys=np.array(tkeys),
                        state=np.array([[s["state"][k] for k in skeys] for s in scen]),
                        target=np.array([[s["target"][k] for k in tkeys] for s in scen]))
    res = np.load(os.path.join(ex, "evaluations", "eval_res_PID_none.npy"), allow_pickle=True).item()
    lens = np.array([len(r) for r in res["rewards"]])
    flat = np.concatenate([np.asarray(r, dtype=np.float64) for r in res["rewards"]])
    np.savez_compressed(os.path.join(GOLDEN, "eval_res_PID_none_rewards.npz"), lengths=lens, rewards=flat)
    print("wrote", sorted(SOURCES), "and golden test set / PID reward trace")


if __name__ == "__main__":
    main()
