"""CPU ORACLE (test infrastructure): restatement of the env half of the hot path, FixedWingAircraft.

/root/reference does not travel to the GPU box, so the GPU parity tests and bench.py's cpu_baseline need an oracle
that does.  This file restates `gym_fixed_wing/fixed_wing.py` (FixedWingAircraft: __init__ :14-212,
set_curriculum_level :224-285, reset :287-336, step :338-437, linear_action_scaling :439-459, sample_target
:461-521, get_reward :674-774, get_observation :776-846, _get_error/_get_angle_dist :890-914, _get_goal_status
:916-931, _get_next_target :933-991) over the restated PyFly (oracle/pyfly_restated.py).  Quirks are kept on purpose
(SURVEY App. A): value-target sign for wrapped states, observation computed before the history is rebuilt at reset,
goal_achieved latch never cleared, float32 accumulation of the action-delta observation, noise "var" used as a std.

It is PINNED against the reference's own code: tests/test_oracle_cpu.py (test_restated_env_matches_reference_file) runs the unmodified reference file
(oracle/reference_env.py) and this restatement side by side on identical Philox streams and requires bit-identical
observations / rewards / dones (in this container, where /root/reference exists), and against the committed fixtures
under tests/golden/ (generated from the reference file by oracle/make_golden.py) everywhere.

Out of scope here as in the product: sampler hook, simulator-parameter randomisation, attitude_angular targets,
render/save_history, metrics.
"""
import copy
import json

import numpy as np

from .pyfly_restated import PyFly

F32MAX = np.finfo(np.float32).max
TWO_PI = 2 * np.pi


def _override(node, kws):
    for key, val in kws.items():
        if isinstance(val, dict) or isinstance(node[key], list):
            _override(node[key], val)
        else:
            node[key] = val


def _limit(var, which):
    v = getattr(var, "value_" + which)
    if v is None:
        v = getattr(var, "constraint_" + which)
    if v is None:
        v = F32MAX if which == "max" else -F32MAX
    return v


class RestatedEnv:
    def __init__(self, config_path, sim_config_path=None, sim_parameter_path=None, config_kw=None, sim_config_kw=None):
        with open(config_path) as f:
            self.cfg = json.load(f)
        if config_kw is not None:
            _override(self.cfg, copy.deepcopy(config_kw))
        skw = dict(sim_config_kw or {})
        skw["actuation"] = {"inputs": [a["name"] for a in self.cfg["action"]["states"]]}
        skw["turbulence_sim_length"] = self.cfg["steps_max"]
        kw = {"config_kw": skw}
        if sim_config_path is not None:
            kw["config_path"] = sim_config_path
        if sim_parameter_path is not None:
            kw["parameter_path"] = sim_parameter_path
        self.simulator = PyFly(**kw)
        sim = self.simulator
        self.steps_max = self.cfg["steps_max"]
        self.integration_window = self.cfg.get("integration_window", 0)
        self.history = None
        self.steps_count = None
        self.steps_for_target = None
        self.goal_achieved = False
        self.np_random = np.random.RandomState()
        ocfg = self.cfg["observation"]
        self.obs_norm = ocfg.get("normalize", False)
        for var in ocfg["states"]:
            hi, lo = var.get("high"), var.get("low")
            rad = var.get("convert_to_radians", False)
            hi = _limit(sim.state[var["name"]], "max") if hi is None else (np.radians(hi) if rad else hi)
            lo = _limit(sim.state[var["name"]], "min") if lo is None else (np.radians(lo) if rad else lo)
            bounded = hi != F32MAX and lo != -F32MAX
            if self.obs_norm:
                if var.get("mean") is None:
                    var["mean"] = hi - lo if bounded else 0
                if var.get("var") is None:
                    var["var"] = (hi - lo) / 16 if bounded else 1
        acfg = self.cfg["action"]
        self.act_names = [a["name"] for a in acfg["states"]]
        self.to_low = np.array([_limit(sim.state[n], "min") for n in self.act_names])
        self.to_high = np.array([_limit(sim.state[n], "max") for n in self.act_names])
        self.scale_actions = acfg.get("scale_space", False)
        if acfg.get("bounds_multiplier") is not None:
            self.bounds_max = np.full(3, acfg.get("scale_high", 1)) * acfg["bounds_multiplier"]
            self.bounds_min = np.full(3, acfg.get("scale_low", -1)) * acfg["bounds_multiplier"]
        self.goal_enabled = self.cfg["target"]["success_streak_req"] > 0
        self.target = None
        self.tprops = None
        self.tprops_init = None
        self.prev_shaping = {}
        self.rew_factors_init = copy.deepcopy(self.cfg["reward"]["factors"])   # fixed_wing.py:198 (randomize_scaling)
        self.set_curriculum_level(1)

    # ----------------------------------------------------------------------------------------------- curriculum
    def set_curriculum_level(self, level):
        assert 0 <= level <= 1
        for entry in self.cfg["simulator"].get("states", []):
            entry = dict(entry)
            name = entry.pop("name")
            rad = entry.pop("convert_to_radians", False)
            for prop, val in entry.items():
                if val is not None:
                    if "constraint" not in prop and ("min" in prop or "max" in prop):
                        mid = (entry[prop[:-3] + "max"] + entry[prop[:-3] + "min"]) / 2
                        val = mid - level * (mid - val)
                    if rad:
                        val = np.radians(val)
                setattr(self.simulator.state[name], prop, val)
        init = {"states": {}}
        for key, val in self.cfg["target"].items():
            if key != "states":
                init[key] = val[round(len(val) * level)] if isinstance(val, list) else val
                continue
            for st in val:
                props = {}
                for k, v in st.items():
                    if k == "name":
                        continue
                    if k not in ("bound", "class") and v is not None and not isinstance(v, bool):
                        mid = (st["high"] + v) / 2 if k == "low" else ((v + st["low"]) / 2 if k == "high" else 0)
                        v = mid - level * (mid - v)
                    props[k] = v
                init["states"][st["name"]] = props
        self.tprops_init = init

    # ----------------------------------------------------------------------------------------------- primitives
    def scale(self, a, backward=False):
        lo, hi = self.cfg["action"].get("scale_low"), self.cfg["action"].get("scale_high")
        if backward:
            return np.array(hi - lo) * (a - self.to_low) / (self.to_high - self.to_low) + lo
        return np.array(self.to_high - self.to_low) * (a - lo) / (hi - lo) + self.to_low

    def error(self, name):
        var = self.simulator.state[name]
        if getattr(var, "wrap", False):
            return (var.value - self.target[name] + np.pi) % TWO_PI - np.pi
        return self.target[name] - var.value

    def goal_status(self):
        res = {}
        for name, props in self.tprops.items():
            if props.get("bound") is not None:
                res[name] = np.abs(self.error(name)) <= props["bound"]
        res["all"] = all(res.values())
        return res

    def sample_target(self):
        self.steps_for_target = 0
        self.target, self.tprops = {}, {}
        for name, props in self.tprops_init["states"].items():
            out = {"class": props.get("class", "constant")}
            rad = props.get("convert_to_radians", False)
            lo, hi, delta = props["low"], props["high"], props.get("delta")
            if rad:
                lo, hi = np.radians(lo), np.radians(hi)
                delta = np.radians(delta) if delta is not None else None
            if delta is not None:
                v = self.simulator.state[name].value
                lo = max(lo, v - delta)
                hi = max(min(hi, v + delta), lo)
            first = self.np_random.uniform(lo, hi)
            if out["class"] == "linear":
                out["slope"] = self.np_random.uniform(props["slope_low"], props["slope_high"])
                if self.np_random.uniform() < 0.5:
                    out["slope"] *= -1
                if rad:
                    out["slope"] = np.radians(out["slope"])
            elif out["class"] == "sinusoidal":
                out["amplitude"] = self.np_random.uniform(props["amplitude_low"], props["amplitude_high"])
                if rad:
                    out["amplitude"] = np.radians(out["amplitude"])
                out["period"] = self.np_random.uniform(props.get("period_low", 250), props.get("period_high", 500))
                out["phase"] = self.np_random.uniform(0, TWO_PI) / (TWO_PI / out["period"])
                out["bias"] = first - out["amplitude"] * np.sin(TWO_PI / out["period"] * (self.steps_count + out["phase"]))
            if props.get("bound") is not None:
                out["bound"] = np.radians(props["bound"]) if rad else props["bound"]
            self.target[name] = first
            self.tprops[name] = out

    def next_targets(self):
        res = {}
        dt = self.simulator.dt
        for name, props in self.tprops.items():
            cls, cur = props.get("class", "constant"), self.target[name]
            if cls == "compensate":
                assert name == "Va"
                pcls = self.tprops["pitch"]["class"]
                ptar = self.tprops["pitch"]["bias"] if pcls == "sinusoidal" else self.target["pitch"]
                if ptar <= np.radians(-2.5):
                    end = 28.434 - 40.0841 * ptar
                    slope = 7 * max(0, 1 if cur < end * 0.95 else 1 - cur / (end * 1.5)) if cur <= end else 0
                    val = cur + (slope * (-self.target["pitch"]) - 0.25) * dt
                elif ptar >= np.radians(5):
                    end = 26.27 - 41.2529 * ptar
                    if cur > end:
                        val = cur + (end - cur) * 1 / 150 if self.steps_for_target < 750 else end
                    else:
                        val = cur
                else:
                    val = cur
            elif cls == "linear":
                val = cur + props["slope"] * dt
            elif cls == "sinusoidal":
                val = props["amplitude"] * np.sin(TWO_PI / props["period"] * (self.steps_count + props["phase"])) + props["bias"]
            else:
                val = cur
            if getattr(self.simulator.state[name], "wrap", False) and np.abs(val) > np.pi:
                val = np.sign(val) * (np.abs(val) % np.pi - np.pi)
            res[name] = val
        return res

    # ---------------------------------------------------------------------------------------------- observation
    def observation(self):
        ocfg = self.cfg["observation"]
        length, step, noise = ocfg["length"], ocfg.get("step", 1), ocfg.get("noise")
        W, sim, H = self.integration_window, self.simulator, self.history
        rows = []
        for i in range(1, (length + (1 if step == 1 else 0)) * step, step):
            jitter = None
            if i > self.steps_count:
                i = self.steps_count + 1
                if length > 1:
                    jitter = self.np_random.uniform(-1, 1) * sim.dt
            row = []
            for var in ocfg["states"]:
                name, kind = var["name"], var["type"]
                if kind == "state":
                    val = sim.state[name].history[-i]
                elif kind == "target":
                    how = var["value"]
                    if how == "relative":
                        val = self.error(name) if i == 1 else H["error"][name][-i]
                    elif how == "absolute":
                        val = self.target[name] if i == 1 else H["target"][name][-i]
                    elif how == "integrator":
                        if H is None:
                            val = self.error(name) * W
                        else:
                            val = np.sum(H["error"][name][-W - i:-i])
                            if self.steps_count - i < W:
                                val += (W - (self.steps_count - i)) * H["error"][name][0]
                    else:
                        raise ValueError(how)
                elif kind == "action":
                    k = self.act_names.index(name)
                    if self.steps_count - i < 0:
                        val = sim.state[name].value
                        if self.scale_actions:
                            probe = np.zeros(len([v for v in ocfg["states"] if v["type"] == "action"]))
                            probe[k] = val
                            val = self.scale(probe, backward=True)[k]
                    else:
                        w = var.get("window_size", 1)
                        a, b = -w - i + 1, (None if i == 1 else -(i - 1))
                        if self.scale_actions:
                            seq = [act[k] for act in H["action"][a:b]]
                        else:
                            seq = sim.state[name].history["command"][a:b]
                        val = np.sum(np.abs(np.diff(seq)), dtype=np.float32)
                else:
                    raise ValueError(kind)
                if jitter is not None:
                    val += jitter
                if self.obs_norm and var.get("norm", True):
                    val -= var["mean"]
                    val /= var["var"]
                if noise is not None:
                    val += self.np_random.normal(loc=noise["mean"], scale=noise["var"])
                row.append(val)
            rows.append(row)
        if ocfg["shape"] == "vector":
            return np.array([v for r in rows for v in r])
        return np.array(rows)

    # --------------------------------------------------------------------------------------------------- reward
    def reward(self, action, success):
        rcfg = self.cfg["reward"]
        potential = rcfg.get("form", "absolute") == "potential"
        acc = {t["function_class"]: [0, 0, t["weight"]] for t in rcfg["terms"]}   # plain, shaping, weight
        for comp in rcfg["factors"]:
            cls = comp["class"]
            if cls == "action":
                if comp["type"] == "value":
                    val = np.sum(np.abs(self.history["action"][-1]))
                elif comp["type"] == "delta":
                    if self.steps_count > 1:
                        val = np.sum(np.abs(np.diff(self.history[comp["name"]][-comp["window_size"]:], axis=0)))
                    else:
                        val = 0
                else:
                    over = np.where(action > self.bounds_max, action - self.bounds_max, 0)
                    under = np.where(action < self.bounds_min, action - self.bounds_min, 0)
                    val = np.sum(np.abs(over)) + np.sum(np.abs(under))
            elif cls == "state":
                if comp["type"] == "value":
                    val = self.simulator.state[comp["name"]].value
                elif comp["type"] == "error":
                    val = self.error(comp["name"])
                else:
                    errs = self.history["error"][comp["name"]]
                    val = np.sum(errs[-self.integration_window:])
                    if self.steps_count < self.integration_window:
                        val += (self.integration_window - self.steps_count) * errs[0]
            elif cls == "success":
                val = 0
                if success:
                    val = (self.steps_max - self.steps_count) if comp["value"] == "timesteps" else comp["value"]
            elif cls == "step":
                val = comp["value"]
            elif cls == "goal":
                status = self.goal_status()
                if comp["type"] == "per_state":
                    val = sum(comp["value"] / len(self.target) for k, ok in status.items() if k != "all" and ok)
                else:
                    val = comp["value"] if status["all"] else 0
            else:
                raise ValueError(cls)
            if comp["function_class"] == "linear":
                val = np.clip(np.abs(val) / comp["scaling"], 0, comp.get("max", None))
            else:
                val = val ** 2 / comp["scaling"]
            acc[comp["function_class"]][1 if comp.get("shaping", False) else 0] += val * np.sign(comp.get("sign", -1))
        total = 0
        for fclass, (plain, shaping, weight) in acc.items():
            prev = self.prev_shaping[fclass]
            if fclass == "exponential":
                if potential:
                    val = -1 + np.exp(plain + (shaping - prev)) if prev is not None else -1 + np.exp(plain)
                else:
                    val = -1 + np.exp(plain + shaping)
            else:
                val = plain
                if potential:
                    if prev is not None:
                        val += shaping - prev
                else:
                    val += shaping
            self.prev_shaping[fclass] = shaping
            total += weight * val
        return total

    # ------------------------------------------------------------------------------------------- reset and step
    def sample_simulator_parameters(self):
        """fixed_wing.py:523-570: per-episode randomisation of PyFly's model parameters and attributes."""
        for key, value in self.cfg["simulator"].items():
            if key == "states":
                continue
            elif key == "model":
                dist_type = value.get("distribution", "gaussian")
                for pa in value["parameters"]:
                    orig = pa.get("original", None)
                    if orig is None:
                        orig = self.simulator.params[pa["name"]]
                        pa["original"] = orig
                    if orig == 0:
                        continue
                    var = pa.get("var", value["var"])
                    if value["var_type"] == "relative":
                        var *= np.abs(orig)
                    if dist_type == "gaussian":
                        val = self.np_random.normal(loc=orig, scale=var)
                        clip = pa.get("clip", value.get("clip", None))
                        if clip is not None:
                            if value["var_type"] == "relative":
                                clip *= orig
                            val = np.clip(val, orig - clip, orig + clip)
                    elif dist_type == "uniform":
                        val = self.np_random.uniform(low=orig - var, high=orig + var)
                    else:
                        raise ValueError("Unexpected distribution type {}".format(dist_type))
                    self.simulator.params[pa["name"]] = val
            else:
                if "values" in value:
                    probs = value.get("probabilities", None)
                    val = self.np_random.choice(value["values"], p=None if probs is None else np.array(probs))
                else:
                    val = self.np_random.uniform(value["low"], value["high"])
                    if isinstance(value["low"], bool):
                        val = bool(val)
                setattr(self.simulator, key, val)

    def reset(self, state=None, target=None, **sim_reset_kw):
        self.steps_count = 0
        self.simulator.reset(state, **sim_reset_kw)
        self.sample_simulator_parameters()
        self.sample_target()
        if target is not None:
            for k, v in target.items():
                if self.tprops[k]["class"] not in ("constant", "compensate"):
                    self.tprops[k]["class"] = "constant"
                self.target[k] = v
        obs = self.observation()   # before the history is rebuilt, as in the reference
        self.history = {"action": [], "target": {k: [v] for k, v in self.target.items()},
                        "error": {k: [self.error(k)] for k in self.target}}
        if self.goal_enabled:
            self.history["goal"] = {k: [v] for k, v in self.goal_status().items()}
        for term in self.cfg["reward"]["terms"]:
            self.prev_shaping[term["function_class"]] = None
        if self.cfg["reward"].get("randomize_scaling", False):   # fixed_wing.py:330-334
            for i, fac in enumerate(self.rew_factors_init):
                if isinstance(fac["scaling"], list):
                    self.cfg["reward"]["factors"][i]["scaling"] = self.np_random.uniform(fac["scaling"][0], fac["scaling"][1])
        return obs

    def step(self, action):
        self.history["action"].append(action)
        assert not np.any(np.isnan(action))
        if self.scale_actions:
            action = self.scale(np.clip(action, self.cfg["action"].get("scale_low"), self.cfg["action"].get("scale_high")))
        ok, sim_info = self.simulator.step(list(action))
        self.steps_count += 1
        self.steps_for_target += 1
        info, done = {}, False
        if self.steps_count >= self.steps_max > 0:
            done, info["termination"] = True, "steps"
        if ok:
            resample = first_success = False
            tcfg = self.cfg["target"]
            if self.goal_enabled:
                for k, v in self.goal_status().items():
                    self.history["goal"][k].append(v)
                req = tcfg["success_streak_req"]
                if self.steps_for_target >= req and np.mean(self.history["goal"]["all"][-req:]) >= tcfg["success_streak_fraction"]:
                    first_success = not self.goal_achieved
                    self.goal_achieved = True
                    if tcfg["on_success"] == "done":
                        done, info["termination"] = True, "success"
                    elif tcfg["on_success"] == "new":
                        resample = True
            reward = self.reward(self.history["action"][-1], first_success)
            every = tcfg.get("resample_every", 0)
            if resample or (every and self.steps_for_target >= every):
                self.sample_target()
            for k, v in self.next_targets().items():
                self.target[k] = v
                self.history["target"][k].append(v)
                self.history["error"][k].append(self.error(k))
            obs = self.observation()
        else:
            done = True
            fail = self.cfg["reward"].get("step_fail", 0)
            reward = self.steps_count - self.steps_max if fail == "timesteps" else fail
            info["termination"] = sim_info["termination"]
            obs = self.observation()
        info["target"] = self.target
        return obs, reward, done, info
