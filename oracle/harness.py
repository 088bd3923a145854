"""CPU ORACLE (test infrastructure): drive an oracle env (restated, or the unmodified reference file over shims) on
the same Philox random streams the CUDA kernels consume, with SubprocVecEnv-style auto-reset.

Tick convention (csrc/philox.cuh): every reset()/step() call of an env uses its own tick, counted from seed().
"""
import os

import numpy as np

from . import philox
from .env_restated import RestatedEnv

# PyFly variable order == enum fw_sv in include/fwgym.h (checked by tests/test_oracle_cpu.py::test_capi_exports_every_declared_symbol)
SV_ORDER = ["roll", "pitch", "yaw", "omega_p", "omega_q", "omega_r", "position_n", "position_e", "position_d",
            "velocity_u", "velocity_v", "velocity_w", "Va", "alpha", "beta", "elevator", "aileron", "rudder",
            "throttle", "elevon_left", "elevon_right"]
PARAMS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fixed-wing-gym_b200", "params")


def config_path(name="fixed_wing_config.json"):
    return os.path.join(PARAMS_DIR, name)


def make_env(kind="restated", config=None, config_kw=None, sim_config_kw=None):
    config = config or config_path()
    if kind == "reference":
        from . import reference_env
        return reference_env.make_reference_env(config, config_kw=config_kw, sim_config_kw=sim_config_kw)
    return RestatedEnv(config, config_kw=config_kw, sim_config_kw=sim_config_kw)


class OracleRunner:
    """One oracle env wired to the Philox streams of global env id `env_id` under `seed`."""

    def __init__(self, env, seed=0, env_id=0, auto_reset=True):
        self.env = env
        self.key = philox.PhiloxKey(seed, env_id)
        self.tick = 0
        self.auto_reset = auto_reset
        env.np_random = philox.EnvRandom(self.key)
        self._var_rngs = []
        for name, var in env.simulator.state.items():
            if name == "attitude":
                continue
            var.np_random = philox.VarRandom(self.key, SV_ORDER.index(name))
            self._var_rngs.append(var.np_random)
        self._wind_rng = philox.WindRandom(self.key)
        env.simulator.wind.np_random = self._wind_rng
        self.turbulence = bool(env.simulator.wind.turbulence)
        self.nfev = []          # RHS evaluations per step (attempts = (nfev - 2) / 6)
        self.on_done = None     # optional callback(env, info) at episode end, before the auto-reset
        self.ep_return = 0.0

    def _begin(self):
        t = self.tick
        self.tick += 1
        self.env.np_random.begin(t)
        for r in self._var_rngs:
            r.tick = t
        self._wind_rng.tick, self._wind_rng.n = t, 0
        return t

    def reset(self, state=None, target=None, turbulence_noise=None):
        """turbulence_noise: an explicit [4, T] standard-normal array for this episode (the reference's
        reset(**sim_reset_kw) pass-through, fixed_wing.py:287,308) instead of the env's Philox stream."""
        self.ep_return = 0.0
        t = self._begin()
        kw = {}
        if self.turbulence:
            kw["turbulence_noise"] = (self.key.turbulence_noise(t, self.env.cfg["steps_max"])
                                      if turbulence_noise is None else np.asarray(turbulence_noise, dtype=np.float64))
        return self.env.reset(state=state, target=target, **kw)

    def step(self, action):
        """-> (obs, reward, done, info); on done with auto_reset the returned obs is the post-reset one and
        info["terminal_observation"] the terminal one (SubprocVecEnv semantics, train_rl_controller.py:223)."""
        self._begin()
        evals0 = self.env.simulator.n_rhs_evals
        obs, rew, done, info = self.env.step(np.asarray(action, dtype=np.float64))
        # right-hand-side evaluations of THIS step, counted at the call (sol.nfev is lost when a ConstraintException
        # leaves solve_ivp)
        self.nfev.append(self.env.simulator.n_rhs_evals - evals0)
        self.ep_return += float(rew)         # what a Monitor wrapper would report as episode "r"
        if done and self.on_done is not None:
            self.on_done(self.env, info)     # before the auto-reset wipes the episode's histories
        if done and self.auto_reset:
            info = dict(info)
            info["terminal_observation"] = obs
            obs = self.reset()
        return obs, rew, done, info

    def attempts_last(self):
        """dopri5 step attempts of the last env step.  RK45.__init__ takes 2 evaluations and every attempt 6; an
        attempt cut short by a ConstraintException (raised inside the RHS, counted by scipy before the call) still
        counts, so a failing step reports the attempt it failed in."""
        return -((2 - self.nfev[-1]) // 6)

    def ode_state(self):
        """The 19-vector PyFly would start the next step from + derived values, for state parity checks."""
        sim = self.env.simulator
        y = list(sim.state["attitude"].value)
        y += [sim.state[n].value for n in SV_ORDER[3:12]]
        y += sim.actuation.get_values()
        d = [sim.state[n].value for n in ("roll", "pitch", "yaw", "Va", "alpha", "beta", "elevator", "aileron")]
        return np.array(y + d, dtype=np.float64)
