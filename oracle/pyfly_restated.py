"""CPU ORACLE (test infrastructure, NOT product code): restatement of the PyFly 0.1.2 simulator.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.

What it restates
----------------
The reference (`/root/reference/gym_fixed_wing/fixed_wing.py`) delegates all physics to the third-party
package `pyfly-fixed-wing==0.1.2` (`/root/reference/setup.py:27`; evaluated at commit 21f5b5c8…,
`gym_fixed_wing/examples/README.md:30`).  That package is ABSENT from /root/reference and from this image, so this
file restates its published algorithm from the public repo `eivindeb/pyfly` (pyfly/pyfly.py, pyfly/dryden.py,
pyfly/pid_controller.py).  Structure: SURVEY.md App. B/D.  The reference's call sites that define the surface are
`fixed_wing.py:46` (constructor), `:221` (seed), `:308` (reset), `:358` (step), `:69-87,143-166` (Variable limits),
`:798,828` (histories), `:897,988` (wrap), `:795,959` (dt).

PARITY STATUS: "parity unpinned" at the PyFly boundary — there is no PyFly source, test or fixture in the reference
that pins intermediate simulator states.  All aircraft/actuator constants come from DATA files
(`fixed-wing-gym_b200/params/x8_param.json`, `pyfly_config.json`), never from literals here.  The only golden data
that constrains it is the end-to-end PID reward trace (`examples/evaluations/eval_res_PID_none.npy`); the gap to that
trace is *reported* by tests/test_oracle_cpu.py::test_golden_pid_trace_gap_is_reported, see DESIGN.md.

The integrator is `scipy.integrate.solve_ivp` called literally with defaults (RK45, rtol 1e-3, atol 1e-6), exactly as
PyFly does, so integrator parity is by construction (scipy/integrate/_ivp/rk.py, common.py in this image: 1.18.1).
"""
import copy
import json
import math
import os.path as osp

import numpy as np
import scipy.integrate
import scipy.signal

PARAMS_DIR = osp.join(osp.dirname(osp.abspath(__file__)), "..", "fixed-wing-gym_b200", "params")


class ConstraintException(Exception):
    def __init__(self, variable, value, limit):
        self.message = "Constraint on {} violated ({}/{})".format(variable, value, limit)
        self.variable = variable


class Variable:
    """A scalar simulator state with init range, value clip, hard constraint and optional +-pi wrap."""

    def __init__(self, name, value_min=None, value_max=None, init_min=None, init_max=None, constraint_min=None,
                 constraint_max=None, convert_to_radians=False, unit=None, label=None, wrap=False, **_ignored):
        self.value_min = value_min
        self.value_max = value_max
        self.init_min = init_min if init_min is not None else value_min
        self.init_max = init_max if init_max is not None else value_max
        self.constraint_min = constraint_min
        self.constraint_max = constraint_max
        if convert_to_radians:
            for attr_name, val in list(self.__dict__.items()):
                if val is None:
                    continue
                if attr_name.endswith("min") or attr_name.endswith("max"):
                    setattr(self, attr_name, val * (np.pi / 180))
        self.name = name
        self.value = None
        self.wrap = wrap
        self.unit = unit
        self.label = label if label is not None else name
        self.np_random = None
        self.seed()
        self.history = None

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)

    def reset(self, value=None):
        self.history = []
        if value is None:
            value = self.np_random.uniform(self.init_min, self.init_max)
        else:
            value = self.apply_conditions(value)
        self.value = value
        self.history.append(value)

    def apply_conditions(self, value):
        if self.constraint_min is not None and value < self.constraint_min:
            raise ConstraintException(self.name, value, self.constraint_min)
        if self.constraint_max is not None and value > self.constraint_max:
            raise ConstraintException(self.name, value, self.constraint_max)
        if self.value_min is not None or self.value_max is not None:
            value = np.clip(value, self.value_min, self.value_max)
        if self.wrap and np.abs(value) > np.pi:
            value = np.sign(value) * (np.abs(value) % np.pi - np.pi)
        return value

    def set_value(self, value, save=True):
        value = self.apply_conditions(value)
        if save:
            self.history.append(value)
        self.value = value


class ControlVariable(Variable):
    """Actuator state (value, dot) with first/second order command-following dynamics."""

    def __init__(self, order=None, tau=None, omega_0=None, zeta=None, dot_max=None, disabled=False, **kwargs):
        self.dot_max = dot_max  # before super().__init__ so convert_to_radians sees it
        super().__init__(**kwargs)
        self.order = order
        self.tau = tau
        self.omega_0 = omega_0
        self.zeta = zeta
        if order == 1:
            self.coefs = [[-1 / self.tau, 0, 1 / self.tau], [0, 0, 0]]
        elif order == 2:
            self.coefs = [[0, 1, 0], [-self.omega_0 ** 2, -2 * self.zeta * self.omega_0, self.omega_0 ** 2]]
        self.dot = None
        self.command = None
        self.disabled = disabled
        if self.disabled:
            self.value = 0

    def apply_conditions(self, values):
        try:
            value, dot = values
        except TypeError:
            value, dot = values, 0
        value = super().apply_conditions(value)
        if self.dot_max is not None:
            dot = np.clip(dot, -self.dot_max, self.dot_max)
        return [value, dot]

    def set_command(self, command):
        command = super().apply_conditions(command)
        self.command = command
        self.history["command"].append(command)

    def reset(self, value=None):
        self.history = {"value": [], "dot": [], "command": []}
        if not self.disabled:
            if value is None:
                value = self.np_random.uniform(self.init_min, self.init_max), 0
            else:
                value = self.apply_conditions(value)
            self.value = value[0]
            self.dot = value[1]
            self.command = None
        else:
            self.value, self.dot, self.command = 0, 0, None
        self.history["value"].append(self.value)
        self.history["dot"].append(self.dot)

    def set_value(self, value, save=True):
        value, dot = self.apply_conditions(value)
        self.value = value
        self.dot = dot
        if save:
            self.history["value"].append(value)
            self.history["dot"].append(dot)


class Actuation:
    """Maps model inputs (elevator/aileron/throttle) onto physical actuator dynamics (elevons/throttle)."""

    def __init__(self, model_inputs, actuator_inputs, dynamics):
        self.states = {}
        self.coefficients = [[np.array([]) for _ in range(3)] for __ in range(2)]
        self.elevon_dynamics = False
        self.dynamics = dynamics
        self.inputs = actuator_inputs
        self.model_inputs = model_inputs
        self.input_indices = {s: i for i, s in enumerate(actuator_inputs)}
        self.dynamics_indices = {s: i for i, s in enumerate(dynamics)}

    def add_state(self, state):
        self.states[state.name] = state
        if state.name in self.dynamics:
            for i in range(2):
                for j in range(3):
                    self.coefficients[i][j] = np.append(self.coefficients[i][j], state.coefs[i][j])

    def finalize(self):
        if "elevon_left" in self.dynamics or "elevon_right" in self.dynamics:
            assert "elevon_left" in self.dynamics and "elevon_right" in self.dynamics
            assert not ("aileron" in self.dynamics or "elevator" in self.dynamics)
            self.elevon_dynamics = True
        # coefficient rows were appended in config-variable order; re-order to `dynamics` order
        order = [n for n in self.states if n in self.dynamics]
        perm = [order.index(n) for n in self.dynamics]
        for i in range(2):
            for j in range(3):
                self.coefficients[i][j] = self.coefficients[i][j][perm]

    def set_states(self, values, save=True):
        n = len(self.dynamics)
        for i, state in enumerate(self.dynamics):
            self.states[state].set_value((values[i], values[n + i]), save=save)
        if self.elevon_dynamics:
            elevator, aileron = self._map_elevon_to_elevail(er=self.states["elevon_right"].value,
                                                            el=self.states["elevon_left"].value)
            self.states["aileron"].set_value((aileron, 0), save=save)
            self.states["elevator"].set_value((elevator, 0), save=save)

    def get_values(self):
        return [self.states[s].value for s in self.dynamics] + [self.states[s].dot for s in self.dynamics]

    def rhs(self, setpoints=None):
        if setpoints is None:
            setpoints = [self.states[s].command for s in self.dynamics]
        states = [self.states[s].value for s in self.dynamics]
        dots = [self.states[s].dot for s in self.dynamics]
        c = self.coefficients
        dot = np.multiply(states, c[0][0]) + np.multiply(setpoints, c[0][2]) + np.multiply(dots, c[0][1])
        ddot = np.multiply(states, c[1][0]) + np.multiply(setpoints, c[1][2]) + np.multiply(dots, c[1][1])
        return np.concatenate((dot, ddot))

    def set_and_constrain_commands(self, commands):
        dynamics_commands = {}
        if self.elevon_dynamics and "elevator" in self.inputs and "aileron" in self.inputs:
            elev_c, ail_c = commands[self.input_indices["elevator"]], commands[self.input_indices["aileron"]]
            er_c, el_c = self._map_elevail_to_elevon(elev=elev_c, ail=ail_c)
            dynamics_commands = {"elevon_right": er_c, "elevon_left": el_c}
        for state in self.dynamics:
            if state in self.input_indices:
                state_command = commands[self.input_indices[state]]
            else:
                state_command = dynamics_commands[state]
            self.states[state].set_command(state_command)
            dynamics_commands[state] = self.states[state].command
        if self.elevon_dynamics:
            elev_c, ail_c = self._map_elevon_to_elevail(er=dynamics_commands["elevon_right"],
                                                        el=dynamics_commands["elevon_left"])
            self.states["elevator"].set_command(elev_c)
            self.states["aileron"].set_command(ail_c)
        for state, i in self.input_indices.items():
            commands[i] = self.states[state].command
        return commands

    def reset(self, state_init=None):
        for state in self.dynamics:
            init = None
            if state_init is not None and state in state_init:
                init = state_init[state]
            self.states[state].reset(value=init)
        if self.elevon_dynamics:
            elev, ail = self._map_elevon_to_elevail(er=self.states["elevon_right"].value,
                                                    el=self.states["elevon_left"].value)
            self.states["elevator"].reset(value=elev)
            self.states["aileron"].reset(value=ail)

    @staticmethod
    def _map_elevail_to_elevon(elev, ail):
        er = -1 * ail + elev
        el = ail + elev
        return er, el

    @staticmethod
    def _map_elevon_to_elevail(er, el):
        ail = (-er + el) / 2
        elev = (er + el) / 2
        return elev, ail


class AttitudeQuaternion:
    def __init__(self):
        self.quaternion = None
        self.history = None

    def seed(self, seed):
        return

    def reset(self, euler_init):
        self._from_euler_angles(euler_init)
        self.history = [self.quaternion]

    def as_euler_angle(self, angle="all", timestep=-1):
        e0, e1, e2, e3 = self.history[timestep]
        res = {}
        if angle == "roll" or angle == "all":
            res["roll"] = np.arctan2(2 * (e0 * e1 + e2 * e3), e0 ** 2 + e3 ** 2 - e1 ** 2 - e2 ** 2)
        if angle == "pitch" or angle == "all":
            res["pitch"] = np.arcsin(2 * (e0 * e2 - e1 * e3))
        if angle == "yaw" or angle == "all":
            res["yaw"] = np.arctan2(2 * (e0 * e3 + e1 * e2), e0 ** 2 + e1 ** 2 - e2 ** 2 - e3 ** 2)
        return res if angle == "all" else res[angle]

    @property
    def value(self):
        return self.quaternion

    def _from_euler_angles(self, euler):
        phi, theta, psi = euler
        cphi, sphi = np.cos(phi / 2), np.sin(phi / 2)
        cth, sth = np.cos(theta / 2), np.sin(theta / 2)
        cpsi, spsi = np.cos(psi / 2), np.sin(psi / 2)
        e0 = cpsi * cth * cphi + spsi * sth * sphi
        e1 = cpsi * cth * sphi - spsi * sth * cphi
        e2 = cpsi * sth * cphi + spsi * cth * sphi
        e3 = spsi * cth * cphi - cpsi * sth * sphi
        self.quaternion = (e0, e1, e2, e3)

    def set_value(self, quaternion, save=True):
        self.quaternion = quaternion
        if save:
            self.history.append(self.quaternion)


# ------------------------------------------------------------------------------------------------ Dryden turbulence
FT_PER_M = 3.28084
KNOT_FTPS = 1.6878098571  # W20 is given in knots


def dryden_filters(b, h=100.0, V_a=25.0, intensity=None):
    """Transfer-function (num, den) pairs of the six MIL-F-8785C low-altitude Dryden shaping filters as PyFly's
    DrydenGustModel builds them: feet inside, W20 = 15 / 30 / 45 knots, linear gusts back in m/s.

    The first restatement used `W_20 = 15 * FT_PER_M` and the ft/s filter outputs as m/s: gusts of 8-9 m/s rms
    ("moderate") instead of the specification's ~2 m/s.  Evidence against it, from the reference's own numbers: its PID
    controller, replayed on 25 scenarios under that "moderate" turbulence, succeeds once and is destroyed (body-rate
    constraint) 15 times, where the README reports 93 % success and a control variation of 0.70; with the units below
    it succeeds 25 / 25 with a control variation of 0.63 (the published sets add steady wind)."""
    h = h * FT_PER_M
    b = b * FT_PER_M
    V_a = V_a * FT_PER_M
    if intensity is None or intensity == "light":
        W_20 = 15 * KNOT_FTPS
    elif intensity == "moderate":
        W_20 = 30 * KNOT_FTPS
    elif intensity == "severe":
        W_20 = 45 * KNOT_FTPS
    else:
        raise Exception("Unsupported intensity type")
    L_u = h / (0.177 + 0.000823 * h) ** 1.2
    L_v = L_u
    L_w = h
    sigma_w = 0.1 * W_20
    sigma_u = sigma_w / (0.177 + 0.000823 * h) ** 0.4
    sigma_v = sigma_u
    K_u = sigma_u * math.sqrt((2 * L_u) / (math.pi * V_a))
    K_v = sigma_v * math.sqrt(L_v / (math.pi * V_a))
    K_w = sigma_w * math.sqrt(L_w / (math.pi * V_a))
    T_u = L_u / V_a
    T_v1 = math.sqrt(3.0) * L_v / V_a
    T_v2 = L_v / V_a
    T_w1 = math.sqrt(3.0) * L_w / V_a
    T_w2 = L_w / V_a
    K_p = sigma_w * math.sqrt(0.8 / V_a) * ((math.pi / (4 * b)) ** (1 / 6)) / (L_w ** (1 / 3))
    K_q = 1 / V_a
    K_r = K_q
    T_p = 4 * b / (math.pi * V_a)
    T_q = T_p
    T_r = 3 * b / (math.pi * V_a)
    m = 1.0 / FT_PER_M    # ft/s -> m/s for the three linear gust components
    return {
        "H_u": ([K_u * m], [T_u, 1]),
        "H_v": ([K_v * T_v1 * m, K_v * m], [T_v2 ** 2, 2 * T_v2, 1]),
        "H_w": ([K_w * T_w1 * m, K_w * m], [T_w2 ** 2, 2 * T_w2, 1]),
        "H_p": ([K_p], [T_p, 1]),
        "H_q": ([-K_w * K_q * T_w1, -K_w * K_q, 0],
                [T_q * T_w2 ** 2, T_w2 ** 2 + 2 * T_q * T_w2, T_q + 2 * T_w2, 1]),
        "H_r": ([K_v * K_r * T_v1, K_v * K_r, 0],
                [T_r * T_v2 ** 2, T_v2 ** 2 + 2 * T_r * T_v2, T_r + 2 * T_v2, 1]),
    }


# which of the four white-noise streams drives each filter (u,v,w | p,q,r)
DRYDEN_NOISE_STREAM = {"H_u": 0, "H_v": 1, "H_w": 2, "H_p": 3, "H_q": 1, "H_r": 2}


class _Filter:
    def __init__(self, num, den):
        self.filter = scipy.signal.lti(num, den)
        self.x = None

    def simulate(self, u, t):
        x0 = None if self.x is None else self.x[-1]
        _, y, self.x = scipy.signal.lsim(self.filter, U=u, T=t, X0=x0)
        if self.x.ndim == 1:
            self.x = self.x[:, None]
        return y

    def reset(self):
        self.x = None


class DrydenGustModel:
    def __init__(self, dt, b, h=100, V_a=25, intensity=None):
        self.filters = {k: _Filter(*v) for k, v in dryden_filters(b, h, V_a, intensity).items()}
        self.np_random = None
        self.seed()
        self.dt = dt
        self.sim_length = None
        self.noise = None
        self.vel_lin = None
        self.vel_ang = None

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)

    def _generate_noise(self, size):
        return np.sqrt(np.pi / self.dt) * self.np_random.standard_normal(size=(4, size))

    def reset(self, noise=None):
        self.vel_lin = None
        self.vel_ang = None
        self.sim_length = 0
        for f in self.filters.values():
            f.reset()
        if noise is not None:
            assert len(noise.shape) == 2 and noise.shape[0] == 4
            noise = noise * math.sqrt(math.pi / self.dt)
        self.noise = noise

    def simulate(self, length):
        t_span = [self.sim_length, self.sim_length + length]
        t = np.linspace(t_span[0] * self.dt, t_span[1] * self.dt, length)
        if self.noise is None:
            noise = self._generate_noise(t.shape[0])
        else:
            if self.noise.shape[-1] >= t_span[1]:
                noise = self.noise[:, t_span[0]:t_span[1]]
            else:
                idx = np.arange(t_span[0], t_span[1]) % self.noise.shape[-1]
                noise = self.noise[:, idx]
        f = self.filters
        vel_lin = np.array([f["H_u"].simulate(noise[0], t), f["H_v"].simulate(noise[1], t),
                            f["H_w"].simulate(noise[2], t)])
        vel_ang = np.array([f["H_p"].simulate(noise[3], t), f["H_q"].simulate(noise[1], t),
                            f["H_r"].simulate(noise[2], t)])
        if self.vel_lin is None:
            self.vel_lin, self.vel_ang = vel_lin, vel_ang
        else:
            self.vel_lin = np.concatenate((self.vel_lin, vel_lin), axis=1)
            self.vel_ang = np.concatenate((self.vel_ang, vel_ang), axis=1)
        self.sim_length += length


class Wind:
    def __init__(self, turbulence, mag_min=None, mag_max=None, b=None, turbulence_intensity=None, sim_length=250,
                 dt=None):
        self.turbulence = turbulence
        self.mag_min = mag_min
        self.mag_max = mag_max
        self.steady = None
        self.turbulence_sim_length = sim_length
        self.dryden = DrydenGustModel(dt, b, intensity=turbulence_intensity) if turbulence else None
        self.np_random = None
        self.seed()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        if self.turbulence:
            self.dryden.seed(seed)

    def reset(self, value=None, noise=None):
        if value is None:
            magnitude = self.np_random.uniform(self.mag_min, self.mag_max)
            w_n = self.np_random.uniform(-magnitude, magnitude)
            w_e_max = np.sqrt(magnitude ** 2 - w_n ** 2)
            w_e = self.np_random.uniform(-w_e_max, w_e_max)
            w_d = np.sqrt(magnitude ** 2 - w_n ** 2 - w_e ** 2)
            value = [w_n, w_e, w_d]
        if self.turbulence:
            self.dryden.reset(noise)
            self.dryden.simulate(self.turbulence_sim_length)
        self.steady = np.array(value, dtype=np.float64)

    def get_turbulence_linear(self, timestep):
        return self._get_turbulence(timestep, "linear")

    def get_turbulence_angular(self, timestep):
        return self._get_turbulence(timestep, "angular")

    def _get_turbulence(self, timestep, component):
        if timestep >= self.dryden.sim_length:
            self.dryden.simulate(self.turbulence_sim_length)
        if component == "linear":
            return self.dryden.vel_lin[:, timestep]
        return self.dryden.vel_ang[:, timestep]


# ------------------------------------------------------------------------------------------------------ the simulator
class PyFly:
    REQUIRED_VARIABLES = ["alpha", "beta", "roll", "pitch", "yaw", "omega_p", "omega_q", "omega_r", "position_n",
                          "position_e", "position_d", "velocity_u", "velocity_v", "velocity_w", "Va",
                          "elevator", "aileron", "rudder", "throttle"]

    def __init__(self, config_path=osp.join(PARAMS_DIR, "pyfly_config.json"),
                 parameter_path=osp.join(PARAMS_DIR, "x8_param.json"), config_kw=None):
        def set_config_attrs(parent, kws):
            for attr, val in kws.items():
                if isinstance(val, dict):
                    set_config_attrs(parent[attr], val)
                else:
                    parent[attr] = val

        with open(parameter_path) as f:
            self.params = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
        p = self.params
        self.I = np.array([[p["Jx"], 0, -p["Jxz"]], [0, p["Jy"], 0], [-p["Jxz"], 0, p["Jz"]]])
        I = self.I
        g0 = I[0, 0] * I[2, 2] - I[0, 2] ** 2
        self.gammas = [g0,
                       (np.abs(I[0, 2]) * (I[0, 0] - I[1, 1] + I[2, 2])) / g0,
                       (I[2, 2] * (I[2, 2] - I[1, 1]) + I[0, 2] ** 2) / g0,
                       I[2, 2] / g0,
                       np.abs(I[0, 2]) / g0,
                       (I[2, 2] - I[0, 0]) / I[1, 1],
                       I[0, 2] / I[1, 1],
                       ((I[0, 0] - I[1, 1]) * I[0, 0] + I[0, 2] ** 2) / g0,
                       I[0, 0] / g0]
        p["ar"] = p["b"] ** 2 / p["S_wing"]

        with open(config_path) as f:
            self.cfg = json.load(f)
        if config_kw is not None:
            set_config_attrs(self.cfg, copy.deepcopy(config_kw))

        self.state = {}
        self.attitude_states = ["roll", "pitch", "yaw"]
        self.actuator_states = ["elevator", "aileron", "rudder", "throttle", "elevon_left", "elevon_right"]
        self.model_inputs = ["elevator", "aileron", "rudder", "throttle"]
        missing = set(self.REQUIRED_VARIABLES) - set(v["name"] for v in self.cfg["variables"])
        if missing:
            raise Exception("Missing required variable(s) in config file: {}".format(",".join(missing)))

        self.dt = self.cfg["dt"]
        self.rho = self.cfg["rho"]
        self.g = self.cfg["g"]
        self.drag_model = self.cfg.get("drag_model", "induced")
        self.wind = Wind(mag_min=self.cfg["wind_magnitude_min"], mag_max=self.cfg["wind_magnitude_max"],
                         turbulence=self.cfg["turbulence"], turbulence_intensity=self.cfg["turbulence_intensity"],
                         sim_length=self.cfg.get("turbulence_sim_length", 250), dt=self.cfg["dt"], b=p["b"])

        self.state["attitude"] = AttitudeQuaternion()
        self.attitude_states_with_constraints = []
        self.actuation = Actuation(model_inputs=self.model_inputs, actuator_inputs=self.cfg["actuation"]["inputs"],
                                   dynamics=self.cfg["actuation"]["dynamics"])
        for v in self.cfg["variables"]:
            if v["name"] in self.attitude_states and any(
                    v.get(a, None) is not None for a in ["constraint_min", "constraint_max", "value_min", "value_max"]):
                self.attitude_states_with_constraints.append(v["name"])
            if v["name"] in self.actuator_states:
                self.state[v["name"]] = ControlVariable(**v)
                self.actuation.add_state(self.state[v["name"]])
            else:
                self.state[v["name"]] = Variable(**v)
        self.actuation.finalize()
        self.plots = []
        self.cur_sim_step = None
        self.n_rhs_evals = 0          # oracle-only instrumentation (not in PyFly)
        self.last_step_nfev = 0

    # -- surface used by fixed_wing.py -----------------------------------------------------------------------------
    def seed(self, seed):
        for i, var in enumerate(self.state.values()):
            var.seed(seed + i)
        self.wind.seed(seed)

    def reset(self, state=None, turbulence_noise=None):
        self.cur_sim_step = 0
        for name, var in self.state.items():
            if name in ["Va", "alpha", "beta", "attitude"] or "wind" in name or isinstance(var, ControlVariable):
                continue
            var_init = state[name] if state is not None and name in state else None
            var.reset(value=var_init)
        self.actuation.reset(state)
        wind_init = None
        if state is not None:
            if "wind" in state:
                wind_init = state["wind"]
            elif all(comp in state for comp in ["wind_n", "wind_e", "wind_d"]):
                wind_init = [state["wind_n"], state["wind_e"], state["wind_d"]]
        self.wind.reset(wind_init, turbulence_noise)
        Theta = self.get_states_vector(["roll", "pitch", "yaw"])
        vel = np.array(self.get_states_vector(["velocity_u", "velocity_v", "velocity_w"]))
        Va, alpha, beta = self._calculate_airspeed_factors(Theta, vel)
        self.state["Va"].reset(Va)
        self.state["alpha"].reset(alpha)
        self.state["beta"].reset(beta)
        self.state["attitude"].reset(Theta)

    def get_states_vector(self, states, attribute="value"):
        return [getattr(self.state[s], attribute) for s in states]

    def step(self, commands):
        success = True
        info = {}
        self.actuation.set_and_constrain_commands(commands)
        y0 = list(self.state["attitude"].value)
        y0.extend(self.get_states_vector(["omega_p", "omega_q", "omega_r", "position_n", "position_e", "position_d",
                                          "velocity_u", "velocity_v", "velocity_w"]))
        y0.extend(self.actuation.get_values())
        y0 = np.array(y0, dtype=np.float64)
        try:
            sol = scipy.integrate.solve_ivp(fun=lambda t, y: self._dynamics(t, y), t_span=(0, self.dt), y0=y0)
            self.last_step_nfev = sol.nfev
            self._set_states_from_ode_solution(sol.y[:, -1], save=True)
            Theta = self.get_states_vector(["roll", "pitch", "yaw"])
            vel = np.array(self.get_states_vector(["velocity_u", "velocity_v", "velocity_w"]))
            Va, alpha, beta = self._calculate_airspeed_factors(Theta, vel)
            self.state["Va"].set_value(Va)
            self.state["alpha"].set_value(alpha)
            self.state["beta"].set_value(beta)
        except ConstraintException as e:
            success = False
            info = {"termination": e.variable}
        self.cur_sim_step += 1
        return success, info

    def render(self, *a, **k):
        raise NotImplementedError("rendering is out of scope (SURVEY §2 #16)")

    # -- dynamics ----------------------------------------------------------------------------------------------------
    def _dynamics(self, t, y, control_sp=None):
        self.n_rhs_evals += 1
        if t > 0:
            self._set_states_from_ode_solution(y, save=False)
        attitude = y[:4]
        omega = self.get_states_vector(["omega_p", "omega_q", "omega_r"])
        vel = np.array(self.get_states_vector(["velocity_u", "velocity_v", "velocity_w"]))
        u_states = self.get_states_vector(self.model_inputs)
        f, tau = self._forces(attitude, omega, vel, u_states)
        return np.concatenate([self._f_attitude_dot(t, attitude, omega), self._f_omega_dot(t, omega, tau),
                               self._f_p_dot(t, vel, attitude), self._f_v_dot(t, vel, omega, f),
                               self._f_u_dot(t, control_sp)])

    def _forces(self, attitude, omega, vel, controls):
        elevator, aileron, rudder, throttle = controls
        p, q, r = omega
        if self.wind.turbulence:
            p_w, q_w, r_w = self.wind.get_turbulence_angular(self.cur_sim_step)
            p, q, r = p - p_w, q - q_w, r - r_w
        Va, alpha, beta = self._calculate_airspeed_factors(attitude, vel)
        Va = self.state["Va"].apply_conditions(Va)
        alpha = self.state["alpha"].apply_conditions(alpha)
        beta = self.state["beta"].apply_conditions(beta)
        P = self.params
        pre_fac = 0.5 * self.rho * Va ** 2 * P["S_wing"]
        e0, e1, e2, e3 = attitude
        fg_b = P["mass"] * self.g * np.array([2 * (e1 * e3 - e2 * e0), 2 * (e2 * e3 + e1 * e0),
                                              e3 ** 2 + e0 ** 2 - e1 ** 2 - e2 ** 2])
        C_L_alpha_lin = P["C_L_0"] + P["C_L_alpha"] * alpha
        a_0, M, e, ar = P["a_0"], P["M"], P["e"], P["ar"]
        C_D_p, C_m_fp, C_m_alpha, C_m_0 = P["C_D_p"], P["C_m_fp"], P["C_m_alpha"], P["C_m_0"]
        sigma = (1 + np.exp(-M * (alpha - a_0)) + np.exp(M * (alpha + a_0))) / (
            (1 + np.exp(-M * (alpha - a_0))) * (1 + np.exp(M * (alpha + a_0))))
        C_L_alpha = (1 - sigma) * C_L_alpha_lin + sigma * (2 * np.sign(alpha) * np.sin(alpha) ** 2 * np.cos(alpha))
        f_lift_s = pre_fac * (C_L_alpha + P["C_L_q"] * P["c"] / (2 * Va) * q + P["C_L_delta_e"] * elevator)
        if self.drag_model == "induced":
            C_D_alpha = C_D_p + (1 - sigma) * C_L_alpha_lin ** 2 / (np.pi * e * ar) + sigma * (
                2 * np.sign(alpha) * math.pow(np.sin(alpha), 3))
        else:  # "polynomial"
            C_D_alpha = P["C_D_0"] + P["C_D_alpha1"] * alpha + P["C_D_alpha2"] * alpha ** 2
        C_D_beta = P["C_D_beta1"] * beta + P["C_D_beta2"] * beta ** 2
        f_drag_s = pre_fac * (C_D_alpha + C_D_beta + P["C_D_q"] * P["c"] / (2 * Va) * q
                              + P["C_D_delta_e"] * elevator ** 2)
        C_m = (1 - sigma) * (C_m_0 + C_m_alpha * alpha) + sigma * (C_m_fp * np.sign(alpha) * np.sin(alpha) ** 2)
        m = pre_fac * P["c"] * (C_m + P["C_m_q"] * P["b"] / (2 * Va) * q + P["C_m_delta_e"] * elevator)
        b2Va = P["b"] / (2 * Va)
        f_y = pre_fac * (P["C_Y_0"] + P["C_Y_beta"] * beta + P["C_Y_p"] * b2Va * p + P["C_Y_r"] * b2Va * r
                         + P["C_Y_delta_a"] * aileron + P["C_Y_delta_r"] * rudder)
        l = pre_fac * P["b"] * (P["C_l_0"] + P["C_l_beta"] * beta + P["C_l_p"] * b2Va * p + P["C_l_r"] * b2Va * r
                                + P["C_l_delta_a"] * aileron + P["C_l_delta_r"] * rudder)
        n = pre_fac * P["b"] * (P["C_n_0"] + P["C_n_beta"] * beta + P["C_n_p"] * b2Va * p + P["C_n_r"] * b2Va * r
                                + P["C_n_delta_a"] * aileron + P["C_n_delta_r"] * rudder)
        f_aero = np.dot(self._rot_b_v(np.array([0, alpha, beta])), np.array([-f_drag_s, f_y, -f_lift_s]))
        tau_aero = np.array([l, m, n])
        Vd = Va + throttle * (P["k_motor"] - Va)
        f_prop = np.array([0.5 * self.rho * P["S_prop"] * P["C_prop"] * Vd * (Vd - Va), 0, 0])
        tau_prop = np.array([-P["k_T_P"] * (P["k_Omega"] * throttle) ** 2, 0, 0])
        return f_prop + fg_b + f_aero, tau_aero + tau_prop

    def _f_attitude_dot(self, t, attitude, omega):
        p, q, r = omega
        T = np.array([[0, -p, -q, -r], [p, 0, r, -q], [q, -r, 0, p], [r, q, -p, 0]])
        return 0.5 * np.dot(T, attitude)

    def _f_omega_dot(self, t, omega, tau):
        g = self.gammas
        return np.array([
            g[1] * omega[0] * omega[1] - g[2] * omega[1] * omega[2] + g[3] * tau[0] + g[4] * tau[2],
            g[5] * omega[0] * omega[2] - g[6] * (omega[0] ** 2 - omega[2] ** 2) + tau[1] / self.I[1, 1],
            g[7] * omega[0] * omega[1] - g[1] * omega[1] * omega[2] + g[4] * tau[0] + g[8] * tau[2]])

    def _f_v_dot(self, t, v, omega, f):
        m = self.params["mass"]
        return np.array([omega[2] * v[1] - omega[1] * v[2] + f[0] / m,
                         omega[0] * v[2] - omega[2] * v[0] + f[1] / m,
                         omega[1] * v[0] - omega[0] * v[1] + f[2] / m])

    def _f_p_dot(self, t, v, attitude):
        e0, e1, e2, e3 = attitude
        T = np.array([[e1 ** 2 + e0 ** 2 - e2 ** 2 - e3 ** 2, 2 * (e1 * e2 - e3 * e0), 2 * (e1 * e3 + e2 * e0)],
                      [2 * (e1 * e2 + e3 * e0), e2 ** 2 + e0 ** 2 - e1 ** 2 - e3 ** 2, 2 * (e2 * e3 - e1 * e0)],
                      [2 * (e1 * e3 - e2 * e0), 2 * (e2 * e3 + e1 * e0), e3 ** 2 + e0 ** 2 - e1 ** 2 - e2 ** 2]])
        return np.dot(T, v)

    def _f_u_dot(self, t, setpoints):
        return self.actuation.rhs(setpoints)

    def _rot_b_v(self, attitude):
        if len(attitude) == 3:
            phi, th, psi = attitude
            return np.array([
                [np.cos(th) * np.cos(psi), np.cos(th) * np.sin(psi), -np.sin(th)],
                [np.sin(phi) * np.sin(th) * np.cos(psi) - np.cos(phi) * np.sin(psi),
                 np.sin(phi) * np.sin(th) * np.sin(psi) + np.cos(phi) * np.cos(psi), np.sin(phi) * np.cos(th)],
                [np.cos(phi) * np.sin(th) * np.cos(psi) + np.sin(phi) * np.sin(psi),
                 np.cos(phi) * np.sin(th) * np.sin(psi) - np.sin(phi) * np.cos(psi), np.cos(phi) * np.cos(th)]])
        e0, e1, e2, e3 = attitude
        return np.array([[-1 + 2 * (e0 ** 2 + e1 ** 2), 2 * (e1 * e2 + e3 * e0), 2 * (e1 * e3 - e2 * e0)],
                         [2 * (e1 * e2 - e3 * e0), -1 + 2 * (e0 ** 2 + e2 ** 2), 2 * (e2 * e3 + e1 * e0)],
                         [2 * (e1 * e3 + e2 * e0), 2 * (e2 * e3 - e1 * e0), -1 + 2 * (e0 ** 2 + e3 ** 2)]])

    def _calculate_airspeed_factors(self, attitude, vel):
        if self.wind.turbulence:
            turbulence = self.wind.get_turbulence_linear(self.cur_sim_step)
        else:
            turbulence = np.zeros(3)
        wind_vec = np.dot(self._rot_b_v(attitude), self.wind.steady) + turbulence
        airspeed_vec = vel - wind_vec
        Va = np.linalg.norm(airspeed_vec)
        alpha = np.arctan2(airspeed_vec[2], airspeed_vec[0])
        beta = np.arcsin(airspeed_vec[1] / Va)
        return Va, alpha, beta

    def _set_states_from_ode_solution(self, ode_sol, save):
        self.state["attitude"].set_value(ode_sol[:4] / np.linalg.norm(ode_sol[:4]), save=save)
        if save:
            euler = self.state["attitude"].as_euler_angle()
            self.state["roll"].set_value(euler["roll"], save=save)
            self.state["pitch"].set_value(euler["pitch"], save=save)
            self.state["yaw"].set_value(euler["yaw"], save=save)
        else:
            for state in self.attitude_states_with_constraints:
                self.state[state].set_value(self.state["attitude"].as_euler_angle(state), save=save)
        names = ["omega_p", "omega_q", "omega_r", "position_n", "position_e", "position_d",
                 "velocity_u", "velocity_v", "velocity_w"]
        for i, n in enumerate(names):
            self.state[n].set_value(ode_sol[4 + i], save=save)
        self.actuation.set_states(ode_sol[13:], save=save)


class PIDController:
    """Restatement of pyfly/pid_controller.py (RECALLED): the controller behind the PID golden traces
    (evaluate_controller.py:6,141-151,202)."""

    def __init__(self, dt=0.01):
        self.k_p_V, self.k_i_V = 0.5, 0.1
        self.k_p_phi, self.k_i_phi, self.k_d_phi = 1, 0, 0.5
        self.k_p_theta, self.k_i_theta, self.k_d_theta = -4, -0.75, -0.1
        self.delta_a_min, self.delta_a_max = np.radians(-30), np.radians(30)
        self.delta_e_min, self.delta_e_max = np.radians(-30), np.radians(35)
        self.dt = dt
        self.va_r = self.phi_r = self.theta_r = None
        self.int_va = self.int_roll = self.int_pitch = 0

    def set_reference(self, phi, theta, va):
        self.va_r, self.phi_r, self.theta_r = va, phi, theta

    def reset(self):
        self.int_va = self.int_roll = self.int_pitch = 0

    def get_action(self, phi, theta, va, omega):
        e_V_a = va - self.va_r
        e_phi = phi - self.phi_r
        e_theta = theta - self.theta_r
        self.int_va = self.int_va + self.dt * e_V_a
        self.int_roll = self.int_roll + self.dt * e_phi
        self.int_pitch = self.int_pitch + self.dt * e_theta
        delta_t = 0 - self.k_p_V * e_V_a - self.k_i_V * self.int_va
        delta_a = - self.k_p_phi * e_phi - self.k_i_phi * self.int_roll - self.k_d_phi * omega[0]
        delta_e = 0 - self.k_p_theta * e_theta - self.k_i_theta * self.int_pitch - self.k_d_theta * omega[1]
        delta_t = np.clip(delta_t, 0, 1.0)
        delta_a = np.clip(delta_a, self.delta_a_min, self.delta_a_max)
        delta_e = np.clip(delta_e, self.delta_e_min, self.delta_e_max)
        return np.asarray([delta_e, delta_a, delta_t])
