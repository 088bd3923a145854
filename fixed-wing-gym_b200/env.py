"""Single-env facade with the reference's constructor / step / reset signatures (fixed_wing.py:14,287,338): it is
just N = 1 of the batched VecEnv, returning numpy values like the reference does."""
import numpy as np
import torch

from .vec_env import FixedWingVecEnv, term_name


class FixedWingAircraft:
    def __init__(self, config_path=None, sampler=None, sim_config_path=None, sim_parameter_path=None, config_kw=None,
                 sim_config_kw=None, device="cuda:0", precision="fp64"):
        self._vec = FixedWingVecEnv(config_path, 1, device, sampler, sim_config_path, sim_parameter_path, config_kw,
                                    sim_config_kw, precision=precision, auto_reset=False)
        self._vec.enable_f64_outputs(True)
        self.cfg = self._vec.cfg
        self.observation_space = self._vec.observation_space
        self.action_space = self._vec.action_space
        self.steps_max = self.cfg["steps_max"]
        self.training = True

    @property
    def simulator(self):
        return self._vec.get_attr("simulator")[0]

    @property
    def target(self):
        return self._vec.get_attr("target")[0]

    @property
    def steps_count(self):
        return self._vec.get_attr("steps_count")[0]

    def seed(self, seed=None):
        return [self._vec.seed(seed)[0]]

    def set_curriculum_level(self, level):
        self._vec.set_curriculum_level(level)

    def _obs(self):
        return self._vec._obs64[0].cpu().numpy().reshape(self._vec.cc.obs_shape).copy()

    def reset(self, state=None, target=None, **sim_reset_kw):
        noise = sim_reset_kw.pop("turbulence_noise", None)   # PyFly.reset's only other keyword (fixed_wing.py:308)
        if sim_reset_kw:
            raise TypeError("reset() got unexpected simulator keywords %s" % sorted(sim_reset_kw))
        self._vec.reset(state=state, target=target, turbulence_noise=noise)
        return self._obs()

    def step(self, action):
        action = np.asarray(action, dtype=np.float64)
        assert not np.any(np.isnan(action))
        a = torch.as_tensor(action.reshape(1, 3), dtype=torch.float64, device=self._vec.device)
        _, _, done, term = self._vec.step_tensors(a)
        info = {}
        d = bool(done[0].item())
        if d:
            info["termination"] = term_name(int(term[0].item()))
        info["target"] = self.target
        return self._obs(), float(self._vec._rew64[0].item()), d, info

    def close(self):
        self._vec.close()
