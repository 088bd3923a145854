"""Oracle for the device-side training plumbing: numpy twins of stable-baselines' RunningMeanStd /
VecNormalize (the reference trains over `VecNormalize(SubprocVecEnv(...))`, train_rl_controller.py:223) and of PPO2's GAE
recursion, against fwgym_b200.ppo.RunningMeanStd / DeviceVecNormalize / compute_gae.  The torch classes are device
agnostic, so this runs on the CPU here and on cuda:0 in tests/test_gpu_fullsize.py::test_vecnormalize_and_gae_on_device."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class NumpyRunningMeanStd:
    """stable_baselines/common/running_mean_std.py: parallel-variance update from batch moments, count starts at 1e-4."""

    def __init__(self, shape):
        self.mean, self.var, self.count = np.zeros(shape), np.ones(shape), 1e-4

    def update(self, x):
        bm, bv, bc = x.mean(axis=0), x.var(axis=0), x.shape[0]
        delta, tot = bm - self.mean, self.count + bc
        m2 = self.var * self.count + bv * bc + np.square(delta) * self.count * bc / tot
        self.mean, self.var, self.count = self.mean + delta * bc / tot, m2 / tot, tot


class NumpyVecNormalize:
    """stable_baselines VecNormalize.step_wait / reset (norm_obs, norm_reward, clip 10, gamma 0.99, epsilon 1e-8)."""

    def __init__(self, n, obs_dim, gamma=0.99, clip=10.0, eps=1e-8):
        self.obs_rms, self.ret_rms = NumpyRunningMeanStd((obs_dim,)), NumpyRunningMeanStd(())
        self.ret, self.gamma, self.clip, self.eps = np.zeros(n), gamma, clip, eps

    def obs(self, o):
        self.obs_rms.update(o)
        return np.clip((o - self.obs_rms.mean) / np.sqrt(self.obs_rms.var + self.eps), -self.clip, self.clip)

    def step(self, o, r, d):
        self.ret = self.ret * self.gamma + r
        on = self.obs(o)
        self.ret_rms.update(self.ret)
        rn = np.clip(r / np.sqrt(self.ret_rms.var + self.eps), -self.clip, self.clip)
        self.ret[d] = 0
        return on, rn


def numpy_gae(rew, val, done, last_val, gamma, lam):
    """PPO2 runner (stable_baselines/ppo2/ppo2.py): done[t] = the episode ended in step t."""
    T = rew.shape[0]
    adv, last = np.zeros_like(rew), 0
    for t in reversed(range(T)):
        nv = last_val if t == T - 1 else val[t + 1]
        nonterm = 1.0 - done[t]
        delta = rew[t] + gamma * nv * nonterm - val[t]
        adv[t] = last = delta + gamma * lam * nonterm * last
    return adv


class ScriptedVecEnv:
    """Stands in for FixedWingVecEnv: replays scripted observations / rewards / dones."""

    def __init__(self, obs, rew, done, device):
        self.device, self.num_envs, self.obs_dim = torch.device(device), obs.shape[1], obs.shape[2]
        self._o = torch.as_tensor(obs, dtype=torch.float32, device=device)
        self._r = torch.as_tensor(rew, dtype=torch.float32, device=device)
        self._d = torch.as_tensor(done.astype(np.uint8), device=device)
        self.t = 0

    def reset(self):
        self.t = 0
        return self._o[0]

    def step_tensors(self, actions):
        self.t += 1
        return self._o[self.t], self._r[self.t - 1], self._d[self.t - 1], None


def check_vecnormalize_and_gae(device):
    from fwgym_b200 import ppo
    rng = np.random.RandomState(3)
    T, n, od = 40, 257, 12
    obs = (rng.standard_normal((T + 1, n, od)) * rng.uniform(0.1, 30, od) + rng.uniform(-5, 5, od)).astype(np.float32)
    rew = (rng.standard_normal((T, n)) * 0.5 - 0.3).astype(np.float32)
    done = rng.uniform(size=(T, n)) < 0.05
    venv = ScriptedVecEnv(obs, rew, done, device)
    dvn, ref = ppo.DeviceVecNormalize(venv), NumpyVecNormalize(n, od)
    o_dev = dvn.reset()
    o_ref = ref.obs(obs[0].astype(np.float64))
    assert np.abs(o_dev.cpu().numpy() - o_ref).max() < 1e-5
    for t in range(T):
        o_dev, r_dev, d_dev, raw = dvn.step(None)
        o_ref, r_ref = ref.step(obs[t + 1].astype(np.float64), rew[t].astype(np.float64), done[t])
        assert np.abs(o_dev.cpu().numpy() - o_ref).max() < 1e-5 and np.abs(r_dev.cpu().numpy() - r_ref).max() < 1e-5
        assert np.array_equal(d_dev.cpu().numpy(), done[t])
    for dev_rms, ref_rms in ((dvn.obs_rms, ref.obs_rms), (dvn.ret_rms, ref.ret_rms)):
        assert np.allclose(dev_rms.mean.cpu().numpy(), ref_rms.mean, rtol=1e-11, atol=1e-12)
        assert np.allclose(dev_rms.var.cpu().numpy(), ref_rms.var, rtol=1e-11, atol=1e-12)
        assert abs(float(dev_rms.count) - ref_rms.count) < 1e-9
    assert np.allclose(dvn.ret.cpu().numpy(), ref.ret, rtol=1e-12, atol=1e-12)
    # GAE
    val = rng.standard_normal((T, n))
    last_val = rng.standard_normal(n)
    want = numpy_gae(rew.astype(np.float64), val, done.astype(np.float64), last_val, 0.99, 0.95)
    tt = lambda a: torch.as_tensor(a, dtype=torch.float64, device=device)
    got = ppo.compute_gae(tt(rew), tt(val), tt(done.astype(np.float64)), tt(last_val), 0.99, 0.95,
                          torch.zeros((T, n), dtype=torch.float64, device=device))
    assert np.allclose(got.cpu().numpy(), want, rtol=1e-12, atol=1e-12)


def test_vecnormalize_and_gae_match_stable_baselines_semantics():
    check_vecnormalize_and_gae("cpu")


def test_curriculum_callback_follows_the_reference_rule():
    """train_rl_controller.py:80-87: cooldown first, then level = min(2 * success, 1) when success > level, 15 callbacks
    of cooldown after a bump, nothing once the level is 1."""
    from fwgym_b200 import ppo

    class Env:
        levels = []

        def env_method(self, name, level):
            assert name == "set_curriculum_level"
            self.levels.append(level)

    env = Env()
    cb = ppo.CurriculumCallback(env, level=0.25, cooldown=2)
    rec = lambda i, s: {"iter": i, "episodes": 50, "success_rate": s}
    for i in range(2):
        cb(rec(i, 0.9))                    # cooling down
    assert cb.level == 0.25
    cb(rec(2, 0.2))                        # below the level: nothing
    assert cb.level == 0.25
    cb(rec(3, 0.4))                        # 0.4 > 0.25 -> 0.8
    assert cb.level == 0.8 and env.levels == [0.25, 0.8]
    for i in range(15):
        cb(rec(4 + i, 0.95))
    assert cb.level == 0.8
    cb(rec(19, 0.95))
    assert cb.level == 1.0 and cb.bumps == [(3, 0.8), (19, 1.0)]
    cb(rec(20, 1.0))
    assert env.levels == [0.25, 0.8, 1.0]
