"""CPU ORACLE (test infrastructure): numpy twin of fixed-wing-gym_b200/csrc/philox.cuh.

Philox4x32-10 (Salmon et al., SC'11; Random123 known-answer vectors are pinned in tests/test_oracle_cpu.py::test_philox_known_answers) and the draw
conventions both sides share:
    counter = (global env id, tick, stream, index), key = 64-bit seed (lo, hi)
    uniform : 53-bit double from words 0,1
    normal  : Box-Muller on words (0,1) -> u1 in (0,1], (2,3) -> u2; two normals per block (cos, sin branch)
It replaces the reference's numpy RandomState draws (fixed_wing.py:57,220,334,492-507,795,837 and PyFly's per-variable
RandomState) so that oracle and CUDA consume identical random numbers.
"""
import math

import numpy as np

RS_INIT, RS_WIND, RS_TURB, RS_ENV_U, RS_ENV_N = 0, 1, 2, 3, 4
_M0, _M1, _W0, _W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = [int(x) & _MASK for x in ctr]
    k0, k1 = [int(x) & _MASK for x in key]
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & _MASK, p1 >> 32, p1 & _MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & _MASK, lo1, (hi0 ^ c3 ^ k1) & _MASK, lo0
        k0, k1 = (k0 + _W0) & _MASK, (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def u53(a, b):
    return ((a >> 5) * 67108864.0 + (b >> 6)) * (1.0 / 9007199254740992.0)


class PhiloxKey:
    def __init__(self, seed, env):
        self.k = (seed & _MASK, (seed >> 32) & _MASK)
        self.env = env & _MASK

    def words(self, tick, stream, idx):
        return philox4x32_10((self.env, tick & _MASK, stream, idx & _MASK), self.k)

    def uniform01(self, tick, stream, idx):
        w = self.words(tick, stream, idx)
        return u53(w[0], w[1])

    def uniform(self, tick, stream, idx, lo, hi):
        return lo + (hi - lo) * self.uniform01(tick, stream, idx)

    def normal2(self, tick, stream, idx):
        w = self.words(tick, stream, idx)
        u1 = 1.0 - u53(w[0], w[1])
        u2 = u53(w[2], w[3])
        r = math.sqrt(-2.0 * math.log(u1))
        return r * math.cos(2.0 * math.pi * u2), r * math.sin(2.0 * math.pi * u2)

    def turbulence_noise(self, episode_tick, length):
        """The [4, length] standard-normal array PyFly's Dryden model would draw for the episode (unscaled)."""
        out = np.empty((4, length))
        for s in range(length):
            out[0, s], out[1, s] = self.normal2(episode_tick, RS_TURB, 2 * s)
            out[2, s], out[3, s] = self.normal2(episode_tick, RS_TURB, 2 * s + 1)
        return out


class EnvRandom:
    """Stand-in for FixedWingAircraft.np_random: env-side uniform/normal draws in call order within a tick."""

    def __init__(self, key):
        self.key = key
        self.begin(0)

    def begin(self, tick):
        self.tick, self.n_u, self.n_n, self._z = tick, 0, 0, None

    def uniform(self, low=0.0, high=1.0):
        v = self.key.uniform(self.tick, RS_ENV_U, self.n_u, low, high)
        self.n_u += 1
        return v

    def normal(self, loc=0.0, scale=1.0):
        if self.n_n & 1:
            z = self._z
        else:
            z, self._z = self.key.normal2(self.tick, RS_ENV_N, self.n_n >> 1)
        self.n_n += 1
        return loc + scale * z


class VarRandom:
    """Stand-in for a PyFly Variable's np_random: uniform(init_min, init_max) keyed by the variable id."""

    def __init__(self, key, sv, stream=RS_INIT):
        self.key, self.sv, self.stream, self.tick = key, sv, stream, 0

    def uniform(self, low=0.0, high=1.0):
        return self.key.uniform(self.tick, self.stream, self.sv, low, high)


class WindRandom:
    def __init__(self, key):
        self.key, self.tick, self.n = key, 0, 0

    def uniform(self, low=0.0, high=1.0):
        v = self.key.uniform(self.tick, RS_WIND, self.n, low, high)
        self.n += 1
        return v
