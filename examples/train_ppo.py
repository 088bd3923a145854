#!/usr/bin/env python
"""BASELINE configs[3]: PPO (PPO2 defaults, MLP 2x64 tanh) on a 16384-env FixedWingVecEnv, everything on the GPU.
    python examples/train_ppo.py --envs 16384 --steps 8388608
Prints env-steps/s inside training and the env / policy / update split (fwgym_b200.ppo.train)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=16384 * 128 * 4)
    ap.add_argument("--config", default="fixed_wing_config_examples.json")
    ap.add_argument("--curriculum", type=float, default=0.25)     # train_rl_controller.py:162
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from fwgym_b200 import FixedWingVecEnv, ppo
    from fwgym_b200.config import DEFAULT_ENV_CONFIG
    cfg = os.path.join(os.path.dirname(DEFAULT_ENV_CONFIG), a.config)
    env = FixedWingVecEnv(cfg, a.envs, seed=0)
    env.env_method("set_curriculum_level", a.curriculum)
    model, norm, stats = ppo.train(env, a.steps, log=lambda r: print(json.dumps(r), flush=True))
    stats.pop("history")
    print(json.dumps(stats))
