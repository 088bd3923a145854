#!/usr/bin/env python
"""BASELINE configs[3]: PPO (PPO2 defaults, MLP 2x64 tanh) on a 16384-env FixedWingVecEnv, everything on the GPU, with
the reference's curriculum rule (train_rl_controller.py:80-87, start level 0.25 as :162).
    python examples/train_ppo.py --envs 16384 --steps 268435456 --log profiles/ppo_run.jsonl
Prints one JSON record per iteration (episodes ended, success rate, mean episode return / length, curriculum level,
losses, rollout / update milliseconds) and a final summary with env-steps/s inside training and the time split."""
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
    ap.add_argument("--no-curriculum", action="store_true")      # --disable-curriculum of the reference: level 1
    ap.add_argument("--minibatches", type=int, default=4)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--lr", type=float, default=2.5e-4)
    ap.add_argument("--n-steps", type=int, default=128)
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--log", default=None, help="also append the records to this JSON-lines file")
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from fwgym_b200 import FixedWingVecEnv, ppo
    from fwgym_b200.config import DEFAULT_ENV_CONFIG
    cfg = os.path.join(os.path.dirname(DEFAULT_ENV_CONFIG), a.config)
    env = FixedWingVecEnv(cfg, a.envs, seed=0, precision=a.precision)
    cb = ppo.CurriculumCallback(env, level=1.0 if a.no_curriculum else a.curriculum)
    logf = open(a.log, "w") if a.log else None

    def log(rec):
        line = json.dumps(rec)
        print(line, flush=True)
        if logf:
            logf.write(line + "\n")

    model, norm, stats = ppo.train(env, a.steps, n_steps=a.n_steps, n_minibatches=a.minibatches, n_epochs=a.epochs, lr=a.lr,
                                   log=log, callback=cb)
    hist = stats.pop("history")
    stats.update(envs=a.envs, config=a.config, precision=a.precision, minibatches=a.minibatches, epochs=a.epochs, lr=a.lr,
                 curriculum_bumps=cb.bumps, final_curriculum_level=cb.level,
                 first_iterations={k: hist[min(1, len(hist) - 1)][k] for k in ("success_rate", "mean_episode_return")},
                 last_iterations={k: sum(h[k] for h in hist[-5:]) / len(hist[-5:]) for k in ("success_rate", "mean_episode_return")})
    log({"summary": stats})
