"""Which part of the PPO rollout step cannot be captured into a CUDA graph?  (GPU box)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fwgym_b200 import FixedWingVecEnv, ppo
from fwgym_b200.config import DEFAULT_ENV_CONFIG
cfg = os.path.join(os.path.dirname(DEFAULT_ENV_CONFIG), "fixed_wing_config_examples.json")
env = FixedWingVecEnv(cfg, 2048, seed=4)
norm = ppo.DeviceVecNormalize(env)
model = ppo.ActorCritic(env.obs_dim, 3).cuda()
obs = norm.reset().clone()
act0 = torch.zeros((2048, 3), device="cuda")
def try_capture(name, fn):
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay()
        torch.cuda.synchronize()
        print(name, "OK")
    except Exception as e:
        print(name, "FAILED:", str(e).splitlines()[0])
        torch.cuda.synchronize()
with torch.no_grad():
    try_capture("policy forward", lambda: model.pi(obs))
    try_capture("dist+sample", lambda: model.dist(obs).sample())
    try_capture("env step only", lambda: env.step_tensors(act0))
    try_capture("norm.step", lambda: norm.step(act0))
    try_capture("norm._obs", lambda: norm._obs(obs))
