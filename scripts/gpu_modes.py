"""Step time of the simulator in its different kernel instantiations at 65536 envs (GPU box): shipped fp64, forced
generic, per-env randomised parameters (FwSpecRand), fp32."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fwgym_b200 import FixedWingVecEnv
from fwgym_b200.config import DEFAULT_ENV_CONFIG
n = 65536
P = os.path.dirname(DEFAULT_ENV_CONFIG)
def run(name, cfg, **kw):
    vec = FixedWingVecEnv(cfg, n, config_kw=bench.CONFIG_KW, sim_config_kw=bench.SIM_KW, seed=1, **kw)
    vec.reset()
    acts = torch.rand((40, n, 3), device="cuda") * 2 - 1
    for i in range(5):
        vec.step_tensors(acts[i])
    vec.reset_counters()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30):
        vec.step_tensors(acts[5 + i])
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    c = vec.counters()
    print("%-28s %-34s %.1f us/step  %.3g env-steps/s  k=%.2f" % (name, vec.kernel_variant(), us, n / us * 1e6, c["attempts"] / c["env_steps"]))
    vec.close()
mode = sys.argv[1] if len(sys.argv) > 1 else "all"
if mode == "generic":
    run("fp64 forced generic", DEFAULT_ENV_CONFIG)
else:
    run("fp64 shipped", DEFAULT_ENV_CONFIG)
    run("fp64 randomised parameters", os.path.join(P, "fixed_wing_config_randomised.json"))
    run("fp32 shipped", DEFAULT_ENV_CONFIG, precision="fp32")
    subprocess.run([sys.executable, __file__, "generic"], env=dict(os.environ, FWGYM_FORCE_GENERIC="1"))
