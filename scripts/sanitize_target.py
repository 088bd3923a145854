"""Target of the compute-sanitizer passes (scripts/gpu_sanitize.sh): every kernel of the step path on small batches —
reset, init / attempt / env kernels with the programmatic-dependent-launch overlap, auto-reset, episode metrics,
parameter randomisation, the host-buffer pipeline, state export / import, the PID kernel."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fwgym_b200 import FixedWingVecEnv, HostStepper  # noqa: E402
from fwgym_b200.config import DEFAULT_ENV_CONFIG  # noqa: E402

PARAMS = os.path.dirname(DEFAULT_ENV_CONFIG)


def run(cfg, n, steps, config_kw=None, sim_kw=None, **kw):
    vec = FixedWingVecEnv(os.path.join(PARAMS, cfg), n, config_kw=config_kw, sim_config_kw=sim_kw, seed=11,
                          keep_terminal_obs=True, **kw)
    vec.reset()
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    for _ in range(steps):
        a = torch.rand((n, 3), generator=g, device="cuda") * 2.6 - 1.3
        vec.step_tensors(a)
    st = vec.get_state()
    vec.set_state(st)
    vec.step_tensors(torch.zeros((n, 3), device="cuda"))
    torch.cuda.synchronize()
    c = vec.counters()
    assert c["watchdog"] == 0, c
    return vec


def main():
    turb = {"turbulence": True, "turbulence_intensity": "moderate"}
    v = run("fixed_wing_config.json", 16, 3, sim_kw={"turbulence": False})
    v.close()
    # ragged last chunk, several env blocks per SM, overlap on (default), turbulence + noise, auto-resets (steps_max 12)
    v = run("fixed_wing_config.json", 1000 + 77, 16, {"observation": {"noise": {"mean": 0, "var": 0.1}}, "steps_max": 12},
            turb, metrics=True)
    hs = HostStepper(v, depth=2)
    acts = (torch.rand((4, v.num_envs, 3)) * 2 - 1).pin_memory()
    pend = []
    for i in range(4):
        pend.append(hs.submit(acts[i]))
        if len(pend) == 2:
            hs.wait(pend.pop(0))
    while pend:
        hs.wait(pend.pop(0))
    hs.close()
    v.close()
    # generic env kernel + history rings + integrator + resample
    v = run("fixed_wing_config_dev.json", 96, 30,
            {"integration_window": 10, "steps_max": 25,
             "observation": {"length": 5, "step": 1, "shape": "matrix",
                             "states": {6: {"value": "integrator"}, 7: {"value": "relative"}}},
             "target": {"resample_every": 10}}, {"turbulence": False})
    v.close()
    # per-env model parameters (FwSpecRand) and the fp32 instantiation
    v = run("fixed_wing_config_randomised.json", 200, 20, {"steps_max": 9}, {"turbulence": False})
    v.close()
    v = run("fixed_wing_config.json", 333, 8, None, turb, precision="fp32")
    v.close()
    print("sanitize target ok")


if __name__ == "__main__":
    main()
