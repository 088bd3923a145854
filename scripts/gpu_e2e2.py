"""e2e host pipeline: legacy default stream vs a side stream; and back-to-back device loop at the same episode position."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fwgym_b200 import FixedWingVecEnv, HostStepper
from fwgym_b200.config import DEFAULT_ENV_CONFIG
n = 65536
vec = FixedWingVecEnv(DEFAULT_ENV_CONFIG, n, config_kw=bench.CONFIG_KW, sim_config_kw=bench.SIM_KW, seed=20261017)
acts = (torch.rand((64, n, 3)) * 2 - 1).pin_memory()
dacts = acts.cuda()
def e2e(depth, count=30, warm=5):
    vec.reset()
    hs = HostStepper(vec, depth=depth)
    def run(first, cnt):
        pend = []
        for i in range(cnt):
            pend.append(hs.submit(acts[(first + i) % 64]))
            if len(pend) == depth:
                hs.wait(pend.pop(0))
        while pend:
            hs.wait(pend.pop(0))
    run(0, warm)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(warm, count)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    hs.close()
    return dt / count * 1e6
def dev(count=30, warm=5):
    vec.reset()
    for i in range(warm):
        vec.step_tensors(dacts[i])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(count):
        vec.step_tensors(dacts[warm + i])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / count * 1e6
for rep in range(2):
    print("legacy stream: device loop %.1f us/step, e2e depth2 %.1f, depth3 %.1f" % (dev(), e2e(2), e2e(3)))
    with torch.cuda.stream(torch.cuda.Stream()):
        print("side stream:   device loop %.1f us/step, e2e depth2 %.1f, depth3 %.1f" % (dev(), e2e(2), e2e(3)))
