"""e2e host-pipeline sweep over depth (scripts/, GPU box): env-steps/s through HostStepper."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fwgym_b200 import FixedWingVecEnv, HostStepper
from fwgym_b200.config import DEFAULT_ENV_CONFIG
n = 65536
vec = FixedWingVecEnv(DEFAULT_ENV_CONFIG, n, config_kw=bench.CONFIG_KW, sim_config_kw=bench.SIM_KW, seed=1)
vec.reset()
acts = (torch.rand((64, n, 3)) * 2 - 1).pin_memory()
for depth in [int(x) for x in sys.argv[1:]] or [1, 2, 3, 4]:
    hs = HostStepper(vec, depth=depth)
    def run(first, count):
        pend = []
        for i in range(count):
            pend.append(hs.submit(acts[(first + i) % 64]))
            if len(pend) == depth:
                hs.wait(pend.pop(0))
        while pend:
            hs.wait(pend.pop(0))
    run(0, 8)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(8, 100)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    # host cost of submit alone
    t1 = time.perf_counter()
    s = hs.submit(acts[0])
    t2 = time.perf_counter()
    hs.wait(s)
    print("depth %d: %.4g env-steps/s, %.1f us/step; submit() host time %.1f us" % (depth, n * 100 / dt, dt / 100 * 1e6, (t2 - t1) * 1e6))
    hs.close()
# device-resident loop without L2 flush (what a policy living on the GPU sees): 100 back-to-back steps
dacts = acts.cuda()
for _ in range(8):
    vec.step_tensors(dacts[0])
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for i in range(100):
    vec.step_tensors(dacts[i % 64])
e1.record()
torch.cuda.synchronize()
print("device loop, no flush: %.1f us/step (%s, overlap env %s)" % (e0.elapsed_time(e1) * 10, vec.kernel_variant(), os.environ.get("FWGYM_OVERLAP", "1")))
