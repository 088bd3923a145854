"""Which aircraft make single env steps expensive?  Runs a configuration at N envs, records per step the largest dopri5
attempt count and the state of that aircraft BEFORE the step, and saves the worst cases for a CPU replay
(scripts/replay_straggler.py).   python scripts/gpu_straggler_probe.py CONFIG_KEY N STEPS"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from baseline_report import CONFIGS  # noqa: E402
from fwgym_b200 import FixedWingVecEnv  # noqa: E402
from fwgym_b200.config import DEFAULT_ENV_CONFIG  # noqa: E402

key, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
c = CONFIGS[key]
vec = FixedWingVecEnv(os.path.join(os.path.dirname(DEFAULT_ENV_CONFIG), c["config"]), n, config_kw=c["config_kw"],
                      sim_config_kw=c["sim_kw"], seed=7)
vec.reset()
g = torch.Generator(device="cuda"); g.manual_seed(1)
acts = torch.rand((16, n, 3), generator=g, device="cuda") * 2 - 1
rows = vec.state_rows()
cases, hist = [], np.zeros(64, dtype=np.int64)
for t in range(steps):
    before = vec.get_state()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, _, done, term = vec.step_tensors(acts[t % 16])
    e1.record()
    k = vec.last_attempts()
    kmax, imax = int(k.max()), int(k.argmax())
    torch.cuda.synchronize()
    hist += np.bincount(np.minimum(k.cpu().numpy(), 63), minlength=64)
    nonfinite = int((~torch.isfinite(before[:27])).any(0).sum())
    if kmax >= 30 or t % 40 == 0:
        print("step %d: %.0f us, max k %d (env %d, term %d, steps_count %d), envs with k>=16: %d, non-finite states: %d"
              % (t, e0.elapsed_time(e1) * 1e3, kmax, imax, int(term[imax]), int(before[rows.index("steps_count"), imax]),
                 int((k >= 16).sum()), nonfinite), flush=True)
    if kmax >= 30 and len(cases) < 12:
        cases.append({"step": t, "env": imax, "k": kmax, "term": int(term[imax]), "action": acts[t % 16][imax].cpu().tolist(),
                      "state": {r: float(before[i, imax]) for i, r in enumerate(rows) if r not in ("ring", "param")}})
print("attempt histogram (bins 0..62, 63+):", hist.tolist())
json.dump({"config": key, "cases": cases}, open(os.path.join(ROOT, "gpurun_out", "stragglers_%s.json" % key.replace("'", "p")), "w"), indent=1)
