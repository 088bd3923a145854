#!/usr/bin/env python
"""BASELINE.md §3 executed on the GPU box: the CPU baseline of the hot path on the host cores (configs 1, 3', 5') next
to the GPU numbers at 4096 / 16 384 / 65 536 envs per GPU for the same configurations.

    python scripts/baseline_report.py [--cpu-seconds 30] [--out profiles/r2_baseline_report.json]

CPU arm: "restated CPU path" (PyFly 0.1.2 is not importable on the box; the oracle port of
FixedWingAircraft + PyFly over scipy solve_ivp), one OS process per host core, one env per process, pre-generated
U(-1,1)^3 actions, auto-reset on done, a 5 s warm-up then a timed run; env-steps/s per core and total, mean RHS
evaluations per step.  GPU arm: device-resident env-steps/s (256 MiB L2 flush between steps, CUDA events), achieved
FP64 FLOP/s = (1080 steps + 3660 attempts) / dynamics-kernel time with attempts counted on the device, fraction of the
measured DFMA peak, lane efficiency of the attempt kernel."""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    "1": dict(config="fixed_wing_config.json", config_kw=None, sim_kw={"turbulence": False},
              what="fixed_wing_config.json, turbulence off"),
    "3'": dict(config="fixed_wing_config.json", config_kw={"observation": {"noise": {"mean": 0, "var": 0.1}}},
               sim_kw={"turbulence": True, "turbulence_intensity": "moderate"},
               what="+ Dryden turbulence moderate + observation noise std 0.1 (the bench workload)"),
    "5'": dict(config="fixed_wing_config_dev.json",
               config_kw={"integration_window": 100,
                          "observation": {"length": 5, "step": 1, "shape": "matrix", "states": {6: {"value": "integrator"}}},
                          "target": {"resample_every": 500}},
               sim_kw={"turbulence": False},
               what="dev config: 5-row matrix observation, integrator target, integration_window 100, resample_every 500"),
}


def _cpu_worker(args):
    wid, key, warm_s, run_s = args
    import numpy as np
    from oracle import harness
    c = CONFIGS[key]
    env = harness.make_env("restated", harness.config_path(c["config"]), c["config_kw"], c["sim_kw"])
    run = harness.OracleRunner(env, seed=4321, env_id=wid)
    run.reset()
    acts = np.random.RandomState(wid).uniform(-1, 1, (4096, 3))
    i = 0
    t_end = time.perf_counter() + warm_s
    while time.perf_counter() < t_end:
        run.step(acts[i % 4096]); i += 1
    n0, e0 = i, sum(run.nfev)
    t0 = time.perf_counter()
    t_end = t0 + run_s
    while time.perf_counter() < t_end:
        run.step(acts[i % 4096]); i += 1
    dt = time.perf_counter() - t0
    return i - n0, dt, sum(run.nfev) - e0


def cpu_arm(seconds, warm):
    cores = os.cpu_count() or 1
    out = {}
    ctx = mp.get_context("fork")
    for key, c in CONFIGS.items():
        with ctx.Pool(cores) as pool:
            res = pool.map(_cpu_worker, [(w, key, warm, seconds) for w in range(cores)], chunksize=1)
        steps = sum(r[0] for r in res)
        rate = sum(r[0] / r[1] for r in res)
        out[key] = {"config": c["what"], "cores": cores, "env_steps_per_s_total": rate, "env_steps_per_s_per_core": rate / cores,
                    "timed_seconds": seconds, "warmup_seconds": warm, "env_steps": steps,
                    "mean_rhs_evals_per_step": sum(r[2] for r in res) / max(1, steps),
                    "kind": "restated CPU path (oracle port of FixedWingAircraft + PyFly, scipy solve_ivp RK45)"}
        print("CPU config %s: %.0f env-steps/s on %d cores (%.1f / core), %.1f RHS evals / step"
              % (key, rate, cores, rate / cores, out[key]["mean_rhs_evals_per_step"]), flush=True)
    return out


GPU_VARIANTS = [("1", {}), ("1", {"dopri5_max_attempts": 64}), ("3'", {}), ("5'", {}), ("5'", {"dopri5_max_attempts": 64})]


def gpu_arm(sizes, steps=40, warm=5, burn=200):
    import torch
    import __graft_entry__ as ge
    ge.build()
    from fwgym_b200 import FixedWingVecEnv, _capi
    from fwgym_b200.config import DEFAULT_ENV_CONFIG
    dev = torch.device("cuda", 0)
    fl = ctypes.c_double()
    _capi.check(_capi.lib().fw_dfma_peak(0, ctypes.byref(fl), None))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    params = os.path.dirname(DEFAULT_ENV_CONFIG)
    out = {"dfma_peak_tflops": fl.value / 1e12, "rows": []}
    for key, extra in GPU_VARIANTS:
        c = CONFIGS[key]
        for n in sizes:
            vec = FixedWingVecEnv(os.path.join(params, c["config"]), n, device=dev, config_kw=c["config_kw"],
                                  sim_config_kw=dict(c["sim_kw"], **extra), seed=7)
            vec.reset()
            g = torch.Generator(device=dev); g.manual_seed(1)
            acts = torch.rand((16, n, 3), generator=g, device=dev) * 2 - 1
            for i in range(burn + warm):
                vec.step_tensors(acts[i % 16])
            vec.reset_counters()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            kmax = torch.zeros((), dtype=torch.int32, device=dev)
            for k in range(steps):
                flush.zero_()
                ev[k][0].record(); vec.step_tensors(acts[k % 16]); ev[k][1].record()
                kmax = torch.maximum(kmax, vec.last_attempts().max())
            torch.cuda.synchronize(dev)
            per = sorted(a.elapsed_time(b) for a, b in ev)
            ms = sum(per)
            ctr = vec.counters()
            vec.set_profiling(True)
            for k in range(10):
                flush.zero_(); vec.step_tensors(acts[k % 16])
            dyn_ms, env_ms, ps = vec.profile()
            c2 = vec.counters()
            flops = 1080.0 * (c2["env_steps"] - ctr["env_steps"]) + 3660.0 * (c2["attempts"] - ctr["attempts"])
            sums = vec.metric_sums()
            row = {"config": key + (" + %s" % json.dumps(extra) if extra else ""), "envs_per_gpu": n,
                   "us_per_step_p50": 1e3 * per[len(per) // 2], "us_per_step_max": 1e3 * per[-1],
                   "max_attempts_of_one_env_step": int(kmax.item()),
                   "episodes_ended": int(sums[0]), "constraint_failures": int(sums[4]), "env_steps_per_s": n * steps / (ms * 1e-3), "us_per_step": 1e3 * ms / steps,
                   "mean_attempts": ctr["attempts"] / max(1, ctr["env_steps"]),
                   "lane_efficiency": ctr["warp_steps"] / max(1.0, 32.0 * ctr["warp_max_attempts"]),
                   "dyn_kernels_us": 1e3 * dyn_ms / ps, "env_kernel_us": 1e3 * env_ms / ps,
                   "achieved_fp64_tflops": flops / (dyn_ms * 1e-3) / 1e12,
                   "frac_of_dfma_peak": flops / (dyn_ms * 1e-3) / fl.value, "kernels": vec.kernel_variant()}
            out["rows"].append(row)
            print("GPU config %s, %6d envs: %.4g env-steps/s (%.1f us / step, p50 %.1f, max %.1f), k mean %.2f max %d, FP64 %.2f "
                  "TFLOP/s = %.3f of peak" % (row["config"], n, row["env_steps_per_s"], row["us_per_step"], row["us_per_step_p50"],
                                              row["us_per_step_max"], row["mean_attempts"], row["max_attempts_of_one_env_step"],
                                              row["achieved_fp64_tflops"], row["frac_of_dfma_peak"]), flush=True)
            vec.close()
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-seconds", type=float, default=30.0)
    ap.add_argument("--cpu-warmup", type=float, default=5.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "baseline_report.json"))
    ap.add_argument("--reuse-cpu", default=None, help="take the CPU arm from an earlier report instead of timing it again")
    a = ap.parse_args()
    if a.reuse_cpu:
        with open(a.reuse_cpu) as f:
            rep = {"cpu": json.load(f)["cpu"]}
    else:
        rep = {"cpu": cpu_arm(a.cpu_seconds, a.cpu_warmup)}      # CPU first: no CUDA context exists in the forking parent
    rep["gpu_1x_b200"] = gpu_arm([4096, 16384, 65536])
    with open(a.out, "w") as f:
        json.dump(rep, f, indent=1)
    print("wrote", a.out)
