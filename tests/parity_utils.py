"""Shared driver for the CUDA-vs-oracle parity tests: N oracle envs (one per global env id) stepped next to the
batched CUDA env on identical actions and identical Philox streams."""
import numpy as np

from oracle import harness

STATE_ROWS = ["q0", "q1", "q2", "q3", "omega_p", "omega_q", "omega_r", "position_n", "position_e", "position_d",
              "velocity_u", "velocity_v", "velocity_w", "elevon_left", "elevon_right", "throttle", "elevon_left_dot",
              "elevon_right_dot", "throttle_dot", "roll", "pitch", "yaw", "Va", "alpha", "beta", "elevator", "aileron"]


def rel_err(a, b, floor=1e-6):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def make_oracles(n, config, config_kw, sim_config_kw, seed, env_offset=0, kind="restated"):
    return [harness.OracleRunner(harness.make_env(kind, config, config_kw, sim_config_kw), seed, env_offset + i)
            for i in range(n)]


def gpu_state(vec):
    rows = vec.state_rows()
    st = vec.get_state().cpu().numpy()
    return np.stack([st[rows.index(r)] for r in STATE_ROWS], axis=1)   # [N, 27]


def run_parity(vec, oracles, actions, check_state=True):
    """Step both for len(actions) steps.  Returns dict of per-step max errors and integer mismatches."""
    import torch
    n = len(oracles)
    vec.enable_f64_outputs(True)
    obs_g = vec.reset()
    obs_o = np.stack([np.asarray(o.reset(), dtype=np.float64).ravel() for o in oracles])
    out = {"obs": [rel_err(vec._obs64.cpu().numpy(), obs_o, 1e-3).max()], "rew": [], "state": [], "done_mismatch": 0,
           "k_mismatch": 0, "term_mismatch": 0, "dones": 0, "k_sum": 0}
    for a in actions:
        at = torch.as_tensor(a, dtype=torch.float64, device=vec.device)
        _, _, done_g, term_g = vec.step_tensors(at)
        res = [o.step(a[i]) for i, o in enumerate(oracles)]
        obs_o = np.stack([np.asarray(r[0], dtype=np.float64).ravel() for r in res])
        rew_o = np.array([r[1] for r in res], dtype=np.float64)
        done_o = np.array([r[2] for r in res])
        k_o = np.array([o.attempts_last() for o in oracles])
        k_g = vec.last_attempts().cpu().numpy()
        done_gh = done_g.cpu().numpy().astype(bool)
        out["obs"].append(rel_err(vec._obs64.cpu().numpy(), obs_o, 1e-3).max())
        out["rew"].append(rel_err(vec._rew64.cpu().numpy(), rew_o, 1e-3).max())
        out["done_mismatch"] += int((done_gh != done_o).sum())
        # auto-reset zeroes last_attempts of finished envs on the GPU side
        live = ~done_o
        out["k_mismatch"] += int((k_g[live] != k_o[live]).sum())
        out["k_sum"] += int(k_o.sum())
        out["dones"] += int(done_o.sum())
        if check_state:
            so = np.stack([o.ode_state() for o in oracles])
            out["state"].append(rel_err(gpu_state(vec), so, 1e-3).max())
    return out
