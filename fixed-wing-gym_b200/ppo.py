"""PPO training loop with a torch policy on the GPU and the batched VecEnv (BASELINE configs[3]).

The reference trains with stable-baselines PPO2 over `VecNormalize(SubprocVecEnv([...]))`
(examples/train_rl_controller.py:223-231: MlpPolicy, PPO2 defaults).  This is the same loop with everything resident
on the device: observations never leave HBM, `DeviceVecNormalize` is the VecNormalize equivalent (running mean / var
of observations and discounted returns, clip 10, gamma 0.99), the policy is SB2's MlpPolicy shape (separate 2x64 tanh
networks for pi and vf, state-independent log-std), hyper-parameters are PPO2's defaults.  torch is used for the
policy math (library GEMMs) — the env side is the CUDA hot path of this repo; `train()` reports how the wall time
splits between the two.  Under torch.distributed (one process per GPU) gradients are averaged with an all-reduce
(a ~10 k-parameter MLP: negligible over NVLink) and every rank steps its own shard of the envs.
"""
import time

import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn


class RunningMeanStd:
    """Batched Welford update on the device (stable-baselines common/running_mean_std.py semantics)."""

    def __init__(self, shape, device):
        self.mean = torch.zeros(shape, dtype=torch.float64, device=device)
        self.var = torch.ones(shape, dtype=torch.float64, device=device)
        self.count = torch.full((), 1e-4, dtype=torch.float64, device=device)   # a tensor: no host state, capturable

    def update(self, x):
        """In place (static storage), so the same update can sit in a CUDA graph."""
        x = x.to(torch.float64)
        bm, bv, bc = x.mean(0), x.var(0, unbiased=False), float(x.shape[0])
        delta = bm - self.mean
        tot = self.count + bc
        m2 = self.var * self.count + bv * bc + delta * delta * self.count * bc / tot
        self.mean.add_(delta * bc / tot)
        self.var.copy_(m2 / tot)
        self.count.copy_(tot)


class DeviceVecNormalize:
    """VecNormalize(norm_obs=True, norm_reward=True, clip_obs=10, clip_reward=10, gamma=0.99) over device tensors."""

    def __init__(self, venv, gamma=0.99, clip_obs=10.0, clip_reward=10.0, epsilon=1e-8):
        self.venv, self.gamma, self.clip_obs, self.clip_reward, self.eps = venv, gamma, clip_obs, clip_reward, epsilon
        self.num_envs, self.device = venv.num_envs, venv.device
        self.obs_rms = RunningMeanStd((venv.obs_dim,), venv.device)
        self.ret_rms = RunningMeanStd((), venv.device)
        self.ret = torch.zeros(venv.num_envs, dtype=torch.float64, device=venv.device)
        self.training = True

    def _obs(self, obs):
        o = obs.reshape(self.num_envs, -1)
        if self.training:
            self.obs_rms.update(o)
        o = (o.to(torch.float64) - self.obs_rms.mean) / torch.sqrt(self.obs_rms.var + self.eps)
        return torch.clamp(o, -self.clip_obs, self.clip_obs).to(torch.float32)

    def reset(self):
        self.ret.zero_()
        return self._obs(self.venv.reset())

    def step(self, actions):
        obs, rew, done, term = self.venv.step_tensors(actions)
        self.ret.mul_(self.gamma).add_(rew.to(torch.float64))
        if self.training:
            self.ret_rms.update(self.ret)
        r = torch.clamp(rew.to(torch.float64) / torch.sqrt(self.ret_rms.var + self.eps), -self.clip_reward,
                        self.clip_reward).to(torch.float32)
        d = done.bool()
        self.ret.masked_fill_(d, 0.0)
        return self._obs(obs), r, d, rew


class ActorCritic(nn.Module):
    """SB2 MlpPolicy: pi and vf are separate 2 x 64 tanh MLPs; diagonal Gaussian with a state-independent log-std."""

    def __init__(self, obs_dim, act_dim, hidden=64):
        super().__init__()
        mlp = lambda out: nn.Sequential(nn.Linear(obs_dim, hidden), nn.Tanh(), nn.Linear(hidden, hidden), nn.Tanh(),
                                        nn.Linear(hidden, out))
        self.pi, self.vf = mlp(act_dim), mlp(1)
        self.log_std = nn.Parameter(torch.zeros(act_dim))
        for net, gain in ((self.pi, 0.01), (self.vf, 1.0)):
            for i, m in enumerate(net):
                if isinstance(m, nn.Linear):
                    nn.init.orthogonal_(m.weight, gain if i == len(net) - 1 else np.sqrt(2))
                    nn.init.zeros_(m.bias)

    def dist(self, obs):
        return torch.distributions.Normal(self.pi(obs), self.log_std.exp(), validate_args=False)   # no host sync

    def sample(self, obs):
        """-> (action, log-probability).  mu + std * randn instead of Normal.sample(): torch.normal(tensor, tensor)
        checks std >= 0 on the HOST (a device synchronisation, illegal while a CUDA graph is being captured)."""
        mu, std = self.pi(obs), self.log_std.exp()
        act = mu + std * torch.randn_like(mu)
        return act, torch.distributions.Normal(mu, std, validate_args=False).log_prob(act).sum(-1)

    def value(self, obs):
        return self.vf(obs).squeeze(-1)


def load_sb2_parameters(model, params):
    """Fill an ActorCritic from a stable-baselines 2 PPO2 parameter set (`parameters` inside the saved model zip,
    e.g. examples/models/mlp_controller/model.pkl; keys model/pi_fc0/w:0 ...).  `params`: mapping with the keys
    pi_fc0_w, pi_fc0_b, pi_fc1_*, pi_*, vf_fc0_*, vf_fc1_*, vf_*, pi_logstd (tf layout [in, out]), e.g. the committed
    tests/golden/mlp_controller.npz (oracle/make_golden_policy.py)."""
    def fill(lin, name):
        with torch.no_grad():
            lin.weight.copy_(torch.as_tensor(np.asarray(params[name + "_w"])).t())
            lin.bias.copy_(torch.as_tensor(np.asarray(params[name + "_b"])))
    fill(model.pi[0], "pi_fc0"); fill(model.pi[2], "pi_fc1"); fill(model.pi[4], "pi")
    fill(model.vf[0], "vf_fc0"); fill(model.vf[2], "vf_fc1"); fill(model.vf[4], "vf")
    with torch.no_grad():
        model.log_std.copy_(torch.as_tensor(np.asarray(params["pi_logstd"])).reshape(-1))
    return model


def compute_gae(rew, val, done, last_val, gamma, lam, adv_out):
    """Generalised advantage estimation over a rollout [T, N] (PPO2's runner: `done[t]` says the episode ended IN step
    t, so nothing is bootstrapped across it); writes adv_out in place, row by row."""
    T = rew.shape[0]
    gae = torch.zeros_like(last_val)
    for t in reversed(range(T)):
        nv = last_val if t == T - 1 else val[t + 1]
        nonterm = 1.0 - done[t]
        delta = rew[t] + gamma * nv * nonterm - val[t]
        gae = delta + gamma * lam * nonterm * gae
        adv_out[t].copy_(gae)
    return adv_out


class CurriculumCallback:
    """The curriculum rule of the reference's training callback (train_rl_controller.py:80-87): while the level is below
    1 and the cooldown has run out, a success rate above the current level raises the level to min(2 * success, 1) on
    every env and starts a cooldown of 15 callbacks.  `success` = fraction of the episodes that ended since the last
    callback whose info["success"]["all"] is true (the device's episode-metric sums)."""

    def __init__(self, venv, level=0.25, cooldown=25, cooldown_after_bump=15):
        self.venv, self.level, self.cooldown, self.after = venv, float(level), int(cooldown), int(cooldown_after_bump)
        self.bumps = []
        venv.env_method("set_curriculum_level", self.level)

    def __call__(self, rec):
        if self.level < 1 and rec.get("episodes", 0) > 0:
            if self.cooldown <= 0:
                if rec["success_rate"] > self.level:
                    self.level = min(rec["success_rate"] * 2, 1)
                    self.venv.env_method("set_curriculum_level", self.level)
                    self.cooldown = self.after
                    self.bumps.append((rec["iter"], self.level))
            else:
                self.cooldown -= 1
        rec["curriculum_level"] = self.level


def _allreduce_grads(model):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        w = dist.get_world_size()
        for p in model.parameters():
            if p.grad is not None:
                dist.all_reduce(p.grad)
                p.grad /= w


def train(venv, total_env_steps, n_steps=128, n_minibatches=4, n_epochs=4, lr=2.5e-4, gamma=0.99, lam=0.95,
          clip_range=0.2, ent_coef=0.01, vf_coef=0.5, max_grad_norm=0.5, seed=0, model=None, log=None, cuda_graph=True,
          callback=None):
    """PPO2-default training on a FixedWingVecEnv.  Returns (model, normalizer, stats); stats has env-steps/s inside
    training and the split of the device time (CUDA events): eager mode env / policy_forward / ppo_update, graph mode
    rollout / ppo_update (the rollout is one graph launch, so it has no inner split).
    cuda_graph: capture the rollout + GAE of an iteration once (at the second iteration; the first runs eagerly and
    warms every allocation up) and replay it afterwards.
    callback(rec): called after every iteration with its record (episodes that ended in it, their success rate / mean
    return / mean length from the device's metric sums, losses, timings), e.g. CurriculumCallback; it may change the env
    configuration (the rollout graph is re-captured when it does)."""
    dev = venv.device
    torch.manual_seed(seed)
    norm = DeviceVecNormalize(venv, gamma=gamma)
    n, od = venv.num_envs, venv.obs_dim
    model = model or ActorCritic(od, 3).to(dev)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    graph_update = bool(cuda_graph) and not distributed       # (the gradient all-reduce stays outside a captured graph)
    opt = torch.optim.Adam(model.parameters(), lr=lr, eps=1e-5, capturable=graph_update)
    obs = norm.reset().clone()          # static: the rollout reads and overwrites it in place
    B = n_steps * n
    buf = dict(obs=torch.zeros((n_steps, n, od), device=dev), act=torch.zeros((n_steps, n, 3), device=dev),
               logp=torch.zeros((n_steps, n), device=dev), val=torch.zeros((n_steps, n), device=dev),
               rew=torch.zeros((n_steps, n), device=dev), done=torch.zeros((n_steps, n), device=dev))
    adv = torch.zeros((n_steps, n), device=dev)
    ret = torch.zeros((n_steps, n), device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    act_low = torch.as_tensor(venv.action_space.low, dtype=torch.float32, device=dev)
    act_high = torch.as_tensor(venv.action_space.high, dtype=torch.float32, device=dev)

    def rollout(marks=None):
        """n_steps x (policy forward, sample, env step, normalise) + GAE into the static buffers."""
        with torch.no_grad():
            for t in range(n_steps):
                if marks is not None:
                    a0, a1, a2 = ev(), ev(), ev()
                    a0.record()
                act, logp = model.sample(obs)   # the buffer keeps the unclipped action and its log-probability
                buf["obs"][t].copy_(obs)
                buf["act"][t].copy_(act)
                buf["logp"][t].copy_(logp)
                buf["val"][t].copy_(model.value(obs))
                if marks is not None:
                    a1.record()
                # PPO2's runner clips to the action space before stepping (finite for actuator-limited spaces)
                o, rew, done, _ = norm.step(torch.max(torch.min(act, act_high), act_low))
                obs.copy_(o)
                buf["rew"][t].copy_(rew)
                buf["done"][t].copy_(done.float())
                if marks is not None:
                    a2.record()
                    marks.append((a0, a1, a2))
            compute_gae(buf["rew"], buf["val"], buf["done"], model.value(obs), gamma, lam, adv)
            torch.add(adv, buf["val"], out=ret)

    flat = {k: v.reshape((B,) + v.shape[2:]) for k, v in buf.items()}
    f_adv, f_ret = adv.reshape(B), ret.reshape(B)
    mb = B // n_minibatches
    mb_idx = torch.zeros(mb, dtype=torch.long, device=dev)      # static: the captured update reads its minibatch from here
    loss_out = torch.zeros(2, device=dev)
    upd_graph = None

    def minibatch_update():
        """One PPO2 gradient step on the minibatch mb_idx points at (clipped surrogate + clipped value loss + entropy)."""
        idx = mb_idx
        d = model.dist(flat["obs"][idx])
        logp = d.log_prob(flat["act"][idx]).sum(-1)
        a = f_adv[idx]
        a = (a - a.mean()) / (a.std() + 1e-8)
        ratio = (logp - flat["logp"][idx]).exp()
        pg = torch.max(-a * ratio, -a * torch.clamp(ratio, 1 - clip_range, 1 + clip_range)).mean()
        v = model.value(flat["obs"][idx])
        vclip = flat["val"][idx] + torch.clamp(v - flat["val"][idx], -clip_range, clip_range)
        vl = 0.5 * torch.max((v - f_ret[idx]) ** 2, (vclip - f_ret[idx]) ** 2).mean()
        loss = pg - ent_coef * d.entropy().sum(-1).mean() + vf_coef * vl
        opt.zero_grad(set_to_none=False)        # gradients keep their storage
        loss.backward()
        _allreduce_grads(model)
        nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
        opt.step()
        loss_out[0].copy_(loss.detach())
        loss_out[1].copy_(vl.detach())

    t_env = t_pol = t_upd = t_roll = 0.0
    iters = max(1, int(total_env_steps) // B)
    stats = {"iterations": iters, "batch": B, "history": [], "cuda_graph": bool(cuda_graph)}
    graph, graph_version = None, -1
    sums_prev = np.asarray(venv.metric_sums(), dtype=np.float64)
    torch.cuda.synchronize(dev)
    wall0 = time.perf_counter()
    for it in range(iters):
        e = [ev() for _ in range(3)]
        marks = []
        e[0].record()
        if not cuda_graph:
            rollout(marks)
        elif it == 0:
            rollout()                                   # eager: warms up allocations, cuBLAS handles, the env
        else:
            # seed / configuration are kernel PARAMETERS frozen into a captured graph: after venv.seed() or
            # set_curriculum_level() (train_rl_controller.py:80-87 changes the curriculum during training) re-capture
            if graph is not None and graph_version != venv.config_version:
                graph = None
            if graph is None:
                graph_version = venv.config_version
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):           # capture only; nothing runs here
                    rollout()
            graph.replay()
        e[1].record()
        for ep in range(n_epochs):
            perm = torch.randperm(B, device=dev)
            for k in range(n_minibatches):
                mb_idx.copy_(perm[k * mb:(k + 1) * mb])
                if not graph_update or it == 0:
                    minibatch_update()                  # eager (also the warm-up of the captured version)
                else:
                    if upd_graph is None:               # one gradient step = one graph: ~100 small kernels per replay
                        torch.cuda.synchronize(dev)
                        upd_graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(upd_graph):
                            minibatch_update()
                    upd_graph.replay()
        loss, vl = loss_out[0], loss_out[1]
        e[2].record()
        torch.cuda.synchronize(dev)
        roll_ms, upd_ms = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
        pol_ms = sum(a0.elapsed_time(a1) for a0, a1, _ in marks)
        env_ms = sum(a1.elapsed_time(a2) for _, a1, a2 in marks)
        t_env, t_pol, t_upd, t_roll = t_env + env_ms, t_pol + pol_ms, t_upd + upd_ms, t_roll + roll_ms
        sums = np.asarray(venv.metric_sums(), dtype=np.float64)     # episodes, successes, sum_return, sum_length, ...
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            from .parallel import allreduce_sum
            sums = allreduce_sum(sums, dev)
        d_ep = sums - sums_prev
        sums_prev = sums
        ep = max(d_ep[0], 1.0)
        rec = {"iter": it, "loss": float(loss.detach()), "mean_norm_reward": float(buf["rew"].mean()),
               "value_loss": float(vl.detach()), "rollout_ms": roll_ms, "env_ms": env_ms, "policy_ms": pol_ms,
               "update_ms": upd_ms, "episodes": int(d_ep[0]), "success_rate": float(d_ep[1] / ep),
               "mean_episode_return": float(d_ep[2] / ep), "mean_episode_length": float(d_ep[3] / ep),
               "failure_rate": float(d_ep[4] / ep)}
        if callback is not None:
            callback(rec)
        stats["history"].append(rec)
        if log:
            log(rec)
    torch.cuda.synchronize(dev)
    wall = time.perf_counter() - wall0
    if cuda_graph:
        tot = t_roll + t_upd
        frac = {"rollout": t_roll / tot, "ppo_update": t_upd / tot}
    else:
        tot = t_env + t_pol + t_upd
        frac = {"env": t_env / tot, "policy_forward": t_pol / tot, "ppo_update": t_upd / tot}
    stats.update(env_steps=iters * B, wall_s=wall, env_steps_per_s=iters * B / wall, time_fraction=frac,
                 episode_metrics=venv.metric_sums().tolist())
    return model, norm, stats
