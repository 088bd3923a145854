"""Multi-GPU plumbing: the env batch shards trivially — one process per GPU, rank r owns the contiguous
block of GLOBAL env ids [r*n, (r+1)*n).  There is NO collective on the step path; torch.distributed (NCCL over
NVLink on the GPU box, gloo in the CPU tests) only all-reduces small vectors at logging boundaries: the episode-metric
sums of `FixedWingVecEnv.metric_sums()` (the reference aggregates the same quantities per process through
Monitor / TensorBoard callbacks, train_rl_controller.py:41-74) and timing maxima.

RNG streams are keyed by the global env id (csrc/philox.cuh), so a job sharded over any number of ranks produces the
same per-env trajectories as a single-GPU job of the same total size.
"""
import numpy as np
import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard(total_envs, rank=None, world=None):
    """-> (env_offset, n_local) of this rank for `total_envs` global envs, contiguous blocks, remainder to the
    lowest ranks."""
    if rank is None or world is None:
        rank, world = rank_world()
    base, rem = divmod(int(total_envs), int(world))
    n = base + (1 if rank < rem else 0)
    off = rank * base + min(rank, rem)
    return off, n


def make_sharded_vec_env(config_path, total_envs, device=None, **kw):
    """FixedWingVecEnv over this rank's block of the global batch (env_offset set so RNG streams are global)."""
    from .vec_env import FixedWingVecEnv
    rank, world = rank_world()
    off, n = shard(total_envs, rank, world)
    if device is None:
        device = "cuda:%d" % (rank % max(1, torch.cuda.device_count()))
    return FixedWingVecEnv(config_path, n, device=device, env_offset=off, **kw)


def allreduce_sum(vec, device=None):
    """Sum a small float64 vector over all ranks (no-op without an initialised process group) -> numpy."""
    t = torch.as_tensor(np.asarray(vec, dtype=np.float64))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def allreduce_metric_sums(env_or_sums, device=None):
    """Global episode metrics from the per-rank sums: dict with episodes, success_rate, mean_return, mean_length,
    failure_rate, goal_step_fraction denominators included."""
    sums = env_or_sums.metric_sums() if hasattr(env_or_sums, "metric_sums") else env_or_sums
    g = allreduce_sum(sums, device)
    names = ("episodes", "successes", "sum_return", "sum_length", "failures", "steps_term", "success_term", "goal_steps")
    out = dict(zip(names, g.tolist()))
    ep = max(out["episodes"], 1.0)
    out.update(success_rate=out["successes"] / ep, mean_return=out["sum_return"] / ep,
               mean_length=out["sum_length"] / ep, failure_rate=out["failures"] / ep)
    return out


# ---- host placement: the host-buffer pipeline (HostStepper) is PCIe + host-DRAM traffic ------------------------------
def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(index):
    """NUMA node the GPU's PCIe root hangs off (sysfs), or -1 when the platform does not say."""
    try:
        p = torch.cuda.get_device_properties(index)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            return int(f.read().strip())
    except Exception:
        return -1


def bind_to_gpu_numa_node(local_rank, n_local_ranks):
    """Pin the calling process to host cores on the NUMA node of GPU `local_rank`, sharing that node's cores with the
    other local ranks whose GPU sits on the same node.  Call it BEFORE allocating pinned host memory (cudaHostAlloc
    places pages on the calling thread's node): results then travel GPU -> PCIe root -> local DRAM instead of crossing
    the socket interconnect.  Returns a dict describing what was done (for the bench line); a no-op where sysfs has no
    topology."""
    import os
    info = {"numa_node": -1, "cores": None}
    if not hasattr(os, "sched_setaffinity"):
        return info
    allowed = sorted(os.sched_getaffinity(0))
    nodes = [gpu_numa_node(i) for i in range(n_local_ranks)]
    mine = nodes[local_rank]
    info["numa_node"] = mine
    pool = allowed
    if mine >= 0:
        try:
            with open("/sys/devices/system/node/node%d/cpulist" % mine) as f:
                local = [c for c in _parse_cpulist(f.read()) if c in set(allowed)]
            if local:
                pool = local
        except OSError:
            pass
    peers = [r for r in range(n_local_ranks) if nodes[r] == mine] if pool is not allowed else list(range(n_local_ranks))
    per = max(1, len(pool) // max(1, len(peers)))
    k = peers.index(local_rank)
    cores = pool[k * per:(k + 1) * per] or pool
    try:
        os.sched_setaffinity(0, cores)
        info["cores"] = [cores[0], cores[-1]]
    except OSError:
        pass
    return info
