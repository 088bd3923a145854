"""World-size-2 tests of the multi-GPU plumbing on CPU (gloo): sharding arithmetic, the metric-sum all-reduce, and
sharding invariance of the per-env random streams (checked with the CPU oracle: the same global env ids produce the
same trajectories whether one rank owns them all or two ranks own half each)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import harness


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _trajectory(env_ids, seed, steps, actions):
    out = []
    for e in env_ids:
        run = harness.OracleRunner(harness.make_env("restated", harness.config_path(), None,
                                                    {"turbulence": True, "turbulence_intensity": "light"}), seed, e)
        obs = [np.asarray(run.reset(), dtype=np.float64).ravel()]
        for t in range(steps):
            obs.append(np.asarray(run.step(actions[t, e])[0], dtype=np.float64).ravel())
        out.append(np.stack(obs))
    return np.stack(out)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fwgym_b200 import parallel
    off, n = parallel.shard(5, rank, world)
    assert (off, n) == ((0, 3) if rank == 0 else (3, 2))
    # metric sums: rank r contributes r+1 episodes with return 10*(r+1) each
    local = np.array([rank + 1, rank, 10.0 * (rank + 1) ** 2, 7.0 * (rank + 1), 0, 0, 0, 3], dtype=np.float64)
    g = parallel.allreduce_metric_sums(local)
    assert g["episodes"] == 3 and g["successes"] == 1 and abs(g["mean_return"] - 50.0 / 3) < 1e-12
    # sharded trajectories of global envs [0, 4): this rank's block
    off, n = parallel.shard(4, rank, world)
    acts = np.random.RandomState(0).uniform(-1, 1, (3, 4, 3))
    traj = _trajectory(range(off, off + n), 77, 3, acts)
    gathered = [None] * world
    dist.all_gather_object(gathered, traj)
    if rank == 0:
        q.put(np.concatenate(gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sharded = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    acts = np.random.RandomState(0).uniform(-1, 1, (3, 4, 3))
    whole = _trajectory(range(4), 77, 3, acts)
    assert np.array_equal(whole, sharded)   # bitwise: the global env id keys every stream


def test_shard_covers_everything():
    from fwgym_b200 import parallel
    for total, world in ((65536 * 8, 8), (10, 3), (7, 8)):
        blocks = [parallel.shard(total, r, world) for r in range(world)]
        assert sum(n for _, n in blocks) == total
        assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
