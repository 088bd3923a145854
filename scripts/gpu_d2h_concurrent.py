"""Concurrent device -> pinned-host copy bandwidth, one worker PROCESS per GPU (the bench's topology): 4.26 MB per copy
(one step's results of 65 536 envs), timed alone and with all GPUs copying at once, with the worker (a) unpinned,
(b) pinned to the cores of its GPU's NUMA node before the pinned buffer is allocated.  DESIGN.md 5."""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
NBYTES = 65536 * 65
REPS = 400


def worker(idx, n_gpus, bind, barrier, out):
    import torch
    torch.cuda.set_device(idx)
    info = None
    if bind:
        from fwgym_b200.parallel import bind_to_gpu_numa_node
        info = bind_to_gpu_numa_node(idx, n_gpus)
    dev = torch.empty(NBYTES, dtype=torch.uint8, device="cuda")
    host = torch.empty(NBYTES, dtype=torch.uint8).pin_memory()
    for _ in range(20):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    res = {}
    for phase in ("alone", "all"):
        if phase == "alone":
            for turn in range(n_gpus):
                barrier.wait()
                if turn == idx:
                    t0 = time.perf_counter()
                    for _ in range(REPS):
                        host.copy_(dev, non_blocking=True)
                        torch.cuda.synchronize()
                    res["alone"] = NBYTES * REPS / (time.perf_counter() - t0) / 1e9
        else:
            barrier.wait()
            t0 = time.perf_counter()
            for _ in range(REPS):
                host.copy_(dev, non_blocking=True)
                torch.cuda.synchronize()
            res["all"] = NBYTES * REPS / (time.perf_counter() - t0) / 1e9
    out.put((idx, res, info))


def main():
    import torch
    n = torch.cuda.device_count()
    ctx = mp.get_context("spawn")
    for bind in (False, True):
        barrier = ctx.Barrier(n)
        out = ctx.Queue()
        ps = [ctx.Process(target=worker, args=(i, n, bind, barrier, out)) for i in range(n)]
        for p in ps:
            p.start()
        got = sorted(out.get() for _ in ps)
        for p in ps:
            p.join()
        print("bind_to_numa=%s" % bind)
        for idx, res, info in got:
            print("  gpu %d: alone %.1f GB/s, all %d concurrently %.1f GB/s  %s" % (idx, res["alone"], n, res["all"], info or ""))


if __name__ == "__main__":
    main()
