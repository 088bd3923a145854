"""PCIe copy times for the e2e payloads (pinned host memory), GPU box."""
import torch, time
n = 65536
for name, nbytes in (("obs 14xf32", n * 56), ("rew", n * 4), ("done", n), ("term", n * 4), ("all packed", n * 65), ("actions", n * 12)):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for direction in ("d2h", "h2d"):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(5):
                (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
            s.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            for _ in range(50):
                (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
            e1.record(s)
            s.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 50
        print("%-12s %s %8d B: %.1f us  (%.1f GB/s)" % (name, direction, nbytes, us, nbytes / us / 1e3))
