"""Measures pinned host<->device copy bandwidth (one direction and both at once)
on the box: the floor of bench.py's e2e number, which ships ~1 GB each way."""
import json
import time
import torch

n = 512 << 20
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name in ("h2d", "d2h", "both"):
    best = 1e9
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if name in ("h2d", "both"):
            with torch.cuda.stream(s1):
                d_a.copy_(h_in, non_blocking=True)
        if name in ("d2h", "both"):
            with torch.cuda.stream(s2):
                h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    res[name + "_GBps_per_dir"] = n / best / 1e9
t0 = time.perf_counter()
x = torch.empty(n, dtype=torch.uint8, pin_memory=True)
res["pin_alloc_512MiB_ms"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(res))
