"""Pure write / read / copy bandwidth for a buffer the size of cfg1's pooled
output (822 MB): the forward kernel is write-dominated, so its ceiling is the
write figure, not the copy figure in MEASURED_PEAKS.json."""
import json
import torch

n = 4096 * 256 * 14 * 14
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def t(fn, reps=10):
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


res = {}
ms = t(lambda: a.fill_(1.0)); res["fill_ms"] = ms; res["fill_GBps"] = n * 4 / ms / 1e6
ms = t(lambda: a.zero_()); res["memset_ms"] = ms; res["memset_GBps"] = n * 4 / ms / 1e6
ms = t(lambda: b.copy_(a)); res["copy_ms"] = ms; res["copy_GBps_rw"] = 2 * n * 4 / ms / 1e6
ms = t(lambda: a.sum()); res["read_sum_ms"] = ms; res["read_GBps"] = n * 4 / ms / 1e6
print(json.dumps(res))
