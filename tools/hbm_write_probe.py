#!/usr/bin/env python
"""Write-only HBM bandwidth of this GPU next to the copy figure in MEASURED_PEAKS.json: the window
generator is a pure store stream, so a fill of the bench bank's size is its natural ceiling.
Prints one JSON line (GB/s; best and median of 20 fills of 1 GiB, CUDA events)."""
import json
import statistics

import torch

torch.cuda.set_device(0)
n = 1 << 28
buf = torch.empty(n, dtype=torch.int32, device="cuda")
src = torch.empty(n, dtype=torch.int32, device="cuda")
res = {}
for name, fn in (("fill_gbs", lambda: buf.fill_(7)), ("zero_gbs", lambda: buf.zero_()), ("copy_gbs_rw", lambda: buf.copy_(src))):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    mult = 2 if name.startswith("copy") else 1
    res[name] = {"best": round(mult * n * 4 / min(ts) / 1e6, 1), "median": round(mult * n * 4 / statistics.median(ts) / 1e6, 1)}
print(json.dumps(res))
