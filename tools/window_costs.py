#!/usr/bin/env python
"""Measured cost of every window of the config-5 sweep on its own (plan resident, tables kept):
the data the cost model of bhw_shard_range_cost is fitted to.  One JSON line per window."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import blackman_harris_win_b200 as bhw
import cases
torch.cuda.set_device(0)
out = torch.empty(1 << 26, dtype=torch.int32, device="cuda")
for v in range(1, 11):
    for pw in range(8, 27):
        d = bhw.variant_desc(v, pw, cases.VARIANT_DW[v])
        plan = bhw.Plan([d])
        for _ in range(3):
            plan.execute(out=out)
        torch.cuda.synchronize()
        reps = 20 if pw < 22 else 6
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            plan.execute(out=out)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / reps * 1e3
        print(json.dumps({"variant": v, "m": d.win_type, "pw": pw, "dw": d.dat_width, "us": round(us, 2),
                          "ps_per_sample": round(us * 1e6 / (1 << pw), 3)}))
        plan.destroy()
