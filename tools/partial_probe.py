#!/usr/bin/env python
"""A 2^26-point window generated whole and as 8 equal ranges (what 8 ranks of a sample-range shard each do):
time of one range vs 1/8 of the whole.  One JSON line per variant."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import blackman_harris_win_b200 as bhw
import cases
torch.cuda.set_device(0)
out = torch.empty(1 << 26, dtype=torch.int32, device="cuda")
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for v in (1, 3, 6, 8, 10):
    d = bhw.variant_desc(v, 26, cases.VARIANT_DW[v])
    plan = bhw.Plan([d])
    n = 1 << 26
    whole = t(lambda: plan.execute(out=out))
    part = [t(lambda r=r: plan.execute(r * (n // 8) + 100, n // 8 - 200, out=out)) for r in (0, 3, 7)]
    print(json.dumps({"variant": v, "m": d.win_type, "dw": d.dat_width, "whole_us": round(whole, 1),
                      "eighth_us (ranges 0, 3, 7; ragged by 100 samples)": [round(x, 1) for x in part]}))
    plan.destroy()
