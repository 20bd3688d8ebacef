#!/usr/bin/env python
"""Single long windows (2^22 .. 2^26 points) over tables too large for shared memory, through a resident
plan; includes ~15 us of Python/ctypes call overhead per execute.  One JSON object."""
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, blackman_harris_win_b200 as bhw, cases
def t(fn, reps=30):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
out = torch.empty(1 << 26, dtype=torch.int32, device="cuda")
shapes = [(10, pw, 32, bhw.SIN_CORDIC) for pw in (23, 24, 25, 26)] + [(9, pw, 24, bhw.SIN_CORDIC) for pw in (22, 23, 24)] + \
         [(6, pw, 24, bhw.SIN_CORDIC) for pw in (23, 24)] + [(10, 24, 32, bhw.SIN_CORDIC48)]
res = []
for v, pw, dw, st in shapes:
    d = bhw.variant_desc(v, pw, dw, sin_type=st)
    plan = bhw.Plan([d])
    ref = None
    us = t(lambda: plan.execute(out=out))
    chk = int(out[: 1 << pw].to(torch.int64).sum().item())
    res.append({"v": v, "m": d.win_type, "pw": pw, "dw": dw, "st": st, "us": round(us, 1), "frac_hbm": round((4 << pw) / us / 1e3 / 6554.6, 3), "sum": chk})
    plan.destroy()
print(json.dumps({"res": res}))
