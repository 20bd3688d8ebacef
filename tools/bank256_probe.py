#!/usr/bin/env python
"""256 MB banks of same-shape windows whose trig table stays in L2 (TAB_GLOBAL): BASELINE config 3 as a bank
of 64, and neighbours.  One JSON line per shape (device time per execute, plan resident, tables kept)."""
import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, blackman_harris_win_b200 as bhw, cases
cfg = cases.baseline_configs()
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
shapes = [("cfg3 dds48 x64", cfg["cfg3_bh7_n1m_dw32_dds48"], 64), ("cfg3 dds x64", cfg["cfg3_bh7_n1m_dw32_dds"], 64),
          ("bh5 dw24 pw20 x64", bhw.variant_desc(9, 20, 24), 64), ("bh4 dw17 pw20 x64", bhw.variant_desc(6, 20, 17), 64),
          ("bh7 dw32 pw18 x256", bhw.variant_desc(10, 18, 32), 256), ("bh7 dw32 pw22 x16", bhw.variant_desc(10, 22, 32), 16),
          ("bh7 dw32 pw24 x4", bhw.variant_desc(10, 24, 32), 4), ("hamming dw24 pw22 x16", bhw.variant_desc(1, 22, 24), 16)]
for name, d, nwin in shapes:
    descs = [d.copy(aa=[int(a) - (i % 7) if k == 0 else int(a) for k, a in enumerate(d.aa)]) for i in range(nwin)]
    plan = bhw.Plan(descs)
    out = torch.empty(plan.total, dtype=torch.int32, device="cuda")
    us = t(lambda: plan.execute(out=out))
    print(json.dumps({"shape": name, "us": round(us, 1), "gsamples_per_s": round(plan.total / us / 1e3, 1),
                      "frac_hbm": round(plan.total * 4 / us / 1e3 / 6554.6, 3)}))
    plan.destroy()
