#!/usr/bin/env python
"""Small fixed workloads for `ncu -k regex:<kernel> -s <skip> -c 1` captures (round 2).
  python tools/ncu_targets.py <target>
targets: sweep | win <variant> <pw> [dw] [sin_type] | bank7 | bank7_dds48 | cfg4 | direct <variant> <pw> <dw> | apply <variant> <pw>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402
import bench  # noqa: E402

DW = bench.VARIANT_DW
t = sys.argv[1]
reps = 3
if t == "sweep":
    descs = bench.sweep_descs()
    out = torch.empty(bhw.batch_total(descs), dtype=torch.int32, device="cuda")
    plan = bhw.Plan(descs)
    bhw.set_table_cache(False)
    bhw.set_side_streams(0)
    for _ in range(reps):
        plan.execute(out=out)
elif t == "win":
    v, pw = int(sys.argv[2]), int(sys.argv[3])
    dw = int(sys.argv[4]) if len(sys.argv) > 4 else DW[v]
    st = int(sys.argv[5]) if len(sys.argv) > 5 else bhw.SIN_CORDIC
    plan = bhw.Plan([bhw.variant_desc(v, pw, dw, sin_type=st)])
    out = torch.empty(1 << pw, dtype=torch.int32, device="cuda")
    bhw.set_table_cache(False)
    for _ in range(reps):
        plan.execute(out=out)
elif t in ("bank7", "bank7_dds48", "bank7_dw24"):
    st = bhw.SIN_CORDIC48 if t == "bank7_dds48" else bhw.SIN_CORDIC
    dw = 24 if t == "bank7_dw24" else 32
    pw, nwin = (16, 4096) if t == "bank7_dw24" else (20, 64)
    base = bhw.variant_desc(10, pw, dw, sin_type=st)
    descs = [base.copy(aa=[int(a) - (i % 7) if k == 0 else int(a) for k, a in enumerate(base.aa)]) for i in range(nwin)]
    plan = bhw.Plan(descs)
    out = torch.empty(plan.total, dtype=torch.int32, device="cuda")
    bhw.set_table_cache(False)
    for _ in range(reps):
        plan.execute(out=out)
elif t == "cfg4":
    d = bhw.make_desc(3, 24, 24, [7046424, 8388600, 1342176], sin_type=bhw.SIN_TAYLOR, lut_size=9)
    out = torch.empty(1 << 24, dtype=torch.int32, device="cuda")
    for _ in range(reps):
        bhw.generate(d, out=out)
elif t == "direct":
    v, pw, dw = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    d = bhw.variant_desc(v, pw, dw, algo=bhw.ALGO_DIRECT)
    out = torch.empty(1 << pw, dtype=torch.int64 if dw > 32 else torch.int32, device="cuda")
    for _ in range(reps):
        bhw.generate(d, out=out)
elif t == "apply":
    v, pw = int(sys.argv[2]), int(sys.argv[3])
    d = bhw.variant_desc(v, pw, DW[v])
    x = torch.randint(-(1 << 15), 1 << 15, (4, 1 << pw), dtype=torch.int32, device="cuda")
    for _ in range(reps):
        bhw.apply(d, x, bhw.APPLY_EXACT)
torch.cuda.synchronize()
