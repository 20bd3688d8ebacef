#!/usr/bin/env python
"""One window through a resident plan, a few executes: the target for `ncu -k regex:k_synth_bank -s 2 -c 1`.
  python tools/one_window.py <variant 1..10> <PHI_WIDTH> <DAT_WIDTH> [sin_type]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402

v, pw, dw = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
st = int(sys.argv[4]) if len(sys.argv) > 4 else bhw.SIN_CORDIC
plan = bhw.Plan([bhw.variant_desc(v, pw, dw, sin_type=st)])
out = torch.empty(1 << pw, dtype=torch.int32, device="cuda")
for _ in range(4):
    plan.execute(out=out)
torch.cuda.synchronize()
plan.destroy()
