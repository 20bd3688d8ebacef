#!/usr/bin/env python
"""Small, fast exercise of every kernel family against the oracle (also the input for
compute-sanitizer where the pool allows it - it is closed on this one):
  compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402
import cases  # noqa: E402
import harness as H  # noqa: E402

torch.cuda.set_device(0)
n_ok = 0
# one-shot windows: direct32, direct taylor, generic direct (int64), table path with staged / global tables
for d in [bhw.variant_desc(1, 10, 16), bhw.variant_desc(6, 12, 17, algo=bhw.ALGO_DIRECT),
          bhw.variant_desc(3, 14, 24, sin_type=bhw.SIN_TAYLOR), bhw.variant_desc(10, 9, 40),
          bhw.variant_desc(10, 12, 32, sin_type=bhw.SIN_CORDIC48, algo=bhw.ALGO_TABLE),
          bhw.variant_desc(8, 13, 24, sin_type=bhw.SIN_CORDIC_SCALED, algo=bhw.ALGO_TABLE),
          bhw.variant_desc(2, 18, 16, algo=bhw.ALGO_TABLE), bhw.variant_desc(6, 18, 17, algo=bhw.ALGO_TABLE),
          bhw.variant_desc(6, 11, 17, model=bhw.MODEL_HLS, algo=bhw.ALGO_TABLE)]:
    got = bhw.generate(d).cpu().numpy().astype(np.int64)
    assert np.array_equal(got, H.orc_window(d, threads=4)), d
    n_ok += 1
# a mixed plan (bank runs + general kernel + side streams), ragged range, stream offset
descs = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v], stream_offset=pw & 1) for v in (1, 3, 6, 8, 10) for pw in (4, 7, 9, 12)]
plan = bhw.Plan(descs)
total = plan.total
out = plan.execute(5, total - 9).cpu().numpy().astype(np.int64)
assert np.array_equal(out, H.orc_batch(descs, 5, total - 9))
plan.destroy()
# the group kernel in its three table placements (mixed PHI_WIDTHs, a cut window), the apply step
for variants, dw, pws in (((1, 3), 16, (9, 13, 17)), ((6,), 17, (10, 17, 18)), ((9, 10), 24, (9, 14, 18))):
    gd = [bhw.variant_desc(v, pw, dw, stream_offset=pw & 1) for pw in pws for v in variants]
    gt = bhw.batch_total(gd)
    want = H.orc_batch(gd, 0, gt)
    assert np.array_equal(bhw.generate_batch(gd).cpu().numpy().astype(np.int64), want)
    assert np.array_equal(bhw.generate_batch(gd, 777, gt - 5000).cpu().numpy().astype(np.int64), want[777:gt - 4223])
    n_ok += 2
xa = torch.randint(-(1 << 15), 1 << 15, (3, 1 << 12), dtype=torch.int32)
da = bhw.variant_desc(6, 12, 17)
assert np.array_equal(bhw.apply(da, xa.cuda(), bhw.APPLY_EXACT).cpu().numpy(), H.orc_apply(da, xa.numpy(), 0))
assert np.array_equal(bhw.apply(da, xa.cuda(), bhw.APPLY_ROUNDED).cpu().numpy().astype(np.int64), H.orc_apply(da, xa.numpy(), 1))
# host entry points, sincos, atan2
assert np.array_equal(bhw.generate_batch_host(descs).astype(np.int64), H.orc_batch(descs, 0, total))
d = bhw.make_desc(2, 12, 20, sin_type=bhw.SIN_CORDIC48)
s, c = bhw.sincos(d)
os_, oc = H.orc_sincos(d)
assert np.array_equal(s.cpu().numpy().astype(np.int64), os_) and np.array_equal(c.cpu().numpy().astype(np.int64), oc)
x = torch.randint(-(1 << 23), 1 << 23, (4099,), dtype=torch.int32)
y = torch.randint(-(1 << 23), 1 << 23, (4099,), dtype=torch.int32)
assert np.array_equal(bhw.atan2(x.cuda(), y.cuda(), 24, 24, 1).cpu().numpy(), H.orc_atan2(24, 24, 1, x.numpy(), y.numpy()))
assert np.array_equal(bhw.atan2(x.cuda(), y.cuda(), 32, 32, 1).cpu().numpy(), H.orc_atan2(32, 32, 1, x.numpy(), y.numpy()))
bhw.cache_clear()
print(f"sanitize_smoke ok ({n_ok} windows + plan + host + sincos + atan2)")
