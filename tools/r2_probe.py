#!/usr/bin/env python
"""Round-2 probe: where the config-5 sweep spends its time.  For one representative variant of
every window class (terms x DAT_WIDTH) and PHI_WIDTH 20..26: a single window through a resident plan
(tables kept / rebuilt, kernel-only device times from the library's events) and a 1 GiB bank of the
same shape (steady state, no ramp).  One JSON object per line.

  python tools/r2_probe.py [--pws 20,22,24,25,26] [--variants 1,3,6,9,10] [--sweep]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402
import cases  # noqa: E402

PEAK = 6554.6


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def kernel_times(fn, reps=4):
    bhw.timing_enable(True)
    bhw.timing_reset()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    kt = {k: [v[0] // reps, round(v[1] / reps * 1e3, 2)] for k, v in bhw.timing_read().items() if v[0]}
    bhw.timing_enable(False)
    return kt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pws", default="20,22,24,25,26")
    ap.add_argument("--variants", default="1,3,6,9,10")
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--no-bank", action="store_true")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    big = torch.empty(1 << 28, dtype=torch.int32, device="cuda")
    for v in [int(x) for x in args.variants.split(",") if x]:
        for pw in [int(x) for x in args.pws.split(",") if x]:
            d = bhw.variant_desc(v, pw, cases.VARIANT_DW[v])
            n = 1 << pw
            plan = bhw.Plan([d])
            bhw.set_table_cache(True)
            us_kept = timed(lambda: plan.execute(out=big[:n]), 20)
            kt_kept = kernel_times(lambda: plan.execute(out=big[:n]))
            bhw.set_table_cache(False)
            us_reb = timed(lambda: plan.execute(out=big[:n]), 20)
            kt_reb = kernel_times(lambda: plan.execute(out=big[:n]))
            bhw.set_table_cache(True)
            plan.destroy()
            line = {"what": "single", "variant": v, "m": d.win_type, "dw": d.dat_width, "pw": pw,
                    "us_tables_kept": round(us_kept, 2), "us_tables_rebuilt": round(us_reb, 2),
                    "roof_us": round(4 * n / PEAK / 1e3, 2),
                    "frac_kept": round(4 * n / PEAK / 1e3 / us_kept, 3),
                    "kernels_kept": kt_kept, "kernels_rebuilt": kt_reb}
            print(json.dumps(line), flush=True)
            if not args.no_bank:
                nwin = (1 << 28) >> pw
                descs = [d.copy(aa=[int(a) - (i % 7) if k == 0 else int(a) for k, a in enumerate(d.aa)]) for i in range(nwin)]
                plan = bhw.Plan(descs)
                us = timed(lambda: plan.execute(out=big), 10)
                kt = kernel_times(lambda: plan.execute(out=big))
                plan.destroy()
                print(json.dumps({"what": "bank_1GiB", "variant": v, "m": d.win_type, "dw": d.dat_width, "pw": pw,
                                  "nwin": nwin, "us": round(us, 2), "frac": round((4 << 28) / PEAK / 1e3 / us, 3),
                                  "kernels": kt}), flush=True)
    if args.sweep:
        descs = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 27)]
        total = bhw.batch_total(descs)
        del big
        out = torch.empty(total, dtype=torch.int32, device="cuda")
        plan = bhw.Plan(descs)
        for cache in (True, False):
            bhw.set_table_cache(cache)
            for side in (0, 4):
                bhw.set_side_streams(side)
                us = timed(lambda: plan.execute(out=out), 10)
                kt = kernel_times(lambda: plan.execute(out=out))
                print(json.dumps({"what": "sweep", "tables_kept": cache, "side_streams": side, "us": round(us, 1),
                                  "frac": round(4 * total / PEAK / 1e3 / us, 3), "kernels": kt}), flush=True)
        bhw.set_side_streams(4)
        bhw.set_table_cache(True)
        plan.destroy()


if __name__ == "__main__":
    main()
