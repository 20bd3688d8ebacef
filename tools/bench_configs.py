#!/usr/bin/env python
"""Per-config timing of the BASELINE.json configurations (single windows and small batches), both
evaluation strategies.  Not the contract bench (that is bench.py) - this fills the table in
DESIGN.md / profiles/.  Prints one JSON object per line.

  python tools/bench_configs.py [--reps 50] [--sweep]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402
import cases  # noqa: E402


def time_calls(fn, reps, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps  # ms per call


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--only", default="", help="substring filter on the config name")
    ap.add_argument("--no-batch", action="store_true", help="skip the plan/batch line of each config")
    ap.add_argument("--sweep", action="store_true", help="also run the config-5 sweep (10 variants x PHI_WIDTH 4..26)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    configs = dict(cases.baseline_configs())
    # throughput shapes for the register-resident direct kernel (not BASELINE configs)
    configs["x_bh4_n16m_dw17 (direct32 throughput)"] = bhw.make_desc(4, 24, 17, [47022, 64001, 18518, 1531])
    configs["x_hamming_n16m_dw16 (direct32 throughput)"] = bhw.make_desc(2, 24, 16, [17808, 14959])
    for name, d in configs.items():
        if args.only and args.only not in name:
            continue
        n = 1 << d.phi_width
        esz = bhw.elem_bytes(d)
        out = torch.empty(n, dtype=torch.int64 if esz == 8 else torch.int32, device="cuda")
        for algo_name, algo in (("auto", bhw.ALGO_AUTO), ("direct", bhw.ALGO_DIRECT)):
            dd = d.copy(algo=algo)
            reps = args.reps if n <= (1 << 22) or algo == bhw.ALGO_AUTO else 5
            bhw.timing_enable(True)
            bhw.timing_reset()
            ms = time_calls(lambda: bhw.generate(dd, out=out), reps)
            kt = {k: v for k, v in bhw.timing_read().items() if v[0]}
            bhw.timing_enable(False)
            print(json.dumps({"config": name, "algo": algo_name, "samples": n, "ms_per_window": round(ms, 5),
                              "gsamples_per_s": round(n / ms / 1e6, 3), "write_gbs": round(n * esz / ms / 1e6, 2),
                              "frac_of_hbm_peak": round(n * esz / ms / 1e6 / peak, 4),
                              "kernels_ms_per_call": {k: round(v[1] / v[0], 5) for k, v in kt.items()},
                              "launches_per_call": {k: v[0] / (reps + 3) for k, v in kt.items()}}))
        # batch of identical-shape windows through a plan (device-resident), 256 MB per step
        if esz == 4 and n <= (1 << 22) and not args.no_batch:
            nwin = max(1, (1 << 26) // n)
            descs = [d.copy(aa=[int(a) - (i % 7) if k == 0 else int(a) for k, a in enumerate(d.aa)]) for i in range(nwin)]
            plan = bhw.Plan(descs)
            big = torch.empty(plan.total, dtype=torch.int32, device="cuda")
            bhw.set_table_cache(False)
            ms = time_calls(lambda: plan.execute(out=big), 20)
            bhw.set_table_cache(True)
            print(json.dumps({"config": name, "algo": "auto, plan of %d windows, tables rebuilt per step" % nwin,
                              "samples": plan.total, "ms_per_step": round(ms, 5),
                              "gsamples_per_s": round(plan.total / ms / 1e6, 3),
                              "frac_of_hbm_peak": round(plan.total * 4 / ms / 1e6 / peak, 4)}))
            plan.destroy()
            del big
    if args.sweep:
        # config 5: all 10 variants x PHI_WIDTH 4..26, one batch per element size, sharded 1-way here
        descs = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 27)]
        total = bhw.batch_total(descs)
        t0 = time.perf_counter()
        plan = bhw.Plan(descs)
        t_plan = time.perf_counter() - t0
        out = torch.empty(total, dtype=torch.int32, device="cuda")
        bhw.set_table_cache(False)
        bhw.timing_enable(True)
        bhw.timing_reset()
        ms = time_calls(lambda: plan.execute(out=out), 5)
        kt = {k: v for k, v in bhw.timing_read().items() if v[0]}
        bhw.timing_enable(False)
        bhw.set_table_cache(True)
        print(json.dumps({"config": "cfg5_sweep_10_variants_pw4_26", "algo": "auto, tables rebuilt per step",
                          "samples": total, "ms_per_step": round(ms, 4), "gsamples_per_s": round(total / ms / 1e6, 3),
                          "frac_of_hbm_peak": round(total * 4 / ms / 1e6 / peak, 4), "plan_create_ms": round(1e3 * t_plan, 2),
                          "kernels_ms_per_step": {k: round(v[1] / 8, 4) for k, v in kt.items()},
                          "launches_per_step": {k: v[0] / 8 for k, v in kt.items()}}))
        plan.destroy()


if __name__ == "__main__":
    main()
