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


_SCALED_SIZE = [15, 15, 15, 18, 21, 22, 23, 26, 30, 31, 32, 33, 38, 38, 38, 42, 42, 45, 47, 47, 47, 48, 48, 48, 48]


def alg_int_ops_per_sample(d):
    """Algorithmic integer work of the one-thread-per-sample formulation in 32-bit-op units, with
    the reference's own accounting of 3 additions + 2 shifts per CORDIC stage
    (src/cordic_dds.vhd:39-43); SURVEY.md 8(d):  (M-1)*[S*5*L + 7] + (M-1)*4*L' + (M-1) + 3."""
    m, dw = d.win_type, d.dat_width
    if d.model == bhw.MODEL_HLS:
        stages, width = dw, dw + 2
    elif d.sin_type == bhw.SIN_CORDIC:
        stages, width = dw - 1, dw + max(d.precision, 1)
    elif d.sin_type == bhw.SIN_CORDIC48:
        stages, width = dw, 48
    elif d.sin_type == bhw.SIN_CORDIC_SCALED:
        stages, width = dw, _SCALED_SIZE[dw - 8]
    else:                       # TAYLOR: ROM look-up + 1 narrow and 2 wide multiplies + ~12 (SURVEY 8d cfg 4)
        stages, width = 4, 32   # 4 * 5 = 20 ops per unit
    L = 1 if width <= 32 else 2
    Lp = 1 if 2 * dw <= 32 else (2 if dw <= 32 else 4)
    return (m - 1) * (stages * 5 * L + 7) + (m - 1) * 4 * Lp + (m - 1) + 3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--only", default="", help="substring filter on the config name")
    ap.add_argument("--no-batch", action="store_true", help="skip the plan/batch line of each config")
    ap.add_argument("--entities", action="store_true",
                    help="also time a 1 GiB bank of N = 65536 windows of every entity / sin-cos source at its BASELINE width")
    ap.add_argument("--atan2", action="store_true", help="also time the cordic_atan2 kernel (16M pairs)")
    ap.add_argument("--sweep", action="store_true", help="also run the config-5 sweep (10 variants x PHI_WIDTH 4..26)")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    int_peak = None
    try:
        int_peak = json.load(open(os.path.join(ROOT, "profiles", "int_peak.json")))["int32_mix_tops"] * 1e12
    except Exception:
        pass
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
            line = {"config": name, "algo": algo_name, "samples": n, "ms_per_window": round(ms, 5),
                    "gsamples_per_s": round(n / ms / 1e6, 3), "write_gbs": round(n * esz / ms / 1e6, 2),
                    "frac_of_hbm_peak": round(n * esz / ms / 1e6 / peak, 4),
                    "kernels_ms_per_call": {k: round(v[1] / v[0], 5) for k, v in kt.items()},
                    "launches_per_call": {k: v[0] / (reps + 3) for k, v in kt.items()}}
            if algo == bhw.ALGO_DIRECT and "k_direct_window" in kt and int_peak:
                # integer-ALU roofline of the direct kernel: algorithmic ops / kernel time vs the
                # measured alu+fma issue peak of this GPU (tools/int_peak.cu -> profiles/int_peak.json).
                # Timed on a sub-range (all but 16 samples): a whole-window request takes sample pairs
                # (n, n + N/2) from one evaluation, which is half the algorithmic work per sample.
                ops = alg_int_ops_per_sample(d)
                if n >= 64:
                    bhw.timing_enable(True)
                    bhw.timing_reset()
                    time_calls(lambda: bhw.generate(dd, 8, n - 16, out=out[:n - 16]), reps)
                    kt2 = {k: v for k, v in bhw.timing_read().items() if v[0]}
                    bhw.timing_enable(False)
                    k_ms = kt2["k_direct_window"][1] / kt2["k_direct_window"][0]
                    line["kernel_ms_sub_range (one evaluation per sample and harmonic)"] = round(k_ms, 5)
                else:
                    k_ms = kt["k_direct_window"][1] / kt["k_direct_window"][0]
                line["int_roofline"] = {"alg_ops_per_sample": ops, "achieved_tops": round(ops * n / k_ms / 1e9, 3),
                                        "peak_tops": round(int_peak / 1e12, 2),
                                        "frac": round(ops * n / (k_ms * 1e-3) / int_peak, 4),
                                        "peak_source": "measured, 1:1 SHF/IMAD mix on both pipes (profiles/int_peak.json)"}
            print(json.dumps(line))
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
    if args.entities:
        # one line per window entity and per sin/cos source (a7-a11): 4096 windows of
        # N = 65536, every window with its own AA0, plan resident, tables rebuilt in every step
        shapes = [("hamming_win DW16 cordic_dds", 1, 16, bhw.SIN_CORDIC), ("bh_win_3term DW16 cordic_dds", 3, 16, bhw.SIN_CORDIC),
                  ("bh_win_4term DW17 cordic_dds", 6, 17, bhw.SIN_CORDIC), ("bh_win_5term DW24 cordic_dds", 8, 24, bhw.SIN_CORDIC),
                  ("bh_win_7term DW32 cordic_dds", 10, 32, bhw.SIN_CORDIC), ("bh_win_7term DW24 cordic_dds", 10, 24, bhw.SIN_CORDIC),
                  ("bh_win_7term DW32 cordic_dds48", 10, 32, bhw.SIN_CORDIC48), ("bh_win_4term DW17 cordic_dds_scaled", 6, 17, bhw.SIN_CORDIC_SCALED),
                  ("hamming_win DW16 taylor", 1, 16, bhw.SIN_TAYLOR), ("bh_win_3term DW24 taylor", 3, 24, bhw.SIN_TAYLOR),
                  ("HLS model type 4 NW17", 6, 17, None)]
        nwin = 4096
        big = torch.empty(nwin << 16, dtype=torch.int32, device="cuda")
        bhw.set_table_cache(False)
        for name, v, dw, st in shapes:
            if st is None:
                base = bhw.variant_desc(v, 16, dw, model=bhw.MODEL_HLS)
            else:
                base = bhw.variant_desc(v, 16, dw, sin_type=st)
            descs = [base.copy(aa=[int(a) - (i % 1021) if k == 0 else int(a) for k, a in enumerate(base.aa)]) for i in range(nwin)]
            plan = bhw.Plan(descs)
            ms_plain = time_calls(lambda: plan.execute(out=big), 20)
            bhw.timing_enable(True)
            bhw.timing_reset()
            time_calls(lambda: plan.execute(out=big), 5)
            kt = {k: round(v_[1] / v_[0], 5) for k, v_ in bhw.timing_read().items() if v_[0]}
            bhw.timing_enable(False)
            print(json.dumps({"config": "bank 4096 x N=65536: " + name, "samples": plan.total, "ms_per_step": round(ms_plain, 5),
                              "gsamples_per_s": round(plan.total / ms_plain / 1e6, 1),
                              "frac_of_hbm_peak": round(plan.total * 4 / ms_plain / 1e6 / peak, 4), "kernel_ms": kt}))
            plan.destroy()
        bhw.set_table_cache(True)
        del big
    if args.atan2:
        n = 1 << 24
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.randint(-(1 << 23), 1 << 23, (n,), generator=g, dtype=torch.int32, device="cuda")
        y = torch.randint(-(1 << 23), 1 << 23, (n,), generator=g, dtype=torch.int32, device="cuda")
        out = torch.empty_like(x)
        for aw, iw, prec in ((16, 16, 1), (24, 24, 1), (32, 32, 1)):
            bhw.timing_enable(True)
            bhw.timing_reset()
            ms = time_calls(lambda: bhw.atan2(x, y, iw, aw, prec, out=out), 20)
            kt = {k: v for k, v in bhw.timing_read().items() if v[0]}
            bhw.timing_enable(False)
            k_ms = kt["k_atan2"][1] / kt["k_atan2"][0]
            ops = (aw - 1) * 5 * (1 if aw + prec <= 32 else 2) + 10   # 3 add + 2 shift per stage (src/cordic_atan2.vhd:24-28)
            line = {"config": f"cordic_atan2 ANGLE_WIDTH {aw} INPUT_WIDTH {iw} PRECISION {prec}, 16M pairs", "samples": n,
                    "ms_per_call": round(ms, 5), "kernel_ms": round(k_ms, 5), "gsamples_per_s": round(n / k_ms / 1e6, 2),
                    "hbm_gbs (8 B read + 4 B written per pair)": round(12 * n / k_ms / 1e6, 1),
                    "frac_of_hbm_peak": round(12 * n / k_ms / 1e6 / peak, 4)}
            if int_peak:
                line["int_roofline"] = {"alg_ops_per_sample": ops, "frac": round(ops * n / (k_ms * 1e-3) / int_peak, 4)}
            print(json.dumps(line))
    if args.sweep:
        # config 5: all 10 variants x PHI_WIDTH 4..26, one batch per element size, sharded 1-way here
        descs = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 27)]
        total = bhw.batch_total(descs)
        t0 = time.perf_counter()
        plan = bhw.Plan(descs)
        t_plan = time.perf_counter() - t0
        out = torch.empty(total, dtype=torch.int32, device="cuda")
        bhw.set_table_cache(False)
        for side in (0, 4):
            bhw.set_side_streams(side)
            ms_plain = time_calls(lambda: plan.execute(out=out), 10)   # without the per-launch timing events
            bhw.timing_enable(True)
            bhw.timing_reset()
            ms = time_calls(lambda: plan.execute(out=out), 5)
            kt = {k: v for k, v in bhw.timing_read().items() if v[0]}
            bhw.timing_enable(False)
            print(json.dumps({"config": "cfg5_sweep_10_variants_pw4_26",
                              "algo": "auto, tables rebuilt per step, %d side streams" % side,
                              "samples": total, "ms_per_step": round(ms_plain, 4),
                              "gsamples_per_s": round(total / ms_plain / 1e6, 3),
                              "frac_of_hbm_peak": round(total * 4 / ms_plain / 1e6 / peak, 4),
                              "plan_create_ms": round(1e3 * t_plan, 2),
                              "kernels_ms_per_step (sum of per-launch spans; they overlap with side streams)":
                                  {k: round(v[1] / 8, 4) for k, v in kt.items()},
                              "launches_per_step": {k: v[0] / 8 for k, v in kt.items()}}))
        bhw.set_side_streams(4)
        bhw.set_table_cache(True)
        ms_cached = time_calls(lambda: plan.execute(out=out), 10)      # the default: a plan keeps its tables
        print(json.dumps({"config": "cfg5_sweep_10_variants_pw4_26", "algo": "auto, tables kept by the plan (default), 4 side streams",
                          "samples": total, "ms_per_step": round(ms_cached, 4), "gsamples_per_s": round(total / ms_cached / 1e6, 3),
                          "frac_of_hbm_peak": round(total * 4 / ms_cached / 1e6 / peak, 4)}))
        plan.destroy()


if __name__ == "__main__":
    main()
