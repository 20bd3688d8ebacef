#!/usr/bin/env python
"""CPU baselines of the BASELINE.json configurations on this box's host cores ("CPU
baseline, same run").  One independent window (or slice of one) per core, one process per core -
the HLS cordic() rewrites a static table on every call (hls/windows/win_function.cpp:74-80), so
it cannot share a process.  Only the generate loop is timed.

  kind "reference" : the reference's own unmodified C++ (oracle/_ref/*.so, built by
                     oracle/build_ref.sh) - exists for the HLS window model and the cpp CORDIC
  kind "port"      : oracle/bhw_oracle.c, the CPU restatement of the RTL entities ("restatement,
                     not reference": the RTL has no software model, TAYLOR has none at all)

This is the checker/baseline leg (it executes oracle/); nothing here is on the product path.
Prints one JSON object per line.

  python tools/cpu_configs.py [--budget 4.0]
"""
import argparse
import ctypes as C
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(args):
    kind, so, what, n0, count, reps = args
    import numpy as np
    out = np.empty(count, dtype=np.int64)
    p = out.ctypes.data_as(C.POINTER(C.c_longlong))
    L = C.CDLL(so)
    if kind == "reference":
        L.ref_hls_window.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.POINTER(C.c_longlong)]
        t0 = time.perf_counter()
        for _ in range(reps):
            L.ref_hls_window(what, n0, count, p)
        return time.perf_counter() - t0
    from blackman_harris_win_b200.api import BhwDesc
    d = BhwDesc.from_buffer_copy(what)
    L.orc_window.argtypes = [C.POINTER(BhwDesc), C.c_uint64, C.c_uint64, C.POINTER(C.c_longlong)]
    L.orc_window.restype = C.c_int
    t0 = time.perf_counter()
    for _ in range(reps):
        st = L.orc_window(C.byref(d), n0, count, p)
        assert st == 0, st
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget", type=float, default=4.0, help="seconds of wall time per line (about)")
    args = ap.parse_args()
    import blackman_harris_win_b200 as bhw
    import cases
    cores = len(os.sched_getaffinity(0))
    pool = mp.get_context("spawn").Pool(cores)
    port = os.path.join(ROOT, "oracle", "libbhw_oracle.so")
    cfgs = cases.baseline_configs()
    # (config, kind, library, selector, samples per call, model description)
    lines = []
    for name, d in cfgs.items():
        lines.append((name, "port", port, bytes(d), d.phi_width, "oracle/bhw_oracle.c restatement of the RTL entity"))
    for name, np_, nw, t in (("cfg1_hamming_n1024_dw16", 10, 16, 1), ("cfg1_hann_n1024_dw16", 10, 16, 2),
                             ("cfg2_bh4_n65536_dw17", 16, 17, 4), ("cfg3_bh7_n1m_dw32", 20, 32, 7)):
        so = os.path.join(ROOT, "oracle", "_ref", f"hls_win_np{np_}_nw{nw}.so")
        if os.path.exists(so):
            lines.append((name, "reference", so, t, np_,
                          f"unmodified hls/windows/win_function.cpp type {t}, NPHASE {np_} / NWIDTH {nw}, g++ -O2, ap_int stand-in"))
    for name, kind, so, what, pw, desc in lines:
        n = 1 << pw
        count = min(n, 1 << 16)                      # a bounded slice of long windows
        pool.map(_worker, [(kind, so, what, 0, min(count, 256), 1)] * cores)   # load the library
        t1 = max(pool.map(_worker, [(kind, so, what, 0, count, 1)] * cores))
        reps = max(1, min(1 << 14, int(args.budget / max(t1, 1e-4))))
        t0 = time.perf_counter()
        pool.map(_worker, [(kind, so, what, 0, count, reps)] * cores)
        wall = time.perf_counter() - t0
        samples = cores * reps * count
        print(json.dumps({"config": name, "kind": kind, "model": desc, "cores": cores,
                          "sample": f"{reps} x {count} samples per core (window N = {n}), one process per core",
                          "msamples_per_s_per_core": round(samples / wall / cores / 1e6, 4),
                          "gsamples_per_s": round(samples / wall / 1e9, 6), "wall_s": round(wall, 2)}), flush=True)
    pool.close()
    pool.join()


if __name__ == "__main__":
    main()
