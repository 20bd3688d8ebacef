#!/usr/bin/env python
"""bench.py - window-generation throughput on B200 (and the reference's CPU model beside it).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own C++ model, host cores

Workload (BASELINE.json configs[1]): a bank of bh_win_4term windows, PHI_WIDTH 16 (N = 65536),
DAT_WIDTH 17, cordic_dds.  One *step* generates the whole bank once: WINDOWS_PER_GPU windows per
GPU (weak scaling; the global bank is the concatenation over ranks, rank r owns flat slice r of it
- bhw_shard_range - and no data-path collective exists).  Every window of the bank has its own
AA0..AA3 port values (the BH4 set with a per-window perturbation), so no two windows are equal.
Trig tables are rebuilt inside every step (bhw_set_table_cache(0)): nothing computed in one step
is reused by the next.

Prints ONE JSON line (rank 0).  `value` = Gsamples/s of bhw_plan_execute with the plan (the resolved
per-window records) and the output resident in HBM (CUDA events on the launching stream, max over
ranks); `e2e` = the same bank through the host-buffer entry point
(bhw_generate_batch_host: descriptors in host memory, result in pinned host memory, all copies in
the timed region); `roofline` = k_synth's algorithmic store bytes / its device time (library-side
CUDA events) against the measured HBM copy bandwidth; `cpu_baseline` = the reference HLS model
timed on the host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

PHI_WIDTH = 16
DAT_WIDTH = 17
WINDOWS_PER_GPU = 4096          # 4096 x 65536 x 4 B = 1.07 GB per GPU per step (>> 126 MB L2)
BH4_AA = (47022, 64001, 18518, 1531)   # round(a_k * (2^17 - 1)), src/tb/tb_windows.vhd:103-111
KERNEL_TIMING_STRIDE = 8        # steps of the timed region whose launches carry per-kernel events: 1 in 8
METRIC = "window_gsamples_per_s"
UNIT = "Gsamples/s"
WORKLOAD = "bank of bh_win_4term windows, N=65536 (PHI_WIDTH 16), DAT_WIDTH 17, cordic_dds, RTL model"


# ---- helpers shared with tests/test_shard_gloo.py ------------------------------------------------
def max_over_ranks(x: float) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rank_workload(rank: int, world: int, nwin_per_gpu: int = WINDOWS_PER_GPU):
    """-> (flat_begin, flat_count, total): rank's contiguous slice of the global bank."""
    import blackman_harris_win_b200 as bhw
    total = (world * nwin_per_gpu) << PHI_WIDTH
    b, c = bhw.shard_range(total, rank, world)
    return b, c, total


def bank_descs(nwin: int, algo: int = 0):
    """The global bank: window i = BH4 with AA_k nudged by a per-window amount (distinct port
    values per window, all inside DAT_WIDTH bits)."""
    import blackman_harris_win_b200 as bhw
    arr = (bhw.BhwDesc * nwin)()
    proto = bhw.make_desc(4, PHI_WIDTH, DAT_WIDTH, BH4_AA, algo=algo)
    raw = bytes(proto)
    same = os.environ.get("BHW_BENCH_IDENTICAL_WINDOWS") == "1"   # tuning experiments only
    for i in range(nwin):
        C.memmove(C.byref(arr, i * C.sizeof(bhw.BhwDesc)), raw, len(raw))
        if same:
            continue
        d = arr[i]
        d.aa[0] = BH4_AA[0] - (i % 1021)
        d.aa[1] = BH4_AA[1] - (i % 509)
        d.aa[2] = BH4_AA[2] + (i % 251)
        d.aa[3] = BH4_AA[3] + (i % 127)
    return arr


# ---- clocks during the timed region ---------------------------------------------------------------
class ClockSampler:
    """Polls NVML for SM clock and throttle reasons of one GPU while the timed region runs."""

    def __init__(self, device_index: int):
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = torch.cuda.get_device_properties(device_index).uuid
            self._nv = pynvml
            try:
                self._h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._h = None

    def _run(self):
        nv = self._nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join(2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- the reference's CPU model -----------------------------------------------------------------------
def _ref_worker(args):
    """One process = one core = independent windows (the HLS cordic() rewrites a static table on
    every call, hls/windows/win_function.cpp:74-80, so it is run in processes, not threads)."""
    so, kind, nwin, desc_bytes = args
    import numpy as np
    n = 1 << PHI_WIDTH
    out = np.empty(n, dtype=np.int32)
    if kind == "reference":
        L = C.CDLL(so)
        L.ref_hls_window_i32.argtypes = [C.c_int, C.c_longlong, C.c_longlong, C.POINTER(C.c_int)]
        t0 = time.perf_counter()
        for _ in range(nwin):
            L.ref_hls_window_i32(4, 0, n, out.ctypes.data_as(C.POINTER(C.c_int)))   # type 4 = BH4
        return time.perf_counter() - t0
    from blackman_harris_win_b200.api import BhwDesc
    L = C.CDLL(so)
    d = BhwDesc.from_buffer_copy(desc_bytes)
    L.orc_window_i32.argtypes = [C.POINTER(BhwDesc), C.c_uint64, C.c_uint64, C.POINTER(C.c_int32)]
    t0 = time.perf_counter()
    for _ in range(nwin):
        L.orc_window_i32(C.byref(d), 0, n, out.ctypes.data_as(C.POINTER(C.c_int32)))
    return time.perf_counter() - t0


class CpuModel:
    """The reference's own CPU implementation of the path: oracle/_ref (unmodified
    hls/windows/win_function.cpp compiled for NPHASE 16 / NWIDTH 17) when present, else the
    oracle port of the RTL.  Executing oracle/ here is the checker/baseline leg only."""

    def __init__(self):
        import multiprocessing as mp
        ref = os.path.join(ROOT, "oracle", "_ref", f"hls_win_np{PHI_WIDTH}_nw{DAT_WIDTH}.so")
        port = os.path.join(ROOT, "oracle", "libbhw_oracle.so")
        if os.path.exists(ref):
            self.kind, self.so = "reference", ref
            self.what = ("unmodified hls/windows/win_function.cpp type 4 (BH4), NPHASE 16 / NWIDTH 17, "
                         "g++ -O2, ap_int stand-in")
        elif os.path.exists(port):
            self.kind, self.so = "port", port
            self.what = "oracle/bhw_oracle.c restatement of bh_win_4term + cordic_dds"
        else:
            raise RuntimeError("neither oracle/_ref nor oracle/libbhw_oracle.so is built")
        import blackman_harris_win_b200 as bhw
        self.desc_bytes = bytes(bhw.make_desc(4, PHI_WIDTH, DAT_WIDTH, BH4_AA))
        self.cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.pool = mp.get_context("spawn").Pool(self.cores)
        self.pool.map(_ref_worker, [(self.so, self.kind, 0, self.desc_bytes)] * self.cores)   # spin up

    def run(self, windows_per_core: int) -> float:
        """All cores generate `windows_per_core` windows each; -> wall seconds."""
        t0 = time.perf_counter()
        self.pool.map(_ref_worker, [(self.so, self.kind, windows_per_core, self.desc_bytes)] * self.cores)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_baseline(budget_s: float = 12.0):
    m = CpuModel()
    try:
        t1 = m.run(1)
        reps = max(1, min(4096, int(budget_s / max(t1, 1e-3))))
        t = m.run(reps)
        samples = reps * m.cores << PHI_WIDTH
        return {"value": samples / t / 1e9, "unit": UNIT, "cores": m.cores, "kind": m.kind,
                "sample": f"{reps} window(s) of 65536 samples per core, one process per core, {m.what}; "
                          f"{samples} samples in {t:.2f} s"}
    finally:
        m.close()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    m = CpuModel()
    try:
        for _ in range(min(args.warmup, 2)):
            m.run(1)
        per_step = []
        for _ in range(args.steps):
            per_step.append(m.run(1))
        t = sum(per_step)
        samples_per_step = m.cores << PHI_WIDTH
        v = samples_per_step * args.steps / t / 1e9
        line = {
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic (the descriptor is the input; no RNG)",
            "config": {"workload": WORKLOAD, "step": f"{m.cores} windows of 65536 samples, one per host core "
                                                      "(bounded sample of the GPU arm's bank)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": m.cores, "kind": m.kind,
                             "sample": f"{args.steps} steps x {m.cores} windows x 65536 samples; {m.what}"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
    finally:
        m.close()
    return 0


# ---- the CUDA arm ---------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import blackman_harris_win_b200 as bhw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the window generator has no CPU path "
                         "(use --impl reference for the CPU model)")
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL's own banner / debug lines go to stderr: stdout carries the one JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus and rank == 0:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}; using {world}", file=sys.stderr)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    algo = {"auto": bhw.ALGO_AUTO, "direct": bhw.ALGO_DIRECT, "table": bhw.ALGO_TABLE}[args.algo]
    nwin = world * args.windows_per_gpu
    descs = bank_descs(nwin, algo)                       # the global bank (host memory)
    begin, count, total = rank_workload(rank, world, args.windows_per_gpu)
    first, touched, local = bhw.shard_windows(descs, begin, count)   # this rank's windows
    mine = (bhw.BhwDesc * touched).from_address(C.addressof(descs) + first * C.sizeof(bhw.BhwDesc))
    out = torch.empty(count, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    L = bhw.lib()
    bhw.set_table_cache(False)          # every step rebuilds its trig tables
    t_plan = time.perf_counter()
    plan = bhw.Plan(mine)               # per-window records resident in HBM before the timed region
    t_plan = time.perf_counter() - t_plan

    def step_device():
        st = L.bhw_plan_execute(plan._h, local, count, out.data_ptr(), stream)
        if st:
            raise bhw.BhwError(st, "bhw_plan_execute")

    # ---- device-resident timing -------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    bhw.timing_reset()
    clocks = ClockSampler(local)
    launches0 = bhw.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clocks.start()
    barrier()
    e0.record()
    for i in range(args.steps):
        # per-kernel CUDA events (library side, on the launching stream) bracket the launches of
        # every KERNEL_TIMING_STRIDE-th step of the timed region; the other steps run bare
        bhw.timing_enable(i % KERNEL_TIMING_STRIDE == 0)
        step_device()
    e1.record()
    barrier()
    clocks.stop()
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    launches = bhw.launch_count() - launches0
    ktimes = bhw.timing_read()
    bhw.timing_enable(False)
    value = (total * args.steps) / (dev_ms * 1e-3) / 1e9

    # sanity: the bank really was written (first and last window of this rank vs a second, single-window call)
    chk = bhw.generate(bhw.BhwDesc.from_buffer_copy(bytes(descs[begin >> PHI_WIDTH])))
    if not torch.equal(chk, out[: 1 << PHI_WIDTH]):
        raise SystemExit("bench.py: bank output differs from the single-window call")

    # ---- end to end through the host-buffer entry point ----------------------------------------
    host = torch.empty(count, dtype=torch.int32, pin_memory=True)

    def step_host():
        st = L.bhw_generate_batch_host(mine, touched, local, count, host.data_ptr())
        if st:
            raise bhw.BhwError(st, "bhw_generate_batch_host")

    e2e = None
    if args.e2e_steps > 0:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_host()
        barrier()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        if not torch.equal(host[: 1 << PHI_WIDTH], chk.cpu()):
            raise SystemExit("bench.py: host-path output differs from the device path")
        e2e = {"value": (total * e2e_steps) / e2e_s / 1e9, "unit": UNIT, "steps": e2e_steps,
               "d2h_gbs_per_gpu": count * 4 * e2e_steps / e2e_s / 1e9,
               "h2d_bytes_per_step": _meta_bytes(touched), "d2h_bytes_per_step": count * 4,
               "api": "bhw_generate_batch_host (pinned host output)"}

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    dom = max(ktimes, key=lambda k: ktimes[k][1])
    n_l, ms_l = ktimes[dom]
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic = prof.get(dom, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roof = None
    if n_l:
        achieved = (count * 4) / (ms_l / n_l * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_per_launch": count * 4, "avg_launch_ms": ms_l / n_l, "launches_timed": n_l,
                "kernel_ms_timed": {k: v[1] for k, v in ktimes.items() if v[0]}}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic (the descriptor is the input; no RNG)",
            "config": {"workload": WORKLOAD, "windows_per_gpu": args.windows_per_gpu,
                       "samples_per_step": total, "bytes_per_step_per_gpu": count * 4, "algo": args.algo,
                       "l2": "output per step (1.07 GB/GPU) is larger than the 126 MB L2; no flush needed",
                       "tables": "rebuilt every step (table cache off)",
                       "plan_create_ms": round(1e3 * t_plan, 3),
                       "sharding": f"flat sample range, {world} rank(s), no collective"},
            "roofline": roof, "clocks": clocks.summary(),
            "e2e": e2e,
            "gpu_launches": int(launches),
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(args.cpu_budget)
        print(json.dumps(line))
    plan.destroy()
    if world > 1:
        dist.destroy_process_group()
    return 0


def _meta_bytes(nwin: int) -> int:
    """Host->device bytes of one host-API call: the per-window records the library uploads
    (224 B each), the window->record index (4 B each) and one table-job header; the descriptors
    themselves are read on the host."""
    return nwin * (224 + 4) + 256


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--algo", default="auto", choices=["auto", "direct", "table"])
    ap.add_argument("--windows-per-gpu", type=int, default=WINDOWS_PER_GPU)
    ap.add_argument("--e2e-steps", type=int, default=10, help="0 skips the host-buffer leg")
    ap.add_argument("--cpu-budget", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
