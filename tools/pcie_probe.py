#!/usr/bin/env python
"""What bounds the host-buffer entry point (bench.py `e2e`): the device->host link.

Times (a) a plain pinned cudaMemcpy D2H of the bench bank's size, (b) bhw_generate_batch_host on
the bench bank, (c) host-side planning alone (bhw_plan_create/destroy).  Prints one JSON line.

  python tools/pcie_probe.py [--windows 4096]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402
import bench  # noqa: E402


def bind_to_gpu_numa_node(dev: int):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (first-touch then places the
    pinned buffer there).  -> node id or None."""
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devid = torch.cuda.get_device_properties(dev).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{devid:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return None
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return node
    except Exception as e:   # noqa: BLE001
        return f"failed: {e}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--windows", type=int, default=bench.WINDOWS_PER_GPU)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--numa", action="store_true", help="bind to the GPU's NUMA node before allocating pinned memory")
    ap.add_argument("--start-at", type=float, default=0.0, help="unix time to start the timed part at (to line up concurrent probes)")
    args = ap.parse_args()
    torch.cuda.set_device(args.device)
    numa = bind_to_gpu_numa_node(args.device) if args.numa else None
    nwin = args.windows
    count = nwin << bench.PHI_WIDTH
    dev = torch.empty(count, dtype=torch.int32, device="cuda")
    host = torch.empty(count, dtype=torch.int32, pin_memory=True)
    res = {"bytes": count * 4, "device": args.device, "numa_node": numa}
    while time.time() < args.start_at:
        time.sleep(0.001)
    # (a) plain D2H, whole buffer and 64 MiB pieces
    for name, piece in (("d2h_whole_gbs", count), ("d2h_64mib_pieces_gbs", (64 << 20) // 4), ("d2h_16mib_pieces_gbs", (16 << 20) // 4)):
        for rep in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                for o in range(0, count, piece):
                    host[o:o + piece].copy_(dev[o:o + piece], non_blocking=True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / args.reps
        res[name] = round(count * 4 / dt / 1e9, 2)
    # H2D for symmetry
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    res["h2d_whole_gbs"] = round(count * 4 / ((time.perf_counter() - t0) / args.reps) / 1e9, 2)
    # (b) the host entry point
    descs = bench.bank_descs(nwin)
    L = bhw.lib()
    for _ in range(2):
        st = L.bhw_generate_batch_host(descs, nwin, 0, count, host.data_ptr())
        assert st == 0, st
    t0 = time.perf_counter()
    for _ in range(args.reps):
        L.bhw_generate_batch_host(descs, nwin, 0, count, host.data_ptr())
    dt = (time.perf_counter() - t0) / args.reps
    res["generate_batch_host_ms"] = round(1e3 * dt, 3)
    res["generate_batch_host_gbs"] = round(count * 4 / dt / 1e9, 2)
    # (c) planning alone
    t0 = time.perf_counter()
    for _ in range(args.reps):
        p = bhw.Plan(descs)
        p.destroy()
    res["plan_create_destroy_ms"] = round(1e3 * (time.perf_counter() - t0) / args.reps, 3)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
