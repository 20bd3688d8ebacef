#!/usr/bin/env python
"""Device -> host link ceiling for bench.py's `e2e`, by destination kind: torch pinned memory (cudaHostAlloc
default), write-combined pinned memory (cudaHostAllocWriteCombined), cudaHostRegister'ed 2 MB-aligned
memory (madvise MADV_HUGEPAGE).  One JSON line.   python tools/d2h_probe.py [--mb 1024]"""
import argparse
import ctypes as C
import json
import mmap
import time

import torch

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=int, default=1024)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
n = a.mb << 20
rt = C.CDLL("libcudart.so.12")
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
dev.fill_(7)
torch.cuda.synchronize()


def time_copy(dst_ptr):
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rt.cudaMemcpy(dst_ptr, dev.data_ptr(), n, 2)
    best = 0.0
    for _ in range(a.reps):
        t0 = time.perf_counter()
        rt.cudaMemcpy(dst_ptr, dev.data_ptr(), n, 2)
        best = max(best, n / (time.perf_counter() - t0) / 1e9)
    return round(best, 2)


res = {"bytes": n}
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
res["pinned_default_gbs"] = time_copy(h.data_ptr())
del h
p = C.c_void_p()
rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
if rt.cudaHostAlloc(C.byref(p), n, 0x04) == 0:          # cudaHostAllocWriteCombined
    res["pinned_write_combined_gbs"] = time_copy(p.value)
    rt.cudaFreeHost.argtypes = [C.c_void_p]
    rt.cudaFreeHost(p)
try:
    m = mmap.mmap(-1, n + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    addr = C.addressof(C.c_char.from_buffer(m))
    aligned = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
    libc = C.CDLL("libc.so.6")
    libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
    res["madvise_hugepage_rc"] = libc.madvise(aligned, n, 14)
    C.memset(aligned, 0, n)
    rt.cudaHostRegister.argtypes = [C.c_void_p, C.c_size_t, C.c_uint]
    if rt.cudaHostRegister(aligned, n, 0) == 0:
        res["registered_hugepage_gbs"] = time_copy(aligned)
        rt.cudaHostUnregister.argtypes = [C.c_void_p]
        rt.cudaHostUnregister(aligned)
except Exception as ex:   # noqa: BLE001
    res["registered_hugepage_error"] = str(ex)
print(json.dumps(res))
