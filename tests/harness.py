"""Test-side loaders: the CPU oracle (oracle/), the compiled reference models (oracle/_ref/), and
the host build of the kernel bodies (tests/hostcheck/).  TEST INFRASTRUCTURE ONLY - nothing in
the product imports this."""
from __future__ import annotations

import ctypes as C
import functools
import hashlib
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from blackman_harris_win_b200.api import BhwDesc  # noqa: E402  (struct layout only)

ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
HOSTCHECK_DIR = os.path.join(ROOT, "tests", "hostcheck")

P = C.POINTER
I64P = P(C.c_int64)


def _build(cmd, cwd):
    subprocess.run(cmd, cwd=cwd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


@functools.lru_cache(None)
def oracle():
    path = os.path.join(ORACLE_DIR, "libbhw_oracle.so")
    if not os.path.exists(path):
        _build(["make", "oracle"], ORACLE_DIR)
    L = C.CDLL(path)
    D = P(BhwDesc)
    L.orc_validate.argtypes = [D]
    L.orc_window.argtypes = [D, C.c_uint64, C.c_uint64, I64P]
    L.orc_window_mt.argtypes = [D, C.c_uint64, C.c_uint64, I64P, C.c_int]
    L.orc_window_i32.argtypes = [D, C.c_uint64, C.c_uint64, P(C.c_int32)]
    L.orc_sincos.argtypes = [D, C.c_uint64, C.c_uint64, I64P, I64P]
    L.orc_apply.argtypes = [D, C.c_int, P(C.c_int32), C.c_uint64, I64P]
    L.orc_atan2.argtypes = [C.c_int, C.c_int, C.c_int, P(C.c_int32), P(C.c_int32), P(C.c_int32), C.c_uint64]
    L.orc_atan2_validate.argtypes = [C.c_int, C.c_int, C.c_int]
    L.orc_atan2_stream.argtypes = [C.c_int, C.c_int, C.c_int, P(C.c_int32), P(C.c_int32), P(C.c_int32), C.c_uint64]
    L.orc_cordic_atan2.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64]
    L.orc_cordic_atan2.restype = C.c_int64
    L.orc_cordic_dds.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, I64P, I64P]
    L.orc_cordic_dds.restype = None
    L.orc_quantize.argtypes = [C.c_int, C.c_int, C.c_int, I64P, P(C.c_int32)]
    L.orc_taylor_rom.argtypes = [C.c_int, C.c_int, I64P, I64P]
    return L


def _i64(n):
    return np.empty(int(n), dtype=np.int64)


def _p(a):
    return a.ctypes.data_as(I64P)


def orc_window(d: BhwDesc, n0=0, count=None, threads=1) -> np.ndarray:
    if count is None:
        count = (1 << d.phi_width) - n0
    out = _i64(count)
    if threads > 1:
        st = oracle().orc_window_mt(C.byref(d), n0, count, _p(out), threads)
    else:
        st = oracle().orc_window(C.byref(d), n0, count, _p(out))
    if st:
        raise ValueError(f"orc_window status {st}")
    return out


def orc_atan2(iw: int, aw: int, prec: int, x: np.ndarray, y: np.ndarray, stream: bool = False) -> np.ndarray:
    """stream=True: PHI_DT as the entity streams it (pair t with the quadrant of pair t+1, see orc_cordic_atan2_q)."""
    x = np.ascontiguousarray(x, dtype=np.int32)
    y = np.ascontiguousarray(y, dtype=np.int32)
    out = np.empty(x.shape, np.int32)
    p32 = P(C.c_int32)
    fn = oracle().orc_atan2_stream if stream else oracle().orc_atan2
    st = fn(iw, aw, prec, x.ctypes.data_as(p32), y.ctypes.data_as(p32), out.ctypes.data_as(p32), x.size)
    if st:
        raise ValueError(f"orc_atan2: status {st}")
    return out


def orc_apply(d: BhwDesc, x: np.ndarray, mode: int) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.int32)
    n = 1 << d.phi_width
    assert x.size % n == 0
    y = _i64(x.size)
    st = oracle().orc_apply(C.byref(d), mode, x.ctypes.data_as(P(C.c_int32)), x.size // n, _p(y))
    if st:
        raise ValueError(f"orc_apply status {st}")
    return y.reshape(x.shape)


def orc_window_status(d: BhwDesc) -> int:
    return oracle().orc_validate(C.byref(d))


def orc_sincos(d: BhwDesc, n0=0, count=None):
    if count is None:
        count = (1 << d.phi_width) - n0
    s, c = _i64(count), _i64(count)
    st = oracle().orc_sincos(C.byref(d), n0, count, _p(s), _p(c))
    if st:
        raise ValueError(f"orc_sincos status {st}")
    return s, c


def orc_quantize(variant, rule, dw):
    aa = (C.c_int64 * 11)()
    wt = C.c_int32(0)
    st = oracle().orc_quantize(variant, rule, dw, aa, C.byref(wt))
    if st:
        raise ValueError(f"orc_quantize status {st}")
    return list(aa), wt.value


def orc_batch(descs, flat_begin, flat_count) -> np.ndarray:
    """Oracle for the flat-range batch contract: concatenation of full windows."""
    out = _i64(flat_count)
    off = 0
    end = flat_begin + flat_count
    for d in descs:
        n = 1 << d.phi_width
        lo, hi = max(off, flat_begin), min(off + n, end)
        if lo < hi:
            out[lo - flat_begin: hi - flat_begin] = orc_window(d, lo - off, hi - lo)
        off += n
    return out


# ---- compiled reference (oracle/_ref) ---------------------------------------------------------
def ref_path(kind: str, a: int, b: int) -> str:
    name = {"hls_win": f"hls_win_np{a}_nw{b}.so", "hls_cordic": f"hls_cordic_np{a}_nw{b}.so",
            "cpp": f"cpp_cordic_pw{a}_dw{b}.so"}[kind]
    return os.path.join(REF_DIR, name)


def ref_configs(kind: str):
    """(a, b) pairs for which a compiled reference object exists."""
    if not os.path.isdir(REF_DIR):
        return []
    out = []
    pre = {"hls_win": "hls_win_np", "hls_cordic": "hls_cordic_np", "cpp": "cpp_cordic_pw"}[kind]
    for f in sorted(os.listdir(REF_DIR)):
        if f.startswith(pre) and f.endswith(".so"):
            a, b = f[len(pre):-3].replace("_nw", " ").replace("_dw", " ").split()
            out.append((int(a), int(b)))
    return out


@functools.lru_cache(None)
def ref_lib(kind: str, a: int, b: int):
    path = ref_path(kind, a, b)
    if not os.path.exists(path):
        return None
    L = C.CDLL(path)
    if kind == "hls_win":
        L.ref_hls_window.argtypes = [C.c_int, C.c_longlong, C.c_longlong, P(C.c_longlong)]
        L.ref_hls_window_i32.argtypes = [C.c_int, C.c_longlong, C.c_longlong, P(C.c_int)]
        L.ref_hls_cordic.argtypes = [C.c_longlong, C.c_longlong, P(C.c_longlong), P(C.c_longlong)]
    elif kind == "hls_cordic":
        L.ref_hls_cordic.argtypes = [C.c_longlong, C.c_longlong, P(C.c_longlong), P(C.c_longlong)]
    else:
        L.ref_cpp_cordic.argtypes = [C.c_longlong, C.c_longlong, P(C.c_int), P(C.c_int)]
    return L


def ref_hls_window(np_, nw, win_type, n0=0, count=None):
    L = ref_lib("hls_win", np_, nw)
    if count is None:
        count = (1 << np_) - n0
    out = _i64(count)
    L.ref_hls_window(win_type, n0, count, out.ctypes.data_as(P(C.c_longlong)))
    return out


def ref_hls_cordic(np_, nw, n0=0, count=None, which="hls_cordic"):
    L = ref_lib(which, np_, nw)
    if count is None:
        count = (1 << np_) - n0
    s, c = _i64(count), _i64(count)
    L.ref_hls_cordic(n0, count, s.ctypes.data_as(P(C.c_longlong)), c.ctypes.data_as(P(C.c_longlong)))
    return s, c


def ref_cpp_cordic(pw, dw, n0=0, count=None):
    L = ref_lib("cpp", pw, dw)
    if count is None:
        count = (1 << pw) - n0
    s, c = np.empty(count, np.int32), np.empty(count, np.int32)
    L.ref_cpp_cordic(n0, count, s.ctypes.data_as(P(C.c_int)), c.ctypes.data_as(P(C.c_int)))
    return s.astype(np.int64), c.astype(np.int64)


# ---- host build of the kernel bodies ----------------------------------------------------------
@functools.lru_cache(None)
def hostcheck():
    path = os.path.join(HOSTCHECK_DIR, "libbhw_hostcheck.so")
    srcs = [os.path.join(HOSTCHECK_DIR, "hostcheck.cpp")] + [
        os.path.join(ROOT, "blackman_harris_win_b200", "csrc", f)
        for f in ("bhw_resolve.cpp", "bhw_plan.cpp", "bhw_device.cuh", "bhw_group.cuh", "bhw_internal.h", "bhw_plan.h")]
    stale = not os.path.exists(path) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(path) for s in srcs)
    if stale:
        _build(["make", "-s"], HOSTCHECK_DIR)
    L = C.CDLL(path)
    D = P(BhwDesc)
    L.hc_direct.argtypes = [D, C.c_uint64, C.c_uint64, I64P]
    L.hc_table.argtypes = [D, C.c_uint64, C.c_uint64, I64P, C.c_int]
    L.hc_direct32.argtypes = [D, C.c_uint64, C.c_uint64, I64P]
    L.hc_direct_taylor.argtypes = [D, C.c_uint64, C.c_uint64, I64P]
    L.hc_taylor_quad_ok.argtypes = [D]
    from blackman_harris_win_b200.api import BhwAtan2Desc
    L.hc_atan2.argtypes = [P(BhwAtan2Desc), P(C.c_int32), P(C.c_int32), P(C.c_int32), C.c_uint64]
    L.hc_sincos.argtypes = [D, C.c_uint64, C.c_uint64, I64P, I64P]
    L.hc_table_cos.argtypes = [D, I64P, C.c_int]
    L.hc_bank.argtypes = [D, I64P, C.c_uint64, C.c_int, C.c_int]
    L.hc_source_antisymmetric.argtypes = [D]
    L.hc_group.argtypes = [D, C.c_int, I64P, C.c_int, C.c_int]
    L.hc_tail_mode.argtypes = [D]
    L.hc_walk.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32]
    L.hc_lin_tiles.argtypes = [C.c_int]
    L.hc_lin_tiles.restype = C.c_uint64
    L.hc_inq_exceptions.argtypes = [C.c_int]
    L.hc_inq_exceptions.restype = C.c_uint64
    return L


# ---- hashing in the format of the SURVEY known answers ----------------------------------------
def sha_lines(values) -> str:
    """sha256 of one decimal per line ('%d\\n')."""
    return hashlib.sha256(("\n".join(str(int(v)) for v in values) + "\n").encode()).hexdigest()


def sha_pairs(s, c) -> str:
    """sha256 of 's c\\n' lines (cpp/cordic_sincos.cpp:138 output format)."""
    return hashlib.sha256("".join(f"{int(a)} {int(b)}\n" for a, b in zip(s, c)).encode()).hexdigest()


# ---- tests/golden/rtl_sim_cases.json helpers (vectors made by executing the reference VHDL, oracle/vhdl_sim.py)
_RTL_TERMS = {"hamming_win": 2, "bh_win_3term": 3, "bh_win_4term": 4, "bh_win_5term": 5, "bh_win_7term": 7,
              "HAMMING": 2, "BH3TERM": 3, "BH4TERM": 4, "BH5TERM": 5, "BH7TERM": 7}


def rtl_case_terms(case) -> int:
    return _RTL_TERMS[case["generics"]["WIN_TYPE"] if case["entity"] == "win_selector" else case["entity"]]


def rtl_case_is_taylor(case) -> bool:
    return case["generics"].get("SIN_TYPE", "CORDIC") == "TAYLOR" and rtl_case_terms(case) <= 3


def rtl_case_desc(case, aa, stream_offset=0) -> BhwDesc:
    """The descriptor that names what a simulated window entity was elaborated with."""
    import blackman_harris_win_b200 as bhw
    g, m = case["generics"], rtl_case_terms(case)
    tay = rtl_case_is_taylor(case)
    return bhw.make_desc(m, g["PHI_WIDTH"], g["DAT_WIDTH"], [int(a) for a in aa[:m]], stream_offset=stream_offset,
                         sin_type=bhw.SIN_TAYLOR if tay else bhw.SIN_CORDIC, lut_size=g["LUT_SIZE"] if tay else 0)


def rtl_case_first_vld(case) -> int:
    """Clock index (ENABLE raised at clock 0) of the first DT_VLD = the entity's ADD_DELAY + 1
    (src/hamming_win.vhd find_delay, src/bh_win_3term.vhd:120-140, +1 / +2 in the 4/5- and 7-term entities)."""
    g, m = case["generics"], rtl_case_terms(case)
    if rtl_case_is_taylor(case):
        if g["PHI_WIDTH"] - g["LUT_SIZE"] <= 2:
            return 11
        return 14 if g["DAT_WIDTH"] < 19 else 17
    return g["DAT_WIDTH"] + {2: 8, 3: 8, 4: 9, 5: 9, 7: 10}[m]


def rtl_taylor_latency(pw: int, dw: int, lut: int) -> int:
    """Clocks from the phase counter value to OUT_SIN / OUT_COS (src/taylor_sincos.vhd:170-215)."""
    if pw - lut <= 2:
        return 3
    return 6 if dw < 19 else 9
