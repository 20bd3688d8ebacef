"""ctypes binding of include/bhw.h.  See the package docstring for the reference interfaces."""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# ---- enums (include/bhw.h) -------------------------------------------------------------------
SIN_CORDIC, SIN_TAYLOR, SIN_CORDIC48, SIN_CORDIC_SCALED = 0, 1, 2, 3
MODEL_RTL, MODEL_HLS, MODEL_CPP = 0, 1, 2
ALGO_AUTO, ALGO_DIRECT, ALGO_TABLE = 0, 1, 2
OUT_DEFAULT, OUT_INT16 = 0, 1
RULE_TB, RULE_HLS = 0, 1
MAX_TERMS = 11
VARIANT_NAMES = {
    1: "hamming", 2: "hann", 3: "blackman", 4: "blackman_harris_3", 5: "nuttall",
    6: "blackman_harris_4", 7: "blackman_nuttall", 8: "flat_top", 9: "blackman_harris_5",
    10: "blackman_harris_7", 11: "blackman_harris_7_readme", 12: "hamming_alt", 13: "flat_top_normalised",
    # minimum-sidelobe sets of doc/blackman-harris coef.jpg without a reference entity (BHW_WIN_MTERM_*)
    14: "min_sidelobe_6", 15: "min_sidelobe_8", 16: "min_sidelobe_9", 17: "min_sidelobe_10", 18: "min_sidelobe_11",
}
_WIN_TYPE_NAMES = {"HAMMING": 2, "BH3TERM": 3, "BH4TERM": 4, "BH5TERM": 5, "BH7TERM": 7}
_SIN_TYPE_NAMES = {"CORDIC": SIN_CORDIC, "TAYLOR": SIN_TAYLOR, "CORDIC48": SIN_CORDIC48,
                   "CORDIC_SCALED": SIN_CORDIC_SCALED}


class BhwError(RuntimeError):
    def __init__(self, status: int, where: str = ""):
        self.status = status
        msg = strerror(status) if _lib is not None else f"status {status}"
        if status == -12 and _lib is not None:
            msg += f" ({_lib.bhw_last_cuda_error().decode()})"
        super().__init__(f"{where}: {msg} [{status}]" if where else f"{msg} [{status}]")


class BhwAtan2Desc(C.Structure):
    """struct bhw_atan2_desc - the generics of cordic_atan2 (src/cordic_atan2.vhd:64-69)."""
    _fields_ = [("input_width", C.c_int32), ("angle_width", C.c_int32), ("precision", C.c_int32),
                ("stream_quadrant", C.c_int32)]


class BhwDesc(C.Structure):
    """struct bhw_desc - one field per win_selector generic/port (src/win_selector.vhd:60-87)."""
    _fields_ = [
        ("win_type", C.c_int32), ("sin_type", C.c_int32), ("model", C.c_int32),
        ("phi_width", C.c_int32), ("dat_width", C.c_int32), ("precision", C.c_int32),
        ("lut_size", C.c_int32), ("stream_offset", C.c_int32), ("algo", C.c_int32),
        ("out_format", C.c_int32), ("aa", C.c_int64 * MAX_TERMS),
    ]

    def copy(self, **changes) -> "BhwDesc":
        d = BhwDesc.from_buffer_copy(bytes(self))
        for k, v in changes.items():
            if k == "aa":
                for i in range(MAX_TERMS):
                    d.aa[i] = int(v[i]) if i < len(v) else 0
            else:
                setattr(d, k, v)
        return d

    def __repr__(self):
        return ("BhwDesc(win_type=%d, sin_type=%d, model=%d, phi_width=%d, dat_width=%d, precision=%d, "
                "lut_size=%d, stream_offset=%d, algo=%d%s, aa=%s)" % (
                    self.win_type, self.sin_type, self.model, self.phi_width, self.dat_width,
                    self.precision, self.lut_size, self.stream_offset, self.algo,
                    ", out_format=%d" % self.out_format if self.out_format else "",
                    list(self.aa)[: max(self.win_type, 1)]))


_lib = None


def lib_path() -> str:
    return os.environ.get("BHW_LIB", os.path.join(_HERE, "libbhw.so"))


class BhwLaunchRecord(C.Structure):
    """bhw_launch_record (include/bhw.h)."""
    _fields_ = [("kernel_class", C.c_int32), ("tag", C.c_uint32), ("bytes", C.c_uint64), ("ms", C.c_double)]


def lib():
    """Load libbhw.so (built in-tree by csrc/build.sh / __graft_entry__.build()).  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with blackman_harris_win_b200/csrc/build.sh "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(path)
    P = C.POINTER
    D = P(BhwDesc)
    sig = {
        "bhw_strerror": (C.c_char_p, [C.c_int]),
        "bhw_version": (C.c_int, []),
        "bhw_validate": (C.c_int, [D]),
        "bhw_elem_bytes": (C.c_int, [D]),
        "bhw_quantize": (C.c_int, [C.c_int, C.c_int, C.c_int, P(C.c_int64), P(C.c_int32)]),
        "bhw_variant_coeffs": (C.c_int, [C.c_int, C.c_int, P(C.c_double), P(C.c_int32)]),
        "bhw_generate": (C.c_int, [D, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
        "bhw_generate_host": (C.c_int, [D, C.c_void_p, C.c_uint64, C.c_uint64]),
        "bhw_batch_total": (C.c_int, [D, C.c_int, P(C.c_uint64)]),
        "bhw_shard_range": (C.c_int, [C.c_uint64, C.c_int, C.c_int, P(C.c_uint64), P(C.c_uint64)]),
        "bhw_generate_batch": (C.c_int, [D, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
        "bhw_generate_batch_host": (C.c_int, [D, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]),
        "bhw_generate_batch_multi": (C.c_int, [D, C.c_int, C.c_int, P(C.c_void_p)]),
        "bhw_sincos": (C.c_int, [D, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
        "bhw_cache_clear": (C.c_int, []),
        "bhw_set_table_cache": (C.c_int, [C.c_int]),
        "bhw_shard_range_cost": (C.c_int, [C.POINTER(BhwDesc), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64),
                                           C.POINTER(C.c_uint64)]),
        "bhw_atan2_validate": (C.c_int, [C.POINTER(BhwAtan2Desc)]),
        "bhw_atan2": (C.c_int, [C.POINTER(BhwAtan2Desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
        "bhw_atan2_host": (C.c_int, [C.POINTER(BhwAtan2Desc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]),
        "bhw_set_side_streams": (C.c_int, [C.c_int]),
        "bhw_launch_count": (C.c_uint64, []),
        "bhw_last_cuda_error": (C.c_char_p, []),
        "bhw_device_count": (C.c_int, []),
        "bhw_shard_windows": (C.c_int, [D, C.c_int, C.c_uint64, C.c_uint64, P(C.c_int), P(C.c_int), P(C.c_uint64)]),
        "bhw_plan_create": (C.c_int, [D, C.c_int, P(C.c_void_p)]),
        "bhw_plan_execute": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
        "bhw_plan_total": (C.c_int, [C.c_void_p, P(C.c_uint64)]),
        "bhw_plan_destroy": (C.c_int, [C.c_void_p]),
        "bhw_timing_enable": (C.c_int, [C.c_int]),
        "bhw_timing_reset": (C.c_int, []),
        "bhw_timing_read": (C.c_int, [C.c_int, P(C.c_double), P(C.c_uint64)]),
        "bhw_timing_launches": (C.c_int, [P(BhwLaunchRecord), C.c_uint64, P(C.c_uint64)]),
        "bhw_apply": (C.c_int, [D, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]),
        "bhw_generate_repeat": (C.c_int, [D, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_void_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


ABI_SYMBOLS = (
    "bhw_strerror", "bhw_version", "bhw_validate", "bhw_elem_bytes", "bhw_quantize",
    "bhw_variant_coeffs", "bhw_generate", "bhw_generate_host", "bhw_batch_total", "bhw_shard_range", "bhw_shard_range_cost",
    "bhw_generate_batch", "bhw_generate_batch_host", "bhw_generate_batch_multi", "bhw_sincos",
    "bhw_atan2_validate", "bhw_atan2", "bhw_atan2_host", "bhw_cache_clear", "bhw_set_table_cache", "bhw_set_side_streams", "bhw_launch_count", "bhw_last_cuda_error",
    "bhw_device_count", "bhw_timing_enable", "bhw_timing_reset", "bhw_timing_read",
    "bhw_shard_windows", "bhw_plan_create", "bhw_plan_execute", "bhw_plan_total", "bhw_plan_destroy",
    "bhw_timing_launches", "bhw_generate_repeat", "bhw_apply",
)


def _check(st: int, where: str):
    if st != 0:
        raise BhwError(st, where)


def strerror(status: int) -> str:
    return lib().bhw_strerror(int(status)).decode()


# ---- descriptors -----------------------------------------------------------------------------
def make_desc(win_type: int, phi_width: int, dat_width: int, aa: Sequence[int] = (), *,
              sin_type: int = SIN_CORDIC, model: int = MODEL_RTL, precision: int = 0,
              lut_size: int = 0, stream_offset: int = 0, algo: int = ALGO_AUTO, out_format: int = 0) -> BhwDesc:
    d = BhwDesc()
    d.win_type, d.sin_type, d.model = int(win_type), int(sin_type), int(model)
    d.phi_width, d.dat_width, d.precision = int(phi_width), int(dat_width), int(precision)
    d.lut_size, d.stream_offset, d.algo, d.out_format = int(lut_size), int(stream_offset), int(algo), int(out_format)
    for i, v in enumerate(aa):
        d.aa[i] = int(v)
    return d


def desc_array(descs: Iterable[BhwDesc]):
    descs = list(descs)
    arr = (BhwDesc * len(descs))()
    for i, d in enumerate(descs):
        C.memmove(C.byref(arr, i * C.sizeof(BhwDesc)), C.byref(d), C.sizeof(BhwDesc))
    return arr


def validate(d: BhwDesc) -> int:
    return lib().bhw_validate(C.byref(d))


def elem_bytes(d: BhwDesc) -> int:
    return lib().bhw_elem_bytes(C.byref(d))


def quantize(variant: int, rule: int, dat_width: int):
    """-> (aa[11], win_type): the reference's own quantisation rules (src/tb/tb_windows.vhd:75-127,
    hls/windows/win_function.cpp:176-355)."""
    aa = (C.c_int64 * MAX_TERMS)()
    wt = C.c_int32(0)
    _check(lib().bhw_quantize(variant, rule, dat_width, aa, C.byref(wt)), "bhw_quantize")
    return list(aa), wt.value


def variant_coeffs(variant: int, rule: int = RULE_TB):
    a = (C.c_double * MAX_TERMS)()
    m = C.c_int32(0)
    _check(lib().bhw_variant_coeffs(variant, rule, a, C.byref(m)), "bhw_variant_coeffs")
    return list(a)[: m.value]


def variant_desc(variant: int, phi_width: int, dat_width: int, *, model: int = MODEL_RTL,
                 sin_type: int = SIN_CORDIC, rule: Optional[int] = None, **kw) -> BhwDesc:
    """Descriptor of one of the 10 README variants with coefficients quantised by the reference's
    rule for that model (TB rule for RTL, HLS rule for HLS)."""
    if rule is None:
        rule = RULE_HLS if model == MODEL_HLS else RULE_TB
    aa, wt = quantize(variant, rule, dat_width)
    return make_desc(wt, phi_width, dat_width, aa, sin_type=sin_type, model=model, **kw)


def batch_total(descs) -> int:
    arr = descs if isinstance(descs, C.Array) else desc_array(descs)
    t = C.c_uint64(0)
    _check(lib().bhw_batch_total(arr, len(arr), C.byref(t)), "bhw_batch_total")
    return t.value


def shard_range(total: int, rank: int, nranks: int):
    b, c = C.c_uint64(0), C.c_uint64(0)
    _check(lib().bhw_shard_range(total, rank, nranks, C.byref(b), C.byref(c)), "bhw_shard_range")
    return b.value, c.value


def shard_range_cost(descs, rank: int, nranks: int):
    """Cost-balanced contiguous flat slice of rank `rank` (bhw_shard_range_cost)."""
    arr = descs if isinstance(descs, C.Array) else desc_array(descs)
    b, c = C.c_uint64(0), C.c_uint64(0)
    _check(lib().bhw_shard_range_cost(arr, len(arr), rank, nranks, C.byref(b), C.byref(c)), "bhw_shard_range_cost")
    return b.value, c.value


# ---- generation ------------------------------------------------------------------------------
def _torch():
    import torch
    if not torch.cuda.is_available():
        raise BhwError(-13, "no CUDA device: the window generator has no CPU path")
    return torch


def _np_dtype(esz: int):
    return {8: np.int64, 4: np.int32, 2: np.int16}[esz]


def _dev_out(torch, esz, count, out, device):
    dt = {8: torch.int64, 4: torch.int32, 2: torch.int16}[esz]
    if out is None:
        out = torch.empty(count, dtype=dt, device=device if device is not None else "cuda")
    else:
        if not out.is_cuda or out.dtype != dt or not out.is_contiguous() or out.numel() < count:
            raise ValueError("out must be a contiguous CUDA tensor of the element type, >= count long")
    return out


def generate(d: BhwDesc, n0: int = 0, count: Optional[int] = None, out=None, device=None):
    """One window (or the range [n0, n0+count) of it) into device memory, on torch's current
    stream.  Returns a torch int32/int64 CUDA tensor.  Replaces the per-sample loop around
    win_function (hls/windows/window_test.cpp:93,193) / a win_selector instance."""
    torch = _torch()
    if count is None:
        count = (1 << d.phi_width) - n0
    esz = elem_bytes(d)
    out = _dev_out(torch, esz, count, out, device)
    with torch.cuda.device(out.device):
        st = lib().bhw_generate(C.byref(d), out.data_ptr(), n0, count,
                                torch.cuda.current_stream().cuda_stream)
    _check(st, "bhw_generate")
    return out[:count]


def generate_host(d: BhwDesc, n0: int = 0, count: Optional[int] = None, out: Optional[np.ndarray] = None):
    """Same with a HOST output buffer (numpy, or any object exposing ctypes.data); device work and
    the copy back happen inside the call."""
    if count is None:
        count = (1 << d.phi_width) - n0
    esz = elem_bytes(d)
    if out is None:
        out = np.empty(count, dtype=_np_dtype(esz))
    _check(lib().bhw_generate_host(C.byref(d), out.ctypes.data, n0, count), "bhw_generate_host")
    return out


def generate_batch(descs, flat_begin: int = 0, flat_count: Optional[int] = None, out=None, device=None):
    torch = _torch()
    arr = descs if isinstance(descs, C.Array) else desc_array(descs)
    if flat_count is None:
        flat_count = batch_total(arr) - flat_begin
    esz = elem_bytes(arr[0])
    out = _dev_out(torch, esz, flat_count, out, device)
    with torch.cuda.device(out.device):
        st = lib().bhw_generate_batch(arr, len(arr), flat_begin, flat_count, out.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream)
    _check(st, "bhw_generate_batch")
    return out[:flat_count]


def generate_batch_host(descs, flat_begin: int = 0, flat_count: Optional[int] = None, out=None):
    """`out`: numpy array or an int giving a raw host pointer (e.g. of a pinned torch tensor)."""
    arr = descs if isinstance(descs, C.Array) else desc_array(descs)
    if flat_count is None:
        flat_count = batch_total(arr) - flat_begin
    esz = elem_bytes(arr[0])
    ret = out
    if out is None:
        ret = out = np.empty(flat_count, dtype=_np_dtype(esz))
    ptr = out if isinstance(out, int) else out.ctypes.data
    _check(lib().bhw_generate_batch_host(arr, len(arr), flat_begin, flat_count, ptr),
           "bhw_generate_batch_host")
    return ret


def shard_windows(descs, flat_begin: int, flat_count: int):
    """-> (first_win, nwin_touched, local_begin) of the windows a flat range touches."""
    arr = descs if isinstance(descs, C.Array) else desc_array(descs)
    f, n, b = C.c_int(0), C.c_int(0), C.c_uint64(0)
    _check(lib().bhw_shard_windows(arr, len(arr), flat_begin, flat_count, C.byref(f), C.byref(n), C.byref(b)),
           "bhw_shard_windows")
    return f.value, n.value, b.value


class Plan:
    """A batch resolved once and resident on the current device (bhw_plan_*): the elaborated
    entity instances; ``execute`` is the ENABLE burst."""

    def __init__(self, descs, device=None):
        torch = _torch()
        self._arr = descs if isinstance(descs, C.Array) else desc_array(descs)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._h = C.c_void_p(None)
        with torch.cuda.device(self.device):
            _check(lib().bhw_plan_create(self._arr, len(self._arr), C.byref(self._h)), "bhw_plan_create")
        t = C.c_uint64(0)
        _check(lib().bhw_plan_total(self._h, C.byref(t)), "bhw_plan_total")
        self.total = t.value
        self.elem_bytes = elem_bytes(self._arr[0])

    def execute(self, flat_begin: int = 0, flat_count: Optional[int] = None, out=None):
        torch = _torch()
        if flat_count is None:
            flat_count = self.total - flat_begin
        out = _dev_out(torch, self.elem_bytes, flat_count, out, self.device)
        with torch.cuda.device(self.device):
            st = lib().bhw_plan_execute(self._h, flat_begin, flat_count, out.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
        _check(st, "bhw_plan_execute")
        return out[:flat_count]

    def destroy(self):
        if self._h:
            lib().bhw_plan_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def sincos(d: BhwDesc, n0: int = 0, count: Optional[int] = None, device=None):
    """(sin, cos) tables of the DDS entity / C++ cordic() the descriptor names."""
    torch = _torch()
    if count is None:
        count = (1 << d.phi_width) - n0
    esz = elem_bytes(d)
    s = _dev_out(torch, esz, count, None, device)
    c = _dev_out(torch, esz, count, None, device)
    with torch.cuda.device(s.device):
        st = lib().bhw_sincos(C.byref(d), s.data_ptr(), c.data_ptr(), n0, count,
                              torch.cuda.current_stream().cuda_stream)
    _check(st, "bhw_sincos")
    return s, c


def cache_clear():
    _check(lib().bhw_cache_clear(), "bhw_cache_clear")


def atan2(x, y, input_width: int, angle_width: int, precision: int = 1, out=None, stream_quadrant: int = 0):
    """cordic_atan2 over two int32 CUDA tensors (VEC_DX, VEC_DY) -> PHI_DT (int32 CUDA tensor).
    stream_quadrant=1: as the entity streams it (pair t corrected with the quadrant of pair t+1)."""
    torch = _torch()
    assert x.is_cuda and y.is_cuda and x.dtype == torch.int32 and y.dtype == torch.int32 and x.shape == y.shape
    x, y = x.contiguous(), y.contiguous()
    if out is None:
        out = torch.empty_like(x)
    d = BhwAtan2Desc(input_width, angle_width, precision, stream_quadrant)
    with torch.cuda.device(x.device):
        st = lib().bhw_atan2(C.byref(d), x.data_ptr(), y.data_ptr(), out.data_ptr(), x.numel(),
                             torch.cuda.current_stream().cuda_stream)
    _check(st, "bhw_atan2")
    return out


def atan2_host(x: np.ndarray, y: np.ndarray, input_width: int, angle_width: int, precision: int = 1,
               stream_quadrant: int = 0) -> np.ndarray:
    """cordic_atan2 over two int32 numpy arrays (host buffers in, host buffer out)."""
    x = np.ascontiguousarray(x, dtype=np.int32)
    y = np.ascontiguousarray(y, dtype=np.int32)
    assert x.shape == y.shape
    out = np.empty(x.shape, np.int32)
    d = BhwAtan2Desc(input_width, angle_width, precision, stream_quadrant)
    _check(lib().bhw_atan2_host(C.byref(d), x.ctypes.data, y.ctypes.data, out.ctypes.data, x.size), "bhw_atan2_host")
    return out


def set_side_streams(n: int):
    _check(lib().bhw_set_side_streams(int(n)), "bhw_set_side_streams")


def set_table_cache(enabled: bool):
    _check(lib().bhw_set_table_cache(1 if enabled else 0), "bhw_set_table_cache")


KERNEL_TABLE_BUILD, KERNEL_SYNTH, KERNEL_DIRECT, KERNEL_SINCOS = 0, 1, 2, 3
KERNEL_NAMES = ("k_table_build", "k_synth", "k_direct_window", "k_sincos", "k_synth_bank", "k_atan2", "k_synth_group", "k_apply_mul")


def timing_enable(on: bool):
    _check(lib().bhw_timing_enable(1 if on else 0), "bhw_timing_enable")


def timing_reset():
    _check(lib().bhw_timing_reset(), "bhw_timing_reset")


def timing_read():
    """-> {kernel name: (launches, total device ms)} since the last reset."""
    out = {}
    for k, name in enumerate(KERNEL_NAMES):
        ms, n = C.c_double(0), C.c_uint64(0)
        _check(lib().bhw_timing_read(k, C.byref(ms), C.byref(n)), "bhw_timing_read")
        out[name] = (int(n.value), float(ms.value))
    return out


def timing_launches():
    """-> [{kernel, tag fields, bytes, ms}] for every launch recorded since the last reset, oldest first."""
    n = C.c_uint64(0)
    _check(lib().bhw_timing_launches(None, 0, C.byref(n)), "bhw_timing_launches")
    buf = (BhwLaunchRecord * max(1, n.value))()
    _check(lib().bhw_timing_launches(buf, n.value, C.byref(n)), "bhw_timing_launches")
    out = []
    for i in range(n.value):
        r = buf[i]
        t = r.tag
        out.append({"kernel": KERNEL_NAMES[r.kernel_class], "terms": t & 0xFF, "table": (t >> 8) & 0xFF,
                    "paired": (t >> 16) & 1, "spread": (t >> 17) & 1, "level": t >> 24, "bytes": int(r.bytes),
                    "ms": float(r.ms)})
    return out


def generate_repeat(d: BhwDesc, out, reps: int, n0: int = 0, count: Optional[int] = None, out_stride: int = 0,
                    out_slots: int = 0):
    """`reps` back-to-back bhw_generate calls issued from C (bhw_generate_repeat)."""
    torch = _torch()
    if count is None:
        count = (1 << d.phi_width) - n0
    _check(lib().bhw_generate_repeat(C.byref(d), out.data_ptr(), n0, count, int(reps), int(out_stride), int(out_slots),
                                     torch.cuda.current_stream().cuda_stream), "bhw_generate_repeat")


APPLY_EXACT, APPLY_ROUNDED = 0, 1


def apply(d: BhwDesc, x, mode: int = APPLY_EXACT, out=None):
    """y[f, n] = x[f, n] * w[n] (bhw_apply): x an int32 CUDA tensor of shape (frames, N) or (N,); returns
    int64 (APPLY_EXACT) or int32 (APPLY_ROUNDED) of the same shape."""
    torch = _torch()
    n = 1 << d.phi_width
    if x.dtype != torch.int32 or not x.is_cuda or not x.is_contiguous() or x.numel() % n:
        raise ValueError("x must be a contiguous int32 CUDA tensor of whole frames")
    frames = x.numel() // n
    if out is None:
        out = torch.empty(x.shape, dtype=torch.int64 if mode == APPLY_EXACT else torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        _check(lib().bhw_apply(C.byref(d), int(mode), x.data_ptr(), out.data_ptr(), frames,
                               torch.cuda.current_stream().cuda_stream), "bhw_apply")
    return out


def launch_count() -> int:
    return int(lib().bhw_launch_count())


# ---- the reference's own interfaces, by name ---------------------------------------------------
class WinSelector:
    """win_selector with its generics as constructor arguments and the AA ports as call arguments
    (src/win_selector.vhd:60-87).  ``stream(n)`` returns what DT_WIN carries for n enabled clocks.

    >>> w = WinSelector(PHI_WIDTH=16, DAT_WIDTH=17, WIN_TYPE="BH4TERM")
    >>> dt_win = w.stream(AA0=47022, AA1=64001, AA2=18518, AA3=1531)
    """

    def __init__(self, PHI_WIDTH: int = 10, DAT_WIDTH: int = 16, WIN_TYPE: str = "HAMMING",
                 SIN_TYPE: str = "CORDIC", LUT_SIZE: int = 9, XSERIES: str = "ULTRA",
                 dt_vld_order: bool = False, algo: int = ALGO_AUTO):
        if WIN_TYPE not in _WIN_TYPE_NAMES:
            raise BhwError(-2, f"WIN_TYPE {WIN_TYPE!r}")
        if SIN_TYPE not in _SIN_TYPE_NAMES:
            raise BhwError(-3, f"SIN_TYPE {SIN_TYPE!r}")
        self.xseries = XSERIES  # no numeric effect (src/mults/mlt35x27_dsp48e2.vhd:81-90)
        self._proto = make_desc(_WIN_TYPE_NAMES[WIN_TYPE], PHI_WIDTH, DAT_WIDTH,
                                sin_type=_SIN_TYPE_NAMES[SIN_TYPE],
                                lut_size=LUT_SIZE if SIN_TYPE == "TAYLOR" else 0,
                                stream_offset=1 if dt_vld_order else 0, algo=algo)

    def desc(self, AA0=0, AA1=0, AA2=0, AA3=0, AA4=0, AA5=0, AA6=0) -> BhwDesc:
        d = self._proto.copy(aa=[AA0, AA1, AA2, AA3, AA4, AA5, AA6])
        _check(validate(d), "win_selector")
        return d

    def stream(self, n: Optional[int] = None, host: bool = False, **aa):
        d = self.desc(**aa)
        if host:
            return generate_host(d, 0, n)
        return generate(d, 0, n)


def win_function(win_type: int, nphase: int, nwidth: int, i0: int = 0, count: Optional[int] = None):
    """Vector form of the HLS ``win_function(win_type, i, &out)`` for i = i0 .. i0+count-1
    (hls/windows/win_function.cpp:380-422; win_type codes 1,2,3,4,5,7 as there; other codes give
    0 like the reference's default branch)."""
    torch = _torch()
    variant = {1: 1, 2: 2, 3: 3, 4: 6, 5: 9, 7: 10}.get(int(win_type))
    if count is None:
        count = (1 << nphase) - i0
    if variant is None:
        return torch.zeros(count, dtype=torch.int32, device="cuda")
    d = variant_desc(variant, nphase, nwidth, model=MODEL_HLS)
    return generate(d, i0, count)
