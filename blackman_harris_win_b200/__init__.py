"""blackman_harris_win_b200 - B200-native generator of fixed-point window-coefficient tables.

The product is the C-ABI shared library ``libbhw.so`` (sources in ``csrc/``, contract in
``include/bhw.h``).  This package is the thin Python host binding used by the tests and the
benchmark: ctypes over the C ABI, torch only for device memory and streams.  There is no CPU
evaluation path here - if the CUDA library is missing every generating call raises.

Reference interfaces mirrored (paths relative to the reference checkout):
  * ``win_selector`` generics/ports          src/win_selector.vhd:60-87   -> :class:`WinSelector`
  * ``win_function(win_type, i, &out)``      hls/windows/win_function.h:65-69 -> :func:`win_function`
  * ``cordic(phi, &cos, &sin)``              hls/cordic/cordic.h:58-62, cpp/cordic_sincos.cpp:10 -> :func:`sincos`
  * ``cordic_atan2`` generics/ports          src/cordic_atan2.vhd:64-78   -> :func:`atan2`
"""
from .api import (  # noqa: F401
    ALGO_AUTO, ALGO_DIRECT, ALGO_TABLE, OUT_DEFAULT, OUT_INT16, MODEL_CPP, MODEL_HLS, MODEL_RTL, RULE_HLS, RULE_TB,
    SIN_CORDIC, SIN_CORDIC48, SIN_CORDIC_SCALED, SIN_TAYLOR, VARIANT_NAMES, BhwAtan2Desc, BhwDesc, BhwError, Plan,
    WinSelector, atan2, atan2_host, batch_total, cache_clear, desc_array, elem_bytes, generate, generate_batch,
    generate_batch_host, generate_host, launch_count, lib, lib_path, make_desc, quantize,
    set_side_streams, set_table_cache, shard_range, shard_range_cost, shard_windows, sincos, strerror, timing_enable, timing_read, timing_reset,
    timing_launches, generate_repeat, BhwLaunchRecord, apply, APPLY_EXACT, APPLY_ROUNDED,
    validate, variant_coeffs, variant_desc, win_function,
)

__version__ = "0.1.0"
