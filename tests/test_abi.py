"""CPU tier: the C-ABI library loads, exports every symbol include/bhw.h declares, and its
GPU-free entry points (validation, quantisation, sharding arithmetic) behave.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import blackman_harris_win_b200 as bhw
from blackman_harris_win_b200 import api
import harness as H

HDR = os.path.join(H.ROOT, "include", "bhw.h")


def declared_symbols():
    txt = open(HDR).read()
    return sorted(set(re.findall(r"BHW_API\s+[\w\s\*]+?\b(bhw_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = C.CDLL(bhw.lib_path())
    decl = declared_symbols()
    assert len(decl) >= 19
    for name in decl:
        assert hasattr(L, name), f"{name} declared in include/bhw.h but not exported"
    assert sorted(api.ABI_SYMBOLS) == decl


def test_struct_layout_matches_header():
    assert C.sizeof(bhw.BhwDesc) == 10 * 4 + 11 * 8 and api.MAX_TERMS == 11      # BHW_MAX_TERMS
    assert bhw.BhwDesc.aa.offset == 40


def test_version_and_strerror():
    assert bhw.lib().bhw_version() == 0x000100
    assert bhw.strerror(0) == "ok"
    for code in range(-16, 0):
        assert bhw.strerror(code) not in ("", "unknown status")
    assert bhw.strerror(-99) == "unknown status"


def test_quantize_matches_reference_rules():
    # src/tb/tb_windows.vhd:75-127 worked examples
    assert bhw.quantize(1, bhw.RULE_TB, 16) == ([17808, 14959] + [0] * 9, 2)
    assert bhw.quantize(2, bhw.RULE_TB, 16) == ([16384, 16384] + [0] * 9, 2)   # 16383.5 ties up
    assert bhw.quantize(6, bhw.RULE_TB, 17)[0][:4] == [47022, 64001, 18518, 1531]
    assert bhw.quantize(3, bhw.RULE_TB, 24)[0][:3] == [7046424, 8388600, 1342176]
    assert bhw.quantize(10, bhw.RULE_TB, 32)[0] == [582441289, 930815217, 468160289, 141272949, 23110934, 1653590, 29379, 0, 0, 0, 0]
    assert bhw.quantize(8, bhw.RULE_TB, 16)[0][:5] == [16383, 31619, 21134, 6357, 491]
    for v in range(1, 11):
        for rule in (0, 1):
            for dw in (8, 16, 17, 24, 32, 40, 48):
                assert bhw.quantize(v, rule, dw) == H.orc_quantize(v, rule, dw)
    with pytest.raises(bhw.BhwError):
        bhw.quantize(19, 0, 16)
    with pytest.raises(bhw.BhwError):
        bhw.quantize(1, 2, 16)
    assert bhw.variant_coeffs(3, bhw.RULE_HLS) == [0.21, 0.25, 0.04]


def test_validate_rejections():
    ok = bhw.make_desc(4, 16, 17, [47022, 64001, 18518, 1531])
    assert bhw.validate(ok) == 0
    assert bhw.validate(ok.copy(win_type=12)) == -2 and bhw.validate(ok.copy(win_type=1)) == -2   # 6 and 8..11: BHW_WIN_MTERM_*
    assert bhw.validate(ok.copy(sin_type=bhw.SIN_TAYLOR)) == -3          # 4-term has no TAYLOR
    assert bhw.validate(ok.copy(sin_type=9)) == -3
    assert bhw.validate(ok.copy(model=7)) == -4
    assert bhw.validate(ok.copy(model=bhw.MODEL_CPP)) == -4              # cpp/ has no window model
    assert bhw.validate(ok.copy(phi_width=3)) == -5
    assert bhw.validate(ok.copy(dat_width=49)) == -6
    assert bhw.validate(ok.copy(precision=2)) == -7
    assert bhw.validate(ok.copy(aa=[1 << 17, 0, 0, 0])) == -9
    assert bhw.validate(ok.copy(stream_offset=2)) == -16
    assert bhw.validate(ok.copy(out_format=2)) == -16
    assert bhw.validate(ok.copy(out_format=bhw.OUT_INT16)) == -6             # DAT_WIDTH 17 does not fit int16
    ok16 = bhw.make_desc(2, 10, 16, [17808, 14959], out_format=bhw.OUT_INT16)
    assert bhw.validate(ok16) == 0 and bhw.elem_bytes(ok16) == 2 and bhw.elem_bytes(ok) == 4
    t3 = bhw.make_desc(3, 12, 16, [1, 1, 1], sin_type=bhw.SIN_TAYLOR, lut_size=9)
    assert bhw.validate(t3) == -8                                         # PHI_WIDTH - LUT_SIZE == 3
    assert bhw.validate(t3.copy(phi_width=13)) == 0
    assert bhw.validate(bhw.make_desc(2, 26, 16, [1, 1], sin_type=bhw.SIN_TAYLOR, lut_size=7)) == -8  # STAGE > 15
    assert bhw.validate(ok.copy(phi_width=27)) == -5                      # 64M points is the reference's maximum
    assert bhw.validate(bhw.make_desc(2, 10, 33, [1, 1], sin_type=bhw.SIN_TAYLOR)) == -6
    assert bhw.validate(bhw.make_desc(2, 20, 16, [1, 1], model=bhw.MODEL_HLS)) == -5  # NP > NW+2
    assert bhw.lib().bhw_validate(None) == -1


def test_alternative_coefficient_sets():
    """Variants 11-13: the second sets the reference prints (README.md:45-51, src/hamming_win.vhd:21-23,
    src/bh_win_5term.vhd:28-33), quantised by the same rules; library and oracle agree."""
    import harness as H
    assert bhw.quantize(11, bhw.RULE_TB, 32)[0][:7] == [round(a * (2 ** 31 - 1)) for a in (
        0.27105140069342, 0.43329793923448, 0.21812299954311, 0.06592544638803, 0.01081174209837, 0.00077658482522,
        0.00001388721735)]
    assert bhw.quantize(12, bhw.RULE_TB, 16) == ([round(0.5383554 * 32767), round(0.4616446 * 32767)] + [0] * 9, 2)
    assert bhw.quantize(13, bhw.RULE_TB, 24)[1] == 5
    for v in (11, 12, 13):
        for rule in (bhw.RULE_TB, bhw.RULE_HLS):
            for dw in (12, 16, 24, 32):
                assert bhw.quantize(v, rule, dw) == H.orc_quantize(v, rule, dw)
    with pytest.raises(bhw.BhwError):
        bhw.quantize(19, bhw.RULE_TB, 16)


def test_apply_argument_checks_need_no_gpu():
    L = bhw.lib()
    d = bhw.make_desc(4, 16, 17, [47022, 64001, 18518, 1531])
    assert L.bhw_apply(None, 0, None, None, 1, None) == -1
    assert L.bhw_apply(C.byref(d.copy(dat_width=40, aa=[1, 1, 1, 1])), 0, None, None, 1, None) == -6
    assert L.bhw_apply(C.byref(d), 2, None, None, 1, None) == -16
    assert L.bhw_apply(C.byref(d), 0, None, None, 0, None) == 0          # no frames: nothing to do
    assert L.bhw_apply(C.byref(d), 0, None, None, 1, None) == -1         # missing buffers


def test_oracle_apply_step_matches_numpy():
    """orc_apply against a direct numpy restatement of the product, the slice and the rounding."""
    import harness as H
    rng = np.random.default_rng(3)
    for d in (bhw.variant_desc(1, 8, 16), bhw.variant_desc(10, 7, 32), bhw.variant_desc(6, 9, 17, model=bhw.MODEL_HLS)):
        n, dw = 1 << d.phi_width, d.dat_width
        x = rng.integers(-(1 << 31), 1 << 31, size=(2, n), dtype=np.int64).astype(np.int32)
        w = H.orc_window(d)
        a = ((x.astype(np.int64) << (64 - dw)) >> (64 - dw))
        q = a * w[None, :]
        assert np.array_equal(H.orc_apply(d, x, 0), q)
        r = ((q >> (dw - 2)) << (63 - dw)) >> (63 - dw)
        b = (r >> 1) + (r & 1)
        assert np.array_equal(H.orc_apply(d, x, 1), (b << (64 - dw)) >> (64 - dw))


def test_shard_ranges_partition_the_flat_range():
    for total in (0, 1, 3, 4, 5, 1023, 1024, 134217712, (1 << 26) * 10 + 7):
        for r in (1, 2, 3, 4, 8):
            pos = 0
            sizes = []
            for k in range(r):
                b, c = bhw.shard_range(total, k, r)
                assert b == pos
                if k < r - 1:
                    assert b % 4 == 0 and c % 4 == 0     # 128-bit store alignment
                pos += c
                sizes.append(c)
            assert pos == total
            if total >= 4 * r:
                assert max(sizes) - min(sizes) <= 4 + total % 4
    with pytest.raises(bhw.BhwError):
        bhw.shard_range(10, 2, 2)


def test_batch_total():
    ds = [bhw.make_desc(2, pw, 16, [1, 1]) for pw in range(4, 27)]
    assert bhw.batch_total(ds) == 134217712            # sum 2^4..2^26
    with pytest.raises(bhw.BhwError):
        bhw.batch_total([bhw.make_desc(2, 31, 16, [1, 1])])


def test_generate_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(bhw.BhwError):
        bhw.generate(bhw.make_desc(2, 10, 16, [17808, 14959]))
    # the C entry point itself reports the missing device instead of computing on the CPU
    import numpy as np
    out = np.full(1024, 12345, np.int32)
    st = bhw.lib().bhw_generate_host(C.byref(bhw.make_desc(2, 10, 16, [17808, 14959])), out.ctypes.data, 0, 1024)
    assert st in (-12, -13) and (out == 12345).all()


def test_mterm_extension_front_end():
    """BHW_WIN_MTERM_* (6, 8..11 terms): the coefficient sets of doc/blackman-harris coef.jpg as quantize variants
    14..18 (each sums to 1 and nearly cancels at n = 0, which any transcription slip would break), quantised by
    the 7-term testbench rule; library and oracle agree; legal for the RTL model with a CORDIC source only."""
    import harness as H
    for v, m in zip(range(14, 19), (6, 8, 9, 10, 11)):
        a = bhw.variant_coeffs(v, bhw.RULE_TB)
        assert len(a) == m and abs(sum(a) - 1.0) < 1e-14 and 0 < sum((-1) ** k * x for k, x in enumerate(a)) < 2e-6
        assert all(a[k] > a[k + 1] for k in range(1, m - 1))
        for dw in (12, 16, 24, 32, 40):
            aa, wt = bhw.quantize(v, bhw.RULE_TB, dw)
            assert wt == m and (list(aa), wt) == H.orc_quantize(v, bhw.RULE_TB, dw)
            assert list(aa[:m]) == [round(x * ((1 << (dw - 1)) - 1)) for x in a] and not any(aa[m:])
        with pytest.raises(bhw.BhwError):
            bhw.quantize(v, bhw.RULE_HLS, 16)
        d = bhw.variant_desc(v, 10, 24)
        assert bhw.validate(d) == 0 and H.orc_window_status(d) == 0
        for bad in (d.copy(model=bhw.MODEL_HLS), d.copy(sin_type=bhw.SIN_TAYLOR, lut_size=5)):
            assert bhw.validate(bad) == -2 and H.orc_window_status(bad) == -2
        assert bhw.validate(d.copy(sin_type=bhw.SIN_CORDIC48)) == 0 and bhw.validate(d.copy(sin_type=bhw.SIN_CORDIC_SCALED)) == 0
    assert bhw.validate(bhw.make_desc(12, 10, 16, [1] * 11)) == -2 and bhw.validate(bhw.make_desc(1, 10, 16, [1])) == -2
    # the oracle window at the two points the coefficient identities speak about.  The CORDIC amplitude is 2^(DW-2), so
    # the harmonics enter at half weight, and DT_WIN is dsp_pp[DW+1:2], a quarter of the sum (src/bh_win_3term.vhd:
    # 295-306): w[N/2] ~ (a0 + (1 - a0)/2) / 4, w[0] ~ (a0 - (a1 - a2 + ...)/2) / 4
    for v in range(14, 19):
        a = bhw.variant_coeffs(v, bhw.RULE_TB)
        d = bhw.variant_desc(v, 12, 24)
        w = H.orc_window(d)
        S = (1 << 23) - 1
        assert abs(int(w[2048]) - S * (a[0] + (1 - a[0]) / 2) / 4) < 64
        assert abs(int(w[0]) - S * (a[0] - sum((-1) ** (k + 1) * a[k] for k in range(1, len(a))) / 2) / 4) < 64
        assert int(w.argmax()) == 2048


def test_one_container_per_batch():
    """BHW_OUT_INT16 (bhw_desc.out_format): a batch has one element size - int16, int32 or int64 - and says so
    before it looks for a device (BHW_E_ELEM); int16 needs DAT_WIDTH <= 16; the apply step has its own containers."""
    import numpy as np
    d16 = bhw.make_desc(2, 10, 16, [17808, 14959], out_format=bhw.OUT_INT16)
    d32 = bhw.make_desc(2, 10, 16, [17808, 14959])
    d64 = bhw.variant_desc(10, 10, 40)
    out = np.zeros(4096, np.int64)
    for mix in ([d16, d32], [d32, d64], [d16, d64], [d32, d16, d32]):
        arr = bhw.desc_array(mix)
        assert bhw.lib().bhw_generate_batch_host(arr, len(mix), 0, 2048, out.ctypes.data) == -11, mix
        plan = C.c_void_p()
        assert bhw.lib().bhw_plan_create(arr, len(mix), C.byref(plan)) == -11
    assert bhw.lib().bhw_generate_batch_host(bhw.desc_array([d32.copy(dat_width=17, out_format=1)]), 1, 0, 1024, out.ctypes.data) == -6
    assert bhw.lib().bhw_apply(C.byref(d16), 0, None, None, 0, None) == -16
    assert bhw.lib().bhw_sincos(C.byref(d16), None, None, 0, 0, None) == -16
    assert [bhw.elem_bytes(d) for d in (d16, d32, d64)] == [2, 4, 8]
    assert H.orc_window_status(d16) == 0 and H.orc_window_status(d32.copy(dat_width=17, out_format=1)) == -6


def test_win_selector_mirror():
    w = bhw.WinSelector(PHI_WIDTH=16, DAT_WIDTH=17, WIN_TYPE="BH4TERM")
    d = w.desc(AA0=47022, AA1=64001, AA2=18518, AA3=1531)
    assert (d.win_type, d.phi_width, d.dat_width, d.sin_type) == (4, 16, 17, 0)
    with pytest.raises(bhw.BhwError):
        bhw.WinSelector(WIN_TYPE="BH6TERM")
    with pytest.raises(bhw.BhwError):
        bhw.WinSelector(WIN_TYPE="BH4TERM", SIN_TYPE="TAYLOR").desc(AA0=1)


def test_cost_balanced_shards_tile_the_batch():
    """bhw_shard_range_cost: contiguous slices in rank order that tile [0, total), cut at multiples of
    4 samples; equal windows -> (nearly) equal slices; costly windows get shorter slices."""
    import cases
    sweep = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 27)]
    same = [bhw.variant_desc(6, 16, 17) for _ in range(64)]
    for descs in (sweep, same, [bhw.variant_desc(10, 20, 32)]):
        total = bhw.batch_total(descs)
        for world in (1, 2, 3, 8):
            cursor = 0
            for r in range(world):
                b, c = bhw.shard_range_cost(descs, r, world)
                assert b == cursor and b % 4 == 0
                cursor = b + c
            assert cursor == total
    total = bhw.batch_total(same)
    for r in range(8):
        b, c = bhw.shard_range_cost(same, r, 8)
        assert abs(c - total // 8) <= 4
    # the sweep ends with the 7-term 32-bit windows (costly samples, and a 190 us table build for whoever touches the
    # 2^26-point one): the first rank takes far more samples than the ranks that share those windows
    total = bhw.batch_total(sweep)
    counts = [bhw.shard_range_cost(sweep, r, 8)[1] for r in range(8)]
    assert counts[0] > 2 * total // 8 and max(counts[4:]) < total // 16
    # cuts inside a window only for long windows, at multiples of 2^14 samples, never closer than an eighth of the
    # window to one of its ends
    offs = np.cumsum([0] + [1 << d.phi_width for d in sweep])
    for world in (2, 3, 4, 8):
        for r in range(1, world):
            b, _ = bhw.shard_range_cost(sweep, r, world)
            w = int(np.searchsorted(offs, b, side="right") - 1)
            inside = b - int(offs[w])
            if inside and b < total:
                n = 1 << sweep[w].phi_width
                assert n >= 1 << 20 and inside % (1 << 14) == 0 and n // 8 <= inside <= n - n // 8, (world, r, w, inside)
    with pytest.raises(bhw.BhwError):
        bhw.shard_range_cost(sweep, 8, 8)
