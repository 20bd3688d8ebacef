"""CPU tier: the kernels' per-thread bodies (csrc/bhw_device.cuh, compiled for the host by
tests/hostcheck) against the oracle.  This is what can be said about kernel arithmetic without a
GPU; the GPU tier (test_gpu_parity.py) repeats it through the C ABI."""
import ctypes as C

import numpy as np
import pytest

import blackman_harris_win_b200 as bhw
import cases
import harness as H


def hc_window(d, n0, cnt, which, force_generic_core=0):
    out = np.empty(cnt, np.int64)
    hc = H.hostcheck()
    if which == "direct":
        st = hc.hc_direct(C.byref(d), n0, cnt, out.ctypes.data_as(H.I64P))
    else:
        st = hc.hc_table(C.byref(d), n0, cnt, out.ctypes.data_as(H.I64P), force_generic_core)
    return st, out


DIRECT32_SEEN = []
DIRECT_TAYLOR_SEEN = []


def check(d, cnt_cap=2048):
    n = 1 << d.phi_width
    cnt = min(n, cnt_cap)
    n0 = 0 if cnt == n else (n // 2 - cnt // 2)
    want = H.orc_window(d, n0, cnt)
    st, got = hc_window(d, n0, cnt, "direct")
    assert st == 0 and np.array_equal(got, want), ("direct", d)
    got32 = np.empty(cnt, np.int64)
    st = H.hostcheck().hc_direct32(C.byref(d), n0, cnt, got32.ctypes.data_as(H.I64P))
    assert st in (0, 1), d
    if st == 0:
        DIRECT32_SEEN.append(1)
        assert np.array_equal(got32, want), ("direct32", d)
    got_t = np.empty(cnt, np.int64)
    st = H.hostcheck().hc_direct_taylor(C.byref(d), n0, cnt, got_t.ctypes.data_as(H.I64P))
    assert st in (0, 1), d
    if st == 0:
        DIRECT_TAYLOR_SEEN.append(1)
        assert np.array_equal(got_t, want), ("direct_taylor", d)
    if d.dat_width <= 32:
        for force in (0, 1):
            st, got = hc_window(d, n0, cnt, "table", force)
            assert st in (0, 1), d
            if st == 0:
                assert np.array_equal(got, want), ("table", force, d)
        if 7 <= d.phi_width <= 14:   # the bank kernel body, planner's own choice of placement
            full = np.full(n, -(1 << 62), np.int64)
            st = H.hostcheck().hc_bank(C.byref(d), full.ctypes.data_as(H.I64P), 192 * 1024, -1, -1)
            assert st in (0, 1), d
            if st == 0:
                assert np.array_equal(full[n0:n0 + cnt], want), ("bank", d)


def test_rtl_sweep_all_variants_widths_sources():
    descs = cases.rtl_sweep()
    assert len(descs) > 800
    for d in descs:
        check(d)
    assert len(DIRECT32_SEEN) > 100      # the 32-bit direct body covered its share of the sweep
    assert len(DIRECT_TAYLOR_SEEN) > 50  # and so did the TAYLOR one


def test_mterm_extension_bodies():
    """6 and 8..11 terms (BHW_WIN_MTERM_*): the direct, 32-bit direct, table and generic bodies against the oracle;
    the bank / group placements decline these term counts (they stay with the general and direct kernels)."""
    descs = cases.mterm_sweep()
    assert len(descs) > 500
    n32 = len(DIRECT32_SEEN)
    for d in descs:
        check(d, cnt_cap=1024)
        full = np.zeros(1 << d.phi_width, np.int64)
        if d.phi_width >= 8 and d.dat_width <= 32:
            assert H.hostcheck().hc_bank(C.byref(d), full.ctypes.data_as(H.I64P), 192 * 1024, -1, -1) == 1, d
    assert len(DIRECT32_SEEN) - n32 > 100


def test_taylor_quarter_window_body():
    """direct_taylor_quad4 (k_direct_taylor's long-window branch: 16 samples from two ROM words) against the
    per-sample body and the oracle over whole windows - every coefficient sign, DAT_WIDTH 19..32, both entities,
    ROM depths from 4 to PHI_WIDTH-4, and the shapes that must NOT take it (DSP datapath, DT_VLD order, short)."""
    hc = H.hostcheck()
    rng = np.random.default_rng(5)
    taken = 0
    shapes = [(3, 14, 24, 9, 0), (4, 16, 24, 9, 0), (1, 14, 24, 9, 0), (3, 12, 19, 5, 0), (4, 15, 31, 10, 0), (4, 18, 24, 12, 0),
              (1, 16, 32, 9, 0), (3, 13, 16, 5, 0), (3, 14, 24, 9, 1), (1, 11, 24, 5, 0)]
    descs = [(bhw.variant_desc(v, pw, dw, sin_type=bhw.SIN_TAYLOR, lut_size=lut).copy(stream_offset=so), None) for v, pw, dw, lut, so in shapes]
    for it in range(40):
        m, pw, dw = int(rng.integers(2, 4)), int(rng.integers(12, 16)), int(rng.integers(19, 33))
        lut = int(rng.integers(4, pw - 4 + 1))
        lim = 1 << (dw - 1)
        aa = [lim - 1] * m if it % 5 == 0 else [-lim] * m if it % 7 == 0 else [int(x) for x in rng.integers(-lim, lim, m)]
        descs.append((bhw.make_desc(m, pw, dw, aa, sin_type=bhw.SIN_TAYLOR, lut_size=min(lut, 12)), None))
    for d, _ in descs:
        if bhw.validate(d) != 0:
            continue
        n = 1 << d.phi_width
        got = np.empty(n, np.int64)
        st = hc.hc_direct_taylor(C.byref(d), 0, n, got.ctypes.data_as(H.I64P))     # -102: quad body differs
        assert st in (0, 1), (st, d)
        q = hc.hc_taylor_quad_ok(C.byref(d))
        assert q == int(st == 0 and d.dat_width >= 19 and d.stream_offset == 0 and d.phi_width >= 12 and d.phi_width - d.lut_size - 2 - (d.win_type == 3) >= 2), d
        taken += q
        if st == 0:
            assert np.array_equal(got, H.orc_window(d)), d
    assert taken >= 30


def test_validation_agrees_with_oracle():
    n = 0
    for v in range(1, 11):
        for pw in (3, 4, 12, 26, 30, 31):
            for dw in (3, 4, 8, 18, 19, 32, 33, 47, 48, 49):
                for st in range(0, 5):
                    for model in (0, 1, 2, 3):
                        for lut in (0, 1, 9, 12, 16, 17):
                            try:
                                aa, wt = bhw.quantize(v, 0, min(max(dw, 4), 48))
                            except bhw.BhwError:
                                continue
                            d = bhw.make_desc(wt, pw, dw, aa, sin_type=st, model=model, lut_size=lut)
                            a, b = bhw.validate(d), H.orc_window_status(d)
                            assert (a == 0) == (b == 0), (d, a, b)
                            n += 1
    assert n > 5000


def test_hls_model_bodies():
    for (np_, nw) in [(4, 8), (6, 6), (8, 32), (10, 16), (10, 24), (12, 12), (14, 12), (16, 17), (18, 16), (13, 30), (13, 31)]:
        for t, v in cases.HLS_TYPES.items():
            check(bhw.variant_desc(v, np_, nw, model=bhw.MODEL_HLS))


def test_coefficient_edge_cases():
    for d in cases.edge_coeff_descs():
        check(d, 512)
    # the same through the other sources
    for st in (bhw.SIN_CORDIC48, bhw.SIN_CORDIC_SCALED, bhw.SIN_TAYLOR):
        for d in cases.edge_coeff_descs()[::7]:
            d2 = d.copy(sin_type=st)
            if bhw.validate(d2) == 0:
                check(d2, 512)


def test_wide_windows_take_the_32_bit_tail_when_the_sum_provably_fits():
    """DAT_WIDTH 31..32: dsp_pp has more bits than a register, but with real coefficient sets the sum
    stays inside an int32 and the planner keeps the 32-bit tail; full-scale ports, |AAk| >= 2^30 at
    DAT_WIDTH 32 and AAk = -2^(DW-1) do not qualify.  (Bit-exactness of both routes: the sweeps.)"""
    hc = H.hostcheck()
    T32, T64, GEN = 0, 1, 2
    for v in range(1, 11):
        for dw in (31, 32):
            for st in (bhw.SIN_CORDIC, bhw.SIN_CORDIC48, bhw.SIN_CORDIC_SCALED):
                d = bhw.variant_desc(v, 12, dw, sin_type=st)
                if bhw.validate(d) == 0:
                    # 32 bits: the 3-/4-term rules scale to 2^W (AA1 ~ 2^31), Hann has AA1 = 2^30 exactly and
                    # the flat-top coefficients sum to 4.6 * 2^(W-2); 31 bits: 2*AAk always fits, sums are small
                    want = T32 if dw == 31 or v in (1, 9, 10) else T64
                    assert hc.hc_tail_mode(C.byref(d)) == want, (v, dw, st)
    q = 1 << 30
    hi = (1 << 31) - 1
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(4, 12, 32, [hi] * 4))) == T64
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(4, 12, 32, [0, q, 0, 0]))) == T64
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(4, 12, 32, [0, q - 1, 0, 0]))) == T32
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(4, 12, 32, [0, -q, 0, 0]))) == T64
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(4, 12, 32, [0, -q + 1, 0, 0]))) == T32
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(4, 12, 32, [hi, q - 1, q - 1, q - 1]))) == T64   # sum too large
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(4, 12, 32, [0, -(1 << 31), 0, 0]))) == GEN
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(2, 12, 31, [q - 1, q - 1]))) == T32
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(4, 12, 30, [(1 << 29) - 1] * 4))) == T32          # always, DW <= 30
    assert hc.hc_tail_mode(C.byref(bhw.make_desc(4, 12, 40, [1, 2, 3, 4]))) == GEN


def test_bank_kernel_tile_walks_visit_every_tile_once():
    """The spread walk (one long window, warps of a CTA spread over the window) and the window-minor walk
    (a bank over one L2-resident table) are index permutations: replayed on the CPU with the kernel's own
    helpers, every tile of every window must come up exactly once, for grids that do and do not divide the work."""
    hc = H.hostcheck()
    for U in (4440, 4441, 8192, 16384, 32768, 100003):
        for G in (30, 24, 15, 32, 7):
            for grid in (148, 132, 1, 37):
                assert hc.hc_walk(0, U, G, grid) == 0, (U, G, grid)
    for U in (1, 2, 64, 1000):
        for nwin in (2, 3, 5, 33, 64, 257):
            for grid in (148, 1, 37, 7):
                assert hc.hc_walk(1, U, nwin, grid) == 0, (U, nwin, grid)


def test_random_descriptors():
    """Seeded fuzz over entity x widths x source x model x ports (cases.random_descs): every kernel
    body that accepts the descriptor must reproduce the oracle."""
    for d in cases.random_descs(400, seed=20260101):
        check(d, 1024)


def test_stream_offset_is_a_rotation():
    for d in (bhw.make_desc(2, 8, 16, [17808, 14959]), bhw.make_desc(3, 14, 24, [7046424, 8388600, 1342176],
                                                                  sin_type=bhw.SIN_TAYLOR)):
        base = H.orc_window(d)
        d1 = d.copy(stream_offset=1)
        assert np.array_equal(H.orc_window(d1), np.roll(base, -1))
        for which in ("direct", "table"):
            st, got = hc_window(d1, 0, len(base), which)
            assert st == 0 and np.array_equal(got, np.roll(base, -1))


def test_taylor_modes():
    # LESS (pw-lut<2), EQ (=2), DSP (dw<19) and WIDE (dw>18) branches, LUT 1..16
    for pw, lut in [(6, 6), (6, 5), (8, 6), (10, 7), (14, 9), (16, 9), (20, 9), (24, 9), (26, 10), (12, 1), (18, 16)]:
        for dw in (8, 16, 18, 19, 24, 32):
            for wt, aa in ((2, [100, 77]), (3, [90, 100, 17])):
                amp = (1 << (dw - 1)) - 1
                d = bhw.make_desc(wt, pw, dw, [a * amp // 128 for a in aa], sin_type=bhw.SIN_TAYLOR, lut_size=lut)
                if bhw.validate(d) == 0:
                    check(d, 1024)


def test_sincos_bodies_all_sources():
    hc = H.hostcheck()
    for model, st, pw, dw, prec in [(0, 0, 10, 16, 0), (0, 0, 12, 8, 3), (0, 0, 20, 32, 1), (0, 0, 9, 47, 2),
                                    (0, 2, 10, 16, 0), (0, 2, 14, 47, 0), (0, 3, 10, 8, 0), (0, 3, 16, 32, 0),
                                    (0, 3, 26, 12, 0), (0, 1, 14, 16, 0), (0, 1, 16, 24, 0), (1, 0, 10, 16, 0),
                                    (1, 0, 18, 16, 0), (2, 0, 14, 12, 0), (2, 0, 10, 32, 0)]:
        d = bhw.make_desc(2, pw, dw, sin_type=st, model=model, precision=prec)
        cnt = min(1 << pw, 4096)
        s, c = H.orc_sincos(d, 0, cnt)
        gs, gc = np.empty(cnt, np.int64), np.empty(cnt, np.int64)
        assert hc.hc_sincos(C.byref(d), 0, cnt, gs.ctypes.data_as(H.I64P), gc.ctypes.data_as(H.I64P)) == 0
        assert np.array_equal(gs, s) and np.array_equal(gc, c), d


def test_table_equals_harmonic_gather():
    """cos_k[n] = C[(k*n) mod N]: the memoised table is the k=1 source sequence."""
    hc = H.hostcheck()
    for d in (bhw.make_desc(7, 12, 16, [1] * 7), bhw.make_desc(4, 18, 16, [1] * 4),
              bhw.make_desc(5, 12, 24, [1] * 5, sin_type=bhw.SIN_CORDIC_SCALED)):
        n = 1 << d.phi_width
        tab = np.empty(n, np.int64)
        for force in (0, 1):
            assert hc.hc_table_cos(C.byref(d), tab.ctypes.data_as(H.I64P), force) == 0
            _, c = H.orc_sincos(d)
            assert np.array_equal(tab, c)


def hc_bank(d, limit=192 * 1024, mode=-1, pair=-1):
    n = 1 << d.phi_width
    out = np.full(n, -(1 << 62), np.int64)
    st = H.hostcheck().hc_bank(C.byref(d), out.ctypes.data_as(H.I64P), limit, mode, pair)
    return st, out


def test_bank_body_all_modes():
    """k_synth_bank's lane/tile body in every table placement (staged full period, staged half
    period with the sign folded into the coefficient, global) with and without (n, n+N/2) pairing."""
    TAB_FULL, TAB_HALF, TAB_GLOBAL = 0, 1, 2
    seen = set()
    descs = []
    H.hostcheck().hc_lin_tiles(1)
    for v in range(1, 11):
        for pw in (8, 9, 10, 12, 14):
            for dw in sorted({cases.VARIANT_DW[v], 12, 16, 24, 30, 31, 32}):
                for st in (bhw.SIN_CORDIC, bhw.SIN_CORDIC48, bhw.SIN_CORDIC_SCALED, bhw.SIN_TAYLOR):
                    d = bhw.variant_desc(v, pw, dw, sin_type=st)
                    if bhw.validate(d) == 0:
                        descs.append(d)
        for (np_, nw) in ((8, 16), (10, 24), (12, 12), (14, 17), (13, 30)):
            if v in cases.HLS_TYPES.values():
                descs.append(bhw.variant_desc(v, np_, nw, model=bhw.MODEL_HLS))
    descs.append(bhw.make_desc(4, 16, 17, [47022, 64001, 18518, 1531]))                 # cfg 2
    descs.append(bhw.make_desc(4, 16, 17, [47022, 64001, 18518, 1531], stream_offset=1))
    descs.append(bhw.make_desc(3, 14, 16, [27518, 32760, 5242], sin_type=bhw.SIN_TAYLOR, stream_offset=1))
    descs.append(bhw.make_desc(7, 12, 16, [-8887, 14203, -7143, 2156, -353, 25, -1]))
    for d in descs:
        want = H.orc_window(d)
        for mode in (-1, TAB_FULL, TAB_HALF, TAB_GLOBAL):
            for pair in (-1, 0, 1, 15, 16, 17):      # +16: every tile on the non-linear path
                st, got = hc_bank(d, mode=mode, pair=pair)
                assert st in (0, 1), d
                if st == 0:
                    seen.add((mode, pair if pair < 15 else pair - 16))
                    assert np.array_equal(got, want), (mode, pair, d)
    assert H.hostcheck().hc_lin_tiles(0) > 10000     # the linear-tile path did run
    # every legal combination was exercised
    for combo in ((-1, -1), (TAB_FULL, 0), (TAB_FULL, 1), (TAB_HALF, 1), (TAB_GLOBAL, 0), (TAB_GLOBAL, 1)):
        assert combo in seen, combo


def test_bank_body_half_table_sign_boundaries():
    """Staged half period: tiles that straddle a half-period boundary of some harmonic take the
    per-lane sign path; all tiles of a long 7-term window must match."""
    d = bhw.variant_desc(10, 16, 16)
    st, got = hc_bank(d, limit=128 * 1024)
    assert st == 0 and np.array_equal(got, H.orc_window(d))
    d = bhw.variant_desc(9, 17, 24).copy(stream_offset=1)
    st, got = hc_bank(d, mode=1)
    assert st == 0 and np.array_equal(got, H.orc_window(d))


def test_bank_body_pairs_the_input_quadrant_cordics():
    """cordic_dds48 / cordic_dds_scaled: the bank body pairs samples through T[i + E/2] == ~T[i] and a patch pass
    recomputes the pairs that read one of the few entries where that fails (src/cordic_dds48.vhd:170-258) - every
    sample must match the oracle, with and without exceptions present, and the relation must really have exceptions
    somewhere (otherwise the patch pass is untested)."""
    hc = H.hostcheck()
    hc.hc_inq_exceptions(1)
    n = 0
    for st in (bhw.SIN_CORDIC48, bhw.SIN_CORDIC_SCALED):
        for v in range(1, 11):
            for pw in (9, 10, 12, 14):
                for dw in sorted({cases.VARIANT_DW[v], 12, 16, 20, 24, 30}):
                    d = bhw.variant_desc(v, pw, dw, sin_type=st).copy(stream_offset=(v + pw) & 1)
                    if bhw.validate(d):
                        continue
                    want = H.orc_window(d)
                    for mode in (-1, 2):
                        st_, got = hc_bank(d, mode=mode, pair=1)
                        assert st_ in (0, 1)
                        if st_ == 0:
                            n += 1
                            assert np.array_equal(got, want), (st, v, pw, dw, mode, int(np.argmax(got != want)))
    assert n > 300
    assert hc.hc_inq_exceptions(0) > 20


def test_bank_body_most_negative_coefficients_every_placement():
    """a_k = -2^(DW-1) pre-shifts to INT32_MIN, which the half-period placement cannot negate: such
    windows (RTL and HLS alike) must take the generic body, whatever placement is asked for."""
    cases_ = [bhw.make_desc(4, 16, 17, [47022, 64001, -65536, 1531], model=bhw.MODEL_HLS),
              bhw.make_desc(4, 12, 17, [47022, -65536, 18518, -65536], model=bhw.MODEL_HLS),
              bhw.make_desc(3, 14, 16, [27518, -32768, 5242], model=bhw.MODEL_HLS),
              bhw.make_desc(4, 16, 17, [47022, 64001, -65536, 1531]),
              bhw.make_desc(7, 12, 24, [1, -(1 << 23), 3, -(1 << 23), 5, 6, -(1 << 23)], model=bhw.MODEL_HLS)]
    for d in cases_:
        assert bhw.validate(d) == 0
        want = H.orc_window(d)
        ran = 0
        for mode in (-1, 0, 1, 2):
            for pair in (-1, 0, 1):
                st, got = hc_bank(d, mode=mode, pair=pair)
                assert st in (0, 1)
                if st == 0:
                    ran += 1
                    assert np.array_equal(got, want), (mode, pair, d.model, list(d.aa))
        # the bank body refuses them (generic record) ...
        assert ran == 0
        # ... and the generic body, which is where they go, matches the oracle
        n = 1 << d.phi_width
        st, got = hc_window(d, 0, min(n, 4096), "direct")
        assert st == 0 and np.array_equal(got, want[:min(n, 4096)])
        assert hc_window(d, 0, 16, "table")[0] == 1      # not eligible for the fast tail


def test_antisymmetry_claim_holds_wherever_the_planner_relies_on_it():
    """Pairing and the half-period table assume T[i + E/2] == -T[i]; check the claim against the
    oracle's cosine sequence for every source the planner declares antisymmetric - and that the
    planner does not declare the sources for which it is false (TAYLOR on the DSP48 branch)."""
    hc = H.hostcheck()
    claimed = refuted = 0
    for st, model in ((bhw.SIN_CORDIC, 0), (bhw.SIN_CORDIC, 1), (bhw.SIN_TAYLOR, 0), (bhw.SIN_CORDIC48, 0),
                      (bhw.SIN_CORDIC_SCALED, 0)):
        for pw in (4, 6, 9, 12, 14):
            for dw in (4, 5, 7, 8, 12, 16, 17, 18, 19, 24, 30):
                for lut in ((0,) if st != bhw.SIN_TAYLOR else (1, 4, 9)):
                    d = bhw.make_desc(2, pw, dw, [1, 1], sin_type=st, model=model, lut_size=lut)
                    if bhw.validate(d):
                        continue
                    _, c = H.orc_sincos(d)
                    half = len(c) // 2
                    holds = bool(np.array_equal(c[:half], -c[half:]))
                    if hc.hc_source_antisymmetric(C.byref(d)) == 1:
                        claimed += 1
                        assert holds, d
                    elif not holds:
                        refuted += 1
    assert claimed > 100 and refuted > 0
