"""CPU tier: the two restatements of the RTL - oracle/bhw_oracle.c (integers + explicit wraps) and
oracle/rtl_bitvec.py (bit vectors with the VHDL's own widths, slices and std_logic_signed `+`) - must
agree.  The RTL has no executable reference here (no VHDL simulator), so this is the anchor that
replaces one: two texts written from the same source in two different styles."""
import ctypes as C
import os
import random
import sys

import numpy as np

import blackman_harris_win_b200 as bhw
import harness as H

sys.path.insert(0, os.path.join(H.ROOT, "oracle"))
import rtl_bitvec as RB  # noqa: E402


def test_cordic_dds_matches_oracle():
    rng = random.Random(1234)
    L = H.oracle()
    s, c = C.c_int64(0), C.c_int64(0)
    n = 0
    for dw in (4, 8, 12, 16, 17, 24, 31, 32, 40, 47):
        for prec in (1, 2, 7) if dw + 7 <= 49 else (1,):
            if dw + prec > 49:
                continue
            for pw in (4, 7, dw - 1, dw, dw + 1, 26):
                if pw < 4:
                    continue
                for _ in range(24):
                    ph = rng.randrange(1 << pw)
                    L.orc_cordic_dds(pw, dw, prec, ph, C.byref(s), C.byref(c))
                    assert RB.cordic_dds(pw, dw, prec, ph) == (s.value, c.value), (pw, dw, prec, ph)
                    n += 1
    assert n > 2000


def test_window_entities_match_oracle():
    """hamming_win / bh_win_{3,4,5,7}term, CORDIC source: DT_WIN of random phases, the README variants
    and adversarial port values (most negative / all-ones coefficients)."""
    rng = random.Random(99)
    n = 0
    for m in (2, 3, 4, 5, 7):
        for dw in (8, 13, 16, 17, 24, 31, 32):
            for pw in (4, 9, dw, 20):
                lo, hi = -(1 << (dw - 1)), (1 << (dw - 1)) - 1
                sets = [[rng.randrange(lo, hi + 1) for _ in range(m)], [hi] * m, [lo] * m,
                        [rng.randrange(0, 1 << dw) for _ in range(m)]]         # raw bits, unsigned reading
                for aa in sets:
                    d = bhw.make_desc(m, pw, dw, aa)
                    if bhw.validate(d):
                        continue
                    N = 1 << pw
                    idx = sorted({0, 1, N // 4, N // 2, N - 1} | {rng.randrange(N) for _ in range(6)})
                    for i in idx:
                        want = int(H.orc_window(d, i, 1)[0])
                        assert RB.window(m, pw, dw, aa, i) == want, (m, pw, dw, aa, i)
                        n += 1
    assert n > 2500


def test_known_answers_from_the_survey():
    """SURVEY 8(c) [derived] anchors, PW=11 / DW=16."""
    assert [RB.cordic_dds(11, 16, 1, i)[1] for i in range(8)] == [16385, 16384, 16384, 16384, 16384, 16380, 16380, 16379]
    assert [RB.window(2, 11, 16, [17808, 14959], i) for i in range(8)] == [5164, 5164, 5164, 5164, 5164, 5165, 5165, 5166]
    assert RB.window(2, 11, 16, [17808, 14959], 1024) == 12644
    assert [RB.window(4, 11, 16, [23511, 32000, 9259, 765], i) for i in range(4)] == [2939, 2940, 2940, 2939]
    assert RB.window(7, 11, 16, [8887, 14203, 7143, 2156, 353, 25, 0], 1024) == 5207


def test_taylor_units_and_windows_match_oracle():
    """taylor_sincos + tay1_order in all four branches (ROM-only LESS / EQ, DSP48 MACC, 35x27 multiplier with
    saturation) and the TAYLOR forms of hamming_win / bh_win_3term (BASELINE config 4 is one of them)."""
    rng = random.Random(2024)
    n = 0
    for pw, lut in [(6, 6), (6, 5), (8, 6), (10, 7), (14, 9), (16, 9), (20, 9), (24, 9), (26, 10), (12, 1), (18, 16)]:
        for dw in (8, 16, 18, 19, 24, 32):
            d0 = bhw.make_desc(2, pw, dw, [1, 1], sin_type=bhw.SIN_TAYLOR, lut_size=lut)
            if bhw.validate(d0):
                continue
            N = 1 << pw
            idx = sorted({0, 1, N // 4 - 1, N // 4, N // 2, 3 * N // 4 + 1, N - 1} | {rng.randrange(N) for _ in range(10)})
            for i in idx:
                s, c = H.orc_sincos(d0, i, 1)
                assert RB.taylor_sincos(pw, dw, lut, i) == (int(s[0]), int(c[0])), (pw, dw, lut, i)
                n += 1
            amp = (1 << (dw - 1)) - 1
            for wt, aa in ((2, [100, 77]), (3, [90, 100, 17])):
                q = [a * amp // 128 for a in aa]
                d = bhw.make_desc(wt, pw, dw, q, sin_type=bhw.SIN_TAYLOR, lut_size=lut)
                if bhw.validate(d):
                    continue
                for i in idx[:8]:
                    assert RB.window_taylor(wt, pw, dw, lut, q, i) == int(H.orc_window(d, i, 1)[0]), (wt, pw, dw, lut, i)
                    n += 1
    assert n > 800
    # BASELINE config 4 itself: Blackman, PHI_WIDTH 24, DAT_WIDTH 24, LUT_SIZE 9
    d = bhw.make_desc(3, 24, 24, [7046424, 8388600, 1342176], sin_type=bhw.SIN_TAYLOR, lut_size=9)
    for i in [0, 1, 4097, (1 << 22) - 1, 1 << 22, (1 << 23) + 12345, (1 << 24) - 1] + [rng.randrange(1 << 24) for _ in range(40)]:
        assert RB.window_taylor(3, 24, 24, 9, [7046424, 8388600, 1342176], i) == int(H.orc_window(d, i, 1)[0]), i


def test_cordic_dds48_and_config3_match_oracle():
    rng = random.Random(48)
    n = 0
    for dw in (4, 8, 16, 24, 32, 40, 47):
        for pw in (4, 10, 20, 26):
            d = bhw.make_desc(2, pw, dw, [1, 1], sin_type=bhw.SIN_CORDIC48)
            assert bhw.validate(d) == 0
            for i in sorted({0, 1, (1 << pw) // 4, (1 << pw) // 2 + 1, (1 << pw) - 1} | {rng.randrange(1 << pw) for _ in range(12)}):
                s, c = H.orc_sincos(d, i, 1)
                assert RB.cordic_dds48(pw, dw, i) == (int(s[0]), int(c[0])), (pw, dw, i)
                n += 1
    assert n > 400
    # BASELINE config 3: bh_win_7term, PHI_WIDTH 20, DAT_WIDTH 32, cordic_dds48 swapped in
    aa = [582441289, 930815217, 468160289, 141272949, 23110934, 1653590, 29379]
    d = bhw.make_desc(7, 20, 32, aa, sin_type=bhw.SIN_CORDIC48)
    for i in [0, 1, 2, 262143, 262144, 524288, 786433, (1 << 20) - 1] + [rng.randrange(1 << 20) for _ in range(40)]:
        assert RB.window_dds48(7, 20, 32, aa, i) == int(H.orc_window(d, i, 1)[0]), i
        # and with the entity's own cordic_dds
        assert RB.window(7, 20, 32, aa, i) == int(H.orc_window(d.copy(sin_type=bhw.SIN_CORDIC), i, 1)[0]), i


def test_cordic_atan2_matches_oracle_and_converges():
    rng = random.Random(7)
    L = H.oracle()
    n = 0
    for aw, iw, prec in [(8, 8, 1), (12, 11, 2), (16, 16, 1), (16, 20, 3), (24, 24, 1), (24, 32, 1), (32, 31, 1),
                         (32, 32, 7), (4, 3, 1), (20, 19, 4)]:
        assert L.orc_atan2_validate(iw, aw, prec) == 0
        for _ in range(60):
            x, y = rng.randrange(1 << iw), rng.randrange(1 << iw)
            want = L.orc_cordic_atan2(iw, aw, prec, x, y)
            assert RB.cordic_atan2(iw, aw, prec, x, y) == want, (aw, iw, prec, x, y)
            n += 1
    assert n == 600
    # the vectoring iterations really do measure the angle: first quadrant, magnitudes well inside the
    # register, PHI_DT = -atan(|y|/|x|) in units of pi = 2^(ANGLE_WIDTH-1) (the entity's sign convention)
    aw, iw, prec = 24, 24, 3
    for _ in range(200):
        x, y = rng.randrange(1 << 16, 1 << 20), rng.randrange(0, 1 << 20)
        got = L.orc_cordic_atan2(iw, aw, prec, x, y)
        want = -np.arctan2(y, x) / np.pi * (1 << (aw - 1))
        assert abs(got - want) < 64, (x, y, got, want)
    # invalid generics
    assert L.orc_atan2_validate(20, 24, 1) != 0      # the entity's own defaults index VEC_DX out of range
    assert L.orc_atan2_validate(24, 33, 1) != 0
    assert L.orc_atan2_validate(24, 24, 8) != 0


def test_atan2_kernel_body_matches_oracle():
    rng = np.random.default_rng(5)
    hc = H.hostcheck()
    p32 = C.POINTER(C.c_int32)
    for aw, iw, prec in [(8, 8, 1), (12, 11, 2), (16, 16, 1), (16, 20, 3), (24, 24, 1), (24, 32, 1), (32, 31, 1),
                         (32, 32, 7), (4, 3, 1), (20, 19, 4), (24, 24, 0)]:
        x = rng.integers(-(1 << 31), 1 << 31, 4096, dtype=np.int64).astype(np.int32)
        y = rng.integers(-(1 << 31), 1 << 31, 4096, dtype=np.int64).astype(np.int32)
        x[:4] = [0, -1, (1 << (iw - 1)) - 1, -(1 << (iw - 1))]
        y[:4] = [0, -1, -(1 << (iw - 1)), (1 << (iw - 1)) - 1]
        d = bhw.BhwAtan2Desc(iw, aw, prec, 0)
        assert bhw.lib().bhw_atan2_validate(C.byref(d)) == 0
        got = np.empty(4096, np.int32)
        assert hc.hc_atan2(C.byref(d), x.ctypes.data_as(p32), y.ctypes.data_as(p32), got.ctypes.data_as(p32), 4096) == 0
        assert np.array_equal(got, H.orc_atan2(iw, aw, prec, x, y)), (aw, iw, prec)
    for bad in (bhw.BhwAtan2Desc(20, 24, 1, 0), bhw.BhwAtan2Desc(24, 33, 1, 0), bhw.BhwAtan2Desc(24, 24, 8, 0),
                bhw.BhwAtan2Desc(24, 24, 1, 5)):
        assert bhw.lib().bhw_atan2_validate(C.byref(bad)) != 0
