"""The reference's own C-sim self-checks, restated: its only executable tests are two tolerance
checks against double-precision math.  Applied to the oracle here (CPU tier) and to the
CUDA path in tests/test_gpu_parity.py::test_reference_selfcheck_criteria_on_the_gpu.

  hls/cordic/cordic_test.cpp:66-93    mean |s - round(2^(NW-2) sin)| < 10 and the same for cos, NPHASE 10 / NWIDTH 16
  hls/windows/window_test.cpp:93-216  sqrt(sum err^2) / N < 10 against round((2^(NW-shift) - 1) * ideal), NPHASE 10 /
                                      NWIDTH 24; shift = 1 for types 1-4, 2 for types 5 and 7; type 3 is checked
                                      against the halved Blackman coefficients 0.21 / 0.25 / 0.04 (:110-116)
"""
import numpy as np
import pytest

import blackman_harris_win_b200 as bhw
import cases
import harness as H

# window_test.cpp:96-186: the double-precision coefficients and the scale shift per win_type
SELFCHECK = {
    1: ([0.5434783, 1.0 - 0.5434783], 1),
    2: ([0.5, 0.5], 1),
    3: ([0.21, 0.25, 0.04], 1),
    4: ([0.35875, 0.48829, 0.14128, 0.01168], 1),
    5: ([0.3232153788877343, 0.4714921439576260, 0.1755341299601972, 0.0284969901061499, 0.0012613570882927], 2),
    7: ([0.271220360585039, 0.433444612327442, 0.218004122892930, 0.065785343295606, 0.010761867305342,
         0.000770012710581, 0.000013680883060], 2),
}


def window_selfcheck_error(win_type, out, np_, nw):
    """acc_err of window_test.cpp:196-212 for the integer window `out`."""
    coe, shift = SELFCHECK[win_type]
    n = 1 << np_
    i = np.arange(n)
    ideal = sum(((-1) ** k) * a * np.cos(2 * k * i * np.pi / n) for k, a in enumerate(coe))
    rnd = np.round((2.0 ** (nw - shift) - 1.0) * ideal)
    rnd = ((rnd.astype(np.int64) + (1 << (nw - 1))) % (1 << nw)) - (1 << (nw - 1))     # the (win_t) cast
    return float(np.sqrt(np.sum((rnd - out.astype(np.float64)) ** 2)) / n)


def cordic_selfcheck_errors(s, c, np_, nw):
    """acc_s, acc_c of cordic_test.cpp:66-85 (mean absolute error against the rounded ideal)."""
    n = 1 << np_
    i = np.arange(n)
    ts = np.round(2.0 ** (nw - 2) * np.sin(2 * i * np.pi / n))
    tc = np.round(2.0 ** (nw - 2) * np.cos(2 * i * np.pi / n))
    return float(np.abs(s - ts).mean()), float(np.abs(c - tc).mean())


def test_hls_cordic_selfcheck_passes_on_the_oracle():
    d = bhw.variant_desc(1, 10, 16, model=bhw.MODEL_HLS)
    s, c = H.orc_sincos(d)
    es, ec = cordic_selfcheck_errors(s, c, 10, 16)
    assert es < 10 and ec < 10, (es, ec)
    assert es < 2 and ec < 2          # what the model actually achieves


@pytest.mark.parametrize("win_type", sorted(SELFCHECK))
def test_hls_window_selfcheck_on_the_oracle(win_type):
    """All types pass the reference's criterion at its own widths except type 2: round(0.5 * (2^23 - 1)) * 2 = 2^23
    wraps the Hann window's centre sample to -2^23, which the criterion sees as an error of 2^24 / N."""
    d = bhw.variant_desc(cases.HLS_TYPES[win_type], 10, 24, model=bhw.MODEL_HLS)
    err = window_selfcheck_error(win_type, H.orc_window(d), 10, 24)
    if win_type == 2:
        out = H.orc_window(d)
        assert out[512] == -(1 << 23)
        assert err > 10
        out = out.copy(); out[512] = (1 << 23) - 1          # without the wrapped sample the criterion holds
        assert window_selfcheck_error(2, out, 10, 24) < 10
    else:
        assert err < 10, err


def test_rtl_entities_track_the_ideal_window():
    """The same criterion applied to the RTL restatement (the reference has no self-check for the RTL):
    with the CORDIC amplitude 2^(DW-2) the entity computes (AA0 - AA1 cos / 2 + ...) / 4 (/ 2 for the
    2-term entity), so the ideal is built from the integer ports themselves."""
    for v in range(1, 11):
        for pw, dw in ((10, 16), (12, 24), (11, 32)):
            for st in (bhw.SIN_CORDIC, bhw.SIN_CORDIC48, bhw.SIN_CORDIC_SCALED):
                d = bhw.variant_desc(v, pw, dw, sin_type=st)
                if bhw.validate(d):
                    continue
                n, m = 1 << pw, d.win_type
                aa = [int(a) for a in d.aa[:m]]
                i = np.arange(n)
                acc = aa[0] + sum(((-1) ** k) * aa[k] * np.cos(2 * k * i * np.pi / n) / 2 for k in range(1, m))
                ideal = acc / (2 if m == 2 else 4)
                ideal = ((np.round(ideal).astype(np.int64) + (1 << (dw - 1))) % (1 << dw)) - (1 << (dw - 1))
                out = H.orc_window(d)
                err = float(np.sqrt(np.sum((ideal - out.astype(np.float64)) ** 2)) / n)
                assert err < 10, (v, pw, dw, st, err)
                assert np.abs(ideal - out).max() < 64, (v, pw, dw, st)
