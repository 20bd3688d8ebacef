"""CPU tier for the group path (k_synth_group, bhw_group.cuh): the kernel's lane/tile body and the
half-period pyramid builder, compiled for the host by tests/hostcheck, against the oracle - for groups that
mix PHI_WIDTHs and ports over one family, in every table placement, paired (whole windows) and unpaired
(windows cut by the requested range)."""
import ctypes as C

import numpy as np
import pytest

import blackman_harris_win_b200 as bhw
import cases
import harness as H

G_HALF32, G_Q16, G_GLOBAL = 0, 1, 2


def hc_group(descs, tab=-1, unpaired=0):
    arr = bhw.desc_array(descs)
    total = sum(1 << d.phi_width for d in descs)
    out = np.full(total, -(1 << 62), np.int64)
    st = H.hostcheck().hc_group(arr, len(descs), out.ctypes.data_as(H.I64P), tab, unpaired)
    return st, out


def want_for(descs):
    return np.concatenate([H.orc_window(d) for d in descs])


def family(variants, dw, pws, model=bhw.MODEL_RTL, offsets=(0,)):
    out = []
    for i, pw in enumerate(pws):
        for v in variants:
            d = bhw.variant_desc(v, pw, dw, model=model)
            out.append(d.copy(stream_offset=offsets[(i + v) % len(offsets)]))
    return out


@pytest.mark.parametrize("variants,dw", [((1, 2), 16), ((3, 4), 16), ((5, 6, 7), 17), ((8, 9), 24), ((10,), 32),
                                         ((10,), 24), ((6,), 12), ((9,), 20), ((1,), 8), ((10,), 31)])
def test_group_body_mixed_phi_widths(variants, dw):
    """Windows of PHI_WIDTH 9..18 (below, at and above the table resolution) in one group."""
    pws = [p for p in (9, 10, 11, 12, 13, 15, 16, 17, 18) if p <= 18]
    descs = family(variants, dw, pws, offsets=(0, 1))
    want = want_for(descs)
    ran = set()
    for tab in (-1, G_HALF32, G_Q16, G_GLOBAL):
        for unpaired in (0, 1):
            st, got = hc_group(descs, tab, unpaired)
            assert st in (0, 1), (st, variants, dw, tab)
            if st == 0:
                ran.add(tab)
                assert np.array_equal(got, want), (variants, dw, tab, unpaired, int(np.argmax(got != want)))
    assert -1 in ran and G_GLOBAL in ran and G_HALF32 in ran
    assert (G_Q16 in ran) == (dw <= 17)


def test_group_body_hls_family():
    for nw, pws in ((16, (9, 12, 16, 17, 18)), (17, (10, 14, 17, 19)), (24, (9, 13, 16)), (12, (9, 12, 13, 14))):
        for v in (1, 3, 6, 9, 10):
            descs = [bhw.variant_desc(v, pw, nw, model=bhw.MODEL_HLS) for pw in pws]
            want = want_for(descs)
            for tab in (-1, G_GLOBAL, G_Q16):
                st, got = hc_group(descs, tab)
                assert st in (0, 1)
                if st == 0:
                    assert np.array_equal(got, want), (nw, v, tab)


def test_group_body_ports_edge_cases():
    """Negative / unsigned-reading / zero / all-ones ports in groups; windows that need the generic body or
    the 64-bit tail are refused (they keep the older paths)."""
    n_ok = n_refused = 0
    for d in cases.edge_coeff_descs():
        if d.phi_width < 9:
            d = d.copy(phi_width=10)
        for pws in ((10, 13), (9,)):
            descs = [d.copy(phi_width=p) for p in pws]
            st, got = hc_group(descs)
            assert st in (0, 1)
            if st == 0:
                n_ok += 1
                assert np.array_equal(got, want_for(descs)), (d.win_type, d.dat_width, list(d.aa))
            else:
                n_refused += 1
    assert n_ok > 100 and n_refused > 10


def test_group_refuses_foreign_windows():
    taylor = bhw.make_desc(3, 14, 24, [7046424, 8388600, 1342176], sin_type=bhw.SIN_TAYLOR, lut_size=9)
    assert hc_group([taylor])[0] == 1
    assert hc_group([bhw.variant_desc(10, 12, 32, sin_type=bhw.SIN_CORDIC48)])[0] == 1
    assert hc_group([bhw.variant_desc(6, 8, 17)])[0] == 1                                 # shorter than a tile pair
    assert hc_group([bhw.variant_desc(6, 12, 17), bhw.variant_desc(6, 12, 16)])[0] == 1   # two families
    assert hc_group([bhw.variant_desc(6, 12, 17), bhw.variant_desc(9, 12, 17)])[0] == 1   # two entities
    assert hc_group([bhw.variant_desc(6, 12, 17, algo=bhw.ALGO_DIRECT)])[0] == 1


def test_group_body_random_descriptors():
    rng = np.random.default_rng(20261018)
    n = 0
    for d in cases.random_descs(300, seed=77):
        if d.dat_width > 32 or d.sin_type != bhw.SIN_CORDIC:
            continue
        pws = sorted(set(int(x) for x in rng.integers(9, 15, size=3)))
        if d.model == bhw.MODEL_HLS:
            pws = [p for p in pws if p <= d.dat_width + 2]
        descs = [d.copy(phi_width=p) for p in pws if bhw.validate(d.copy(phi_width=p)) == 0]
        if not descs:
            continue
        for tab in (-1, G_GLOBAL):
            st, got = hc_group(descs, tab, int(rng.integers(0, 2)))
            assert st in (0, 1)
            if st == 0:
                n += 1
                assert np.array_equal(got, want_for(descs)), (d.win_type, d.dat_width, d.model, pws, tab)
    assert n > 60
