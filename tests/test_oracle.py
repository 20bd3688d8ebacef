"""CPU tier: pin the oracle (oracle/bhw_oracle.c) against
  (1) the committed golden vectors and hashes produced by the reference's own C++ (tests/golden/),
  (2) the compiled reference itself when oracle/_ref is present (live cross-check),
  (3) the [derived] RTL anchors of SURVEY 8(c) (parity unpinned - regression anchor only)."""
import json
import os

import numpy as np
import pytest

import blackman_harris_win_b200 as bhw
import cases
import harness as H

GOLD = os.path.join(os.path.dirname(__file__), "golden")
VEC = np.load(os.path.join(GOLD, "reference_vectors.npz"))
KAT = json.load(open(os.path.join(GOLD, "reference_kat.json")))
RTL = json.load(open(os.path.join(GOLD, "rtl_kat.json")))


MTERM = json.load(open(os.path.join(GOLD, "mterm_kat.json")))


@pytest.mark.parametrize("e", MTERM["windows"], ids=lambda e: f"v{e['variant']}_pw{e['phi_width']}_dw{e['dat_width']}_s{e['sin_type']}")
def test_mterm_definition_is_frozen(e):
    """6 and 8..11 terms have no reference entity: tests/golden/mterm_kat.json freezes the oracle's answer (a definition,
    not parity) together with the ports bhw_quantize produced for it."""
    d = bhw.variant_desc(e["variant"], e["phi_width"], e["dat_width"], sin_type=e["sin_type"])
    assert d.win_type == e["win_type"] and list(d.aa)[:d.win_type] == e["aa"]
    w = H.orc_window(d)
    for i, v in e["values"].items():
        assert int(w[int(i)]) == v
    assert H.sha_lines(w) == e["sha256"]


def hls_desc(np_, nw, t):
    return bhw.variant_desc(cases.HLS_TYPES[t], np_, nw, model=bhw.MODEL_HLS)


@pytest.mark.parametrize("key", [k for k in VEC.files if k.startswith("hls_win_")])
def test_hls_window_golden_vectors(key):
    np_, nw, t = (int(x) for x in key.replace("hls_win_np", "").replace("_nw", " ").replace("_t", " ").split())
    assert np.array_equal(H.orc_window(hls_desc(np_, nw, t)), VEC[key])


@pytest.mark.parametrize("key", [k for k in VEC.files if k.startswith("hls_cordic_")])
def test_hls_cordic_golden_vectors(key):
    np_, nw = (int(x) for x in key.replace("hls_cordic_np", "").replace("_nw", " ").split())
    s, c = H.orc_sincos(bhw.make_desc(2, np_, nw, model=bhw.MODEL_HLS))
    assert np.array_equal(s, VEC[key][0]) and np.array_equal(c, VEC[key][1])


@pytest.mark.parametrize("key", [k for k in VEC.files if k.startswith("cpp_")])
def test_cpp_cordic_golden_vectors(key):
    pw, dw = (int(x) for x in key.replace("cpp_pw", "").replace("_dw", " ").split())
    s, c = H.orc_sincos(bhw.make_desc(2, pw, dw, model=bhw.MODEL_CPP))
    assert np.array_equal(s, VEC[key][0]) and np.array_equal(c, VEC[key][1])


@pytest.mark.parametrize("e", [e for e in KAT["hls_win"] if e["nphase"] <= 18],
                         ids=lambda e: f"np{e['nphase']}_nw{e['nwidth']}_t{e['type']}")
def test_hls_window_hashes(e):
    w = H.orc_window(hls_desc(e["nphase"], e["nwidth"], e["type"]))
    assert [int(x) for x in w[:4]] == e["first"]
    assert int(w[len(w) // 2]) == e["mid"]
    assert H.sha_lines(w) == e["sha256"]


def test_hls_window_hash_cfg3_bh7_n1m_dw32():
    """BASELINE config 3 widths through the reference HLS model (full 2^20 table, 8 threads)."""
    e = next(x for x in KAT["hls_win"] if (x["nphase"], x["nwidth"], x["type"]) == (20, 32, 7))
    w = H.orc_window(hls_desc(20, 32, 7), threads=8)
    assert H.sha_lines(w) == e["sha256"]


@pytest.mark.parametrize("e", [e for e in KAT["cpp"] if e["phase_width"] <= 16],
                         ids=lambda e: f"pw{e['phase_width']}_dw{e['data_width']}")
def test_cpp_hashes(e):
    s, c = H.orc_sincos(bhw.make_desc(2, e["phase_width"], e["data_width"], model=bhw.MODEL_CPP))
    assert H.sha_pairs(s, c) == e["sha256_s_c"]


def test_survey_reference_kats():
    """Hashes captured during the survey with an independent ap_int stand-in."""
    want = {(10, 24, 1): "0fc423980e44f78269178a813f950b55f2835811e91ea4dcf3808c0b2a99bec5",
            (10, 24, 7): "742166b9bd777736faf47872260777a6d0b7f0aecac3a224e14231b1a6b3b363",
            (10, 16, 2): "b76bacd3c2d6f5e07f35c3526336a868593b0177c3902e6b6280737f589ba87c",
            (16, 17, 4): "c327569466cad2e948c0cd2e15d299c54ecd92ecd79506254cffc2983911d4f6"}
    for (np_, nw, t), sha in want.items():
        assert H.sha_lines(H.orc_window(hls_desc(np_, nw, t))) == sha
    s, c = H.orc_sincos(bhw.make_desc(2, 14, 12, model=bhw.MODEL_CPP))
    assert H.sha_pairs(s, c) == "966e930a791ecdaa0b4f27e7277d3a0b5736f24f2708952380718777292b1d9e"
    assert (int(s[4096]), int(c[4096])) == (1024, -1) and (int(s[12288]), int(c[12288])) == (-1025, 0)
    # Hann wraps at the centre in the HLS model
    assert int(H.orc_window(hls_desc(10, 24, 2))[512]) == -8388608


# ---- live cross-check against the compiled reference (present here and on the GPU box) ---------
@pytest.mark.parametrize("cfg", H.ref_configs("hls_win"), ids=lambda c: f"np{c[0]}_nw{c[1]}")
def test_hls_window_vs_compiled_reference(cfg):
    np_, nw = cfg
    n = 1 << np_
    for t in cases.HLS_TYPES:
        for n0, cnt in ((0, min(n, 4096)), (max(0, n - 2048), min(n, 2048)), (n // 2 - 8 if n >= 16 else 0, 16)):
            assert np.array_equal(H.orc_window(hls_desc(np_, nw, t), n0, cnt),
                                  H.ref_hls_window(np_, nw, t, n0, cnt)), (np_, nw, t, n0)


@pytest.mark.parametrize("cfg", H.ref_configs("cpp"), ids=lambda c: f"pw{c[0]}_dw{c[1]}")
def test_cpp_vs_compiled_reference(cfg):
    pw, dw = cfg
    cnt = min(1 << pw, 1 << 15)
    s, c = H.orc_sincos(bhw.make_desc(2, pw, dw, model=bhw.MODEL_CPP), 0, cnt)
    rs, rc = H.ref_cpp_cordic(pw, dw, 0, cnt)
    assert np.array_equal(s, rs) and np.array_equal(c, rc)


# ---- RTL anchors (parity unpinned) ---------------------------------------------------------------
@pytest.mark.parametrize("e", RTL["windows"], ids=lambda e: e["name"])
def test_rtl_window_anchors(e):
    d = bhw.make_desc(e["win_type"], e["phi_width"], e["dat_width"], e["aa"], sin_type=e["sin_type"],
                      lut_size=e["lut_size"])
    w = H.orc_window(d)
    for i, v in e["values"].items():
        if int(i) < len(w):
            assert int(w[int(i)]) == v, (e["name"], i)
    if "min" in e:
        assert int(w.min()) == e["min"]
    assert H.sha_lines(w) == e["sha256"]


@pytest.mark.parametrize("e", RTL["sincos"], ids=lambda e: e["name"])
def test_rtl_sincos_anchors(e):
    d = bhw.make_desc(2, e["phi_width"], e["dat_width"], sin_type=e["sin_type"])
    _, c = H.orc_sincos(d, 0, 8)
    assert [int(x) for x in c[: len(e["cos_first"])]] == e["cos_first"]


def test_taylor_rom_anchors():
    import ctypes as C
    for e in RTL["taylor_rom"]:
        n = 1 << e["lut_size"]
        rc, rs = np.empty(n, np.int64), np.empty(n, np.int64)
        H.oracle().orc_taylor_rom(e["dat_width"], e["lut_size"], rc.ctypes.data_as(H.I64P), rs.ctypes.data_as(H.I64P))
        assert (int(rc[e["index"]]), int(rs[e["index"]])) == (e["cos"], e["sin"])
        assert int(rc[0]) == (1 << (e["dat_width"] - 1)) - 1 and int(rs[0]) == 0


def test_dds48_sin_is_negated_and_amplitude():
    """cordic_dds48 outputs -A*sin, amplitude 2^(DW-2), max error ~1.5 LSB."""
    d = bhw.make_desc(2, 10, 16, sin_type=bhw.SIN_CORDIC48)
    s, c = H.orc_sincos(d)
    ph = 2 * np.pi * np.arange(1024) / 1024
    assert np.abs(c - 16384 * np.cos(ph)).max() < 2.5
    assert np.abs(s + 16384 * np.sin(ph)).max() < 2.5


def test_window_shape_sanity():
    """RTL Hamming with CORDIC ~ (AA0 - AA1*cos/2)/2 (SURVEY 2.1 D consequence)."""
    d = bhw.make_desc(2, 10, 16, [17808, 14959])
    w = H.orc_window(d).astype(float)
    ideal = (17808 - 14959 * np.cos(2 * np.pi * np.arange(1024) / 1024) / 2) / 2
    assert np.abs(w - ideal).max() < 4
