#!/usr/bin/env python
"""Golden vectors of the RTL entities, produced by EXECUTING the reference's own VHDL (oracle/vhdl_sim.py reads
/root/reference/src/*.vhd where they lie and clocks the elaborated entities).  Runs in the container that has the
reference checkout; the vectors travel in tests/golden/rtl_sim_vectors.npz (+ rtl_sim_cases.json) and pin
oracle/bhw_oracle.c (tests/test_rtl_vhdl_sim.py) and the CUDA path (tests/test_gpu_parity.py) wherever the reference
itself is absent.

  python tests/golden/make_rtl_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vhdl_sim as V  # noqa: E402
import blackman_harris_win_b200 as bhw  # noqa: E402  (bhw_quantize only: the testbench's coefficient rules)


def phases_for(pw, rng, n=192):
    N = 1 << pw
    if N <= 1024:
        return list(range(N))
    fixed = [0, 1, 2, 3, N // 4 - 1, N // 4, N // 4 + 1, N // 2 - 1, N // 2, N // 2 + 1, 3 * N // 4 - 1, 3 * N // 4, N - 2, N - 1]
    return sorted(set(fixed + [int(x) for x in rng.integers(0, N, n)]))


def swapped_cases(lib, arrays, cases):
    """BASELINE config 3 names a composition the reference never elaborates: a window entity with cordic_dds48 (or
    cordic_dds_scaled) in place of cordic_dds.  The three DDS entities have the same port list, so the swap is made
    here by binding the name cordic_dds to the other entity and clocking the unmodified window entity."""
    import copy
    ent_of = {2: "hamming_win", 3: "bh_win_3term", 4: "bh_win_4term", 5: "bh_win_5term", 7: "bh_win_7term"}
    cases["windows_swapped"] = []
    for dds in ("cordic_dds48", "cordic_dds_scaled"):
        l2 = copy.copy(lib)
        l2.units = dict(lib.units)
        l2.units["cordic_dds"] = lib.units[dds]
        for v, pw, dw in ((10, 6, 32), (10, 8, 16), (10, 7, 24), (6, 7, 17), (1, 6, 16), (9, 6, 24), (3, 6, 12), (8, 7, 20)):
            aa, m = bhw.quantize(v, bhw.RULE_TB, dw)
            gen = {"PHI_WIDTH": pw, "DAT_WIDTH": dw}
            if m <= 3:
                gen["SIN_TYPE"] = "CORDIC"
            clocks = 2 * (1 << pw) + dw + 30
            out = V.run_window(l2, ent_of[m], gen, [int(a) for a in aa[:m]], clocks)
            vld = [o[1] for o in out]
            key = f"winswap/{dds}/{ent_of[m]}/pw{pw}_dw{dw}_variant{v}"
            arrays[key + "/aa"] = np.array([int(a) for a in aa[:m]], np.int64)
            arrays[key + "/dt_win_per_clock"] = np.array([o[0] for o in out], np.int64)
            arrays[key + "/dt_vld_per_clock"] = np.array(vld, np.int8)
            cases["windows_swapped"].append({"dds": dds, "entity": ent_of[m], "generics": gen, "key": key, "clocks": clocks,
                                             "first_dt_vld_clock": vld.index(1)})
            print(key, "first DT_VLD clock", vld.index(1), flush=True)


TB_FILE = os.path.join(V.REF_SRC, "tb", "tb_windows.vhd")
# constant prefix in the testbench -> (README variant whose real-valued set it spells, number of terms)
TB_SETS = {"cnt7": (10, 7), "cnt5": (8, 5), "cnt4": (6, 4), "cnt3": (3, 3), "cnt2": (1, 2)}


def tb_constants(width):
    """The integer port constants CNT*_STD* of the reference's testbench (src/tb/tb_windows.vhd:64-127), obtained by
    ELABORATING the declarative part of its architecture with the simulator - the reals, 2.0**(CONST_WIDTH-1)-1.0 and
    friends, INTEGER(real) and conv_std_logic_vector are evaluated from the reference's own text.  The statement part
    (clock / reset processes with `wait for`, the one live taylor_sincos instance) is not needed for constants and is
    left out, as are the two `time` constants; `width` replaces the literal 16 of `constant CONST_WIDTH : integer:=16`
    (the line a user of the testbench edits).  -> {prefix: [signed ints]}"""
    import re
    src = open(TB_FILE, encoding="latin-1").read()
    m = re.search(r"architecture\s+testbench\s+of\s+tb_windows\s+is(.*?)\nbegin\b", src, re.S | re.I)
    decl = "\n".join(l for l in m.group(1).split("\n") if not re.search(r":\s*time\s*:=", l))
    decl, n = re.subn(r"(constant\s+CONST_WIDTH\s*:\s*integer\s*:=\s*)16", r"\g<1>%d" % width, decl)
    assert n == 1
    text = ("library ieee; use ieee.std_logic_1164.all;\nentity tb_consts is end tb_consts;\n"
            "architecture a of tb_consts is\n" + decl + "\nbegin\nend a;\n")
    lib = V.Library([])
    for name, u in V.Parser(text).design_file().items():
        lib.units.setdefault(name, {}).update(u)
    inst = V.Instance(lib, "tb_consts")
    out = {}
    for prefix, (_, terms) in TB_SETS.items():
        out[prefix] = [int(inst.scope.d[f"{prefix}_std{k}"].signed()) for k in range(terms)]
    return out


def tb_constant_cases(cases):
    """CONST_WIDTH 8 .. 32: every product a_k * scale stays inside VHDL's 32-bit INTEGER (the largest, flat-top
    1.93 * (2**30 - 1), is 2.07e9), so a conforming simulator elaborates all of them."""
    cases["tb_constants"] = []
    for w in (8, 12, 16, 17, 20, 24, 30, 31, 32):
        got = tb_constants(w)
        for prefix, (variant, terms) in TB_SETS.items():
            # what the CONST_WIDTH-bit vector holds, read as the signed port value
            cases["tb_constants"].append({"width": w, "set": prefix, "variant": variant, "terms": terms, "aa": got[prefix]})
    print("tb constants:", len(cases["tb_constants"]), "sets", flush=True)


def update_only_tb():
    """python tests/golden/make_rtl_golden.py tb: add the testbench constants to the existing case file."""
    cases = json.load(open(os.path.join(HERE, "rtl_sim_cases.json")))
    tb_constant_cases(cases)
    json.dump(cases, open(os.path.join(HERE, "rtl_sim_cases.json"), "w"), indent=1)


def precision_cases(lib, arrays, cases, rng):
    """cordic_dds with its PRECISION generic off the default (src/cordic_dds.vhd:79: "from 1 to 7"; the window entities
    always leave it at 1, bhw_desc.precision exposes it for bhw_sincos and the windows alike)."""
    cases["dds_precision"] = []
    for pw, dw, prec in ((8, 12, 2), (10, 16, 3), (9, 8, 7), (12, 24, 4), (11, 32, 2), (14, 17, 5), (10, 40, 7), (7, 42, 7)):
        ph = phases_for(pw, rng)
        out, lat = V.run_dds(lib, "cordic_dds", pw, dw, ph, precision=prec)
        key = f"ddsprec/cordic_dds/pw{pw}_dw{dw}_p{prec}"
        arrays[key + "/phases"] = np.array(ph, np.int64)
        arrays[key + "/sin"] = np.array([o[0] for o in out], np.int64)
        arrays[key + "/cos"] = np.array([o[1] for o in out], np.int64)
        cases["dds_precision"].append({"entity": "cordic_dds", "phase_width": pw, "data_width": dw, "precision": prec, "key": key,
                                       "dt_val_latency": lat})
        print(key, "latency", lat, flush=True)


def update_only_precision():
    """python tests/golden/make_rtl_golden.py precision: add the PRECISION cases to the existing files."""
    lib = V.reference_library()
    z = np.load(os.path.join(HERE, "rtl_sim_vectors.npz"))
    arrays = {k: z[k] for k in z.files if not k.startswith("ddsprec/")}
    cases = json.load(open(os.path.join(HERE, "rtl_sim_cases.json")))
    precision_cases(lib, arrays, cases, np.random.default_rng(20261019))
    np.savez_compressed(os.path.join(HERE, "rtl_sim_vectors.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "rtl_sim_cases.json"), "w"), indent=1)
    print("wrote", len(arrays), "arrays")


def update_only_swapped():
    """python tests/golden/make_rtl_golden.py swapped: add the swapped compositions to the existing files."""
    lib = V.reference_library()
    z = np.load(os.path.join(HERE, "rtl_sim_vectors.npz"))
    arrays = {k: z[k] for k in z.files if not k.startswith("winswap/")}
    cases = json.load(open(os.path.join(HERE, "rtl_sim_cases.json")))
    swapped_cases(lib, arrays, cases)
    np.savez_compressed(os.path.join(HERE, "rtl_sim_vectors.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "rtl_sim_cases.json"), "w"), indent=1)
    print("wrote", len(arrays), "arrays")


def main():
    lib = V.reference_library()
    rng = np.random.default_rng(20261018)
    arrays, cases = {}, {"dds": [], "windows": [], "atan2": [], "source": "oracle/vhdl_sim.py over /root/reference/src/*.vhd"}

    # ---- the DDS entities: PH_IN -> (DT_SIN, DT_COS), valid flag from DT_VAL
    dds_sets = {"cordic_dds": [(4, 8), (10, 16), (11, 16), (8, 12), (6, 10), (16, 17), (14, 24), (20, 32), (24, 24), (26, 16), (12, 8), (9, 40)],
                "cordic_dds48": [(4, 8), (10, 16), (14, 12), (16, 24), (20, 32), (9, 47), (12, 8)],
                "cordic_dds_scaled": [(4, 8), (10, 16), (14, 12), (16, 24), (20, 32), (12, 8), (26, 17), (9, 21)]}
    for ent, sets in dds_sets.items():
        for pw, dw in sets:
            ph = phases_for(pw, rng)
            out, lat = V.run_dds(lib, ent, pw, dw, ph)
            key = f"dds/{ent}/pw{pw}_dw{dw}"
            arrays[key + "/phases"] = np.array(ph, np.int64)
            arrays[key + "/sin"] = np.array([o[0] for o in out], np.int64)
            arrays[key + "/cos"] = np.array([o[1] for o in out], np.int64)
            cases["dds"].append({"entity": ent, "phase_width": pw, "data_width": dw, "key": key, "dt_val_latency": lat})
            print(key, "latency", lat, flush=True)

    # ---- the window entities: ENABLE held high, DT_WIN / DT_VLD per clock
    def window_case(ent, gen, aa, tag):
        pw, dw = gen["PHI_WIDTH"], gen["DAT_WIDTH"]
        N = 1 << pw
        clocks = 2 * N + dw + 24
        out = V.run_window(lib, ent, gen, aa, clocks)
        win = np.array([o[0] for o in out], np.int64)
        vld = [o[1] for o in out]
        key = f"win/{ent}/pw{pw}_dw{dw}_{tag}"
        arrays[key + "/aa"] = np.array(aa, np.int64)
        arrays[key + "/dt_win_per_clock"] = win
        arrays[key + "/dt_vld_per_clock"] = np.array(vld, np.int8)
        cases["windows"].append({"entity": ent, "generics": gen, "key": key, "clocks": clocks, "first_dt_vld_clock": vld.index(1)})
        print(key, "first DT_VLD clock", vld.index(1), flush=True)

    ent_of = {2: "hamming_win", 3: "bh_win_3term", 4: "bh_win_4term", 5: "bh_win_5term", 7: "bh_win_7term"}
    shapes = {1: [(7, 16), (10, 16), (5, 12)], 2: [(7, 16), (8, 8)], 3: [(8, 16), (6, 24)], 4: [(8, 16)], 5: [(8, 17)], 6: [(8, 17), (10, 17), (6, 31)],
              7: [(7, 17)], 8: [(8, 24), (7, 32)], 9: [(8, 24), (9, 16)], 10: [(7, 32), (8, 24), (7, 16), (6, 40)],
              11: [(7, 32)], 12: [(7, 16)], 13: [(7, 24)]}
    for v, lst in shapes.items():
        for pw, dw in lst:
            aa, m = bhw.quantize(v, bhw.RULE_TB, dw)
            gen = {"PHI_WIDTH": pw, "DAT_WIDTH": dw}
            if m <= 3:
                gen["SIN_TYPE"] = "CORDIC"
            window_case(ent_of[m], gen, [int(a) for a in aa[:m]], f"variant{v}")
    # port edge cases: negative, most negative, all ones (unsigned reading), zeros
    for m, pw, dw in ((2, 6, 16), (3, 6, 12), (4, 6, 17), (5, 6, 24), (7, 6, 32), (4, 6, 32), (7, 6, 31)):
        lo, hi = -(1 << (dw - 1)), (1 << (dw - 1)) - 1
        gen = {"PHI_WIDTH": pw, "DAT_WIDTH": dw}
        if m <= 3:
            gen["SIN_TYPE"] = "CORDIC"
        for tag, aa in (("hi", [hi] * m), ("lo", [lo] * m), ("mixed", [lo if k & 1 else hi for k in range(m)]),
                        ("small", [-3, 5, -7, 11, -13, 17, -19][:m]), ("zero", [0] * m)):
            window_case(ent_of[m], gen, aa, tag)
    # the selector on top (src/win_selector.vhd:93-199)
    for wt, m, pw, dw, v in (("HAMMING", 2, 6, 16, 1), ("BH3TERM", 3, 6, 16, 4), ("BH4TERM", 4, 6, 17, 6), ("BH5TERM", 5, 6, 24, 9), ("BH7TERM", 7, 6, 32, 10)):
        aa, _ = bhw.quantize(v, bhw.RULE_TB, dw)
        window_case("win_selector", {"PHI_WIDTH": pw, "DAT_WIDTH": dw, "WIN_TYPE": wt, "SIN_TYPE": "CORDIC"}, [int(a) for a in aa[:7]], f"sel_{wt.lower()}")

    # ---- cordic_atan2: one pair per clock, PHI_DT / PHI_VL per clock
    for iw, aw, prec in ((16, 16, 1), (24, 24, 1), (15, 16, 2), (32, 32, 1), (12, 12, 3), (20, 12, 1), (31, 32, 7)):
        n = 160
        lim = 1 << (iw - 1)
        x = rng.integers(-lim, lim, n)
        y = rng.integers(-lim, lim, n)
        x[:8] = [0, lim - 1, -lim, 1, -1, 0, lim - 1, -lim]
        y[:8] = [0, 0, 0, -1, 1, lim - 1, lim - 1, -lim]
        out = V.run_atan2(lib, iw, aw, prec, [(int(a), int(b)) for a, b in zip(x, y)])
        key = f"atan2/iw{iw}_aw{aw}_p{prec}"
        arrays[key + "/x"] = x.astype(np.int64)
        arrays[key + "/y"] = y.astype(np.int64)
        arrays[key + "/phi_dt_per_clock"] = np.array([o[0] for o in out], np.int64)
        arrays[key + "/phi_vl_per_clock"] = np.array([o[1] for o in out], np.int8)
        cases["atan2"].append({"input_width": iw, "angle_width": aw, "precision": prec, "key": key})
        print(key, flush=True)

    # ---- taylor_sincos (src/taylor_sincos.vhd + tay1_order.vhd + mults/*.vhd, DSP48 primitives modelled by
    #      vhdl_sim.Dsp48): PHI_ENA held high, OUT_SIN / OUT_COS per clock; `start` = deposited phase counter
    cases["taylor"] = []
    for pw, dw, lut, xs, start, clocks in (
            (10, 16, 5, "ULTRA", 0, 1100), (10, 12, 5, "7SERIES", 0, 1100), (11, 18, 9, "ULTRA", 0, 2100), (10, 16, 9, "ULTRA", 0, 1100),
            (9, 16, 9, "ULTRA", 0, 600), (8, 8, 4, "ULTRA", 0, 300), (12, 16, 7, "ULTRA", 0, 4200), (12, 24, 6, "ULTRA", 0, 4200),
            (11, 24, 6, "7SERIES", 0, 2100), (11, 32, 5, "ULTRA", 0, 2100), (10, 19, 5, "ULTRA", 0, 1100), (10, 18, 5, "7SERIES", 0, 1100),
            (14, 16, 9, "ULTRA", 0, 700), (14, 16, 9, "ULTRA", (1 << 12) - 300, 700), (14, 24, 9, "ULTRA", (1 << 13) - 300, 700),
            (20, 24, 9, "ULTRA", 3 * (1 << 18) - 300, 700), (24, 16, 10, "ULTRA", (1 << 24) - 300, 700), (16, 17, 9, "7SERIES", (1 << 15) - 200, 500),
            (26, 16, 11, "ULTRA", (1 << 24) - 200, 500), (13, 32, 8, "ULTRA", (1 << 11) - 200, 500)):
        out = V.run_taylor(lib, pw, dw, lut, xs, clocks, start)
        key = f"taylor/pw{pw}_dw{dw}_lut{lut}_{xs.lower()}_s{start}"
        arrays[key + "/sin_per_clock"] = np.array([o[0] for o in out], np.int64)
        arrays[key + "/cos_per_clock"] = np.array([o[1] for o in out], np.int64)
        cases["taylor"].append({"phase_width": pw, "data_width": dw, "lut_size": lut, "xseries": xs, "start": start, "clocks": clocks, "key": key})
        print(key, flush=True)

    # ---- the window entities with SIN_TYPE = "TAYLOR"
    for ent, v, pw, dw, lut, xs in (("hamming_win", 1, 10, 16, 5, "ULTRA"), ("hamming_win", 1, 8, 16, 9, "ULTRA"), ("hamming_win", 2, 10, 24, 5, "ULTRA"),
                                    ("hamming_win", 1, 11, 32, 6, "7SERIES"), ("hamming_win", 12, 9, 12, 5, "7SERIES"), ("bh_win_3term", 3, 10, 16, 5, "ULTRA"),
                                    ("bh_win_3term", 4, 10, 24, 5, "7SERIES"), ("bh_win_3term", 3, 11, 12, 6, "ULTRA"), ("bh_win_3term", 4, 9, 32, 4, "ULTRA"),
                                    ("bh_win_3term", 3, 7, 16, 9, "ULTRA")):
        aa, m = bhw.quantize(v, bhw.RULE_TB, dw)
        window_case(ent, {"PHI_WIDTH": pw, "DAT_WIDTH": dw, "SIN_TYPE": "TAYLOR", "LUT_SIZE": lut, "XSERIES": xs}, [int(a) for a in aa[:m]], f"taylor_lut{lut}_{xs.lower()}_variant{v}")
    for wt, m, pw, dw, v, lut in (("HAMMING", 2, 9, 16, 1, 5), ("BH3TERM", 3, 9, 16, 4, 4)):
        aa, _ = bhw.quantize(v, bhw.RULE_TB, dw)
        window_case("win_selector", {"PHI_WIDTH": pw, "DAT_WIDTH": dw, "WIN_TYPE": wt, "SIN_TYPE": "TAYLOR", "LUT_SIZE": lut, "XSERIES": "ULTRA"},
                    [int(a) for a in aa[:7]], f"sel_{wt.lower()}_taylor_lut{lut}")

    # ---- int_multNxN_dsp48: DAT_Q per clock
    cases["mult"] = []
    for dtw in (8, 16, 17, 24, 32):
        lim = 1 << (dtw - 1)
        a = rng.integers(-lim, lim, 64)
        b = rng.integers(-lim, lim, 64)
        a[:6] = [lim - 1, -lim, -lim, 0, -1, lim - 1]
        b[:6] = [lim - 1, -lim, lim - 1, -lim, -1, -1]
        q = V.run_mult(lib, dtw, [(int(x), int(y)) for x, y in zip(a, b)])
        key = f"mult/dtw{dtw}"
        arrays[key + "/a"] = a.astype(np.int64)
        arrays[key + "/b"] = b.astype(np.int64)
        arrays[key + "/q_per_clock"] = np.array(q, np.int64)
        cases["mult"].append({"dtw": dtw, "key": key})
        print(key, flush=True)

    swapped_cases(lib, arrays, cases)
    tb_constant_cases(cases)
    precision_cases(lib, arrays, cases, np.random.default_rng(20261019))
    np.savez_compressed(os.path.join(HERE, "rtl_sim_vectors.npz"), **arrays)
    json.dump(cases, open(os.path.join(HERE, "rtl_sim_cases.json"), "w"), indent=1)
    print("wrote", len(arrays), "arrays")


if __name__ == "__main__":
    if sys.argv[1:] == ["swapped"]:
        update_only_swapped()
    elif sys.argv[1:] == ["tb"]:
        update_only_tb()
    elif sys.argv[1:] == ["precision"]:
        update_only_precision()
    else:
        main()
