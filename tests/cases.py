"""Shared descriptor sets for the parity tests (CPU hostcheck tier and GPU tier)."""
import blackman_harris_win_b200 as bhw

VARIANT_DW = {1: 16, 2: 16, 3: 16, 4: 16, 5: 17, 6: 17, 7: 17, 8: 24, 9: 24, 10: 32}

# BASELINE.json configs 1-4: name -> descriptor factory
def baseline_configs():
    return {
        "cfg1_hamming_n1024_dw16": bhw.make_desc(2, 10, 16, [17808, 14959]),
        "cfg1_hann_n1024_dw16": bhw.make_desc(2, 10, 16, [16384, 16384]),
        "cfg2_bh4_n65536_dw17": bhw.make_desc(4, 16, 17, [47022, 64001, 18518, 1531]),
        "cfg3_bh7_n1m_dw32_dds48": bhw.make_desc(
            7, 20, 32, [582441289, 930815217, 468160289, 141272949, 23110934, 1653590, 29379],
            sin_type=bhw.SIN_CORDIC48),
        "cfg3_bh7_n1m_dw32_dds": bhw.make_desc(
            7, 20, 32, [582441289, 930815217, 468160289, 141272949, 23110934, 1653590, 29379]),
        "cfg4_blackman_taylor_n16m_dw24": bhw.make_desc(
            3, 24, 24, [7046424, 8388600, 1342176], sin_type=bhw.SIN_TAYLOR, lut_size=9),
    }


def rtl_sweep(pws=(4, 5, 7, 10, 13), extra_dws=(8, 12, 20, 31, 32, 33, 47),
              sin_types=(bhw.SIN_CORDIC, bhw.SIN_CORDIC48, bhw.SIN_CORDIC_SCALED, bhw.SIN_TAYLOR)):
    """Every variant x phase width x data width x sin source the library accepts."""
    out = []
    for v in range(1, 11):
        for pw in pws:
            for dw in sorted(set((VARIANT_DW[v],) + tuple(extra_dws))):
                for st in sin_types:
                    d = bhw.variant_desc(v, pw, dw, sin_type=st)
                    if bhw.validate(d) == 0:
                        out.append(d)
    return out


def mterm_sweep(pws=(4, 6, 9, 12), dws=(8, 16, 17, 24, 31, 32, 33, 40),
                sin_types=(bhw.SIN_CORDIC, bhw.SIN_CORDIC48, bhw.SIN_CORDIC_SCALED)):
    """The 6- and 8..11-term extension (BHW_WIN_MTERM_*, quantize variants 14..18) over widths and CORDIC sources,
    plus port edge cases."""
    out = []
    for v in range(14, 19):
        for pw in pws:
            for dw in dws:
                for st in sin_types:
                    d = bhw.variant_desc(v, pw, dw, sin_type=st)
                    if bhw.validate(d) == 0:
                        out.append(d)
    for m in (6, 8, 9, 10, 11):
        for dw in (12, 24, 32):
            lo, hi = -(1 << (dw - 1)), (1 << (dw - 1)) - 1
            for aa in ([hi] * m, [lo] * m, [lo if k & 1 else hi for k in range(m)], [(-1) ** k * (k + 3) for k in range(m)], [0] * m):
                out.append(bhw.make_desc(m, 8, dw, aa))
                out.append(bhw.make_desc(m, 8, dw, aa, stream_offset=1))
    return out


def hls_sweep(cfgs):
    out = []
    for (np_, nw) in cfgs:
        for t, v in HLS_TYPES.items():
            out.append((t, bhw.variant_desc(v, np_, nw, model=bhw.MODEL_HLS)))
    return out


HLS_TYPES = {1: 1, 2: 2, 3: 3, 4: 6, 5: 9, 7: 10}  # win_function win_type -> README variant


def edge_coeff_descs():
    """Coefficient edge cases: negative, unsigned reading of the port bits, the most negative
    value (b_k can wrap), zeros, all-ones."""
    out = []
    for dw in (8, 16, 17, 24, 30, 31, 32):
        lo, hi = -(1 << (dw - 1)), (1 << (dw - 1)) - 1
        for m in (2, 3, 4, 5, 7):
            out.append(bhw.make_desc(m, 9, dw, [hi] * m))
            out.append(bhw.make_desc(m, 9, dw, [lo] * m))
            out.append(bhw.make_desc(m, 9, dw, [lo if k & 1 else hi for k in range(m)]))
            out.append(bhw.make_desc(m, 9, dw, [(1 << dw) - 1] * m))   # raw bits, unsigned reading
            out.append(bhw.make_desc(m, 9, dw, [0] * m))
            out.append(bhw.make_desc(m, 9, dw, [-3, 5, -7, 11, -13, 17, -19][:m]))
    # DAT_WIDTH 31..32: ports around the limits of the 32-bit tail (|AAk| < 2^30 at DW 32, and the
    # bound on the sum that keeps dsp_pp inside an int32) - both sides of each limit
    q = 1 << 30
    for dw in (31, 32):
        for m in (2, 3, 4, 5, 7):
            for a0 in (0, q - 1, -q, (1 << (dw - 1)) - 1):
                for ak in (q - 1, q, -q + 1, -q, q // 2, q // 3, -(q // 5)):
                    if -(1 << (dw - 1)) <= ak < (1 << (dw - 1)):
                        out.append(bhw.make_desc(m, 9, dw, [a0] + [ak] * (m - 1)))
            out.append(bhw.make_desc(m, 9, dw, [q + q // 2] + [q - 1] + [q // 8] * (m - 2)))
    return out


def random_descs(count, seed, max_pw=12, models=("rtl", "hls"), terms=(2, 3, 4, 5, 7)):
    """Seeded random descriptors over everything the library accepts: entity, widths, source, model,
    stream offset, strategy, and ports drawn from {full-range random, small, real-window-like,
    extremes} so that every tail route (32-bit, bounded 32-bit at DAT_WIDTH 31..32, 64-bit, generic)
    is hit."""
    import random
    rng = random.Random(seed)
    out = []
    while len(out) < count:
        m = rng.choice(terms)
        pw = rng.randint(4, max_pw)
        model = rng.choice(models)
        if model == "hls":
            dw = rng.choice((8, 12, 16, 17, 24, 30, 31, 32))
            st, mdl = bhw.SIN_CORDIC, bhw.MODEL_HLS
        else:
            dw = rng.choice((8, 9, 12, 16, 17, 18, 19, 24, 30, 31, 32, 33, 40, 47))
            st, mdl = rng.choice((bhw.SIN_CORDIC, bhw.SIN_CORDIC, bhw.SIN_CORDIC48, bhw.SIN_CORDIC_SCALED, bhw.SIN_TAYLOR)), bhw.MODEL_RTL
        lo, hi = -(1 << (dw - 1)), (1 << (dw - 1)) - 1
        kind = rng.choice(("full", "small", "window", "extreme", "quarter"))
        if kind == "full":
            aa = [rng.randint(lo, hi) for _ in range(m)]
        elif kind == "small":
            aa = [rng.randint(-1000, 1000) for _ in range(m)]
        elif kind == "window":       # positive, decaying, summing to about the CORDIC full scale
            aa = [int(hi * 0.45 / (k + 1) ** 1.5) + rng.randint(-3, 3) for k in range(m)]
        elif kind == "quarter":      # around the 2^(DW-2) limit of the bounded 32-bit tail
            q = 1 << (dw - 2)
            aa = [rng.choice((q - 1, q, -q, -q + 1, q // 2, 0)) for _ in range(m)]
        else:
            aa = [rng.choice((lo, hi, 0, -1, 1, lo + 1)) for _ in range(m)]
        d = bhw.make_desc(m, pw, dw, aa, sin_type=st, model=mdl, stream_offset=rng.randint(0, 1),
                          lut_size=rng.choice((0, 0, 4, 7, 9)) if st == bhw.SIN_TAYLOR else 0,
                          algo=rng.choice((bhw.ALGO_AUTO, bhw.ALGO_AUTO, bhw.ALGO_TABLE, bhw.ALGO_DIRECT)))
        if bhw.validate(d) == 0:
            out.append(d)
    return out
