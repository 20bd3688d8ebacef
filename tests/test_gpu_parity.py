"""GPU tier (-m gpu): the CUDA path, called through the C ABI, against the oracle - bit-exact.
Small/medium sizes compare every sample; the BASELINE full sizes are covered by (i) the oracle on
all host threads where that takes seconds, (ii) the compiled reference models (oracle/_ref), and
(iii) size-independent properties: the two independent evaluation strategies agree, any split of
the range reproduces the whole, batches equal their windows, the DT_VLD order is a rotation."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import blackman_harris_win_b200 as bhw
import cases
import harness as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def gpu_window(d, n0=0, count=None):
    return bhw.generate(d, n0, count).cpu().numpy().astype(np.int64)


def both_algos(d):
    yield "auto", d
    yield "direct", d.copy(algo=bhw.ALGO_DIRECT)
    if d.dat_width <= 32:
        yield "table", d.copy(algo=bhw.ALGO_TABLE)


def test_library_is_the_cuda_one():
    import torch
    assert torch.cuda.is_available()
    assert os.path.exists(bhw.lib_path())
    n0 = bhw.launch_count()
    gpu_window(bhw.make_desc(2, 10, 16, [17808, 14959]))
    assert bhw.launch_count() > n0


def test_rtl_sweep_all_variants_widths_sources():
    """Every variant x PHI_WIDTH x DAT_WIDTH x sin source, both strategies, every sample."""
    descs = cases.rtl_sweep(pws=(4, 5, 7, 10, 13))
    for d in descs:
        want = H.orc_window(d)
        for name, dd in both_algos(d):
            assert np.array_equal(gpu_window(dd), want), (name, d)


def test_hls_model_vs_oracle_and_compiled_reference():
    cfgs = H.ref_configs("hls_win")
    assert cfgs, "oracle/_ref is missing (it is built by __graft_entry__.build() and ships with the repo)"
    for (np_, nw) in cfgs:
        for t, v in cases.HLS_TYPES.items():
            d = bhw.variant_desc(v, np_, nw, model=bhw.MODEL_HLS)
            n = 1 << np_
            cnt = min(n, 1 << 15)
            ref = H.ref_hls_window(np_, nw, t, 0, cnt)
            for name, dd in both_algos(d):
                assert np.array_equal(gpu_window(dd, 0, cnt), ref), (name, np_, nw, t)
            # the bhw.win_function mirror of the HLS entry point
            assert np.array_equal(bhw.win_function(t, np_, nw, 0, cnt).cpu().numpy(), ref)


def test_reference_hashes_full_size():
    kat = json.load(open(os.path.join(GOLD, "reference_kat.json")))
    for e in kat["hls_win"]:
        d = bhw.variant_desc(cases.HLS_TYPES[e["type"]], e["nphase"], e["nwidth"], model=bhw.MODEL_HLS)
        w = gpu_window(d)
        assert [int(x) for x in w[:4]] == e["first"] and int(w[len(w) // 2]) == e["mid"]
        if e["nphase"] <= 16 or (e["nphase"], e["nwidth"], e["type"]) == (20, 32, 7):
            assert H.sha_lines(w) == e["sha256"], e


def test_rtl_anchor_hashes():
    rtl = json.load(open(os.path.join(GOLD, "rtl_kat.json")))
    for e in rtl["windows"]:
        d = bhw.make_desc(e["win_type"], e["phi_width"], e["dat_width"], e["aa"], sin_type=e["sin_type"],
                          lut_size=e["lut_size"])
        for name, dd in both_algos(d):
            assert H.sha_lines(gpu_window(dd)) == e["sha256"], (name, e["name"])


@pytest.mark.parametrize("name", list(cases.baseline_configs()))
def test_baseline_configs_full_size(name):
    """BASELINE.json configs 1-4 at their full sizes, every sample, oracle on all host threads."""
    d = cases.baseline_configs()[name]
    n = 1 << d.phi_width
    want = H.orc_window(d, 0, n, threads=os.cpu_count() or 8)
    got_auto = gpu_window(d)
    assert np.array_equal(got_auto, want)
    # the other strategy must agree too (direct is slow for 64-bit-limb sources: subsample there)
    dd = d.copy(algo=bhw.ALGO_DIRECT)
    if n <= 1 << 20:
        assert np.array_equal(gpu_window(dd), want)
    else:
        for n0 in (0, n // 4 - 1000, n // 2 - 4096, n - 8192):
            assert np.array_equal(gpu_window(dd, n0, 8192), want[n0:n0 + 8192])
    # DT_VLD-gated order = rotation by one
    assert np.array_equal(gpu_window(d.copy(stream_offset=1)), np.roll(want, -1))


def test_large_windows_properties():
    """N = 2^24 .. 2^26: strategies agree, splits reproduce the whole, spot ranges match the oracle."""
    import torch
    for v, pw, dw, st in ((1, 26, 16, bhw.SIN_CORDIC), (6, 25, 17, bhw.SIN_CORDIC), (9, 24, 24, bhw.SIN_CORDIC),
                          (10, 24, 32, bhw.SIN_CORDIC), (10, 23, 32, bhw.SIN_CORDIC48)):
        d = bhw.variant_desc(v, pw, dw, sin_type=st)
        n = 1 << pw
        full = bhw.generate(d)
        # spot ranges against the oracle; for the 7-term windows also around the seams between the
        # shares of the 30 warps a CTA spreads over the window (tile U*j/30 of U tiles of 256 samples,
        # or of U tile pairs (n, n + N/2) for the paired sources)
        spots = [0, 12345, n // 2 - 2048, n - 4096]
        if d.win_type == 7:
            for tiles in (n // 256, n // 512):
                spots += [max(0, (tiles * j // 30) * 256 - 2048) for j in (1, 7, 16, 29)]
        for n0 in spots:
            assert np.array_equal(full[n0:n0 + 4096].cpu().numpy().astype(np.int64), H.orc_window(d, n0, 4096))
        # ragged 3-way split equals the whole
        cuts = [0, n // 3 + 1, n // 3 + 1 + 777, n]
        parts = [bhw.generate(d, cuts[i], cuts[i + 1] - cuts[i]) for i in range(3)]
        assert torch.equal(torch.cat(parts), full)
        # direct strategy on sub-ranges
        dd = d.copy(algo=bhw.ALGO_DIRECT)
        for n0 in (0, n - 65536):
            assert torch.equal(bhw.generate(dd, n0, 65536), full[n0:n0 + 65536])
        # checksum-of-checksums identity between two layouts of the same data
        assert int(full.to(torch.int64).sum()) == sum(int(p.to(torch.int64).sum()) for p in parts)
        del full, parts


def test_phi_width_beyond_the_reference_range_is_rejected():
    """The reference documents 16 .. 64M points (README.md:2): PHI_WIDTH 27 and up is an argument error on
    every entry point, not an untested code path."""
    d = bhw.variant_desc(6, 26, 17)
    d27 = d.copy(phi_width=27)
    assert bhw.validate(d) == 0 and bhw.validate(d27) == -5
    with pytest.raises(bhw.BhwError):
        bhw.generate(d27, 0, 16)
    with pytest.raises(bhw.BhwError):
        bhw.Plan([d, d27])
    with pytest.raises(bhw.BhwError):
        bhw.sincos(d27, 0, 16)


def test_bench_bank_shape_against_the_oracle():
    """The instantiation bench.py times for BASELINE config 2 - k_synth_bank<4, staged half period, paired> on
    bh_win_4term, PHI_WIDTH 16, DAT_WIDTH 17, every window with its own AA0..AA3 - sample by sample against
    the oracle (src/bh_win_4term.vhd:225-280), whole bank and a ragged sub-range of it."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    nwin = 256
    arr = bench.bank_descs(nwin)
    descs = [bhw.BhwDesc.from_buffer_copy(bytes(arr[i])) for i in range(nwin)]
    assert len({tuple(d.aa) for d in descs}) == nwin
    want = H.orc_batch(descs, 0, nwin << 16)
    n0 = bhw.launch_count()
    plan = bhw.Plan(descs)
    bhw.timing_enable(True)
    bhw.timing_reset()
    got = plan.execute().cpu().numpy().astype(np.int64)
    kt = bhw.timing_read()
    bhw.timing_enable(False)
    assert kt["k_synth_bank"][0] == 1 and kt["k_synth"][0] == 0 and kt["k_synth_group"][0] == 0   # the bench kernel ran
    assert np.array_equal(got, want)
    b, c = 3 * 65536 + 1001, 200 * 65536 + 4444
    assert np.array_equal(plan.execute(b, c).cpu().numpy().astype(np.int64), want[b:b + c])
    plan.destroy()
    assert bhw.launch_count() > n0


@pytest.mark.parametrize("m", [2, 3, 4, 5, 7])
def test_bank_kernel_every_placement_against_the_oracle(m):
    """k_synth_bank on uniform banks, one instantiation per table placement and tail per entity: table staged
    whole (short windows), staged half period (N = 65536 at DAT_WIDTH 16), read from L2 (DAT_WIDTH 24), the
    64-bit tail (DAT_WIDTH 32 with full-scale ports), the input-quadrant CORDICs (paired through the ones'-complement
    relation, exceptions patched) and TAYLOR."""
    v = {2: 1, 3: 4, 4: 6, 5: 9, 7: 10}[m]
    hi = (1 << 31) - 1
    shapes = [("staged full", bhw.variant_desc(v, 12, 16), 64),
              ("staged half", bhw.variant_desc(v, 16, 16), 4),
              ("global", bhw.variant_desc(v, 17, 24), 2),
              ("64-bit tail", bhw.make_desc(m, 13, 32, [hi - 3 * k for k in range(m)]), 32),
              ("cordic_dds48, paired by ones' complement + patch pass", bhw.variant_desc(v, 13, 24, sin_type=bhw.SIN_CORDIC48), 32),
              ("cordic_dds_scaled, paired, table in L2", bhw.variant_desc(v, 17, 17, sin_type=bhw.SIN_CORDIC_SCALED), 3),
              ("cordic_dds48 DW 32, table staged whole", bhw.variant_desc(v, 12, 32, sin_type=bhw.SIN_CORDIC48), 40)]
    if m <= 3:
        shapes.append(("taylor", bhw.variant_desc(v, 14, 24, sin_type=bhw.SIN_TAYLOR), 16))
    for name, d, nwin in shapes:
        descs = [d.copy(aa=[int(a) - (5 * i + k) % 97 if k < m else 0 for k, a in enumerate(d.aa)], stream_offset=i & 1)
                 for i in range(nwin)]
        want = H.orc_batch(descs, 0, nwin << d.phi_width)
        plan = bhw.Plan(descs)
        bhw.timing_enable(True)
        bhw.timing_reset()
        got = plan.execute().cpu().numpy().astype(np.int64)
        kt = bhw.timing_read()
        bhw.timing_enable(False)
        plan.destroy()
        assert kt["k_synth_bank"][0] >= 1, (name, kt)
        assert np.array_equal(got, want), (m, name)


def test_input_quadrant_cordic_banks_pair_through_ones_complement():
    """cordic_dds48 / cordic_dds_scaled banks of >= 2^23 samples: k_synth_bank pairs samples through
    T[i + E/2] == ~T[i], k_inq_exceptions lists the entries where that fails and k_inq_patch recomputes the pairs
    that read them (src/cordic_dds48.vhd:170-258) - every sample against the oracle."""
    for v, pw, dw, st, nwin in ((1, 17, 24, bhw.SIN_CORDIC48, 64), (6, 17, 17, bhw.SIN_CORDIC_SCALED, 64),
                                (10, 14, 32, bhw.SIN_CORDIC48, 512)):
        base = bhw.variant_desc(v, pw, dw, sin_type=st)
        descs = [base.copy(aa=[int(a) - (3 * i + k) % 61 for k, a in enumerate(base.aa)], stream_offset=i & 1) for i in range(nwin)]
        plan = bhw.Plan(descs)
        bhw.timing_enable(True)
        bhw.timing_reset()
        got = plan.execute().cpu().numpy().astype(np.int64)
        recs = bhw.timing_launches()
        bhw.timing_enable(False)
        plan.destroy()
        banks = [r for r in recs if r["kernel"] == "k_synth_bank"]
        assert len(banks) == 1 and banks[0]["paired"] == 1, recs
        assert sum(r["kernel"] == "k_synth" for r in recs) == 1          # the patch pass
        want = np.concatenate([H.orc_window(d, threads=8) for d in descs])
        assert np.array_equal(got, want), (v, st, int(np.argmax(got != want)))


def group_descs(variants, dw, pws, model=0):
    out = []
    for i, pw in enumerate(pws):
        for v in variants:
            d = bhw.variant_desc(v, pw, dw, model=model)
            out.append(d.copy(stream_offset=(i + v) & 1, aa=[int(a) - (i if k == 0 else 0) for k, a in enumerate(d.aa)]))
    return out


@pytest.mark.parametrize("variants,dw,pws", [((1, 2), 16, (4, 9, 12, 15, 16, 17, 19)), ((3, 4), 16, (10, 16, 18, 20)),
                                             ((5, 6, 7), 17, (8, 9, 13, 17, 18, 20)), ((8, 9), 24, (9, 14, 18, 19, 20)),
                                             ((10,), 32, (9, 12, 17, 19, 20)), ((10,), 24, (11, 18))])
def test_group_kernel_mixed_phi_widths_against_the_oracle(variants, dw, pws):
    """k_synth_group: windows of one family with different PHI_WIDTHs, entities' ports and stream offsets in one
    plan - every sample against the oracle; then ragged sub-ranges that cut windows (unpaired tile ranges,
    general-kernel ends reading the pyramid), on a side stream, and the same batch one-shot."""
    import torch
    descs = group_descs(variants, dw, pws)
    total = bhw.batch_total(descs)
    want = H.orc_batch(descs, 0, total)
    plan = bhw.Plan(descs)
    bhw.timing_enable(True)
    bhw.timing_reset()
    got = plan.execute().cpu().numpy().astype(np.int64)
    kt = bhw.timing_read()
    bhw.timing_enable(False)
    assert kt["k_synth_group"][0] >= 1, kt
    assert np.array_equal(got, want), int(np.argmax(got != want))
    rng = np.random.default_rng(dw * 100 + len(pws))
    s = torch.cuda.Stream()
    for _ in range(6):
        b = int(rng.integers(0, total - 1))
        c = int(rng.integers(1, total - b + 1))
        with torch.cuda.stream(s):
            part = plan.execute(b, c)
        s.synchronize()
        assert np.array_equal(part.cpu().numpy().astype(np.int64), want[b:b + c]), (b, c)
    plan.destroy()
    assert np.array_equal(bhw.generate_batch(descs).cpu().numpy().astype(np.int64), want)
    b, c = total // 3 + 5, total // 2
    assert np.array_equal(bhw.generate_batch(descs, b, c).cpu().numpy().astype(np.int64), want[b:b + c])


def test_group_kernel_hls_family_and_mixed_families_in_one_plan():
    """HLS-model families (tshift 2, floor products) through the group kernel, and a plan that mixes several
    families, entities, a TAYLOR window, a cordic_dds48 window and a uniform bank: every launch kind at once."""
    hls = [bhw.variant_desc(v, pw, 17, model=bhw.MODEL_HLS) for v in (1, 3, 6) for pw in (9, 14, 17, 19)]
    want = H.orc_batch(hls, 0, bhw.batch_total(hls))
    assert np.array_equal(bhw.generate_batch(hls).cpu().numpy().astype(np.int64), want)
    mix = (group_descs((1, 3, 6), 16, (10, 14)) + [bhw.variant_desc(4, 14, 24, sin_type=bhw.SIN_TAYLOR, lut_size=9)]
           + group_descs((9, 10), 24, (12, 18)) + [bhw.variant_desc(10, 13, 32, sin_type=bhw.SIN_CORDIC48)]
           + [bhw.variant_desc(6, 16, 17).copy(aa=[47022 - i, 64001, 18518, 1531]) for i in range(4)]
           + [bhw.variant_desc(1, 14, 16, sin_type=bhw.SIN_TAYLOR, lut_size=9), bhw.variant_desc(2, 6, 16)]
           + group_descs((5,), 17, (9, 19)))
    total = bhw.batch_total(mix)
    want = H.orc_batch(mix, 0, total)
    plan = bhw.Plan(mix)
    bhw.timing_enable(True)
    bhw.timing_reset()
    got = plan.execute().cpu().numpy().astype(np.int64)
    kt = bhw.timing_read()
    bhw.timing_enable(False)
    assert kt["k_synth_group"][0] >= 4 and kt["k_synth_bank"][0] >= 1 and kt["k_synth"][0] >= 1, kt
    assert np.array_equal(got, want)
    for b, c in ((12345, total - 99999), (total // 2 + 77, 4096), (0, 1 << 14)):
        assert np.array_equal(plan.execute(b, c).cpu().numpy().astype(np.int64), want[b:b + c])
    plan.destroy()


def test_coefficient_edge_cases():
    for d in cases.edge_coeff_descs():
        want = H.orc_window(d)
        for name, dd in both_algos(d):
            assert np.array_equal(gpu_window(dd), want), (name, d)


def test_taylor_modes():
    for pw, lut in [(6, 6), (6, 5), (8, 6), (10, 7), (14, 9), (16, 9), (20, 9), (24, 9), (26, 10), (12, 1), (18, 16)]:
        for dw in (8, 16, 18, 19, 24, 32):
            for wt, aa in ((2, [100, 77]), (3, [90, 100, 17])):
                amp = (1 << (dw - 1)) - 1
                d = bhw.make_desc(wt, pw, dw, [a * amp // 128 for a in aa], sin_type=bhw.SIN_TAYLOR, lut_size=lut)
                if bhw.validate(d):
                    continue
                n = 1 << pw
                cnt = min(n, 1 << 14)
                n0 = (n - cnt) // 2
                want = H.orc_window(d, n0, cnt)
                for name, dd in both_algos(d):
                    assert np.array_equal(gpu_window(dd, n0, cnt), want), (name, d)


def test_sincos_all_sources():
    for model, st, pw, dw, prec in [(0, 0, 10, 16, 0), (0, 0, 12, 8, 3), (0, 0, 20, 32, 1), (0, 0, 9, 47, 2),
                                    (0, 2, 10, 16, 0), (0, 2, 14, 47, 0), (0, 3, 10, 8, 0), (0, 3, 16, 32, 0),
                                    (0, 3, 26, 12, 0), (0, 1, 14, 16, 0), (0, 1, 16, 24, 0), (1, 0, 10, 16, 0),
                                    (1, 0, 18, 16, 0), (2, 0, 14, 12, 0), (2, 0, 10, 32, 0)]:
        d = bhw.make_desc(2, pw, dw, sin_type=st, model=model, precision=prec)
        cnt = min(1 << pw, 1 << 15)
        s, c = bhw.sincos(d, 0, cnt)
        ws, wc = H.orc_sincos(d, 0, cnt)
        assert np.array_equal(s.cpu().numpy().astype(np.int64), ws), d
        assert np.array_equal(c.cpu().numpy().astype(np.int64), wc), d
    # the cpp model against the compiled cpp/cordic_sincos.cpp itself
    for (pw, dw) in H.ref_configs("cpp"):
        cnt = min(1 << pw, 1 << 15)
        s, c = bhw.sincos(bhw.make_desc(2, pw, dw, model=bhw.MODEL_CPP), 0, cnt)
        rs, rc = H.ref_cpp_cordic(pw, dw, 0, cnt)
        assert np.array_equal(s.cpu().numpy(), rs) and np.array_equal(c.cpu().numpy(), rc)


def test_batch_mixed_windows_and_ragged_ranges():
    """win_selector sweep in miniature: all 10 variants x PHI_WIDTH 4..12 in one batch, read back
    in ragged flat ranges (ranges that start/end mid-window, windows shorter than a tile)."""
    descs = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 13)]
    total = bhw.batch_total(descs)
    want = H.orc_batch(descs, 0, total)
    got = bhw.generate_batch(descs).cpu().numpy().astype(np.int64)
    assert np.array_equal(got, want)
    rng = np.random.default_rng(7)
    for _ in range(12):
        b = int(rng.integers(0, total))
        c = int(rng.integers(0, total - b + 1))
        assert np.array_equal(bhw.generate_batch(descs, b, c).cpu().numpy().astype(np.int64), want[b:b + c])
    assert bhw.generate_batch(descs, 5, 0).numel() == 0
    # plans: execute == one-shot, repeatable, with and without the table cache
    plan = bhw.Plan(descs)
    for cache in (True, False, True):
        bhw.set_table_cache(cache)
        for _ in range(2):
            assert np.array_equal(plan.execute().cpu().numpy().astype(np.int64), want)
        assert np.array_equal(plan.execute(1000, 5000).cpu().numpy().astype(np.int64), want[1000:6000])
    plan.destroy()
    bhw.set_table_cache(True)


def test_batch_windows_sharing_one_table_at_different_phi_width():
    """PHI_WIDTH > DAT_WIDTH: cordic_dds only sees the top DAT_WIDTH phase bits, so windows of
    different lengths share one trig table and differ in how many phase bits they drop."""
    for dw, m, v in ((8, 2, 1), (10, 4, 6), (12, 7, 10), (9, 3, 3)):
        descs = [bhw.variant_desc(v, pw, dw) for pw in (6, 9, 10, 11, 12, 13, 14, 15, 9, 16)]
        total = bhw.batch_total(descs)
        want = H.orc_batch(descs, 0, total)
        assert np.array_equal(bhw.generate_batch(descs).cpu().numpy().astype(np.int64), want), (dw, m)
        plan = bhw.Plan(descs)
        off = 0
        for d in descs:      # window by window through the plan
            n = 1 << d.phi_width
            assert np.array_equal(plan.execute(off, n).cpu().numpy().astype(np.int64), want[off:off + n]), d
            off += n
        plan.destroy()
    # HLS model, NPHASE up to NWIDTH + 2
    descs = [bhw.variant_desc(6, np_, 12, model=bhw.MODEL_HLS) for np_ in (8, 11, 12, 13, 14)]
    assert np.array_equal(bhw.generate_batch(descs).cpu().numpy().astype(np.int64),
                          H.orc_batch(descs, 0, bhw.batch_total(descs)))


def test_win_selector_sweep_all_variants_all_lengths():
    """BASELINE config 5: all 10 variants x PHI_WIDTH 4..26 in one plan (1.34 G samples, 5.4 GB);
    checked against the per-window one-shot path on the long windows (an independent route through
    the planner) and against the oracle on every window up to 2^14 samples."""
    import torch
    descs = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in range(4, 27)]
    plan = bhw.Plan(descs)
    out = plan.execute()
    off = 0
    for d in descs:
        n = 1 << d.phi_width
        if n <= 1 << 14:
            assert np.array_equal(out[off:off + n].cpu().numpy().astype(np.int64), H.orc_window(d)), d
        elif d.phi_width in (15, 17, 20, 22, 26):
            assert torch.equal(out[off:off + n], bhw.generate(d)), d
            assert np.array_equal(out[off + n - 4096:off + n].cpu().numpy().astype(np.int64), H.orc_window(d, n - 4096, 4096)), d
        off += n
    # the launches of one execute are fanned out over side streams by default; everything on the
    # caller's stream, and a ragged sub-range on a non-default stream, must give the same samples
    bhw.set_side_streams(0)
    try:
        assert torch.equal(plan.execute(), out)
    finally:
        bhw.set_side_streams(4)
    b, c = (1 << 20) + 12, bhw.batch_total(descs) - (1 << 22)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        part = plan.execute(b, c)
    s.synchronize()
    assert torch.equal(part, out[b:b + c])
    plan.destroy()


def test_batch_int64_windows():
    descs = [bhw.variant_desc(v, pw, 40) for v in (1, 6, 10) for pw in (4, 9)]
    total = bhw.batch_total(descs)
    want = H.orc_batch(descs, 0, total)
    got = bhw.generate_batch(descs)
    assert got.dtype.itemsize == 8 and np.array_equal(got.cpu().numpy(), want)
    assert np.array_equal(bhw.generate_batch(descs, 17, 600).cpu().numpy(), want[17:617])
    with pytest.raises(bhw.BhwError):
        bhw.generate_batch(descs + [bhw.variant_desc(1, 8, 16)])       # mixed element sizes


def test_long_int64_windows_whole_and_in_pieces():
    """DAT_WIDTH > 32 windows long enough for the whole-window route of k_direct_window (four samples
    n + r*N/4 per evaluation): whole window == the same window in unpaired pieces == the oracle on slices,
    one-shot and through a plan, with and without the stream offset."""
    import torch
    for v, pw, dw, st, off in ((10, 18, 40, bhw.SIN_CORDIC, 0), (6, 17, 47, bhw.SIN_CORDIC, 1), (3, 18, 33, bhw.SIN_CORDIC, 1),
                               (10, 17, 36, bhw.SIN_CORDIC48, 0)):
        d = bhw.variant_desc(v, pw, dw, sin_type=st, stream_offset=off)
        n = 1 << pw
        whole = bhw.generate(d)
        assert whole.dtype.itemsize == 8
        pieces = torch.cat([bhw.generate(d, 0, n // 3), bhw.generate(d, n // 3, n - n // 3)])
        assert torch.equal(whole, pieces), (v, pw, dw, st)
        for n0 in (0, n // 4 - 500, n // 2 - 500, 3 * n // 4 - 500, n - 1000):
            assert np.array_equal(whole[n0:n0 + 1000].cpu().numpy(), H.orc_window(d, n0, 1000)), (v, pw, dw, st, n0)
        plan = bhw.Plan([d, d.copy(stream_offset=1 - off)])
        out = plan.execute()
        assert torch.equal(out[:n], whole)
        assert torch.equal(out[n:], torch.roll(whole, -1 if off == 0 else 1))
        plan.destroy()


def test_shards_reassemble():
    """1/2/4/8-way sharding by flat sample range (what each rank of bench.py does) reproduces the
    unsharded batch; shards are planned from their own windows only."""
    import torch
    descs = [bhw.variant_desc(6, 12, 17).copy(aa=[47022 - i, 64001 - i, 18518 + i, 1531 + i]) for i in range(37)]
    total = bhw.batch_total(descs)
    full = bhw.generate_batch(descs)
    assert np.array_equal(full.cpu().numpy().astype(np.int64), H.orc_batch(descs, 0, total))
    arr = bhw.desc_array(descs)
    for world in (1, 2, 4, 8):
        parts = []
        for r in range(world):
            b, c = bhw.shard_range(total, r, world)
            first, touched, local = bhw.shard_windows(arr, b, c)
            parts.append(bhw.generate_batch(descs[first:first + touched], local, c))
        assert torch.equal(torch.cat(parts), full)


def test_cost_balanced_shards_reassemble_and_cut_long_windows():
    """bhw_shard_range_cost on a mixed batch: the cuts fall inside long windows, so each shard runs
    whole tiles of a cut window through the bank kernel (tile range of one window, unpaired shape)
    and the ragged ends through the general kernel; the shards must reassemble to the unsharded
    batch, which is checked against the oracle on its short windows."""
    import torch
    descs = [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in range(1, 11) for pw in (6, 13, 19, 21)]
    arr = bhw.desc_array(descs)
    total = bhw.batch_total(descs)
    full = bhw.generate_batch(descs)
    off = 0
    for d in descs:
        n = 1 << d.phi_width
        if n <= 1 << 13:
            assert np.array_equal(full[off:off + n].cpu().numpy().astype(np.int64), H.orc_window(d)), d
        off += n
    for world in (2, 3, 8):
        parts = []
        for r in range(world):
            b, c = bhw.shard_range_cost(arr, r, world)
            if c == 0:          # a batch dominated by one costly window may leave a rank without work
                continue
            first, touched, local = bhw.shard_windows(arr, b, c)
            parts.append(bhw.generate_batch(descs[first:first + touched], local, c))
        assert torch.equal(torch.cat(parts), full), world
    # ranges that cut one long window of every entity at odd places (plan route)
    for v in (1, 3, 6, 8, 10):
        d = bhw.variant_desc(v, 22, cases.VARIANT_DW[v])
        whole = bhw.generate(d)
        plan = bhw.Plan([d])
        for b, c in ((100, (1 << 19) - 200), ((1 << 21) - 12345, (1 << 20) + 777), ((1 << 22) - 70000, 70000)):
            assert torch.equal(plan.execute(b, c), whole[b:b + c]), (v, b, c)
        plan.destroy()


def test_banks_over_a_table_too_large_for_shared_memory():
    """Runs of same-shape windows whose trig table stays in L2 (TAB_GLOBAL): the bank kernel walks
    them window-minor (same tile of consecutive windows side by side).  Every window must equal its
    one-shot generation, and the oracle on a slice."""
    import torch
    for v, pw, dw, st, nwin in ((10, 18, 32, bhw.SIN_CORDIC, 5), (10, 17, 32, bhw.SIN_CORDIC48, 3),
                                (8, 19, 24, bhw.SIN_CORDIC, 4), (6, 18, 17, bhw.SIN_CORDIC_SCALED, 33),
                                (1, 19, 24, bhw.SIN_CORDIC, 7)):
        base = bhw.variant_desc(v, pw, dw, sin_type=st)
        descs = [base.copy(aa=[int(a) - 3 * i if k == 0 else int(a) + (i if k == 1 else 0) for k, a in enumerate(base.aa)],
                           stream_offset=i & 1) for i in range(nwin)]
        plan = bhw.Plan(descs)
        out = plan.execute()
        n = 1 << pw
        for i, d in enumerate(descs):
            assert torch.equal(out[i * n:(i + 1) * n], bhw.generate(d)), (v, pw, dw, st, i)
        i = nwin - 1
        assert np.array_equal(out[i * n + n // 2 - 1000:i * n + n // 2 + 1000].cpu().numpy().astype(np.int64),
                              H.orc_window(descs[i], n // 2 - 1000, 2000))
        # a sub-range that starts and ends inside windows of the run
        b, c = n // 3 + 5, (nwin - 1) * n
        assert torch.equal(plan.execute(b, c), out[b:b + c])
        plan.destroy()


def test_reference_selfcheck_criteria_on_the_gpu():
    """The reference's own pass/fail criteria (hls/cordic/cordic_test.cpp:66-93,
    hls/windows/window_test.cpp:93-216) applied to what the CUDA path generates at the reference's
    widths - on top of, not instead of, the bit-exact comparisons above."""
    import test_reference_selfchecks as R
    s, c = bhw.sincos(bhw.variant_desc(1, 10, 16, model=bhw.MODEL_HLS))
    es, ec = R.cordic_selfcheck_errors(s.cpu().numpy().astype(np.int64), c.cpu().numpy().astype(np.int64), 10, 16)
    assert es < 10 and ec < 10
    for win_type in sorted(R.SELFCHECK):
        for algo in (bhw.ALGO_AUTO, bhw.ALGO_DIRECT, bhw.ALGO_TABLE):
            d = bhw.variant_desc(cases.HLS_TYPES[win_type], 10, 24, model=bhw.MODEL_HLS, algo=algo)
            out = bhw.generate(d).cpu().numpy().astype(np.int64)
            if win_type == 2:        # the Hann centre sample wraps in the reference model itself
                assert out[512] == -(1 << 23)
                out[512] = (1 << 23) - 1
            assert R.window_selfcheck_error(win_type, out, 10, 24) < 10, (win_type, algo)


def test_random_descriptors_one_shot_and_in_one_plan():
    """Seeded fuzz (cases.random_descs): every descriptor through the one-shot entry point, and all
    32-bit ones of one seed together in a single plan, against the oracle."""
    descs = cases.random_descs(300, seed=20260102)
    for d in descs:
        assert np.array_equal(gpu_window(d), H.orc_window(d)), d
    batch = []
    for d in descs:
        if d.dat_width > 32 or len(batch) == 140:
            continue
        batch.append(d.copy(algo=bhw.ALGO_AUTO))      # TAYLOR windows of any (DAT_WIDTH, LUT_SIZE) mix: a plan owns its ROMs
    assert len({(d.dat_width, d.lut_size) for d in batch if d.sin_type == bhw.SIN_TAYLOR and d.model == bhw.MODEL_RTL}) > 2
    want = H.orc_batch(batch, 0, bhw.batch_total(batch))
    assert np.array_equal(bhw.generate_batch(batch).cpu().numpy().astype(np.int64), want)
    plan = bhw.Plan(batch)
    b, c = 1234, bhw.batch_total(batch) - 5000
    assert np.array_equal(plan.execute(b, c).cpu().numpy().astype(np.int64), want[b:b + c])
    plan.destroy()


def test_random_descriptors_with_the_extensions():
    """The same fuzz over the round-2 additions: 6 and 8..11 terms among the entities, and every DAT_WIDTH <= 16
    descriptor of the seed once more in the int16 container - one shot, in one batch, in ragged ranges of a plan."""
    import torch
    descs = cases.random_descs(260, seed=20261019, max_pw=13, terms=(2, 3, 4, 5, 6, 7, 8, 9, 10, 11))
    assert sum(d.win_type in (6, 8, 9, 10, 11) for d in descs) > 60
    for d in descs:
        assert np.array_equal(gpu_window(d), H.orc_window(d)), d
    small = [d.copy(algo=bhw.ALGO_AUTO, out_format=bhw.OUT_INT16) for d in descs if d.dat_width <= 16]
    assert len(small) > 50
    for d in small[:40]:
        got = bhw.generate(d)
        assert got.dtype == torch.int16 and np.array_equal(got.cpu().numpy().astype(np.int64), H.orc_window(d)), d
    total = bhw.batch_total(small)
    want = H.orc_batch(small, 0, total)
    assert np.array_equal(bhw.generate_batch(small).cpu().numpy().astype(np.int64), want)
    plan = bhw.Plan(small)
    for b, c in ((0, total), (1, total - 2), (total // 3, total // 2), (total - 300, 300)):
        assert np.array_equal(plan.execute(b, c).cpu().numpy().astype(np.int64), want[b:b + c]), (b, c)
    plan.destroy()
    assert np.array_equal(bhw.generate_batch_host(small, 17, total - 40).astype(np.int64), want[17:total - 23])


def test_host_entry_points():
    d = bhw.make_desc(4, 16, 17, [47022, 64001, 18518, 1531])
    want = H.orc_window(d)
    assert np.array_equal(bhw.generate_host(d).astype(np.int64), want)
    assert np.array_equal(bhw.generate_host(d, 1000, 3000).astype(np.int64), want[1000:4000])
    # chunked pipeline: a window larger than one 64 MiB staging chunk
    big = bhw.variant_desc(1, 25, 16)
    host = bhw.generate_host(big)
    dev = bhw.generate(big).cpu().numpy()
    assert np.array_equal(host, dev)
    d64 = bhw.variant_desc(10, 12, 40)
    assert np.array_equal(bhw.generate_host(d64), H.orc_window(d64))
    descs = [bhw.variant_desc(v, 10, cases.VARIANT_DW[v]) for v in range(1, 11)]
    assert np.array_equal(bhw.generate_batch_host(descs).astype(np.int64), H.orc_batch(descs, 0, bhw.batch_total(descs)))


def test_host_pipeline_segments():
    """The host entry point plans the request in segments cut at window boundaries (a short first
    one, then >= 256 MiB); a ragged range over a mixed batch that spans several segments must equal
    the device path sample for sample."""
    import torch
    descs = ([bhw.variant_desc(6, 16, 17, ) for _ in range(40)]            # 10 MiB of small windows
             + [bhw.variant_desc(1 + (i % 10), 20 + (i % 3), cases.VARIANT_DW[1 + (i % 10)]) for i in range(30)]
             + [bhw.variant_desc(3, 26, 16)]                               # one 256 MiB window
             + [bhw.variant_desc(9, 12, 24, stream_offset=1) for _ in range(100)])
    for i, d in enumerate(descs):
        d.aa[0] -= i % 5
    total = bhw.batch_total(descs)
    assert total * 4 > (256 << 20) + (64 << 20)
    dev = bhw.generate_batch(descs).cpu().numpy()
    assert np.array_equal(bhw.generate_batch_host(descs), dev)
    b, c = 12345, total - 12345 - 777
    assert np.array_equal(bhw.generate_batch_host(descs, b, c), dev[b:b + c])
    b = (40 << 16) + (3 << 20) + 5                                         # starts inside a large window
    assert np.array_equal(bhw.generate_batch_host(descs, b, 1 << 22), dev[b:b + (1 << 22)])
    del dev
    torch.cuda.empty_cache()


def test_cuda_graph_capture_and_replay():
    """A plan execute (table build + programmatic dependent bank launch, and a mixed batch with the
    general kernel) captured into a CUDA graph and replayed into the same buffer."""
    import torch
    for descs in ([bhw.make_desc(4, 16, 17, [47022 - i, 64001, 18518, 1531 + i]) for i in range(64)],
                  [bhw.variant_desc(v, pw, cases.VARIANT_DW[v]) for v in (1, 3, 6, 8, 10) for pw in (5, 9, 13, 17)]):
        plan = bhw.Plan(descs)
        want = plan.execute().clone()
        out = torch.zeros_like(want)
        bhw.set_table_cache(False)
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                plan.execute(out=out)                       # warm-up on the capture stream
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=s):
                    plan.execute(out=out)
            for _ in range(3):
                out.zero_()
                g.replay()
                torch.cuda.synchronize()
                assert torch.equal(out, want)
        finally:
            bhw.set_table_cache(True)
        del g
        plan.destroy()


def test_concurrent_host_threads():
    """The ABI is re-entrant: several host threads generate different windows at once (one-shot
    calls on their own streams, plus the host-buffer entry point) and every result is exact."""
    import threading
    import torch
    work = [bhw.variant_desc(1 + (i % 10), 9 + (i % 6), cases.VARIANT_DW[1 + (i % 10)]) for i in range(24)]
    work += [bhw.make_desc(3, 14, 24, [7046424, 8388600, 1342176], sin_type=bhw.SIN_TAYLOR),
             bhw.variant_desc(10, 12, 32, sin_type=bhw.SIN_CORDIC48), bhw.variant_desc(6, 18, 17)]
    want = [H.orc_window(d, threads=4) for d in work]
    got = [None] * len(work)
    errs = []

    def run(tid):
        try:
            torch.cuda.set_device(0)
            s = torch.cuda.Stream()
            for i in range(tid, len(work), 4):
                with torch.cuda.stream(s):
                    for _ in range(3):
                        g = bhw.generate(work[i])
                s.synchronize()
                got[i] = g.cpu().numpy().astype(np.int64)
                if i % 5 == 0:
                    assert np.array_equal(bhw.generate_host(work[i]).astype(np.int64), got[i])
        except Exception as e:   # noqa: BLE001
            errs.append(repr(e))

    ts = [threading.Thread(target=run, args=(t,)) for t in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for i, d in enumerate(work):
        assert np.array_equal(got[i], want[i]), d


def test_stream_offset_long_windows_all_kernels():
    """DT_VLD-gated order (stream_offset = 1) is a rotation by one sample through every kernel:
    bank (staged and global tables), general, direct."""
    import torch
    for d in (bhw.variant_desc(6, 16, 17), bhw.variant_desc(2, 21, 16), bhw.variant_desc(10, 18, 32),
              bhw.variant_desc(3, 18, 24, sin_type=bhw.SIN_TAYLOR), bhw.variant_desc(9, 19, 24, sin_type=bhw.SIN_CORDIC_SCALED)):
        base = bhw.generate(d)
        for algo in (bhw.ALGO_AUTO, bhw.ALGO_TABLE, bhw.ALGO_DIRECT):
            rot = bhw.generate(d.copy(stream_offset=1, algo=algo))
            assert torch.equal(rot, torch.roll(base, -1)), (d, algo)
        n = 1 << d.phi_width
        part = bhw.generate(d.copy(stream_offset=1), n - 1000, 1000)        # ragged tail incl. the wrapped w[0]
        assert torch.equal(part, torch.roll(base, -1)[n - 1000:])


def test_apply_step_fused_and_unfused():
    """bhw_apply: y = x * w through int_multNxN_dsp48 (src/int_multNxN_dsp48.vhd:105), exact (2*DW bits) and as the
    entities slice and round it (src/hamming_win.vhd:195-208) - bit-exact against the oracle for windows the fused
    kernel takes (every table placement, spread walk) and for those that go through scratch memory."""
    import torch
    rng = np.random.default_rng(7)
    hi = (1 << 31) - 1
    shapes = [("fused, staged half period", bhw.variant_desc(1, 12, 16), 3),
              ("fused, staged half period, 3 terms, DT_VLD order", bhw.variant_desc(4, 16, 16).copy(stream_offset=1), 2),
              ("fused, uint16 quarter waves", bhw.variant_desc(6, 18, 17), 2),
              ("fused, pyramid", bhw.variant_desc(9, 19, 24), 2),
              ("fused, pyramid, 7 terms DW 32", bhw.variant_desc(10, 17, 32), 1),
              ("fused, HLS family", bhw.variant_desc(6, 14, 17, model=bhw.MODEL_HLS), 2),
              ("fused, frames split over the grid (short window, many frames)", bhw.variant_desc(1, 12, 16), 301),
              ("fused, frames split, one tile per window", bhw.variant_desc(6, 9, 17), 1000),
              ("fused, frames split, pyramid", bhw.variant_desc(9, 13, 24), 77),
              ("scratch: TAYLOR", bhw.variant_desc(3, 14, 24, sin_type=bhw.SIN_TAYLOR, lut_size=9), 2),
              ("scratch: cordic_dds48", bhw.variant_desc(10, 12, 32, sin_type=bhw.SIN_CORDIC48), 2),
              ("scratch: shorter than a tile pair", bhw.variant_desc(2, 6, 16), 5),
              ("scratch: 64-bit tail", bhw.make_desc(4, 11, 32, [hi, hi - 1, hi - 2, hi - 3]), 2),
              ("scratch: DIRECT", bhw.variant_desc(6, 12, 17, algo=bhw.ALGO_DIRECT), 1)]
    for name, d, frames in shapes:
        n = 1 << d.phi_width
        x = rng.integers(-(1 << 31), 1 << 31, size=(frames, n), dtype=np.int64).astype(np.int32)   # upper bits are not port bits
        x[0, :4] = [(1 << (d.dat_width - 1)) - 1, -(1 << (d.dat_width - 1)), 0, -1]
        xd = torch.from_numpy(x).cuda()
        for mode in (bhw.APPLY_EXACT, bhw.APPLY_ROUNDED):
            bhw.timing_enable(True)
            bhw.timing_reset()
            got = bhw.apply(d, xd, mode)
            kt = bhw.timing_read()
            bhw.timing_enable(False)
            assert got.dtype == (torch.int64 if mode == bhw.APPLY_EXACT else torch.int32)
            want = H.orc_apply(d, x, mode)
            assert np.array_equal(got.cpu().numpy().astype(np.int64), want), (name, mode)
            assert (kt["k_apply_mul"][0] == 0) == name.startswith("fused"), (name, kt)
    # a long window on the spread walk, one frame, spot-checked
    d = bhw.variant_desc(9, 23, 24)
    n = 1 << 23
    xd = torch.randint(-(1 << 23), 1 << 23, (n,), dtype=torch.int32, device="cuda")
    got = bhw.apply(d, xd, bhw.APPLY_EXACT)
    w = bhw.generate(d)
    assert torch.equal(got, xd.to(torch.int64) * w.to(torch.int64))
    assert np.array_equal(w[:4096].cpu().numpy().astype(np.int64), H.orc_window(d, 0, 4096))
    with pytest.raises(bhw.BhwError):
        bhw.apply(bhw.variant_desc(10, 10, 40), torch.zeros(1024, dtype=torch.int32, device="cuda"))


def test_cordic_atan2():
    import torch
    g = torch.Generator(device="cpu").manual_seed(11)
    for aw, iw, prec in [(8, 8, 1), (16, 16, 1), (16, 20, 3), (24, 24, 1), (24, 32, 2), (32, 31, 1), (32, 32, 7)]:
        n = (1 << 20) + 77
        x = torch.randint(-(1 << 31), 1 << 31, (n,), generator=g, dtype=torch.int64).to(torch.int32)
        y = torch.randint(-(1 << 31), 1 << 31, (n,), generator=g, dtype=torch.int64).to(torch.int32)
        got = bhw.atan2(x.cuda(), y.cuda(), iw, aw, prec).cpu().numpy()
        assert np.array_equal(got, H.orc_atan2(iw, aw, prec, x.numpy(), y.numpy())), (aw, iw, prec)
    with pytest.raises(bhw.BhwError):
        bhw.atan2(x.cuda(), y.cuda(), 20, 24, 1)          # INPUT_WIDTH < ANGLE_WIDTH - 1
    # host buffers, more than one staging chunk (4M pairs per chunk)
    n = (9 << 20) + 5
    xh = torch.randint(-(1 << 23), 1 << 23, (n,), generator=g, dtype=torch.int32).numpy()
    yh = torch.randint(-(1 << 23), 1 << 23, (n,), generator=g, dtype=torch.int32).numpy()
    assert np.array_equal(bhw.atan2_host(xh, yh, 24, 24, 2), H.orc_atan2(24, 24, 2, xh, yh))
    # stream_quadrant=1 (the entity's own pairing, tests/test_rtl_vhdl_sim.py): the look-ahead crosses chunk borders
    assert np.array_equal(bhw.atan2_host(xh, yh, 24, 24, 2, stream_quadrant=1), H.orc_atan2(24, 24, 2, xh, yh, stream=True))
    n = (1 << 20) + 77
    for aw, iw, prec in [(16, 16, 1), (24, 32, 2), (32, 32, 7), (12, 20, 1)]:
        got = bhw.atan2(x[:n].cuda(), y[:n].cuda(), iw, aw, prec, stream_quadrant=1).cpu().numpy()
        assert np.array_equal(got, H.orc_atan2(iw, aw, prec, x[:n].numpy(), y[:n].numpy(), stream=True)), (aw, iw, prec)
    with pytest.raises(bhw.BhwError):
        bhw.atan2(x.cuda(), y.cuda(), 16, 16, 1, stream_quadrant=2)


def test_mterm_extension():
    """BHW_WIN_MTERM_* (6 and 8..11 terms, quantize variants 14..18): every strategy against the oracle over widths,
    CORDIC sources and port edge cases; long windows; mixed into batches with the reference's entities; the int16
    container, the 64-bit container and the apply step."""
    import torch
    descs = cases.mterm_sweep()
    assert len(descs) > 500
    for d in descs:
        want = H.orc_window(d)
        for name, dd in both_algos(d):
            assert np.array_equal(gpu_window(dd), want), (name, d)
    for v, pw, dw in ((14, 16, 24), (18, 16, 32), (15, 18, 16), (17, 14, 40), (16, 20, 17)):
        d = bhw.variant_desc(v, pw, dw)
        n = 1 << pw
        got = gpu_window(d)
        for n0 in (0, n // 2 - 2048, n - 4096):
            assert np.array_equal(got[n0:n0 + 4096], H.orc_window(d, n0, 4096)), (d, n0)
        assert np.array_equal(gpu_window(d.copy(stream_offset=1)), np.roll(got, -1)), d
        assert np.array_equal(gpu_window(d, 777, 5000), got[777:5777]), d
    # batches: M-term windows between group, bank-shaped and short windows; ragged ranges; a resident plan
    mix = []
    for pw in (5, 9, 12, 14):
        mix += [bhw.variant_desc(1, pw, 16), bhw.variant_desc(14, pw, 16), bhw.variant_desc(6, pw, 16), bhw.variant_desc(18, pw, 16),
                bhw.variant_desc(10, pw, 16), bhw.variant_desc(16, pw, 16, sin_type=bhw.SIN_CORDIC48)]
    mix += [bhw.variant_desc(15, 10, 16).copy(aa=[int(a) - i if k == 0 else int(a) for k, a in enumerate(bhw.variant_desc(15, 10, 16).aa)])
            for i in range(200)]
    total = bhw.batch_total(bhw.desc_array(mix))
    want = H.orc_batch(mix, 0, total)
    assert np.array_equal(bhw.generate_batch(mix).cpu().numpy().astype(np.int64), want)
    for b, c in ((3, 10001), (total - 70000, 70000)):
        assert np.array_equal(bhw.generate_batch(mix, b, c).cpu().numpy().astype(np.int64), want[b:b + c]), (b, c)
    assert np.array_equal(bhw.generate_batch_host(mix).astype(np.int64), want)
    plan = bhw.Plan(mix)
    assert np.array_equal(plan.execute(5, total - 9).cpu().numpy().astype(np.int64), want[5:total - 4])
    plan.destroy()
    packed = bhw.generate_batch([d.copy(out_format=bhw.OUT_INT16) for d in mix])
    assert packed.dtype == torch.int16 and np.array_equal(packed.cpu().numpy().astype(np.int64), want)
    # the apply step (scratch path: these term counts are outside the fused kernel)
    d = bhw.variant_desc(17, 12, 24)
    x = torch.randint(-(1 << 23), 1 << 23, (3, 4096), dtype=torch.int32, device="cuda")
    for mode in (bhw.APPLY_EXACT, bhw.APPLY_ROUNDED):
        assert np.array_equal(bhw.apply(d, x, mode).cpu().numpy().astype(np.int64), H.orc_apply(d, x.cpu().numpy(), mode))
    with pytest.raises(bhw.BhwError):
        bhw.generate(bhw.make_desc(8, 10, 16, [1] * 8, model=bhw.MODEL_HLS))


def test_taylor_long_window_quarter_body():
    """k_direct_taylor's quarter-window branch (direct_taylor_quad4: 16 samples from two ROM words) on whole TAY_WIDE
    windows, every sample against the oracle; the shapes around it (DT_VLD order, a range, the DSP datapath) keep
    the per-sample / pair branches and are checked the same way."""
    rng = np.random.default_rng(17)
    shapes = [(3, 12, 19, 5), (4, 14, 24, 9), (1, 16, 24, 9), (2, 18, 32, 11), (3, 20, 24, 9), (4, 13, 31, 4), (1, 22, 20, 12)]
    for v, pw, dw, lut in shapes:
        d = bhw.variant_desc(v, pw, dw, sin_type=bhw.SIN_TAYLOR, lut_size=lut, algo=bhw.ALGO_DIRECT)
        n = 1 << pw
        got = gpu_window(d)
        if n <= (1 << 18):
            assert np.array_equal(got, H.orc_window(d)), d
        else:
            for n0 in (0, n // 4 - 4096, n // 2 - 4096, 3 * (n // 4) - 4096, n - 8192):
                assert np.array_equal(got[n0:n0 + 8192], H.orc_window(d, n0, 8192)), (d, n0)
        assert np.array_equal(gpu_window(d.copy(stream_offset=1)), np.roll(got, -1)), d
        assert np.array_equal(gpu_window(d, 1000, 3000), got[1000:4000]), d
    for it in range(12):
        m, pw, dw = int(rng.integers(2, 4)), int(rng.integers(12, 17)), int(rng.integers(19, 33))
        lut = int(rng.integers(4, min(pw - 4, 12) + 1))
        lim = 1 << (dw - 1)
        aa = [lim - 1] * m if it % 4 == 0 else [-lim] * m if it % 5 == 0 else [int(x) for x in rng.integers(-lim, lim, m)]
        d = bhw.make_desc(m, pw, dw, aa, sin_type=bhw.SIN_TAYLOR, lut_size=lut, algo=bhw.ALGO_DIRECT)
        assert np.array_equal(gpu_window(d), H.orc_window(d)), d


def test_packed_int16_output():
    """BHW_OUT_INT16 (bhw_desc.out_format): the same integers in an int16 container, for every window kind whose
    DAT_WIDTH fits - through the group kernel (every table placement, paired / cut / spread walk), the general kernel
    (short windows, TAYLOR and input-quadrant tables, the generic body), one-shot, host and plan entry points, ragged
    ranges.  Reference for each: the int32 output of the same batch, itself oracle-checked elsewhere, plus the oracle
    directly on the small ones."""
    import torch
    P = bhw.OUT_INT16
    small = []
    for pw in (4, 5, 7, 8, 9, 10, 12, 13, 15):
        for v in (1, 2, 3, 4):
            small.append(bhw.variant_desc(v, pw, 16))
        small.append(bhw.variant_desc(6, pw, 14))
        small.append(bhw.variant_desc(9, pw, 16))
        small.append(bhw.variant_desc(10, pw, 12))
        small.append(bhw.variant_desc(1, pw, 16, sin_type=bhw.SIN_CORDIC48))
        small.append(bhw.variant_desc(3, pw, 16, sin_type=bhw.SIN_CORDIC_SCALED).copy(stream_offset=1))
        if pw >= 8:
            small.append(bhw.variant_desc(4, pw, 16, sin_type=bhw.SIN_TAYLOR, lut_size=pw - 4))
            small.append(bhw.variant_desc(6, pw, 16, model=bhw.MODEL_HLS))
    small.append(bhw.make_desc(2, 9, 16, [-32768, -32768]))
    small.append(bhw.make_desc(3, 9, 8, [127, -128, 127]))
    small += [bhw.make_desc(2, 10, 16, [17808 + i, 14959 - i]) for i in range(300)]        # a bank-shaped run
    for d in small:
        assert bhw.validate(d) == 0, d
    packed = [d.copy(out_format=P) for d in small]
    total = bhw.batch_total(bhw.desc_array(small))
    want = bhw.generate_batch(small).cpu().numpy()
    assert np.abs(want).max() < (1 << 15)
    got = bhw.generate_batch(packed)
    assert got.dtype == torch.int16 and got.numel() == total
    assert np.array_equal(got.cpu().numpy().astype(np.int32), want)
    assert np.array_equal(want[:4096].astype(np.int64), H.orc_batch(small, 0, 4096))
    for b, c in ((1, 777), (12345, 100001), (total - 5000, 5000), (3, total - 7)):
        g = bhw.generate_batch(packed, b, c).cpu().numpy()
        assert np.array_equal(g.astype(np.int32), want[b:b + c]), (b, c)
    h = bhw.generate_batch_host(packed)
    assert h.dtype == np.int16 and np.array_equal(h.astype(np.int32), want)
    h = bhw.generate_batch_host(packed, 1001, 54321)
    assert np.array_equal(h.astype(np.int32), want[1001:1001 + 54321])
    plan = bhw.Plan(packed)
    assert plan.elem_bytes == 2
    for b, c in ((0, total), (257, 99999)):
        assert np.array_equal(plan.execute(b, c).cpu().numpy().astype(np.int32), want[b:b + c])
    plan.destroy()
    # one window through bhw_generate / bhw_generate_host
    d = bhw.variant_desc(4, 14, 16)
    w32 = bhw.generate(d).cpu().numpy()
    assert np.array_equal(bhw.generate(d.copy(out_format=P)).cpu().numpy().astype(np.int32), w32)
    assert np.array_equal(bhw.generate(d.copy(out_format=P), 100, 1000).cpu().numpy().astype(np.int32), w32[100:1100])
    assert np.array_equal(bhw.generate_host(d.copy(out_format=P)).astype(np.int32), w32)
    # long windows: pyramid gathers, the spread walk of a 7-term window, a long TAYLOR window
    for d in (bhw.variant_desc(2, 22, 16), bhw.variant_desc(10, 23, 16), bhw.variant_desc(9, 21, 16),
              bhw.variant_desc(3, 20, 16, sin_type=bhw.SIN_TAYLOR, lut_size=9)):
        a = bhw.generate_batch([d.copy(out_format=P)])
        b_ = bhw.generate_batch([d])
        assert a.dtype == torch.int16 and torch.equal(a.to(torch.int32), b_), d
        assert np.array_equal(b_[:2048].cpu().numpy().astype(np.int64), H.orc_window(d, 0, 2048))
    # the launches of a packed batch are the same kernel classes as for int32 (the bank-shaped run through the
    # bank kernel's int16 instantiation), no conversion pass
    bhw.timing_enable(True)
    bhw.timing_reset()
    bhw.generate_batch(packed)
    kt = bhw.timing_read()
    bhw.timing_enable(False)
    assert kt["k_synth_bank"][0] == 1 and kt["k_synth_group"][0] > 0 and kt["k_apply_mul"][0] == 0, kt
    # banks the int16 bank kernel does not take (4 and more terms; an input-quadrant CORDIC) still come out right
    for base in (bhw.variant_desc(6, 10, 16), bhw.variant_desc(10, 9, 16), bhw.variant_desc(1, 11, 16, sin_type=bhw.SIN_CORDIC48),
                 bhw.variant_desc(3, 12, 16), bhw.variant_desc(2, 8, 12)):
        bank = [base.copy(aa=[int(a) - i if k == 0 else int(a) for k, a in enumerate(base.aa)]) for i in range(600)]
        a = bhw.generate_batch([d.copy(out_format=P) for d in bank])
        b_ = bhw.generate_batch(bank)
        assert torch.equal(a.to(torch.int32), b_), base
        assert np.array_equal(b_[:3000].cpu().numpy().astype(np.int64), H.orc_batch(bank[:16], 0, 3000)), base
    with pytest.raises(bhw.BhwError):
        bhw.generate_batch([small[0], packed[1]])                          # one container per batch
    with pytest.raises(bhw.BhwError):
        bhw.apply(packed[0], torch.zeros(16, dtype=torch.int32, device="cuda"))


def test_rtl_golden_vectors_on_gpu():
    """The CUDA path against what the reference's VHDL entities output when executed (tests/golden/rtl_sim_vectors.npz,
    made by tests/golden/make_rtl_golden.py through oracle/vhdl_sim.py) - no oracle in between."""
    import json
    import torch
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z = np.load(os.path.join(gold, "rtl_sim_vectors.npz"))
    cases = json.load(open(os.path.join(gold, "rtl_sim_cases.json")))
    sin_of = {"cordic_dds": bhw.SIN_CORDIC, "cordic_dds48": bhw.SIN_CORDIC48, "cordic_dds_scaled": bhw.SIN_CORDIC_SCALED}
    for c in cases["dds"] + cases["dds_precision"]:
        d = bhw.make_desc(2, c["phase_width"], c["data_width"], sin_type=sin_of[c["entity"]], precision=c.get("precision", 0))
        ph = z[c["key"] + "/phases"]
        if c["phase_width"] <= 20:
            s, co = bhw.sincos(d)
            s, co = s.cpu().numpy().astype(np.int64)[ph], co.cpu().numpy().astype(np.int64)[ph]
        else:
            pairs = [bhw.sincos(d, int(p), 1) for p in ph]
            s = np.array([int(a[0]) for a, _ in pairs], np.int64)
            co = np.array([int(b[0]) for _, b in pairs], np.int64)
        assert np.array_equal(s, z[c["key"] + "/sin"]) and np.array_equal(co, z[c["key"] + "/cos"]), c["key"]
    for c in cases["taylor"]:
        pw, dw, lut = c["phase_width"], c["data_width"], c["lut_size"]
        N, L = 1 << pw, H.rtl_taylor_latency(pw, dw, lut)
        d = bhw.make_desc(2, pw, dw, sin_type=bhw.SIN_TAYLOR, lut_size=lut)
        n = c["clocks"] - L
        if N <= 4096:
            s, co = bhw.sincos(d)
            idx = (c["start"] + np.arange(n)) % N
            s, co = s.cpu().numpy().astype(np.int64)[idx], co.cpu().numpy().astype(np.int64)[idx]
        else:
            first = min(n, N - c["start"])                                      # may run across the end of the period
            parts = [bhw.sincos(d, c["start"], first)] + ([bhw.sincos(d, 0, n - first)] if first < n else [])
            s = np.concatenate([p[0].cpu().numpy().astype(np.int64) for p in parts])
            co = np.concatenate([p[1].cpu().numpy().astype(np.int64) for p in parts])
        assert np.array_equal(s, z[c["key"] + "/sin_per_clock"][L:]), c["key"]
        assert np.array_equal(co, z[c["key"] + "/cos_per_clock"][L:]), c["key"]
    descs, wants = [], []
    for c in cases["windows"]:
        g = c["generics"]
        N = 1 << g["PHI_WIDTH"]
        vld = z[c["key"] + "/dt_vld_per_clock"].astype(bool)
        stream = z[c["key"] + "/dt_win_per_clock"][vld][:N]                    # w[1] ... w[N-1], w[0]
        d = H.rtl_case_desc(c, z[c["key"] + "/aa"])
        for name, dd in both_algos(d):
            assert np.array_equal(gpu_window(dd), np.roll(stream, 1)), (name, c["key"])
        assert np.array_equal(gpu_window(d.copy(stream_offset=1)), stream), c["key"]
        if g["DAT_WIDTH"] <= 32:
            descs.append(d.copy(stream_offset=1))
            wants.append(stream)
    # config 3's composition: the window entities over cordic_dds48 / cordic_dds_scaled (the name cordic_dds bound to
    # the other entity in the simulator); w[0] reaches DT_WIN one clock after DT_VLD has risen
    sin_sw = {"cordic_dds48": bhw.SIN_CORDIC48, "cordic_dds_scaled": bhw.SIN_CORDIC_SCALED}
    for c in cases["windows_swapped"]:
        g = c["generics"]
        m = H.rtl_case_terms(c)
        N = 1 << g["PHI_WIDTH"]
        d = bhw.make_desc(m, g["PHI_WIDTH"], g["DAT_WIDTH"], [int(a) for a in z[c["key"] + "/aa"]], sin_type=sin_sw[c["dds"]])
        L = c["first_dt_vld_clock"] + 1
        want = z[c["key"] + "/dt_win_per_clock"][L:L + N]
        for name, dd in both_algos(d):
            assert np.array_equal(gpu_window(dd), want), (name, c["key"])
        descs.append(d)
        wants.append(want)
    # and all the int32 ones as one plan (group / bank kernels)
    got = bhw.generate_batch(descs).cpu().numpy().astype(np.int64)
    assert np.array_equal(got, np.concatenate(wants))
    for c in cases["atan2"]:
        aw = c["angle_width"]
        x = torch.from_numpy(z[c["key"] + "/x"].astype(np.int32)).cuda()
        y = torch.from_numpy(z[c["key"] + "/y"].astype(np.int32)).cuda()
        want = z[c["key"] + "/phi_dt_per_clock"][aw + 1:aw + 1 + x.numel()]
        got = bhw.atan2(x, y, c["input_width"], aw, c["precision"], stream_quadrant=1).cpu().numpy().astype(np.int64)
        assert np.array_equal(got & ((1 << aw) - 1), want & ((1 << aw) - 1)), c["key"]


def test_win_selector_and_errors():
    w = bhw.WinSelector(PHI_WIDTH=10, DAT_WIDTH=16, WIN_TYPE="HAMMING")
    assert np.array_equal(w.stream(AA0=17808, AA1=14959).cpu().numpy().astype(np.int64),
                          H.orc_window(bhw.make_desc(2, 10, 16, [17808, 14959])))
    wt = bhw.WinSelector(PHI_WIDTH=14, DAT_WIDTH=16, WIN_TYPE="BH3TERM", SIN_TYPE="TAYLOR", dt_vld_order=True)
    d = wt.desc(AA0=27518, AA1=32760, AA2=5242)
    assert np.array_equal(wt.stream(AA0=27518, AA1=32760, AA2=5242, host=True).astype(np.int64), H.orc_window(d))
    good = bhw.make_desc(2, 10, 16, [17808, 14959])
    with pytest.raises(bhw.BhwError):
        bhw.generate(good, 1000, 100)                                     # past the end
    with pytest.raises(bhw.BhwError):
        bhw.generate(good.copy(dat_width=50))
    import torch
    out = torch.full((1024,), 7, dtype=torch.int32, device="cuda")
    st = bhw.lib().bhw_generate(C.byref(good.copy(win_type=12)), out.data_ptr(), 0, 1024, None)
    assert st == -2 and bool((out == 7).all())                            # errors write nothing


def test_multi_gpu_single_process_form():
    import torch
    ng = torch.cuda.device_count()
    descs = bhw.desc_array([bhw.variant_desc(6, 12, 17) for _ in range(9)])
    total = bhw.batch_total(descs)
    outs, ptrs = [], (C.c_void_p * ng)()
    for g in range(ng):
        b, c = bhw.shard_range(total, g, ng)
        outs.append(torch.empty(c, dtype=torch.int32, device=f"cuda:{g}"))
        ptrs[g] = outs[-1].data_ptr()
    assert bhw.lib().bhw_generate_batch_multi(descs, len(descs), ng, ptrs) == 0
    got = np.concatenate([o.cpu().numpy() for o in outs]).astype(np.int64)
    assert np.array_equal(got, H.orc_batch(list(descs), 0, total))
