"""rtl_bitvec.py - a second, independently written restatement of the RTL entities, in the style of
the VHDL itself: every signal is a fixed-width bit vector, every slice / concatenation / `+` is
done on vectors with the widths the source declares (ieee.std_logic_signed semantics: operands are
sign-extended to the longer length, the result keeps that length).

TEST INFRASTRUCTURE ONLY.  Nothing in the product imports this.  It exists because the RTL has no
executable reference in this image (no VHDL simulator): oracle/bhw_oracle.c restates the entities
with integer arithmetic + explicit wraps, this file restates them with bit vectors, and
tests/test_rtl_bitvec.py requires the two to agree on randomised generics, ports and phases.
Agreement of two restatements is a regression anchor, not ground truth; since round 2 the ground truth is
oracle/vhdl_sim.py, which executes the reference's VHDL itself (tests/test_rtl_vhdl_sim.py).

Covered: cordic_dds (src/cordic_dds.vhd:97-249), cordic_dds48 (src/cordic_dds48.vhd:110-259), int_multNxN_dsp48 (src/int_multNxN_dsp48.vhd:105),
the tails of hamming_win (src/hamming_win.vhd:133-231), bh_win_3term (src/bh_win_3term.vhd:151-306),
bh_win_4term (src/bh_win_4term.vhd:125-280), bh_win_5term (src/bh_win_5term.vhd:148-347),
bh_win_7term (src/bh_win_7term.vhd:160-438), cordic_atan2 (src/cordic_atan2.vhd:80-220), and
taylor_sincos + tay1_order (src/taylor_sincos.vhd:86-255, src/tay1_order.vhd:112-640) with the
TAYLOR forms of hamming_win and bh_win_3term.
"""
from __future__ import annotations

ROM_LUT = [  # src/cordic_dds.vhd:104-117 == src/cordic_atan2.vhd:86-95
    0x400000000000, 0x25C80A3B3BE6, 0x13F670B6BDC7, 0x0A2223A83BBB, 0x05161A861CB1, 0x028BAFC2B209,
    0x0145EC3CB850, 0x00A2F8AA23A9, 0x00517CA68DA2, 0x0028BE5D7661, 0x00145F300123, 0x000A2F982950,
    0x000517CC19C0, 0x00028BE60D83, 0x000145F306D6, 0x0000A2F9836D, 0x0000517CC1B7, 0x000028BE60DC,
    0x0000145F306E, 0x00000A2F9837, 0x00000517CC1B, 0x0000028BE60E, 0x00000145F307, 0x000000A2F983,
    0x000000517CC2, 0x00000028BE61, 0x000000145F30, 0x0000000A2F98, 0x0000000517CC, 0x000000028BE6,
    0x0000000145F3, 0x00000000A2FA, 0x00000000517D, 0x0000000028BE, 0x00000000145F, 0x000000000A30,
    0x000000000518, 0x00000000028C, 0x000000000146, 0x0000000000A3, 0x000000000051, 0x000000000029,
    0x000000000014, 0x00000000000A, 0x000000000005, 0x000000000003, 0x000000000001, 0x000000000000]
GAIN48 = 0x4DBA76D421AF  # src/cordic_dds.vhd:97


class BV:
    """std_logic_vector(width-1 downto 0) holding `bits` (an unsigned Python int < 2**width)."""
    __slots__ = ("w", "u")

    def __init__(self, width: int, bits: int):
        assert width >= 1
        self.w = width
        self.u = bits & ((1 << width) - 1)

    def __getitem__(self, idx):          # v[hi:lo] = v(hi downto lo) ; v[i] = v(i)
        if isinstance(idx, slice):
            hi, lo = idx.start, idx.stop
            assert self.w > hi >= lo >= 0, (self.w, hi, lo)
            return BV(hi - lo + 1, self.u >> lo)
        assert 0 <= idx < self.w
        return (self.u >> idx) & 1

    def signed(self) -> int:
        return self.u - (1 << self.w) if self.u >> (self.w - 1) else self.u

    def sext(self, width: int) -> "BV":
        assert width >= self.w
        return BV(width, self.signed())

    def __add__(self, o):                # std_logic_signed "+": result length = max of the two
        if isinstance(o, int):
            return BV(self.w, self.signed() + o)
        w = max(self.w, o.w)
        return BV(w, self.signed() + o.signed())

    def __sub__(self, o):
        w = max(self.w, o.w)
        return BV(w, self.signed() - o.signed())

    def __invert__(self):                # not(v)
        return BV(self.w, ~self.u)


def cat(*parts) -> BV:                   # a & b & c  (leftmost = most significant); ints are single bits
    w, u = 0, 0
    for p in parts:
        if isinstance(p, int):
            p = BV(1, p)
        u = (u << p.w) | p.u
        w += p.w
    return BV(w, u)


def cordic_dds(phase_width: int, data_width: int, precision: int, ph_in: int):
    """-> (dt_sin, dt_cos) as signed ints.  src/cordic_dds.vhd:97-249."""
    PW, DW, W = phase_width, data_width, data_width + precision
    ph = BV(PW, ph_in)
    gain = cat(0, BV(48, GAIN48)[47:48 - W + 1])                        # :98
    rom = []
    for ii in range(DW - 1):                                              # func_atan :123-131
        rom.append(cat(0, BV(48, ROM_LUT[ii])[47:47 - (W - 2)]))
    init_t = cat(BV(2, 0), ph[PW - 3:0])                                  # :179
    if PW >= DW:                                                          # xPHI_LESS :159-162
        init_z = cat(init_t[PW - 1:PW - DW], BV(precision, 0))
    else:                                                                 # xPHI_MORE :163-166
        init_z = cat(init_t[PW - 1:0], BV(DW - PW + precision, 0))
    assert init_z.w == W and gain.w == W
    x, y, z = gain, BV(W, 0), init_z
    for ii in range(DW - 1):                                              # lpXY / lpZ :197-213
        if z[W - 1] == 1:
            xn = x + y[W - 1:ii]
            yn = y - x[W - 1:ii]
            zn = z + rom[ii]
        else:
            xn = x - y[W - 1:ii]
            yn = y + x[W - 1:ii]
            zn = z - rom[ii]
        x, y, z = xn, yn, zn
        assert x.w == W and y.w == W and z.w == W
    dat_sin, dat_cos = y[W - 1:precision], x[W - 1:precision]             # :218-219
    quadrant = (ph[PW - 1] << 1) | ph[PW - 2]                             # :170-172
    if quadrant == 0:
        s, c = dat_sin, dat_cos
    elif quadrant == 1:
        s, c = dat_cos, ~dat_sin + 1
    elif quadrant == 2:
        s, c = ~dat_sin + 1, ~dat_cos + 1
    else:
        s, c = ~dat_cos + 1, dat_sin
    return s.signed(), c.signed()


def _round_product(aa: BV, cos: BV, DW: int) -> BV:
    """mult_p = AAk * cos_k (src/int_multNxN_dsp48.vhd:105, 2*DW bits); dsp_r = mult_p(2DW-2 downto
    DW-2); dsp_b = dsp_r(DW downto 1) [+ 1 when dsp_r(0) = '1'] (src/bh_win_4term.vhd:231-257)."""
    mult_p = BV(2 * DW, aa.signed() * cos.signed())
    dsp_r = mult_p[2 * DW - 2:DW - 2]
    assert dsp_r.w == DW + 1
    b = dsp_r[DW:1]
    return b + 1 if dsp_r[0] else b


def window(win_type: int, phi_width: int, dat_width: int, aa, n: int, precision: int = 1) -> int:
    """DT_WIN for phase counter value n of hamming_win / bh_win_{3,4,5,7}term with the CORDIC source.
    aa = raw AA0.. port values (DAT_WIDTH bits each)."""
    DW, PW, M = dat_width, phi_width, win_type
    AA = [BV(DW, a) for a in aa[:M]]
    b = [AA[0]]                                                           # dsp_b0 <= AA0
    for k in range(1, M):
        ph = (k * n) & ((1 << PW) - 1)                                    # ph_in_k += k each ENABLE (:138-149)
        _, ck = cordic_dds(PW, DW, precision, ph)
        b.append(_round_product(AA[k], BV(DW, ck), DW))
    if M == 2:                                                            # src/hamming_win.vhd:206-228
        dsp_pp = cat(b[0][DW - 1], b[0]) - cat(b[1][DW - 1], b[1])        # DW+1 bits
        assert dsp_pp.w == DW + 1
        out = dsp_pp[DW:1]
        return (out + 1 if dsp_pp[0] else out).signed()

    def x2(v):                                                            # one sign bit: DW+1 bits
        return cat(v[DW - 1], v)

    def x3(v):                                                            # two sign bits: DW+2 bits
        return cat(v[DW - 1], v[DW - 1], v)

    if M == 3:                                                            # src/bh_win_3term.vhd:277-284
        dsp_pp = x3(b[2]) - x3(b[1]) + x3(b[0])
    elif M == 4:                                                          # src/bh_win_4term.vhd:258-265
        p1 = x2(b[2]) - x2(b[3])
        p2 = x2(b[0]) - x2(b[1])
        dsp_pp = cat(p1[DW], p1) + cat(p2[DW], p2)
    elif M == 5:                                                          # src/bh_win_5term.vhd:318-328
        p1 = x3(b[4]) - x3(b[3]) + x3(b[2])
        p2 = x3(b[0]) - x3(b[1])
        dsp_pp = p1 + p2
    else:                                                                 # src/bh_win_7term.vhd:405-423
        p1 = x3(b[0]) - x3(b[1])
        p2 = x3(b[2]) - x3(b[3])
        p3 = x3(b[4]) - x3(b[5])
        pz = x3(b[6])
        dsp_pp = (p1 + p2) + (p3 + pz)
    assert dsp_pp.w == DW + 2
    out = dsp_pp[DW + 1:2]                                                # pr_out: rounds on bit 1
    return (out + 1 if dsp_pp[1] else out).signed()


def cordic_atan2(input_width: int, angle_width: int, precision: int, vec_dx: int, vec_dy: int) -> int:
    """PHI_DT as a signed int.  src/cordic_atan2.vhd:80-220."""
    IW, AW, W = input_width, angle_width, angle_width + precision
    dx, dy = BV(IW, vec_dx), BV(IW, vec_dy)
    rom = [cat(0, BV(48, ROM_LUT[ii])[47:47 - (W - 2)]) for ii in range(AW - 1)]   # :97-108
    ix = iy = 0
    for ii in range(AW - 1):                                              # pr_abs :136-146
        ix |= (dx[ii] ^ dx[IW - 1]) << ii
        iy |= (dy[ii] ^ dy[IW - 1]) << ii
    x, y, z = BV(W, ix), BV(W, iy), BV(W, 0)
    for ii in range(AW - 1):                                              # lpXY / lpZ :166-184
        if y[W - 1] == 0:
            xn = x + y[W - 1:ii]
            yn = y - x[W - 1:ii]
            zn = z - rom[ii]
        else:
            xn = x - y[W - 1:ii]
            yn = y + x[W - 1:ii]
            zn = z + rom[ii]
        x, y, z = xn, yn, zn
    dat_phi = z[W - 1:precision]                                          # :188
    phi_pi = BV(AW, 1 << (AW - 2))                                        # :121
    quadrant = (dx[IW - 1] << 1) | dy[IW - 1]                             # :129-131
    if quadrant == 0:
        o = dat_phi
    elif quadrant == 1:
        o = dat_phi + phi_pi
    elif quadrant == 2:
        o = ~dat_phi + 1
    else:
        o = dat_phi - phi_pi
    return o.signed()


# ---- taylor_sincos + tay1_order ----------------------------------------------------------------------
# Pipeline alignment (checked on the source, not assumed): in xGEN_MORE `addr`/`acnt` are taken from
# `cnt` combinationally, `dpo` (ROM word) and `mpi` (pi * acnt) are both one register later, and
#   DATA_WIDTH < 19 : A (mpx, AREG=1) and B (sin_aa/cos_aa, BREG=1) meet at MREG; C (cos_cc/sin_cc,
#                     one fabric register + CREG) meets the product at PREG          (tay1_order.vhd:180-504)
#   DATA_WIDTH > 18 : the 35x27 multiplier has 4 clocks of latency (mults/mlt35x27_dsp48e2.vhd:23)
#                     and the ROM word waits in the 4-deep cos_del/sin_del           (tay1_order.vhd:524-596)
# so every sample combines its own ROM word with its own correction.
import math


def _taylor_rom(data_width: int, lut_size: int):
    """ROM_ARRAY(ii) = sin & cos words, INTEGER() = round to nearest (src/taylor_sincos.vhd:91-111)."""
    depth = 1 << lut_size
    amp = 2.0 ** (data_width - 1) - 1.0
    rom = []
    for ii in range(depth):
        pi_new = (float(ii) * math.pi) / (2.0 * float(depth))
        re_int = int(math.floor(amp * math.cos(pi_new) + 0.5))
        im_int = int(math.floor(amp * math.sin(pi_new) + 0.5))
        rom.append(cat(BV(data_width, im_int), BV(data_width, re_int)))
    return rom


def taylor_sincos(phase_width: int, data_width: int, lut_size: int, cnt_value: int, _rom_cache={}):
    """-> (out_sin, out_cos) for counter value cnt.  src/taylor_sincos.vhd:86-255, src/tay1_order.vhd:112-640."""
    PW, DW, LUT = phase_width, data_width, lut_size
    key = (DW, LUT)
    if key not in _rom_cache:
        _rom_cache[key] = _taylor_rom(DW, LUT)
    ROM = _rom_cache[key]
    cnt = BV(PW, cnt_value)
    quadrant = (cnt[PW - 1] << 1) | cnt[PW - 2]                               # :141
    if PW - LUT < 2:                                                          # xGEN_LESS :157-161
        addr = cat(cnt[PW - 3:0], BV(LUT - PW + 2, 0))
        dpo = ROM[addr.u]
        mem_sin, mem_cos = dpo[2 * DW - 1:DW], dpo[DW - 1:0]
    elif PW - LUT == 2:                                                       # xGEN_EQ :164-167
        dpo = ROM[cnt[LUT - 1:0].u]
        mem_sin, mem_cos = dpo[2 * DW - 1:DW], dpo[DW - 1:0]
    else:                                                                     # xGEN_MORE :169-217
        STAGE = PW - LUT - 3
        addr = cnt[PW - 3:PW - LUT - 2]
        acnt = cnt[PW - 3 - LUT:0]
        assert acnt.w == STAGE + 1
        rom_dat = ROM[addr.u]
        XSHIFT = 19 + LUT                                                     # tay1_order.vhd:112
        ramb_pi = int(math.floor(math.pi * 2.0 ** (17 - STAGE) + 0.5))        # :131 integer(round(...))
        mpi = BV(24, ramb_pi * acnt.u)                                        # conv_std_logic_vector(ramb_pi*jj, 24) :136
        mpx = cat(BV(6, 0), mpi)                                              # :168-169, 30 bits, non-negative
        sin_w, cos_w = rom_dat[2 * DW - 1:DW], rom_dat[DW - 1:0]
        if DW < 19:                                                           # xWIDTH18 :171-504
            sin_aa, cos_aa = sin_w.sext(18), cos_w.sext(18)
            # *_cc: the ROM word at bit XSHIFT, sign-extended to 48 bits, zeros below (:183-196)
            sin_cc = BV(48, sin_w.signed() << XSHIFT)
            cos_cc = BV(48, cos_w.signed() << XSHIFT)
            cos_prod = BV(48, cos_cc.signed() - mpx.signed() * sin_aa.signed())   # ALUMODE "0011": C - A*B
            sin_prod = BV(48, sin_cc.signed() + mpx.signed() * cos_aa.signed())   # ALUMODE "0000": C + A*B
            mem_cos = cos_prod[XSHIFT + DW - 1:XSHIFT]                        # :501-502
            mem_sin = sin_prod[XSHIFT + DW - 1:XSHIFT]
        else:                                                                 # xWIDTH35 :506-637
            sin_aa, cos_aa = sin_w.sext(35), cos_w.sext(35)
            cos_pp = BV(62, cos_aa.signed() * mpx[26:0].u)                    # 35 x 27 (unsigned mpx: bits 29..24 are zero)
            sin_pp = BV(62, sin_aa.signed() * mpx[26:0].u)
            mlt1_bb = sin_pp[DW + XSHIFT - 1:XSHIFT]                          # :583-586
            mlt2_bb = cos_pp[DW + XSHIFT - 1:XSHIFT]
            cos_pdt = cos_w - mlt1_bb                                         # :595-596 (DW bits, wraps)
            sin_pdt = sin_w + mlt2_bb
            sat = BV(DW, (1 << (DW - 1)) - 1)                                 # ((DW-1) => '0', others => '1')
            mem_cos = cos_pdt if cos_pdt[DW - 1] == 0 else sat                # pr_rnd :602-617
            mem_sin = sin_pdt if sin_pdt[DW - 1] == 0 else sat
    if quadrant == 0:                                                         # pr_quad :237-255
        s, c = mem_sin, mem_cos
    elif quadrant == 1:
        s, c = mem_cos, ~mem_sin + 1
    elif quadrant == 2:
        s, c = ~mem_sin + 1, ~mem_cos + 1
    else:
        s, c = ~mem_cos + 1, mem_sin
    return s.signed(), c.signed()


def window_taylor(win_type: int, phi_width: int, dat_width: int, lut_size: int, aa, n: int) -> int:
    """DT_WIN of hamming_win / bh_win_3term with SIN_TYPE = "TAYLOR": every unit has its own +1 counter,
    the second harmonic of bh_win_3term is a PHASE_WIDTH-1 unit (src/bh_win_3term.vhd:221-233)."""
    DW, PW, M = dat_width, phi_width, win_type
    assert M in (2, 3)
    AA = [BV(DW, a) for a in aa[:M]]
    b = [AA[0]]
    for k in range(1, M):
        pw_k = PW - (k - 1)
        _, ck = taylor_sincos(pw_k, DW, lut_size, n & ((1 << pw_k) - 1))
        b.append(_round_product(AA[k], BV(DW, ck), DW))
    if M == 2:
        dsp_pp = cat(b[0][DW - 1], b[0]) - cat(b[1][DW - 1], b[1])
        out = dsp_pp[DW:1]
        return (out + 1 if dsp_pp[0] else out).signed()
    x3 = lambda v: cat(v[DW - 1], v[DW - 1], v)                               # noqa: E731
    dsp_pp = x3(b[2]) - x3(b[1]) + x3(b[0])
    out = dsp_pp[DW + 1:2]
    return (out + 1 if dsp_pp[1] else out).signed()


# ---- cordic_dds48 --------------------------------------------------------------------------------------
ROM_LUT48 = [  # src/cordic_dds48.vhd:115-128 (pi/4 -> 2^45; independently rounded, not ROM_LUT >> 1)
    0x200000000000, 0x12E4051D9DF3, 0x09FB385B5EE4, 0x051111D41DDE, 0x028B0D430E59, 0x0145D7E15904,
    0x00A2F61E5C28, 0x00517C5511D4, 0x0028BE5346D1, 0x00145F2EBB31, 0x000A2F980092, 0x000517CC14A8,
    0x00028BE60CE0, 0x000145F306C1, 0x0000A2F9836B, 0x0000517CC1B7, 0x000028BE60DC, 0x0000145F306E,
    0x00000A2F9837, 0x00000517CC1B, 0x0000028BE60E, 0x00000145F307, 0x000000A2F983, 0x000000517CC2,
    0x00000028BE61, 0x000000145F30, 0x0000000A2F98, 0x0000000517CC, 0x000000028BE6, 0x0000000145F3,
    0x00000000A2FA, 0x00000000517D, 0x0000000028BE, 0x00000000145F, 0x000000000A30, 0x000000000518,
    0x00000000028C, 0x000000000146, 0x0000000000A3, 0x000000000051, 0x000000000029, 0x000000000014,
    0x00000000000A, 0x000000000005, 0x000000000003, 0x000000000001, 0x000000000001, 0x000000000000]
GAIN48_B = 0x26DD3B6A10D8  # src/cordic_dds48.vhd:110


def cordic_dds48(phase_width: int, data_width: int, ph_in: int):
    """-> (dt_sin, dt_cos).  src/cordic_dds48.vhd:110-259: 48-bit registers, the quadrant is folded into
    the start vector and the phase word, DATA_WIDTH X/Y stages but only DATA_WIDTH-1 Z stages."""
    PW, DW = phase_width, data_width
    ph = BV(PW, ph_in)
    quadrant = ph[PW - 1:PW - 2].u                                            # :167
    gain = BV(48, GAIN48_B)
    if quadrant in (0, 3):                                                    # pr_phi :170-186, pr_xy :189-216
        init_t, init_x, init_y = ph, gain, BV(48, 0)
    elif quadrant == 1:
        init_t, init_x, init_y = cat(BV(2, 0), ph[PW - 3:0]), BV(48, 0), ~gain + 1
    else:
        init_t, init_x, init_y = cat(BV(2, 3), ph[PW - 3:0]), BV(48, 0), gain
    init_z = cat(init_t, BV(48 - PW, 0)) if PW < 48 else init_t               # :164-165
    x, y, z = init_x, init_y, init_z
    for ii in range(DW):                                                      # xl :234-242
        if z[47] == 0:
            xn, yn = x + y[47:ii], y - x[47:ii]
        else:
            xn, yn = x - y[47:ii], y + x[47:ii]
        if ii <= DW - 2:                                                      # xp :244-250
            z = z + BV(48, ROM_LUT48[ii]) if z[47] == 1 else z - BV(48, ROM_LUT48[ii])
        x, y = xn, yn
    return y[47:47 - (DW - 1)].signed(), x[47:47 - (DW - 1)].signed()         # :257-258


def window_dds48(win_type: int, phi_width: int, dat_width: int, aa, n: int) -> int:
    """The window entities with cordic_dds48 swapped in for cordic_dds (same port list; BASELINE
    config 3 names this composition - the reference itself never instantiates it)."""
    DW, PW, M = dat_width, phi_width, win_type
    AA = [BV(DW, a) for a in aa[:M]]
    b = [AA[0]]
    for k in range(1, M):
        _, ck = cordic_dds48(PW, DW, (k * n) & ((1 << PW) - 1))
        b.append(_round_product(AA[k], BV(DW, ck), DW))
    assert M == 7
    x3 = lambda v: cat(v[DW - 1], v[DW - 1], v)                               # noqa: E731
    dsp_pp = ((x3(b[0]) - x3(b[1])) + (x3(b[2]) - x3(b[3]))) + ((x3(b[4]) - x3(b[5])) + x3(b[6]))
    out = dsp_pp[DW + 1:2]
    return (out + 1 if dsp_pp[1] else out).signed()
