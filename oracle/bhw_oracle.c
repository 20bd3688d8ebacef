/*
 * bhw_oracle.c - plain-C restatement of the reference's window-generation path.
 *
 * TEST INFRASTRUCTURE ONLY (see bhw_oracle.h).  Every function follows the
 * reference file:line it cites; nothing here is shared with the CUDA product.
 * All arithmetic is done in int64_t / __int128 with explicit wrap ("sx") to the
 * bit width the reference signal has, so the code reads like the RTL/C++ it
 * restates rather than like an optimised implementation.
 *
 * Parity: HLS and CPP models are pinned against the compiled reference
 * (oracle/_ref).  The RTL model is pinned by vectors recorded while EXECUTING
 * the reference's VHDL sources (src/, every .vhd) in oracle/vhdl_sim.py (tests/golden/rtl_sim_vectors.npz,
 * tests/test_rtl_vhdl_sim.py); for the TAYLOR path that execution relies on our
 * behavioural model of the third-party DSP48E1/E2 primitives and on libm for
 * ieee.math_real.  Second anchors: oracle/rtl_bitvec.py + tests/golden KATs.
 */
#include "bhw_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef __int128 i128;

/* wrap v to a b-bit two's-complement value, 1 <= b <= 64 */
static inline int64_t sx(int64_t v, int b) {
  if (b >= 64) return v;
  uint64_t m = 1ull << (b - 1);
  uint64_t x = (uint64_t)v & ((m << 1) - 1);
  return (int64_t)((x ^ m) - m);
}
static inline int64_t sx128(i128 v, int b) { return sx((int64_t)v, b); } /* b <= 64: low bits suffice */

/* atan ROM with pi/4 -> 2^46: src/cordic_dds.vhd:104-117 (same constants in
 * hls/cordic/cordic.cpp:57-70, hls/windows/win_function.cpp:59-72). */
static const int64_t ROM4[48] = {
    0x400000000000, 0x25C80A3B3BE6, 0x13F670B6BDC7, 0x0A2223A83BBB, 0x05161A861CB1, 0x028BAFC2B209,
    0x0145EC3CB850, 0x00A2F8AA23A9, 0x00517CA68DA2, 0x0028BE5D7661, 0x00145F300123, 0x000A2F982950,
    0x000517CC19C0, 0x00028BE60D83, 0x000145F306D6, 0x0000A2F9836D, 0x0000517CC1B7, 0x000028BE60DC,
    0x0000145F306E, 0x00000A2F9837, 0x00000517CC1B, 0x0000028BE60E, 0x00000145F307, 0x000000A2F983,
    0x000000517CC2, 0x00000028BE61, 0x000000145F30, 0x0000000A2F98, 0x0000000517CC, 0x000000028BE6,
    0x0000000145F3, 0x00000000A2FA, 0x00000000517D, 0x0000000028BE, 0x00000000145F, 0x000000000A30,
    0x000000000518, 0x00000000028C, 0x000000000146, 0x0000000000A3, 0x000000000051, 0x000000000029,
    0x000000000014, 0x00000000000A, 0x000000000005, 0x000000000003, 0x000000000001, 0x000000000000};
/* atan ROM with pi/4 -> 2^45: src/cordic_dds48.vhd:115-128,
 * src/cordic_dds_scaled.vhd:117-130, cpp/cordic_sincos.cpp:97-110.
 * Independently rounded: not ROM4 >> 1. */
static const int64_t ROM2[48] = {
    0x200000000000, 0x12E4051D9DF3, 0x09FB385B5EE4, 0x051111D41DDE, 0x028B0D430E59, 0x0145D7E15904,
    0x00A2F61E5C28, 0x00517C5511D4, 0x0028BE5346D1, 0x00145F2EBB31, 0x000A2F980092, 0x000517CC14A8,
    0x00028BE60CE0, 0x000145F306C1, 0x0000A2F9836B, 0x0000517CC1B7, 0x000028BE60DC, 0x0000145F306E,
    0x00000A2F9837, 0x00000517CC1B, 0x0000028BE60E, 0x00000145F307, 0x000000A2F983, 0x000000517CC2,
    0x00000028BE61, 0x000000145F30, 0x0000000A2F98, 0x0000000517CC, 0x000000028BE6, 0x0000000145F3,
    0x00000000A2FA, 0x00000000517D, 0x0000000028BE, 0x00000000145F, 0x000000000A30, 0x000000000518,
    0x00000000028C, 0x000000000146, 0x0000000000A3, 0x000000000051, 0x000000000029, 0x000000000014,
    0x00000000000A, 0x000000000005, 0x000000000003, 0x000000000001, 0x000000000001, 0x000000000000};
static const int64_t GAIN_A = 0x4DBA76D421AF; /* src/cordic_dds.vhd:97 */
static const int64_t GAIN_B = 0x26DD3B6A10D8; /* src/cordic_dds48.vhd:110, cpp/cordic_sincos.cpp:21 */

/* output-side quadrant fix shared by cordic_dds (src/cordic_dds.vhd:232-246)
 * and taylor_sincos (src/taylor_sincos.vhd:237-255): negation is not(v)+1 in
 * dw bits. */
static void quad_fix(int q, int64_t s, int64_t c, int dw, int64_t* so, int64_t* co) {
  switch (q & 3) {
    case 0: *so = s; *co = c; break;
    case 1: *so = c; *co = sx(~s + 1, dw); break;
    case 2: *so = sx(~s + 1, dw); *co = sx(~c + 1, dw); break;
    default: *so = sx(~c + 1, dw); *co = s; break;
  }
}

/* ---- (A) cordic_dds: src/cordic_dds.vhd:97-249 --------------------------- */
void orc_cordic_dds(int pw, int dw, int prec, uint64_t ph, int64_t* s, int64_t* c) {
  const int w = dw + prec;                      /* sigX/Y/Z width :134-135 */
  const int64_t gain = GAIN_A >> (49 - w);      /* "0" & GAIN48(47 downto 48-w+1) :98 */
  ph &= (1ull << pw) - 1;
  const int q = (int)(ph >> (pw - 2));          /* quadz1/quadz2 :170-172 */
  const uint64_t t = ph & ((1ull << (pw - 2)) - 1); /* init_t :179 */
  int64_t z;
  if (pw >= dw) z = (int64_t)((t >> (pw - dw)) << prec); /* xPHI_LESS :159-162 */
  else z = (int64_t)(t << (dw - pw + prec));             /* xPHI_MORE :163-166 */
  z = sx(z, w);
  int64_t x = gain, y = 0;                      /* :177-178 */
  for (int i = 0; i <= dw - 2; i++) {           /* lpXY / lpZ :197-213 */
    const int64_t rom = ROM4[i] >> (49 - w);    /* func_atan :123-131 */
    int64_t xn, yn, zn;
    if (z < 0) { xn = x + (y >> i); yn = y - (x >> i); zn = z + rom; }
    else       { xn = x - (y >> i); yn = y + (x >> i); zn = z - rom; }
    x = sx(xn, w); y = sx(yn, w); z = sx(zn, w);
  }
  const int64_t ds = sx(y >> prec, dw);         /* dat_sin :218 */
  const int64_t dc = sx(x >> prec, dw);         /* dat_cos :219 */
  quad_fix(q, ds, dc, dw, s, c);
}

/* ---- (B)/(C) cordic_dds48 and cordic_dds_scaled -------------------------- */
/* one algorithm, parameters (size, dwph): src/cordic_dds48.vhd:110-259,
 * src/cordic_dds_scaled.vhd:100-285 */
static void cordic_inq(int pw, int dw, int size, int dwph, uint64_t ph, int64_t* s, int64_t* c) {
  const int64_t g = GAIN_B >> (48 - size);      /* GAIN32 = GAIN48(47 downto 48-size) :110-111 */
  ph &= (1ull << pw) - 1;
  const int q = (int)(ph >> (pw - 2));
  const uint64_t low = ph & ((1ull << (pw - 2)) - 1);
  uint64_t t; int64_t x, y;
  switch (q) {                                  /* pr_phi / pr_xy: dds48 :170-216, scaled :196-242 */
    case 1:  t = low;                         x = 0; y = sx(~g + 1, size); break;
    case 2:  t = low | (3ull << (pw - 2));    x = 0; y = g; break;
    default: t = ph;                          x = g; y = 0; break;
  }
  /* init_z: phase left-aligned in dwph bits (dds48 :164-165; scaled :186-192) */
  int64_t z = sx((int64_t)(t << (dwph - pw)), dwph);
  for (int i = 0; i <= dw - 1; i++) {           /* xl: DW iterations of X/Y (dds48 :234-242) */
    int64_t xn, yn;
    if (z >= 0) { xn = x + (y >> i); yn = y - (x >> i); }
    else        { xn = x - (y >> i); yn = y + (x >> i); }
    if (i <= dw - 2) {                          /* xp: Z advances only DW-1 times (:244-250) */
      const int64_t rom = ROM2[i] >> (48 - dwph);
      z = sx(z < 0 ? z + rom : z - rom, dwph);
    }
    x = sx(xn, size); y = sx(yn, size);
  }
  *s = sx(y >> (size - dw), dw);                /* top DW bits (dds48 :257-258) */
  *c = sx(x >> (size - dw), dw);
}
void orc_cordic_dds48(int pw, int dw, uint64_t ph, int64_t* s, int64_t* c) {
  cordic_inq(pw, dw, 48, 48, ph, s, c);
}
static const int SEL_SIZE[25] = {15, 15, 15, 18, 21, 22, 23, 26, 30, 31, 32, 33, 38,
                                 38, 38, 42, 42, 45, 47, 47, 47, 48, 48, 48, 48}; /* scaled :102-107 */
void orc_cordic_dds_scaled(int pw, int dw, uint64_t ph, int64_t* s, int64_t* c) {
  const int size = SEL_SIZE[dw - 8];
  const int dwph = size < pw ? pw : size;       /* func_width :132-143 */
  cordic_inq(pw, dw, size, dwph, ph, s, c);
}

/* ---- (E) taylor_sincos + tay1_order -------------------------------------- */
/* rom_calculate: src/taylor_sincos.vhd:91-111.  VHDL INTEGER(real) rounds to
 * nearest; llround() differs only on exact .5 ties, which only ii = 0 could
 * produce and that entry is an exact integer. */
void orc_taylor_rom(int dw, int lut, int64_t* rom_cos, int64_t* rom_sin) {
  const int depth = 1 << lut;
  const double amp = ldexp(1.0, dw - 1) - 1.0;
  for (int ii = 0; ii < depth; ii++) {
    const double pi_new = ((double)ii * M_PI) / (2.0 * (double)depth);
    rom_cos[ii] = sx(llround(amp * cos(pi_new)), dw);
    rom_sin[ii] = sx(llround(amp * sin(pi_new)), dw);
  }
}
static void taylor_core(int pw, int dw, int lut, const int64_t* rc, const int64_t* rs, uint64_t cnt,
                        int64_t* s, int64_t* c) {
  cnt &= (1ull << pw) - 1;
  const int q = (int)(cnt >> (pw - 2));                         /* :141 */
  const uint64_t t = cnt & ((1ull << (pw - 2)) - 1);
  const int d = pw - lut;
  int64_t ms, mc;
  if (d < 2) {                                                  /* xGEN_LESS :157-161 */
    const uint64_t addr = (t << (lut - pw + 2)) & ((1ull << lut) - 1);
    mc = rc[addr]; ms = rs[addr];
  } else if (d == 2) {                                          /* xGEN_EQ :164-167 */
    mc = rc[t]; ms = rs[t];
  } else {                                                      /* xGEN_MORE :169-217 */
    const uint64_t addr = t >> (pw - lut - 2);                  /* cnt(PW-3 downto PW-LUT-2) :190 */
    const int64_t acnt = (int64_t)(t & ((1ull << (pw - lut - 2)) - 1)); /* :191 */
    const int stage = pw - lut - 3;                             /* :200 */
    const int64_t c0 = rc[addr], s0 = rs[addr];
    /* tay1_order: ramb_pi, mpi (src/tay1_order.vhd:130-147), XSHIFT (:112) */
    const int64_t ramb_pi = (int64_t)round(M_PI * ldexp(1.0, 17 - stage));
    const int64_t mpi = (ramb_pi * acnt) & 0xFFFFFF;
    const int xs = 19 + lut;
    if (dw < 19) {  /* DSP48 MACC: C -/+ A*B, slice [xs+dw-1:xs] (:180-504) */
      mc = sx128((((i128)c0 << xs) - (i128)mpi * s0) >> xs, dw);
      ms = sx128((((i128)s0 << xs) + (i128)mpi * c0) >> xs, dw);
    } else {        /* wide multiplier + add/sub + saturation (:506-617) */
      const int64_t m1 = sx128(((i128)s0 * mpi) >> xs, dw);     /* mlt1_bb :585 */
      const int64_t m2 = sx128(((i128)c0 * mpi) >> xs, dw);     /* mlt2_bb :586 */
      const int64_t cp = sx(c0 - m1, dw);                       /* cos_pdt :595 */
      const int64_t sp = sx(s0 + m2, dw);                       /* sin_pdt :596 */
      const int64_t sat = ((int64_t)1 << (dw - 1)) - 1;
      mc = cp < 0 ? sat : cp;                                   /* pr_rnd :602-617 */
      ms = sp < 0 ? sat : sp;
    }
  }
  quad_fix(q, ms, mc, dw, s, c);                                /* pr_quad :237-255 */
}
void orc_taylor_sincos(int pw, int dw, int lut, uint64_t cnt, int64_t* s, int64_t* c) {
  const int depth = 1 << lut;
  int64_t* rc = (int64_t*)malloc(sizeof(int64_t) * 2 * depth);
  orc_taylor_rom(dw, lut, rc, rc + depth);
  taylor_core(pw, dw, lut, rc, rc + depth, cnt, s, c);
  free(rc);
}

/* ---- (F) HLS cordic: hls/windows/win_function.cpp:47-156 ------------------ */
/* (identical text in hls/cordic/cordic.cpp:45-154) */
void orc_hls_cordic(int np, int nw, uint64_t phi, int64_t* s, int64_t* c) {
  const int w = nw + 2;                                         /* dat_t: win_function.h:61 */
  const int64_t gain = sx(GAIN_B >> (46 - nw), w);              /* GAIN48 :83 */
  phi &= (1ull << np) - 1;
  const int q = (int)(phi >> (np - 2));                         /* quadrant :86 */
  const int64_t t = (int64_t)(phi & ((1ull << (np - 2)) - 1));  /* init_t :88 */
  int64_t z;
  if (np - 1 < nw) z = sx(t << (nw - np + 2), w);               /* :91-92 */
  else z = sx((t >> (np - nw)) << 2, w);                        /* :94-95 */
  int64_t x = gain, y = 0;
  for (int k = 0; k < nw; k++) {                                /* stg :110-125 */
    /* lut_angle has NW-1 entries (:74-80); the last pass reads one past the
     * end, which only feeds z[NW] - never used - so any value works here. */
    const int64_t lut = k < nw - 1 ? sx((ROM4[k] >> (47 - nw)) & 0xFFFFFFFFFFll, w) : 0;
    int64_t xn, yn, zn;
    if (z < 0) { xn = x + (y >> k); yn = y - (x >> k); zn = z + lut; }
    else       { xn = x - (y >> k); yn = y + (x >> k); zn = z - lut; }
    x = sx(xn, w); y = sx(yn, w); z = sx(zn, w);
  }
  const int64_t oc = x >> 2, os = y >> 2;                       /* :128-129 */
  int64_t dc, ds;
  switch (q) {                                                  /* :135-154, in dat_t */
    case 0: ds = os; dc = oc; break;
    case 1: ds = oc; dc = sx(~os + 1, w); break;
    case 2: ds = sx(~os + 1, w); dc = sx(~oc + 1, w); break;
    default: ds = sx(~oc + 1, w); dc = os; break;
  }
  *c = sx(dc, nw);                                              /* win_t truncation :153-154 */
  *s = sx(ds, nw);
}

/* ---- (G) plain C++ cordic: cpp/cordic_sincos.cpp:10-92 -------------------- */
void orc_cpp_cordic(int pw, int dw, int theta, int* s, int* c) {
  const int precision = 1;                                      /* :12 */
  const long long gain = GAIN_B >> (48 - dw - 2);               /* GAIN32 :21-22 */
  const int q = theta >> (pw - 2);                              /* :25 */
  const long long t = theta & (~(0x3 << (pw - 2)));             /* :27 */
  long long z;
  if (pw - 1 < dw) z = t << (dw - pw + precision);              /* :31-33 */
  else z = (t >> (pw - dw)) << precision;                       /* :34-36 */
  long long x = gain, y = 0;
  for (int k = 0; k < dw; k++) {                                /* :49-63 */
    /* lut_angle has DW-1 entries (:13-18); pass DW-1 reads past the end and
     * only produces z[DW], which is never used. */
    const long long lut = k < dw - 1 ? ((ROM2[k] >> (48 - dw - precision)) & 0xFFFFFFFFFFFFll) : 0;
    long long xn, yn;
    if (z < 0) { xn = x + (y >> k); yn = y - (x >> k); z = z + lut; }
    else       { xn = x - (y >> k); yn = y + (x >> k); z = z - lut; }
    x = xn; y = yn;
  }
  const long long oc = x >> 2, os = y >> 2;                     /* :64-65 */
  long long dc, ds;
  switch (q) {                                                  /* ones' complement, :70-86 */
    case 0: ds = os; dc = oc; break;
    case 1: ds = oc; dc = ~os; break;
    case 2: ds = ~os; dc = ~oc; break;
    default: ds = ~oc; dc = os; break;
  }
  *c = (int)dc; *s = (int)ds;                                   /* :89-90 */
}

/* ---- descriptor handling -------------------------------------------------- */
static int eff_prec(const bhw_desc* d) { return d->precision == 0 ? 1 : d->precision; }
static int eff_lut(const bhw_desc* d) { return d->lut_size == 0 ? 9 : d->lut_size; }

static int validate_source(const bhw_desc* d, int for_window) {
  const int pw = d->phi_width, dw = d->dat_width;
  if (pw < BHW_MIN_PHI_WIDTH || pw > BHW_MAX_PHI_WIDTH) return BHW_E_PHI_WIDTH;   /* 16 .. 64M points (README.md:2) */
  switch (d->model) {
    case BHW_MODEL_RTL:
      switch (d->sin_type) {
        case BHW_SIN_CORDIC: {
          const int p = eff_prec(d);
          if (dw < 4 || dw > 48) return BHW_E_DAT_WIDTH;
          if (p < 1 || p > 7 || dw + p > 49) return BHW_E_PRECISION;
          if (for_window && p != 1) return BHW_E_PRECISION; /* windows never override it: src/hamming_win.vhd:153-157 */
          return BHW_OK;
        }
        case BHW_SIN_CORDIC48:
          if (dw < 4 || dw > 48) return BHW_E_DAT_WIDTH;
          return BHW_OK;
        case BHW_SIN_CORDIC_SCALED:
          if (dw < 8 || dw > 32) return BHW_E_DAT_WIDTH;
          return BHW_OK;
        case BHW_SIN_TAYLOR: {
          const int lut = eff_lut(d);
          if (for_window && d->win_type != BHW_WIN_HAMMING && d->win_type != BHW_WIN_BH3TERM)
            return BHW_E_SIN_TYPE;                              /* src/bh_win_4term.vhd:57-61 */
          if (dw < 4 || dw > 32) return BHW_E_DAT_WIDTH;
          if (lut < 1 || lut > 16) return BHW_E_LUT_SIZE;
          const int nunits = (for_window && d->win_type == BHW_WIN_BH3TERM) ? 2 : 1;
          if (nunits == 2 && pw - lut == 3) return BHW_E_LUT_SIZE; /* src/bh_win_3term.vhd:30-31 */
          for (int u = 0; u < nunits; u++) {
            const int dd = (pw - u) - lut;
            if (dd > 2) {
              if (dd - 3 > 15) return BHW_E_LUT_SIZE;           /* cnt_exp is 16 bits: tay1_order.vhd:116-127 */
              if (dw < 19 && 19 + lut + dw > 48) return BHW_E_LUT_SIZE; /* slice of a 48-bit P */
              if (dw > 18 && 19 + lut + dw > 62) return BHW_E_LUT_SIZE; /* slice of a 62-bit product */
            }
          }
          return BHW_OK;
        }
        default: return BHW_E_SIN_TYPE;
      }
    case BHW_MODEL_HLS:
      if (d->sin_type != BHW_SIN_CORDIC) return BHW_E_SIN_TYPE;
      if (dw < 4 || dw > 32) return BHW_E_DAT_WIDTH;
      if (pw > dw + 2) return BHW_E_PHI_WIDTH;  /* init_t no longer fits dat_t */
      return BHW_OK;
    case BHW_MODEL_CPP:
      if (for_window) return BHW_E_MODEL;
      if (d->sin_type != BHW_SIN_CORDIC) return BHW_E_SIN_TYPE;
      if (dw < 4 || dw > 32) return BHW_E_DAT_WIDTH;
      return BHW_OK;
    default: return BHW_E_MODEL;
  }
}

int orc_validate(const bhw_desc* d) {
  if (!d) return BHW_E_NULL;
  const int m = d->win_type;
  if (m < 2 || m > BHW_MAX_TERMS) return BHW_E_WIN_TYPE;
  /* 6 and 8..11 terms: the structure of bh_win_3term .. bh_win_7term continued (one cordic_dds per harmonic, the
   * same product rounding, the same DW+2-bit sum and rounding on bit 1) for the coefficient sets the reference only
   * tabulates (doc/blackman-harris coef.jpg).  Not a reference entity: RTL model, CORDIC sources only. */
  if (m == 6 || m > 7) {
    if (d->model != BHW_MODEL_RTL || d->sin_type == BHW_SIN_TAYLOR) return BHW_E_WIN_TYPE;
  }
  int st = validate_source(d, 1);
  if (st) return st;
  if (d->stream_offset != 0 && d->stream_offset != 1) return BHW_E_ARG;
  /* the container of the output is not the oracle's business (it returns int64), its legality is */
  if (d->out_format != BHW_OUT_DEFAULT && d->out_format != BHW_OUT_INT16) return BHW_E_ARG;
  if (d->out_format == BHW_OUT_INT16 && d->dat_width > 16) return BHW_E_DAT_WIDTH;
  const int dw = d->dat_width;
  for (int k = 0; k < m; k++) {
    const int64_t v = d->aa[k];
    if (dw < 63) {
      if (v < -((int64_t)1 << (dw - 1)) || v >= ((int64_t)1 << dw)) return BHW_E_COEFF;
      if (d->model == BHW_MODEL_HLS && v >= ((int64_t)1 << (dw - 1))) return BHW_E_COEFF;
    }
  }
  return BHW_OK;
}

typedef struct {
  const bhw_desc* d;
  int64_t* rom[2]; /* Taylor ROMs of unit 1 (pw) and, for 3-term, unit 2 (pw-1): same table, kept once */
} ctx_t;

/* cos of harmonic k at sample n, per the entity's wiring */
static int64_t harmonic_cos(const ctx_t* cx, int k, uint64_t n) {
  const bhw_desc* d = cx->d;
  const int pw = d->phi_width, dw = d->dat_width;
  int64_t s, c;
  if (d->model == BHW_MODEL_HLS) {            /* cordic(k*i, ...): win_function.cpp:361-366 */
    orc_hls_cordic(pw, dw, (uint64_t)k * n, &s, &c);
    return c;
  }
  switch (d->sin_type) {
    case BHW_SIN_CORDIC:                      /* ph_in_k += k: src/bh_win_7term.vhd:176-197 */
      orc_cordic_dds(pw, dw, eff_prec(d), (uint64_t)k * n, &s, &c); break;
    case BHW_SIN_CORDIC48: orc_cordic_dds48(pw, dw, (uint64_t)k * n, &s, &c); break;
    case BHW_SIN_CORDIC_SCALED: orc_cordic_dds_scaled(pw, dw, (uint64_t)k * n, &s, &c); break;
    default: {                                /* TAYLOR: each unit owns a +1 counter; the 2nd
                                                 harmonic is a PHASE_WIDTH-1 unit
                                                 (src/bh_win_3term.vhd:205-234) */
      const int lut = eff_lut(d), depth = 1 << lut;
      taylor_core(pw - (k - 1), dw, lut, cx->rom[0], cx->rom[0] + depth, n, &s, &c);
    }
  }
  return c;
}

static int64_t window_sample(const ctx_t* cx, uint64_t n) {
  const bhw_desc* d = cx->d;
  const int m = d->win_type, dw = d->dat_width;
  if (d->model == BHW_MODEL_HLS) {
    /* m_k = (a_k*c_k) >> (NW-2), out = (win_t)(a0 - m1 + m2 - ...):
     * hls/windows/win_function.cpp:182,197,222-225,271-275,327-332,368-375 */
    i128 acc = d->aa[0];
    for (int k = 1; k < m; k++) {
      const i128 mk = ((i128)d->aa[k] * harmonic_cos(cx, k, n)) >> (dw - 2);
      acc += (k & 1) ? -mk : mk;
    }
    return sx128(acc, dw);
  }
  /* RTL tail, identical in all five entities (src/hamming_win.vhd:192-231,
   * src/bh_win_3term.vhd:258-306, bh_win_4term.vhd:225-280, bh_win_5term.vhd:
   * 281-347, bh_win_7term.vhd:350-438) */
  int64_t b[BHW_MAX_TERMS];
  b[0] = sx(d->aa[0], dw);                                    /* dsp_b0 <= AA0 */
  for (int k = 1; k < m; k++) {
    const i128 p = (i128)sx(d->aa[k], dw) * harmonic_cos(cx, k, n); /* int_multNxN_dsp48.vhd:105 */
    const int64_t r = sx128(p >> (dw - 2), dw + 1);           /* mult_p(2DW-2 downto DW-2) */
    b[k] = sx((r >> 1) + (r & 1), dw);                        /* pr_rnd: +1 when bit0 set */
  }
  if (m == 2) {
    const int64_t pp = sx(b[0] - b[1], dw + 1);               /* hamming_win.vhd:214 */
    return sx((pp >> 1) + (pp & 1), dw);                      /* :220-228 */
  }
  int64_t sum = 0;
  for (int k = 0; k < m; k++) sum += (k & 1) ? -b[k] : b[k];
  const int64_t pp = sx(sum, dw + 2);                         /* dsp_pp is DW+2 bits */
  return sx((pp >> 2) + ((pp >> 1) & 1), dw);                 /* rounds on bit 1: bh_win_3term.vhd:295-306 */
}

static int ctx_init(ctx_t* cx, const bhw_desc* d) {
  cx->d = d; cx->rom[0] = cx->rom[1] = NULL;
  if (d->model == BHW_MODEL_RTL && d->sin_type == BHW_SIN_TAYLOR) {
    const int depth = 1 << eff_lut(d);
    cx->rom[0] = (int64_t*)malloc(sizeof(int64_t) * 2 * depth);
    if (!cx->rom[0]) return BHW_E_ALLOC;
    orc_taylor_rom(d->dat_width, eff_lut(d), cx->rom[0], cx->rom[0] + depth);
  }
  return BHW_OK;
}
static void ctx_free(ctx_t* cx) { free(cx->rom[0]); }

int orc_window(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out) {
  int st = orc_validate(d);
  if (st) return st;
  if (!out && count) return BHW_E_NULL;
  const uint64_t N = 1ull << d->phi_width;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
  ctx_t cx;
  if ((st = ctx_init(&cx, d))) return st;
  for (uint64_t j = 0; j < count; j++)
    out[j] = window_sample(&cx, (n0 + j + (uint64_t)d->stream_offset) & (N - 1));
  ctx_free(&cx);
  return BHW_OK;
}

int orc_window_i32(const bhw_desc* d, uint64_t n0, uint64_t count, int32_t* out) {
  int st = orc_validate(d);
  if (st) return st;
  if (d->dat_width > 32) return BHW_E_ELEM;
  const uint64_t N = 1ull << d->phi_width;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
  ctx_t cx;
  if ((st = ctx_init(&cx, d))) return st;
  for (uint64_t j = 0; j < count; j++)
    out[j] = (int32_t)window_sample(&cx, (n0 + j + (uint64_t)d->stream_offset) & (N - 1));
  ctx_free(&cx);
  return BHW_OK;
}

/* The apply step: y[f*N + n] = x[f*N + n] * w[n] through int_multNxN_dsp48 - DAT_Q <= SIGNED(sig_a) *
 * SIGNED(sig_b), DTW = DAT_WIDTH bits per port, 2*DTW bits out (src/int_multNxN_dsp48.vhd:77-82,105).  x: the
 * low DAT_WIDTH bits are the DAT_A port.  mode 0: DAT_Q; mode 1: the window entities' own use of it,
 * r = DAT_Q[2DW-2 : DW-2] (src/hamming_win.vhd:195), y = r[DW:1] (+1 when r[0]) in DW bits (:198-208). */
int orc_apply(const bhw_desc* d, int mode, const int32_t* x, uint64_t frames, int64_t* y) {
  int st = orc_validate(d);
  if (st) return st;
  if (d->dat_width > 32) return BHW_E_DAT_WIDTH;
  if (mode != 0 && mode != 1) return BHW_E_ARG;
  const int dw = d->dat_width;
  const uint64_t N = 1ull << d->phi_width;
  int64_t* w = (int64_t*)malloc(N * sizeof(int64_t));
  if (!w) return BHW_E_ALLOC;
  st = orc_window(d, 0, N, w);
  for (uint64_t f = 0; !st && f < frames; f++)
    for (uint64_t n = 0; n < N; n++) {
      const int64_t a = (int64_t)((uint64_t)(int64_t)x[f * N + n] << (64 - dw)) >> (64 - dw);   /* DAT_A, signed DW bits */
      const int64_t q = a * w[n];                                                                /* DAT_Q */
      if (mode == 0) { y[f * N + n] = q; continue; }
      const int64_t r = (int64_t)((uint64_t)(q >> (dw - 2)) << (63 - dw)) >> (63 - dw);          /* DW+1 bits */
      const int64_t b = (r >> 1) + (r & 1);
      y[f * N + n] = (int64_t)((uint64_t)b << (64 - dw)) >> (64 - dw);                           /* DW bits */
    }
  free(w);
  return st;
}

typedef struct { const bhw_desc* d; uint64_t n0, count; int64_t* out; int st; } job_t;
static void* mt_worker(void* p) {
  job_t* j = (job_t*)p;
  j->st = orc_window(j->d, j->n0, j->count, j->out);
  return NULL;
}
int orc_window_mt(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256]; job_t jobs[256];
  const uint64_t per = (count + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
  int used = 0;
  for (int i = 0; i < nthreads; i++) {
    const uint64_t b = per * (uint64_t)i;
    if (b >= count) break;
    const uint64_t c = count - b < per ? count - b : per;
    jobs[i] = (job_t){d, n0 + b, c, out + b, 0};
    pthread_create(&th[i], NULL, mt_worker, &jobs[i]);
    used++;
  }
  int st = BHW_OK;
  for (int i = 0; i < used; i++) { pthread_join(th[i], NULL); if (jobs[i].st) st = jobs[i].st; }
  return st;
}

int orc_sincos(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out_sin, int64_t* out_cos) {
  if (!d) return BHW_E_NULL;
  int st = validate_source(d, 0);
  if (st) return st;
  const int pw = d->phi_width, dw = d->dat_width;
  const uint64_t N = 1ull << pw;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
  ctx_t cx;
  if ((st = ctx_init(&cx, d))) return st;
  for (uint64_t j = 0; j < count; j++) {
    const uint64_t n = (n0 + j) & (N - 1);
    int64_t s = 0, c = 0;
    if (d->model == BHW_MODEL_HLS) orc_hls_cordic(pw, dw, n, &s, &c);
    else if (d->model == BHW_MODEL_CPP) { int si, ci; orc_cpp_cordic(pw, dw, (int)n, &si, &ci); s = si; c = ci; }
    else switch (d->sin_type) {
      case BHW_SIN_CORDIC: orc_cordic_dds(pw, dw, eff_prec(d), n, &s, &c); break;
      case BHW_SIN_CORDIC48: orc_cordic_dds48(pw, dw, n, &s, &c); break;
      case BHW_SIN_CORDIC_SCALED: orc_cordic_dds_scaled(pw, dw, n, &s, &c); break;
      default: { const int depth = 1 << eff_lut(d);
                 taylor_core(pw, dw, eff_lut(d), cx.rom[0], cx.rom[0] + depth, n, &s, &c); }
    }
    if (out_sin) out_sin[j] = s;
    if (out_cos) out_cos[j] = c;
  }
  ctx_free(&cx);
  return BHW_OK;
}

/* ---- cordic_atan2: src/cordic_atan2.vhd:80-220 ----------------------------- */
int orc_atan2_validate(int iw, int aw, int prec) {
  if (prec == 0) prec = 1;
  if (aw < 4 || aw > 32) return BHW_E_DAT_WIDTH;
  if (iw > 32 || iw < aw - 1) return BHW_E_PHI_WIDTH;   /* VEC_DX(ii), ii <= ANGLE_WIDTH-2 (:140-141) */
  if (prec < 1 || prec > 7) return BHW_E_PRECISION;
  return BHW_OK;
}

/* (qx, qy): the pair whose sign bits select the quadrant at the output.  In the entity as written the quadrant
 * shift registers quadz1/quadz2 are ANGLE_WIDTH stages long (:127-129) while the data takes ANGLE_WIDTH + 1 clocks
 * to reach dat_phi (init_x/init_y :136-146, sigX(0) :160, ANGLE_WIDTH-1 stages :166-184), so on a stream of pairs
 * PHI_DT for pair t is corrected with the quadrant of pair t+1 (seen by executing the VHDL: oracle/vhdl_sim.py,
 * tests/golden/rtl_sim_vectors.npz).  orc_cordic_atan2 = the aligned reading (qx, qy) = (vx, vy). */
int64_t orc_cordic_atan2_q(int iw, int aw, int prec, int64_t vx, int64_t vy, int64_t qx, int64_t qy);
int64_t orc_cordic_atan2(int iw, int aw, int prec, int64_t vx, int64_t vy) {
  return orc_cordic_atan2_q(iw, aw, prec, vx, vy, vx, vy);
}
int64_t orc_cordic_atan2_q(int iw, int aw, int prec, int64_t vx, int64_t vy, int64_t qx, int64_t qy) {
  if (prec == 0) prec = 1;
  const int w = aw + prec;                                 /* dat_array / phi_array element width :117-118 */
  const int sxb = (int)((vx >> (iw - 1)) & 1), syb = (int)((vy >> (iw - 1)) & 1);
  const int qxb = (int)((qx >> (iw - 1)) & 1), qyb = (int)((qy >> (iw - 1)) & 1);
  int64_t x = 0, y = 0, z = 0;                             /* init_z <= 0 :149 */
  for (int ii = 0; ii <= aw - 2; ii++) {                   /* pr_abs :136-146 */
    x |= (int64_t)(((vx >> ii) & 1) ^ sxb) << ii;
    y |= (int64_t)(((vy >> ii) & 1) ^ syb) << ii;
  }
  for (int ii = 0; ii <= aw - 2; ii++) {                   /* lpXY, lpZ :166-184 */
    /* ROM_TABLE(ii) = '0' & ROM_LUT(ii)(47 downto 47-(w-2))  :102-105 */
    const int64_t rom = ROM4[ii] >> (48 - (w - 1));
    int64_t xn, yn, zn;
    if (y >= 0) { xn = x + (y >> ii); yn = y - (x >> ii); zn = z - rom; }   /* sigY MSB = '0' */
    else        { xn = x - (y >> ii); yn = y + (x >> ii); zn = z + rom; }
    x = sx(xn, w); y = sx(yn, w); z = sx(zn, w);
  }
  const int64_t dat_phi = sx(z >> prec, aw);               /* sigZ(AW-1)(w-1 downto PRECISION) :188 */
  const int64_t phi_pi = (int64_t)1 << (aw - 2);           /* (ANGLE_WIDTH-2 => '1', others => '0') :121 */
  switch ((qxb << 1) | qyb) {                              /* quadrant = quadz1 & quadz2 :129-131; case :203-208 */
    case 0: return dat_phi;
    case 1: return sx(dat_phi + phi_pi, aw);
    case 2: return sx(~dat_phi + 1, aw);
    default: return sx(dat_phi - phi_pi, aw);
  }
}

int orc_atan2(int iw, int aw, int prec, const int32_t* x, const int32_t* y, int32_t* phi, uint64_t count) {
  int st = orc_atan2_validate(iw, aw, prec);
  if (st) return st;
  for (uint64_t j = 0; j < count; j++)
    phi[j] = (int32_t)orc_cordic_atan2(iw, aw, prec, (int64_t)(uint32_t)x[j], (int64_t)(uint32_t)y[j]);
  return BHW_OK;
}

/* PHI_DT as the entity streams it: pair t with the quadrant of pair t+1; after the last pair the inputs are 0. */
int orc_atan2_stream(int iw, int aw, int prec, const int32_t* x, const int32_t* y, int32_t* phi, uint64_t count) {
  int st = orc_atan2_validate(iw, aw, prec);
  if (st) return st;
  for (uint64_t j = 0; j < count; j++) {
    const int64_t qx = j + 1 < count ? (int64_t)(uint32_t)x[j + 1] : 0, qy = j + 1 < count ? (int64_t)(uint32_t)y[j + 1] : 0;
    phi[j] = (int32_t)orc_cordic_atan2_q(iw, aw, prec, (int64_t)(uint32_t)x[j], (int64_t)(uint32_t)y[j], qx, qy);
  }
  return BHW_OK;
}

/* ---- coefficient rules ---------------------------------------------------- */
/* variants per README.md:30-41; values per the entity headers */
static const double COEF[18][BHW_MAX_TERMS] = {
    {0.5434783, 1.0 - 0.5434783},                              /* 1 Hamming   tb :123-124 */
    {0.5, 0.5},                                                /* 2 Hann      hamming_win.vhd:14-16 */
    {0.42, 0.5, 0.08},                                         /* 3 Blackman  tb :114-116 */
    {0.4243801, 0.4973406, 0.0782793},                         /* 4 BH3       bh_win_3term.vhd:20 */
    {0.355768, 0.487396, 0.144323, 0.012604},                  /* 5 Nuttall   bh_win_4term.vhd:16-17 */
    {0.35875, 0.48829, 0.14128, 0.01168},                      /* 6 BH4       tb :103-106 */
    {0.3635819, 0.4891775, 0.1365995, 0.0106411},              /* 7 B-Nuttall bh_win_4term.vhd:18-19 */
    {1.000, 1.930, 1.290, 0.388, 0.030},                       /* 8 Flat-top  tb :90-94 */
    {0.3232153788877343, 0.4714921439576260, 0.1755341299601972, 0.0284969901061499,
     0.0012613570882927},                                      /* 9 BH5       bh_win_5term.vhd:14-19 */
    {0.271220360585039, 0.433444612327442, 0.218004122892930, 0.065785343295606,
     0.010761867305342, 0.000770012710581, 0.000013680883060}, /* 10 BH7      tb :67-73 */
    {0.27105140069342, 0.43329793923448, 0.21812299954311, 0.06592544638803, 0.01081174209837,
     0.00077658482522, 0.00001388721735},                      /* 11 BH7, README.md:45-51 (magnitudes) */
    {0.5383554, 0.4616446},                                    /* 12 Hamming, second set hamming_win.vhd:21-23 */
    {0.215578950, 0.416631580, 0.277263158, 0.083578947, 0.006947368},  /* 13 flat-top normalised bh_win_5term.vhd:28-33 */
    /* 14..18: the 6- and 8..11-term minimum-sidelobe sets the reference only tabulates (doc/blackman-harris coef.jpg,
       "Table 1. Coefficients of minimum sidelobe windows"); BHW_WIN_MTERM_* - no reference entity */
    {2.935578950102797e-001, 4.519357723474506e-001, 2.014164714263962e-001, 4.792610922105837e-002,
     5.026196426859393e-003, 1.375555679558877e-004},
    {2.533176817029088e-001, 4.163269305810218e-001, 2.288396213719708e-001, 8.157508425925879e-002,
     1.773592450349622e-002, 2.096702749032688e-003, 1.067741302205525e-004, 1.280702090361482e-006},
    {2.384331152777942e-001, 4.005545348643820e-001, 2.358242530472107e-001, 9.527918858383112e-002,
     2.537395516617152e-002, 4.152432907505835e-003, 3.685604163298180e-004, 1.384355593917030e-005,
     1.161808358932861e-007},
    {2.257345387130214e-001, 3.860122949150963e-001, 2.401294214106057e-001, 1.070542338664613e-001,
     3.325916184016952e-002, 6.873374952321475e-003, 8.751673238035159e-004, 6.008598932721187e-005,
     1.710716472110202e-006, 1.027272130265191e-008},
    {2.151527506679809e-001, 3.731348357785249e-001, 2.424243358446660e-001, 1.166907592689211e-001,
     4.077422105878731e-002, 1.000904500852923e-002, 1.639806917362033e-003, 1.651660820997142e-004,
     8.884663168541479e-006, 1.938617116029048e-007, 8.482485599330470e-010}};
static const int NTERMS[18] = {2, 2, 3, 3, 4, 4, 4, 5, 5, 7, 7, 2, 5, 6, 8, 9, 10, 11};

int orc_quantize(int variant, int rule, int dw, int64_t aa[BHW_MAX_TERMS], int32_t* win_type) {
  if (variant < 1 || variant > 18 || (rule != BHW_RULE_TB && rule != BHW_RULE_HLS)) return BHW_E_VARIANT;
  if (variant > 13 && rule != BHW_RULE_TB) return BHW_E_VARIANT;
  if (dw < 4 || dw > 48) return BHW_E_DAT_WIDTH;
  const int m = NTERMS[variant - 1];
  double scale;
  if (rule == BHW_RULE_TB) {        /* src/tb/tb_windows.vhd:75-127 */
    switch (m) {
      case 2: case 7: case 6: case 8: case 9: case 10: case 11:
        scale = ldexp(1.0, dw - 1) - 1.0; break;               /* :75-81, :126-127; the M-term extension follows the 7-term rule */
      case 3: scale = ldexp(1.0, dw) - 16.0; break;            /* :118-120 */
      case 4: scale = ldexp(1.0, dw) - 1.0; break;             /* :108-111 */
      default: scale = ldexp(1.0, dw - 2) - 1.0; break;        /* 5-term :96-100 */
    }
  } else {                          /* hls/windows/win_function.cpp:176-355 */
    scale = (m >= 5) ? ldexp(1.0, dw - 2) - 1.0 : ldexp(1.0, dw - 1) - 1.0;
  }
  for (int k = 0; k < BHW_MAX_TERMS; k++) aa[k] = 0;
  for (int k = 0; k < m; k++) {
    double a = COEF[variant - 1][k];
    if (rule == BHW_RULE_HLS && variant == 3) a *= 0.5;        /* 0.21/0.25/0.04: :206-208 */
    if (rule == BHW_RULE_HLS && variant == 1 && k == 1) a = 1 - 0.5434783; /* :174 */
    aa[k] = (int64_t)round(a * scale);                         /* VHDL integer(): nearest; C round() */
  }
  if (win_type) *win_type = m;
  return BHW_OK;
}
