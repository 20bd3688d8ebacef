// bhw_plan.cpp - CUDA-free planning helpers shared by the executor (bhw_api.cu) and by
// tests/hostcheck: table canonicalisation, fast-path eligibility, the 32-bit tail record and
// the Taylor ROM contents.
#include <math.h>
#include <string.h>

#include "bhw_plan.h"

namespace bhw {

// Source as the table builder sees it: phase bits the source ignores are removed, so tables of
// e.g. cordic_dds with PHASE_WIDTH 17..26 and DATA_WIDTH 16 are one and the same table.
SrcParams canonical_source(const SrcParams& sp, uint32_t* drop) {
  SrcParams c = sp;
  *drop = 0;
  if ((sp.kind == SRC_DDS || sp.kind == SRC_HLS) && sp.z_rshift > 0) {
    *drop = (uint32_t)sp.z_rshift;
    c.pw = sp.pw - sp.z_rshift;
    c.z_rshift = 0;
  }
  return c;
}

bool fast32_ok(const SrcParams& sp) {
  // 32-bit registers, no wrap possible (DESIGN.md "no-wrap argument")
  if (sp.kind == SRC_DDS) return sp.dw >= 8 && sp.w <= 32;
  if (sp.kind == SRC_HLS) return sp.dw >= 8 && sp.w <= 32;
  return false;
}

// Which shift-add core the table builder may use (TABCORE_*): the plain 32-bit one, its biased
// form (cordic_dds with DW+PRECISION == 33, i.e. DAT_WIDTH 32 in the window entities;
// out_shift >= 1), the left-aligned 64-bit one, or the generic body.
int table_core32(const SrcParams& sp) {
  if (fast32_ok(sp)) return TABCORE_32;
  if (sp.kind == SRC_DDS && sp.dw >= 8 && sp.w == 33 && sp.out_shift >= 1 && sp.n_xy <= 31) return TABCORE_32BIAS;
  if ((sp.kind == SRC_INQ || sp.kind == SRC_DDS || sp.kind == SRC_HLS) && sp.w >= 8 && sp.w <= 64 && sp.zw >= 8 &&
      sp.zw <= 64 && sp.n_xy <= 48 && sp.pw >= 3 && sp.z_lshift + (64 - sp.zw) <= 63)
    return TABCORE_A64;
  return TABCORE_GENERIC;
}

// Core choice + sliced atan words for the one-thread-per-sample kernels (any source, canonical or not).
void init_src_core(const SrcParams& sp, SrcCore* sc) {
  memset(sc, 0, sizeof(*sc));
  if (sp.kind == SRC_TAYLOR) return;
  sc->core = (uint32_t)table_core32(sp);
  for (int i = 0; i < sp.n_z && i < 48; i++) {
    const int64_t r = (c_atan[sp.rom_sel][i] >> sp.rom_shift) & sp.rom_mask;
    if (i < 32 && (sc->core == TABCORE_32 || sc->core == TABCORE_32BIAS)) sc->rom32[i] = (int32_t)r;
    if (sc->core == TABCORE_A64) sc->rom64[i] = (int64_t)((uint64_t)r << (64 - sp.zw));
  }
}

bool table_build_unrolled_ok(const TabJob& j) {
  if (j.sp.kind == SRC_INQ)   // cordic_dds48 / cordic_dds_scaled at DAT_WIDTH 16, 17, 24, 32
    return j.fast == TABCORE_A64 && (j.sp.n_xy == 16 || j.sp.n_xy == 17 || j.sp.n_xy == 24 || j.sp.n_xy == 32);
  if (j.sp.kind == SRC_TAYLOR) return false;
  if (j.fast == TABCORE_32) return j.sp.n_xy == 15 || j.sp.n_xy == 16 || j.sp.n_xy == 23;
  if (j.fast == TABCORE_32BIAS) return j.sp.n_xy == 31;
  return false;
}

// Everything of a table job except its place in the launch (work_begin) and the Taylor ROM offset.
void init_tab_job(const SrcParams& canon, int32_t* tab, TabJob* j) {
  memset(j, 0, sizeof(*j));
  j->sp = canon;
  j->tab = tab;
  j->entries = 1u << canon.pw;
  j->fast = (uint32_t)table_core32(canon);
  j->tshift = (uint32_t)table_tshift(canon);
  j->work = canon.kind == SRC_INQ ? j->entries : j->entries / 4;
  if (canon.kind == SRC_TAYLOR) return;
  for (int i = 0; i < canon.n_z && i < 48; i++) {
    const int64_t r = (c_atan[canon.rom_sel][i] >> canon.rom_shift) & canon.rom_mask;
    if (i < 32 && (j->fast == TABCORE_32 || j->fast == TABCORE_32BIAS)) j->rom32[i] = (int32_t)r;
    if (j->fast == TABCORE_A64) {
      const uint64_t a = (uint64_t)r << (64 - canon.zw);
      j->rom64[i] = (int64_t)a;
      if (i < 32) j->rom32[i] = (int32_t)(uint32_t)((a + ((a & 0x80000000ull) << 1)) >> 32);
    }
  }
}

// quarter-wave ROM of taylor_sincos, computed as the VHDL does with math_real
// (src/taylor_sincos.vhd:91-111): INTEGER((2^(DW-1)-1) * cos(ii*pi/(2*depth))), round to nearest.
void build_taylor_rom(int dw, int lut, std::vector<I2>& rom) {
  const int depth = 1 << lut;
  rom.resize(depth);
  const double amp = ldexp(1.0, dw - 1) - 1.0;
  for (int ii = 0; ii < depth; ii++) {
    const double a = ((double)ii * M_PI) / (2.0 * (double)depth);
    rom[ii].x = (int32_t)llround(amp * cos(a));
    rom[ii].y = (int32_t)llround(amp * sin(a));
  }
}

// Left shift applied to the entries of a source's trig table (WinRec comment): it lets the
// synthesis kernels take the wanted product bits with a multiply-high.  1 for the RTL tail
// (2 for the HLS tail) whenever the shifted cosine still fits 32 bits, else 0.
int table_tshift(const SrcParams& sp) {
  switch (sp.kind) {
    case SRC_HLS: return sp.dw <= 30 ? 2 : 0;      // |cos| <= 2^(NW-2)+eps
    case SRC_TAYLOR: return sp.dw <= 31 ? 1 : 0;   // |cos| <= 2^(DW-1)-1
    default: return sp.dw <= 31 ? 1 : 0;           // CORDIC entities: |cos| <= 2^(DW-2)+eps (overshoots 2^(DW-2))
  }
}

// DAT_WIDTH 31..32: dsp_pp has DW+2 (DW+1) bits, more than a 32-bit register - unless the ports
// keep it small.  |b_k| <= (|AAk|*cmax + 2^(DW-2)) >> (DW-1), cmax = largest |cos| the source can
// emit, so when |AA0| + sum_k |b_k| + rounding stays below 2^31 the true sum fits an int32: the
// entity's wraps of dsp_pp (DW+2 bits) and DT_WIN (DW bits) are then no-ops and the 32-bit tail
// applies with lsh = 0 (intermediate wraps of the 32-bit accumulator are harmless, the arithmetic is
// modular).  Every real coefficient set qualifies (a0 + (a1+..)/2 < 1 at the CORDIC amplitude
// 2^(DW-2)); full-scale synthetic ports fall back to the 64-bit tail.  With DW = 32 the table holds
// the unshifted cosine (tshift 0) and the coefficient carries the factor 2: A_k = 2*AAk must fit.
static bool narrow_sum_ok(const WinParams& wp, const SrcParams* src) {
  const int dw = wp.dw;
  if (dw < 31 || dw > 32) return false;
  const int t = table_tshift(src[0]);
  for (int u = 1; u < wp.nsrc; u++) if (table_tshift(src[u]) != t) return false;
  if (t != 33 - dw && t != 32 - dw) return false;           // A_k = AAk << (33 - DW - t), shift 0 or 1
  const int ash = 33 - dw - t;
  int64_t bound = (wp.aa[0] < 0 ? -wp.aa[0] : wp.aa[0]) + 4;
  for (int k = 1; k < wp.m; k++) {
    const int64_t a = wp.aa[k] < 0 ? -wp.aa[k] : wp.aa[k];
    if (ash && a >= ((int64_t)1 << 30)) return false;       // 2*AAk and its negative must fit an int32
    const SrcParams& sp = src[wp.term[k - 1].src];
    const int64_t cmax = sp.kind == SRC_TAYLOR ? ((int64_t)1 << (dw - 1)) : ((int64_t)1 << (dw - 2)) + 64;
    bound += ((a * cmax) >> (dw - 1)) + 2;
  }
  return bound < ((int64_t)1 << 31);
}

// Which synthesis tail reproduces the entity for this window.
TailMode fast_tail_mode(const WinParams& wp, const SrcParams* src) {
  if (wp.dw > 32) return TAILMODE_GENERIC;
  const int t = table_tshift(src[0]);
  // AAk = -2^(DW-1) (k >= 1) is left to the generic body: the RTL tail computes b_k there with the
  // entity's own DW-bit wrap, and for both models the pre-shifted coefficient is then INT32_MIN,
  // which the half-period table placement cannot negate (it folds the table's sign into -A_k)
  const int64_t lo = -((int64_t)1 << (wp.dw - 1));
  for (int k = 1; k < wp.m; k++) if (wp.aa[k] == lo) return TAILMODE_GENERIC;
  if (wp.tail == TAIL_HLS) return t == 2 ? TAILMODE_FAST32 : TAILMODE_GENERIC;
  const int dmax = wp.tail == TAIL_RTL2 ? 31 : 30;  // dsp_pp (DW+1 / DW+2 bits) must fit 32 bits
  if (t == 1 && wp.dw <= dmax) return TAILMODE_FAST32;
  return narrow_sum_ok(wp, src) ? TAILMODE_FAST32 : TAILMODE_ACC64;
}

// Fill the tail record of a window (see the derivation above WinRec).
void fill_fast_rec(const WinParams& wp, const SrcParams* src, WinRec& r) {
  const int dw = wp.dw, m = wp.m;
  const TailMode mode = fast_tail_mode(wp, src);
  r.m = (uint32_t)m; r.dw = (uint32_t)dw; r.pw = (uint32_t)wp.pw;
  r.flags = 0;
  r.tshift = (uint32_t)table_tshift(src[0]);
  for (int k = 0; k < m; k++) r.aa[k] = (int32_t)wp.aa[k];
  if (wp.tail == TAIL_HLS) r.flags |= WR_HLS;
  if (wp.tail == TAIL_RTL2) r.flags |= WR_RTL2;
  if (mode == TAILMODE_ACC64) {   // raw coefficients; the kernels apply the 64-bit tail
    r.flags |= WR_ACC64;
    for (int k = 1; k < m; k++) r.A[k] = r.aa[k];
    r.S0 = r.aa[0];
    return;
  }
  if (wp.tail == TAIL_HLS) {
    const int ashift = 32 - dw;
    for (int k = 1; k < m; k++) r.A[k] = (int32_t)((uint32_t)r.aa[k] << ashift);
    r.rc = 0; r.S0 = r.aa[0]; r.lsh = 32 - dw; r.rsh = 32 - dw;
    return;
  }
  // b_k = hi32(A_k * (cos << tshift) + 2^31) needs A_k * 2^tshift == AAk * 2^(33-DW)
  const int ashift = 33 - dw - (int)r.tshift;
  for (int k = 1; k < m; k++) r.A[k] = (int32_t)((uint32_t)r.aa[k] << ashift);
  const int fin = wp.tail == TAIL_RTL2 ? 1 : 2;             // dsp_pp carries DW+fin bits
  r.rc = 0x80000000u;
  r.S0 = (int32_t)((uint32_t)r.aa[0] + (uint32_t)fin);
  if (dw + fin <= 32) { r.lsh = 32 - fin - dw; r.rsh = 32 - dw; }
  else { r.lsh = 0; r.rsh = fin; }                          // narrow_sum_ok(): the sum fits as it is
}

bool direct32_params(const WinParams& wp, const SrcParams* src, Direct32Params* out) {
  if (wp.nsrc != 1 || !fast32_ok(src[0])) return false;
  if (fast_tail_mode(wp, src) != TAILMODE_FAST32) return false;
  const SrcParams& sp = src[0];
  if (sp.n_xy > 32 || sp.pw < 3) return false;
  WinRec r;
  memset(&r, 0, sizeof(r));
  fill_fast_rec(wp, src, r);
  Direct32Params& p = *out;
  memset(&p, 0, sizeof(p));
  p.m = wp.m; p.pw = sp.pw;
  p.n_xy = sp.n_xy; p.n_z = sp.n_z;
  p.z_rshift = sp.z_rshift; p.z_lshift = sp.z_lshift;
  p.out_shift = sp.out_shift;
  p.tshift = (int32_t)r.tshift;
  p.gain = (int32_t)sp.gain;
  p.S0 = r.S0; p.lsh = r.lsh; p.rsh = r.rsh; p.rc = r.rc;
  p.n_first = (uint32_t)wp.stream_offset;
  for (int k = 1; k < wp.m; k++) { p.A[k] = r.A[k]; p.kmul[k] = wp.term[k - 1].kmul; }
  for (int i = 0; i < sp.n_z && i < 32; i++)   // entries past n_z stay 0
    p.rom[i] = (int32_t)((c_atan[sp.rom_sel][i] >> sp.rom_shift) & sp.rom_mask);
  return true;
}

bool direct_taylor_params(const WinParams& wp, const SrcParams* src, DirectTayParams* out) {
  if (wp.m > 3 || wp.nsrc < 1 || wp.nsrc > 2 || wp.nsrc != wp.m - 1) return false;
  for (int u = 0; u < wp.nsrc; u++) if (src[u].kind != SRC_TAYLOR || src[u].pw < 3 || src[u].dw > 32) return false;
  if (fast_tail_mode(wp, src) != TAILMODE_FAST32) return false;
  WinRec r;
  memset(&r, 0, sizeof(r));
  fill_fast_rec(wp, src, r);
  DirectTayParams& p = *out;
  memset(&p, 0, sizeof(p));
  p.m = wp.m; p.dw = wp.dw;
  p.tshift = (int32_t)r.tshift;
  p.S0 = r.S0; p.lsh = r.lsh; p.rsh = r.rsh; p.rc = r.rc;
  p.n_first = (uint32_t)wp.stream_offset;
  for (int k = 1; k < wp.m; k++) {
    const TermParams& t = wp.term[k - 1];
    if (t.kmul != 1 || t.src != k - 1) return false;
    const SrcParams& sp = src[t.src];
    p.A[k] = r.A[k];
    TayUnit& u = p.unit[k - 1];
    u.pw = sp.pw; u.mode = sp.tay_mode; u.ashift = sp.tay_ashift; u.cbits = sp.tay_cbits; u.xs = sp.tay_xs;
    u.pi = (int32_t)sp.tay_pi;
  }
  if (src[0].lut > 12) return false;  // the kernel keeps the ROM in shared memory (32 KB at LUT_SIZE 12)
  p.rom_entries = 1u << src[0].lut;
  // both units share one datapath class (see TMODE_*)
  int tm = -1;
  for (int k = 1; k < wp.m; k++) {
    const int md = p.unit[k - 1].mode;
    const int cls = (md == TAY_LESS || md == TAY_EQ) ? TMODE_ROM : md == TAY_DSP ? TMODE_DSP : TMODE_WIDE;
    if (tm >= 0 && tm != cls) return false;
    tm = cls;
  }
  p.tmode = (uint32_t)tm;
  return true;
}

// direct_taylor_quad4 applies: TAY_WIDE, stream offset 0, the window's units are (PHI_WIDTH[, PHI_WIDTH-1]) and
// four aligned consecutive samples share every unit's ROM word (>= 2 counter bits below the ROM address).
bool direct_taylor_quad_ok(const DirectTayParams& p, int phi_width) {
  if (p.tmode != TMODE_WIDE || p.n_first != 0 || phi_width < 12) return false;
  if (p.unit[0].pw != phi_width || (p.m > 2 && p.unit[1].pw != phi_width - 1)) return false;
  for (int k = 1; k < p.m; k++) {
    const TayUnit& u = p.unit[k - 1];
    if (u.mode != TAY_WIDE || u.cbits < 2 || u.ashift != u.cbits) return false;
  }
  return true;
}

// cordic_atan2 generics -> kernel parameters.  ROM_TABLE(ii) = "0" & ROM_LUT(ii)(47 downto
// 47-(W-2)): the top W-1 bits of the 48-bit word (src/cordic_atan2.vhd:97-108).
int resolve_atan2(const bhw_atan2_desc* d, Atan2Params* p) {
  if (!d) return BHW_E_NULL;
  const int prec = d->precision == 0 ? 1 : d->precision;
  if (d->stream_quadrant != 0 && d->stream_quadrant != 1) return BHW_E_ARG;
  if (d->angle_width < 4 || d->angle_width > 32) return BHW_E_DAT_WIDTH;
  if (d->input_width > 32 || d->input_width < d->angle_width - 1) return BHW_E_PHI_WIDTH;
  if (prec < 1 || prec > 7) return BHW_E_PRECISION;
  if (!p) return BHW_OK;
  memset(p, 0, sizeof(*p));
  p->iw = d->input_width; p->aw = d->angle_width; p->w = d->angle_width + prec;
  for (int i = 0; i <= p->aw - 2; i++) {
    const int64_t r = c_atan[0][i] >> (48 - (p->w - 1));
    p->rom64[i] = (int64_t)((uint64_t)r << (64 - p->w));
    if (p->w <= 32 && i < 32) p->rom32[i] = (uint32_t)((uint64_t)r << (32 - p->w));
  }
  p->fast32 = p->w <= 32 ? 1 : 0;
  p->skew = d->stream_quadrant;
  return BHW_OK;
}

// Which harmonics see the flipped quadrant half a window later (bit k), or false when some harmonic
// does neither "same phase" nor "half a source period further" (then there is no pairing).
bool direct_pair_flip(const WinParams& wp, const SrcParams* src, uint32_t* flip) {
  if (wp.pw < 3) return false;
  const uint64_t half = 1ull << (wp.pw - 1);
  uint32_t f = 0;
  for (int k = 1; k < wp.m; ++k) {
    const TermParams& t = wp.term[k - 1];
    const uint64_t delta = ((uint64_t)t.kmul * half) & t.ph_mask;
    if (delta == 0) continue;
    if (delta != ((uint64_t)t.ph_mask + 1) >> 1 || src[t.src].kind == SRC_INQ) return false;
    f |= 1u << k;
  }
  *flip = f;
  return true;
}

// Quadrants harmonic k's source advances per quarter window (2 bits at position 2k), or false when
// some harmonic does not advance by whole quadrants or its source has no output quadrant mux.
bool direct_quad_adv(const WinParams& wp, const SrcParams* src, uint32_t* adv) {
  if (wp.pw < 5) return false;                       // N >= 32
  const uint64_t quarter = 1ull << (wp.pw - 2);
  uint32_t a = 0;
  for (int k = 1; k < wp.m; ++k) {
    const TermParams& t = wp.term[k - 1];
    if (src[t.src].kind == SRC_INQ) return false;
    const uint64_t period = (uint64_t)t.ph_mask + 1;
    if (period < 4) return false;
    const uint64_t delta = ((uint64_t)t.kmul * quarter) & t.ph_mask;
    if (delta % (period / 4)) return false;
    a |= (uint32_t)(delta / (period / 4)) << (2 * k);
  }
  *adv = a;
  return true;
}

bool source_antisymmetric(const SrcParams& sp) {
  switch (sp.kind) {
    case SRC_DDS:   // |value| <= 2^(DW-2) + a few LSB: far from -2^(DW-1) once DW >= 8; the quadrant
    case SRC_HLS:   // fix negates with a wrap that is then exact (src/cordic_dds.vhd:232-246)
      return sp.dw >= 8;
    case SRC_TAYLOR:
      // ROM-only branches: entries in [0, 2^(DW-1)-1].  DW > 18: tay1_order saturates negatives
      // to 2^(DW-1)-1 (src/tay1_order.vhd:602-617), so again [0, 2^(DW-1)-1].  The DSP48 branch
      // (DW < 19) wraps instead: sin + delta*cos can reach 2^(DW-1) = -2^(DW-1) - not provable.
      return sp.tay_mode != TAY_DSP;
    default:        // cordic_dds48 / cordic_dds_scaled fold the quadrant in at the input
      return false;
  }
}

bool source_inq_complement(const SrcParams& sp) { return sp.kind == SRC_INQ && sp.dw >= 8; }

bool bank_shape(const WinRec& r, const BankTableInfo* tk, size_t smem_limit_bytes, BankShape* sh,
                int* tab_mode, bool* pair, bool allow_pair) {
  if (r.flags & WR_GENERIC) return false;
  if (!kernel_terms((int)r.m)) return false;
  if (r.pw < (uint32_t)kBankTileLog2) return false;  // a window must hold at least one tile
  memset(sh, 0, sizeof(*sh));
  sh->m = r.m; sh->pw = r.pw;
  sh->rc = r.rc; sh->rcn = 0xFFFFFFFFu - r.rc;
  sh->lsh = r.lsh; sh->rsh = r.rsh;
  if (r.flags & WR_ACC64) {           // RTL tail in 64 bits (DAT_WIDTH 31..32)
    sh->acc64 = 1;
    sh->psh = r.dw - 1 + r.tshift;
    if (sh->psh > 31) return false;
    sh->prnd = 1u << (r.dw - 2 + r.tshift);
    sh->prndn = ((sh->psh == 32 ? 0u : (1u << sh->psh)) - 1u) - sh->prnd;
    sh->fin = (r.flags & WR_RTL2) ? 1 : 2;
    sh->fadd = sh->fin;               // (pp>>2)+((pp>>1)&1) == (pp+2)>>2 ; (pp>>1)+(pp&1) == (pp+1)>>1
    sh->flsh = 32 - r.dw;
  }
  bool antisym = true, half_ok = true, inqc = true;
  for (uint32_t k = 1; k < r.m; k++) {
    if (!tk[k].ptr) return false;
    uint32_t u = 0;
    for (; u < sh->ntab; u++) if (sh->tab[u] == tk[k].ptr) break;
    if (u == sh->ntab) {
      if (sh->ntab == 2) return false;
      sh->tab[u] = tk[k].ptr;
      sh->tentries[u] = tk[k].entries;
      sh->ntab++;
    }
    sh->kstep[k] = r.kstep[k];
    sh->idx_rsh[k] = r.idx_rsh[k];
    sh->tsel[k] = u;
    if (!tk[k].antisym) antisym = false;
    if (!tk[k].inq_comp) inqc = false;
    if ((uint64_t)r.kstep[k] * (uint64_t)(kBankTile - 1) >= (1ull << 31)) half_ok = false;
  }
  size_t words = 0;
  for (uint32_t u = 0; u < sh->ntab; u++) words += sh->tentries[u];
  if (sh->ntab == 1) { sh->tab[1] = sh->tab[0]; sh->tentries[1] = sh->tentries[0]; }
  const size_t limit = smem_limit_bytes / sizeof(int32_t);
  // exact antisymmetry, or the ones'-complement relation of the input-quadrant CORDICs (32-bit tail only; the
  // exceptions are patched afterwards, see BankShape::pair_adj)
  const bool comp_pair = !antisym && inqc && !(r.flags & WR_ACC64);
  const bool can_pair = allow_pair && (antisym || comp_pair) && r.pw >= (uint32_t)kBankTileLog2 + 1;
  if (can_pair && comp_pair) { sh->pair_adj = 1u << r.tshift; half_ok = false; }
  if (words <= limit) { *tab_mode = TAB_SMEM_FULL; *pair = can_pair; }
  else if (can_pair && half_ok && words / 2 <= limit) { *tab_mode = TAB_SMEM_HALF; *pair = true; }
  else { *tab_mode = TAB_GLOBAL; *pair = can_pair; }
  const int sh_half = *tab_mode == TAB_SMEM_HALF ? 1 : 0;
  uint32_t off = 0;
  for (uint32_t u = 0; u < sh->ntab; u++) { sh->toff[u] = off; off += sh->tentries[u] >> sh_half; }
  if (sh->ntab == 1) sh->toff[1] = sh->toff[0];
  sh->smem_words = *tab_mode == TAB_GLOBAL ? 0 : off;
  // linear indexing: the harmonic's phase step is a whole number of table entries
  sh->lin = 1;
  for (uint32_t k = 1; k < r.m; k++) {
    const uint32_t entries = sh->tentries[sh->tsel[k]];
    const uint32_t per_entry = 1u << r.idx_rsh[k];           // phase32 units per table entry (idx_rsh >= 2)
    if (r.kstep[k] % per_entry) { sh->lin = 0; break; }
    sh->lin_step[k] = r.kstep[k] / per_entry;
    sh->lin_dmask[k] = (entries >> sh_half) - 1;
    sh->lin_dbit[k] = sh_half ? entries >> 1 : 0;
    if ((uint64_t)sh->lin_step[k] * kBankTile > (entries >> sh_half)) { sh->lin = 0; break; }
  }
  return true;
}


// ---- families and groups --------------------------------------------------------------------------
// term counts the group and bank kernels are instantiated for (the reference's entities); 6 and 8..11 terms
// (BHW_WIN_MTERM_*) go through the general and the direct kernels
bool kernel_terms(int m) { return m == 2 || m == 3 || m == 4 || m == 5 || m == 7; }

bool group_eligible(const bhw_desc& d, const WinParams& wp, const SrcParams* src) {
  if (!kernel_terms(wp.m)) return false;
  if (d.algo == BHW_ALGO_DIRECT || wp.nsrc != 1 || wp.pw < kBankTileLog2 + 1 || wp.elem64) return false;
  const SrcParams& sp = src[0];
  if (sp.kind != SRC_DDS && sp.kind != SRC_HLS) return false;
  if (!source_antisymmetric(sp) || sp.pw != wp.pw) return false;
  if (fast_tail_mode(wp, src) != TAILMODE_FAST32) return false;
  for (int k = 1; k < wp.m; k++) {
    const TermParams& t = wp.term[k - 1];
    if (t.kmul != (uint32_t)k || t.src != 0 || t.ph_mask != (uint32_t)((1ull << wp.pw) - 1)) return false;
  }
  return true;
}

int family_source(const bhw_desc& d, int pw, SrcParams* canon) {
  bhw_desc d2 = d;
  d2.phi_width = pw;
  SrcParams sp;
  int st = resolve_source(&d2, 0, &sp);
  if (st) return st;
  uint32_t drop;
  *canon = canonical_source(sp, &drop);
  return BHW_OK;
}

int group_tab_mode(const SrcParams& canon, uint32_t top, size_t smem_limit_bytes) {
  if (((size_t)4 << (top - 1)) <= smem_limit_bytes) return G_HALF32;
  // uint16 quarter waves: |cos|, |sin| <= 2^(DW-2) + (a few LSB of CORDIC error) must fit 16 bits with the
  // bias on both sides: DAT_WIDTH <= 17 (amplitude 2^15; the error is below DAT_WIDTH LSB, the bias is 1024)
  if (canon.dw <= 17 && ((size_t)4 << (top - 2)) <= smem_limit_bytes) return G_Q16;
  return G_GLOBAL;
}

void group_shape(const WinRec& r, uint32_t top, uint32_t lmin, GroupShape* sh) {
  memset(sh, 0, sizeof(*sh));
  sh->m = r.m;
  sh->top = top;
  sh->lmin = lmin;
  sh->rc = r.rc;
  sh->rcn = 0xFFFFFFFFu - r.rc;
  sh->lsh = r.lsh;
  sh->rsh = r.rsh;
  // light kernels are paced by the store path: the whole GPU sweeps the output front to back; from 4
  // terms up a contiguous share per CTA measured faster (k_synth_bank)
  sh->interleave = r.m <= 3 ? 1u : 0u;
  sh->tmul = 1u << r.tshift;
  sh->tbias = kQ16Bias << r.tshift;
}

void init_pyramid_job(const SrcParams& canon, uint32_t lmin, int32_t* pyr, uint16_t* q16, TabJob* j) {
  init_tab_job(canon, pyr, j);
  j->pyr_lmin = lmin;
  j->q16 = q16;
}

}  // namespace bhw
