// bhw_resolve.cpp - host-side half of the ABI that needs no GPU: descriptor validation, resolution
// of the entity generics into kernel parameter blocks, and the coefficient front end.
#include <math.h>
#include <string.h>

#include <vector>

#include "bhw_internal.h"
#include "bhw_plan.h"

namespace bhw {

static inline int eff_prec(const bhw_desc* d) { return d->precision == 0 ? 1 : d->precision; }
static inline int eff_lut(const bhw_desc* d) { return d->lut_size == 0 ? 9 : d->lut_size; }

// internal width table of cordic_dds_scaled, DATA_WIDTH 8..32 (src/cordic_dds_scaled.vhd:102-107)
static const int kScaledSize[25] = {15, 15, 15, 18, 21, 22, 23, 26, 30, 31, 32, 33, 38,
                                    38, 38, 42, 42, 45, 47, 47, 47, 48, 48, 48, 48};

static const int64_t kGainDds = 0x4DBA76D421AFll;  // src/cordic_dds.vhd:97
static const int64_t kGainInq = 0x26DD3B6A10D8ll;  // src/cordic_dds48.vhd:110, win_function.cpp:83

static int check_taylor_unit(int pw, int dw, int lut) {
  const int d = pw - lut;
  if (pw < 3) return BHW_E_PHI_WIDTH;
  if (d > 2) {
    if (d - 3 > 15) return BHW_E_LUT_SIZE;                      // cnt_exp is 16 bits (tay1_order.vhd:116-127)
    if (dw < 19 && 19 + lut + dw > 48) return BHW_E_LUT_SIZE;   // slice of the 48-bit P (:501-502)
    if (dw > 18 && 19 + lut + dw > 62) return BHW_E_LUT_SIZE;   // slice of the 62-bit product (:585-586)
  }
  return BHW_OK;
}

int validate_desc(const bhw_desc* d, bool for_window) {
  if (!d) return BHW_E_NULL;
  const int pw = d->phi_width, dw = d->dat_width;
  if (for_window) {
    const int m = d->win_type;
    if (m < 2 || m > BHW_MAX_TERMS) return BHW_E_WIN_TYPE;
    // 6 and 8..11 terms (BHW_WIN_MTERM_*) are the RTL structure extended, with a CORDIC source: the HLS model has
    // no such type (win_function.cpp:380-422) and only hamming_win / bh_win_3term take TAYLOR
    const bool entity = m == 2 || m == 3 || m == 4 || m == 5 || m == 7;
    if (!entity && (d->model != BHW_MODEL_RTL || d->sin_type == BHW_SIN_TAYLOR)) return BHW_E_WIN_TYPE;
  }
  if (pw < BHW_MIN_PHI_WIDTH || pw > BHW_MAX_PHI_WIDTH) return BHW_E_PHI_WIDTH;
  switch (d->model) {
    case BHW_MODEL_RTL:
      switch (d->sin_type) {
        case BHW_SIN_CORDIC: {
          const int p = eff_prec(d);
          if (dw < 4 || dw > 48) return BHW_E_DAT_WIDTH;
          if (p < 1 || p > 7 || dw + p > 49) return BHW_E_PRECISION;
          // the window entities never override PRECISION (src/hamming_win.vhd:153-157)
          if (for_window && p != 1) return BHW_E_PRECISION;
          break;
        }
        case BHW_SIN_CORDIC48:
          if (dw < 4 || dw > 48) return BHW_E_DAT_WIDTH;
          break;
        case BHW_SIN_CORDIC_SCALED:
          if (dw < 8 || dw > 32) return BHW_E_DAT_WIDTH;
          break;
        case BHW_SIN_TAYLOR: {
          const int lut = eff_lut(d);
          if (for_window && d->win_type != BHW_WIN_HAMMING && d->win_type != BHW_WIN_BH3TERM)
            return BHW_E_SIN_TYPE;  // 4/5/7-term entities have no SIN_TYPE (src/bh_win_4term.vhd:57-61)
          if (dw < 4 || dw > 32) return BHW_E_DAT_WIDTH;  // ROM built through VHDL INTEGER
          if (lut < 1 || lut > 16) return BHW_E_LUT_SIZE;
          int st = check_taylor_unit(pw, dw, lut);
          if (st) return st;
          if (for_window && d->win_type == BHW_WIN_BH3TERM) {
            // "must set LUT_SIZE < (PHASE_WIDTH - 3)": the two units would sit in different
            // latency branches (src/bh_win_3term.vhd:30-31,221-233)
            if (pw - lut == 3) return BHW_E_LUT_SIZE;
            if ((st = check_taylor_unit(pw - 1, dw, lut))) return st;
          }
          break;
        }
        default:
          return BHW_E_SIN_TYPE;
      }
      break;
    case BHW_MODEL_HLS:
      if (d->sin_type != BHW_SIN_CORDIC) return BHW_E_SIN_TYPE;
      if (dw < 4 || dw > 32) return BHW_E_DAT_WIDTH;
      if (pw > dw + 2) return BHW_E_PHI_WIDTH;  // init_t would be truncated by dat_t (win_function.cpp:88)
      break;
    case BHW_MODEL_CPP:
      if (for_window) return BHW_E_MODEL;  // cpp/ has no window model
      if (d->sin_type != BHW_SIN_CORDIC) return BHW_E_SIN_TYPE;
      if (dw < 4 || dw > 32) return BHW_E_DAT_WIDTH;
      break;
    default:
      return BHW_E_MODEL;
  }
  if (d->algo < BHW_ALGO_AUTO || d->algo > BHW_ALGO_TABLE) return BHW_E_ARG;
  if (!for_window) return BHW_OK;
  if (d->stream_offset != 0 && d->stream_offset != 1) return BHW_E_ARG;
  if (d->out_format != BHW_OUT_DEFAULT && d->out_format != BHW_OUT_INT16) return BHW_E_ARG;
  if (d->out_format == BHW_OUT_INT16 && dw > 16) return BHW_E_DAT_WIDTH;
  for (int k = 0; k < d->win_type; k++) {
    const int64_t v = d->aa[k];
    // an AAk port is DAT_WIDTH raw bits: accept the signed or the unsigned reading of them
    if (v < -((int64_t)1 << (dw - 1)) || v >= ((int64_t)1 << dw)) return BHW_E_COEFF;
    // the HLS a_k are plain non-wrapping integers below 2^(NWIDTH-1)
    if (d->model == BHW_MODEL_HLS && v >= ((int64_t)1 << (dw - 1))) return BHW_E_COEFF;
  }
  return BHW_OK;
}

int resolve_source(const bhw_desc* d, int unit, SrcParams* sp) {
  memset(sp, 0, sizeof(*sp));
  const int pw = d->phi_width - unit, dw = d->dat_width;
  sp->pw = pw;
  sp->dw = dw;
  sp->rom_mask = -1;
  if (d->model == BHW_MODEL_HLS) {
    sp->kind = SRC_HLS;
    sp->w = sp->zw = dw + 2;                 // dat_t
    sp->n_xy = dw;                           // NWIDTH passes
    sp->n_z = dw - 1;                        // lut_angle[NWIDTH-1]
    sp->rom_sel = 0;
    sp->rom_shift = 47 - dw;
    sp->rom_mask = 0xFFFFFFFFFFll;
    sp->gain = kGainInq >> (46 - dw);
    if (pw - 1 < dw) { sp->z_rshift = 0; sp->z_lshift = dw - pw + 2; }
    else { sp->z_rshift = pw - dw; sp->z_lshift = 2; }
    sp->out_shift = 2;
    sp->negw = dw + 2;
    sp->outw = dw;
    return BHW_OK;
  }
  if (d->model == BHW_MODEL_CPP) {
    sp->kind = SRC_CPP;
    sp->w = sp->zw = 64;                     // long long, never wraps
    sp->n_xy = dw;
    sp->n_z = dw - 1;
    sp->rom_sel = 1;
    sp->rom_shift = 47 - dw;
    sp->rom_mask = 0xFFFFFFFFFFFFll;
    sp->gain = kGainInq >> (46 - dw);
    if (pw - 1 < dw) { sp->z_rshift = 0; sp->z_lshift = dw - pw + 1; }
    else { sp->z_rshift = pw - dw; sp->z_lshift = 1; }
    sp->out_shift = 2;
    sp->negw = 0;                            // ones' complement
    sp->outw = 32;                           // int(dat_c)
    return BHW_OK;
  }
  switch (d->sin_type) {
    case BHW_SIN_CORDIC: {
      const int p = eff_prec(d), w = dw + p;
      sp->kind = SRC_DDS;
      sp->w = sp->zw = w;
      sp->n_xy = sp->n_z = dw - 1;
      sp->rom_sel = 0;
      sp->rom_shift = 49 - w;
      sp->gain = kGainDds >> (49 - w);
      if (pw >= dw) { sp->z_rshift = pw - dw; sp->z_lshift = p; }
      else { sp->z_rshift = 0; sp->z_lshift = dw - pw + p; }
      sp->out_shift = p;
      sp->negw = dw;
      sp->outw = dw;
      return BHW_OK;
    }
    case BHW_SIN_CORDIC48:
    case BHW_SIN_CORDIC_SCALED: {
      const int size = d->sin_type == BHW_SIN_CORDIC48 ? 48 : kScaledSize[dw - 8];
      const int dwph = d->sin_type == BHW_SIN_CORDIC48 ? 48 : (size < pw ? pw : size);
      sp->kind = SRC_INQ;
      sp->w = size;
      sp->zw = dwph;
      sp->n_xy = dw;
      sp->n_z = dw - 1;
      sp->rom_sel = 1;
      sp->rom_shift = 48 - dwph;
      sp->gain = kGainInq >> (48 - size);
      sp->z_rshift = 0;
      sp->z_lshift = dwph - pw;
      sp->out_shift = size - dw;
      sp->negw = size;                       // used for the -GAIN start vector
      sp->outw = dw;
      return BHW_OK;
    }
    case BHW_SIN_TAYLOR: {
      const int lut = eff_lut(d), dd = pw - lut;
      sp->kind = SRC_TAYLOR;
      sp->w = sp->zw = dw;
      sp->lut = lut;
      sp->negw = dw;
      sp->outw = dw;
      sp->tay_xs = 19 + lut;
      if (dd < 2) { sp->tay_mode = TAY_LESS; sp->tay_ashift = lut - pw + 2; }
      else if (dd == 2) { sp->tay_mode = TAY_EQ; sp->tay_ashift = 0; }
      else {
        sp->tay_mode = dw < 19 ? TAY_DSP : TAY_WIDE;
        sp->tay_ashift = pw - lut - 2;
        sp->tay_cbits = pw - lut - 2;
        const int stage = pw - lut - 3;
        sp->tay_pi = (int64_t)round(M_PI * ldexp(1.0, 17 - stage));  // tay1_order.vhd:131
      }
      return BHW_OK;
    }
  }
  return BHW_E_SIN_TYPE;
}

TabLookup table_lookup_for(const SrcParams& sp) {
  TabLookup t;
  const uint32_t n = 1u << sp.pw;
  if (sp.kind == SRC_INQ) {  // no output symmetry: one entry per phase
    t.idx_mask = n - 1; t.idx_shift = 0; t.neg_bit = 0; t.entries = n;
    return t;
  }
  // half a period; phases that differ only in bits the source drops share an entry
  int drop = sp.kind == SRC_TAYLOR ? 0 : sp.z_rshift;
  if (sp.kind == SRC_TAYLOR && sp.tay_mode == TAY_LESS) drop = 0;
  t.idx_mask = (n >> 1) - 1;
  t.idx_shift = (uint32_t)drop;
  t.neg_bit = n >> 1;
  t.entries = (n >> 1) >> drop;
  return t;
}

int resolve_window(const bhw_desc* d, WinParams* wp, SrcParams src[2]) {
  int st = validate_desc(d, true);
  if (st) return st;
  memset(wp, 0, sizeof(*wp));
  const int m = d->win_type, dw = d->dat_width, pw = d->phi_width;
  wp->m = m; wp->dw = dw; wp->pw = pw;
  wp->stream_offset = d->stream_offset;
  wp->elem64 = dw > 32;
  wp->tail = d->model == BHW_MODEL_HLS ? TAIL_HLS : (m == 2 ? TAIL_RTL2 : TAIL_RTLM);
  for (int k = 0; k < m; k++) {
    const int sh = 64 - dw;
    wp->aa[k] = (int64_t)((uint64_t)d->aa[k] << sh) >> sh;  // raw bits -> signed
  }
  if ((st = resolve_source(d, 0, &src[0]))) return st;
  wp->nsrc = 1;
  const bool taylor = d->model == BHW_MODEL_RTL && d->sin_type == BHW_SIN_TAYLOR;
  if (taylor && m == 3) {
    if ((st = resolve_source(d, 1, &src[1]))) return st;
    wp->nsrc = 2;
  }
  for (int k = 1; k < m; k++) {
    TermParams& t = wp->term[k - 1];
    if (taylor) {  // every unit owns a +1 counter; harmonic 2 is a PHASE_WIDTH-1 unit (bh_win_3term.vhd:221-233)
      t.kmul = 1; t.src = k - 1; t.ph_mask = (1u << (pw - (k - 1))) - 1;
    } else {       // ph_in_k += k (src/bh_win_7term.vhd:176-197); HLS: cordic(k*i) (win_function.cpp:361-366)
      t.kmul = (uint32_t)k; t.src = 0; t.ph_mask = (1u << pw) - 1;
    }
  }
  return BHW_OK;
}

// ---- coefficient front end -----------------------------------------------------------------
// Real-valued sets as the reference spells them (README.md:30-41 for the list of variants).
static const double kCoef[18][BHW_MAX_TERMS] = {
    {0.5434783, 1.0 - 0.5434783},                        // Hamming       src/tb/tb_windows.vhd:123-124
    {0.5, 0.5},                                          // Hann          src/hamming_win.vhd:14-16
    {0.42, 0.5, 0.08},                                   // Blackman      src/tb/tb_windows.vhd:114-116
    {0.4243801, 0.4973406, 0.0782793},                   // BH 3-term     src/bh_win_3term.vhd:20
    {0.355768, 0.487396, 0.144323, 0.012604},            // Nuttall       src/bh_win_4term.vhd:16-17
    {0.35875, 0.48829, 0.14128, 0.01168},                // BH 4-term     src/tb/tb_windows.vhd:103-106
    {0.3635819, 0.4891775, 0.1365995, 0.0106411},        // Blackman-Nuttall src/bh_win_4term.vhd:18-19
    {1.000, 1.930, 1.290, 0.388, 0.030},                 // Flat-top      src/tb/tb_windows.vhd:90-94
    {0.3232153788877343, 0.4714921439576260, 0.1755341299601972, 0.0284969901061499,
     0.0012613570882927},                                // BH 5-term     src/bh_win_5term.vhd:14-19
    {0.271220360585039, 0.433444612327442, 0.218004122892930, 0.065785343295606, 0.010761867305342,
     0.000770012710581, 0.000013680883060},              // BH 7-term     src/tb/tb_windows.vhd:67-73
    // the alternative sets the reference spells out next to the ones above
    {0.27105140069342, 0.43329793923448, 0.21812299954311, 0.06592544638803, 0.01081174209837,
     0.00077658482522, 0.00001388721735},                // BH 7-term, README.md:45-51 (magnitudes; the entity alternates the signs)
    {0.5383554, 0.4616446},                              // Hamming, second set  src/hamming_win.vhd:21-23
    {0.215578950, 0.416631580, 0.277263158, 0.083578947, 0.006947368},   // Flat-top, normalised  src/bh_win_5term.vhd:28-33
    // 14..18: the 6- and 8..11-term minimum-sidelobe sets the reference only tabulates (doc/blackman-harris coef.jpg,
    // "Table 1. Coefficients of minimum sidelobe windows"); BHW_WIN_MTERM_* - no reference entity
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
static const int kTerms[18] = {2, 2, 3, 3, 4, 4, 4, 5, 5, 7, 7, 2, 5, 6, 8, 9, 10, 11};

static int variant_coeffs(int variant, int rule, double a[BHW_MAX_TERMS], int* nterms) {
  if (variant < 1 || variant > 18) return BHW_E_VARIANT;
  if (rule != BHW_RULE_TB && rule != BHW_RULE_HLS) return BHW_E_VARIANT;
  if (variant > 13 && rule != BHW_RULE_TB) return BHW_E_VARIANT;   // the HLS model has no such window type
  const int m = kTerms[variant - 1];
  for (int k = 0; k < BHW_MAX_TERMS; k++) a[k] = k < m ? kCoef[variant - 1][k] : 0.0;
  if (rule == BHW_RULE_HLS && variant == 3)  // the HLS Blackman uses 0.21/0.25/0.04 (win_function.cpp:206-208)
    for (int k = 0; k < m; k++) a[k] *= 0.5;
  *nterms = m;
  return BHW_OK;
}

// element size of a batch: one container for all its windows (BHW_E_ELEM otherwise)
int batch_elem_bytes(const bhw_desc* descs, int nwin, size_t* esz) {
  bool any64 = false, any32 = false, any16 = false;
  for (int i = 0; i < nwin; i++) {
    if (descs[i].out_format == BHW_OUT_INT16) any16 = true;
    else (descs[i].dat_width > 32 ? any64 : any32) = true;
  }
  if ((int)any64 + (int)any32 + (int)any16 > 1) return BHW_E_ELEM;
  *esz = any16 ? 2 : any64 ? 8 : 4;
  return BHW_OK;
}

}  // namespace bhw

using namespace bhw;

extern "C" {

int bhw_version(void) { return BHW_VERSION; }

const char* bhw_strerror(int s) {
  switch (s) {
    case BHW_OK: return "ok";
    case BHW_E_NULL: return "null pointer";
    case BHW_E_WIN_TYPE: return "win_type must be 2, 3, 4, 5 or 7";
    case BHW_E_SIN_TYPE: return "sin_type unknown or not available for this entity/model";
    case BHW_E_MODEL: return "model unknown or not applicable to this call";
    case BHW_E_PHI_WIDTH: return "phi_width out of range";
    case BHW_E_DAT_WIDTH: return "dat_width out of range for this sin/cos source";
    case BHW_E_PRECISION: return "cordic_dds PRECISION out of range";
    case BHW_E_LUT_SIZE: return "TAYLOR LUT_SIZE invalid for this PHI_WIDTH/DAT_WIDTH";
    case BHW_E_COEFF: return "coefficient does not fit dat_width bits";
    case BHW_E_RANGE: return "sample range outside the window or batch";
    case BHW_E_ELEM: return "output element size mismatch";
    case BHW_E_CUDA: return "CUDA runtime error";
    case BHW_E_NO_DEVICE: return "no usable CUDA device";
    case BHW_E_ALLOC: return "allocation failed";
    case BHW_E_VARIANT: return "unknown window variant or quantisation rule";
    case BHW_E_ARG: return "invalid argument";
    case BHW_E_CAPTURE: return "one-shot batch call on a capturing stream (capture bhw_plan_execute instead)";
    default: return "unknown status";
  }
}

int bhw_validate(const bhw_desc* d) { return validate_desc(d, true); }

int bhw_elem_bytes(const bhw_desc* d) {
  if (d && d->out_format == BHW_OUT_INT16) return 2;
  return d && d->dat_width > 32 ? 8 : 4;
}


int bhw_variant_coeffs(int variant, int rule, double a_out[BHW_MAX_TERMS], int32_t* nterms) {
  if (!a_out) return BHW_E_NULL;
  int m = 0;
  int st = variant_coeffs(variant, rule, a_out, &m);
  if (st) return st;
  if (nterms) *nterms = m;
  return BHW_OK;
}

int bhw_quantize(int variant, int rule, int dat_width, int64_t aa_out[BHW_MAX_TERMS], int32_t* win_type) {
  if (!aa_out) return BHW_E_NULL;
  double a[BHW_MAX_TERMS];
  int m = 0;
  int st = variant_coeffs(variant, rule, a, &m);
  if (st) return st;
  if (dat_width < 4 || dat_width > 48) return BHW_E_DAT_WIDTH;
  double scale;
  if (rule == BHW_RULE_TB) {
    // integer(a * S) with S per entity: src/tb/tb_windows.vhd:75-81 (7-term), :96-100 (5-term),
    // :108-111 (4-term), :118-120 (3-term), :126-127 (2-term)
    if (m == 3) scale = ldexp(1.0, dat_width) - 16.0;
    else if (m == 4) scale = ldexp(1.0, dat_width) - 1.0;
    else if (m == 5) scale = ldexp(1.0, dat_width - 2) - 1.0;
    else scale = ldexp(1.0, dat_width - 1) - 1.0;   // 2 and 7 terms; 6 and 8..11 terms follow the 7-term entity
  } else {
    // round(a * (2^(NW-1)-1)) for types 1-4, 2^(NW-2)-1 for 5 and 7 (win_function.cpp:176-355)
    scale = m >= 5 ? ldexp(1.0, dat_width - 2) - 1.0 : ldexp(1.0, dat_width - 1) - 1.0;
  }
  for (int k = 0; k < BHW_MAX_TERMS; k++) aa_out[k] = k < m ? (int64_t)round(a[k] * scale) : 0;
  if (win_type) *win_type = m;
  return BHW_OK;
}

int bhw_batch_total(const bhw_desc* descs, int nwin, uint64_t* total) {
  if (!descs || !total) return BHW_E_NULL;
  if (nwin < 0) return BHW_E_ARG;
  uint64_t t = 0;
  for (int i = 0; i < nwin; i++) {
    if (descs[i].phi_width < BHW_MIN_PHI_WIDTH || descs[i].phi_width > BHW_MAX_PHI_WIDTH) return BHW_E_PHI_WIDTH;
    t += 1ull << descs[i].phi_width;
  }
  *total = t;
  return BHW_OK;
}

// Contiguous, balanced shards whose boundaries are multiples of 4 samples (one 128-bit store)
// whenever the total allows it.
int bhw_shard_range(uint64_t total, int rank, int nranks, uint64_t* begin, uint64_t* count) {
  if (!begin || !count) return BHW_E_NULL;
  if (nranks < 1 || rank < 0 || rank >= nranks) return BHW_E_ARG;
  const uint64_t quads = total / 4, rem = total % 4;
  const uint64_t q0 = quads * (uint64_t)rank / (uint64_t)nranks;
  const uint64_t q1 = quads * (uint64_t)(rank + 1) / (uint64_t)nranks;
  *begin = q0 * 4;
  *count = (q1 - q0) * 4 + (rank == nranks - 1 ? rem : 0);
  return BHW_OK;
}

// ---- cost model of bhw_shard_range_cost -----------------------------------------------------------
// Estimated device time (us) of generating a flat slice of a batch with every trig table rebuilt, fitted to
// round-2 measurements on B200 (bench.py roofline.by_instantiation, tools/r2_probe.py):
//   * per sample, by how the window is generated: staged table (k_synth_group G_HALF32 0.70 ps, G_Q16 0.89 ps),
//     pyramid gathers through L1/L2 (5 terms 1.2 ps, 7 terms 2.6 ps at 2^26 points, 10 % more per halving of
//     the length), windows outside the groups by the round-1 fit; a window cut by the slice loses sample pairing;
//   * per family and slice, the pyramid build: one CORDIC evaluation per quarter-wave phase of the top level
//     the slice needs, ~0.4 ps per stage - 190 us for the 2^26-point DAT_WIDTH 32 family.  It is paid by
//     EVERY rank that touches a window of that length, which is what bounds the strong scaling of the sweep.
struct SliceCost {
  // per-family state of the slice being grown
  struct Fam { SrcParams key; int top; int stages; bool used; };
  Fam fam[16];
  int nfam = 0;
  double us = 0.0;
  void reset() { nfam = 0; us = 0.0; }
};

static double build_us(int top, int stages) { return 4.0 + ldexp(1.0, top - 2) * (double)stages * 0.4e-6; }

// us per sample of window d when a slice holds `part` of its samples (1.0 = the whole window); *is_group: the
// window belongs to a family (then *key / *res / *stages describe it)
static double sample_cost_ps(const bhw_desc& d, bool whole, bool* is_group, SrcParams* key, int* res, int* stages) {
  *is_group = false;
  const int m = d.win_type, pw = d.phi_width, dw = d.dat_width;
  if (dw > 32) return 60.0 * (m - 1) * 0.7;                  // one-thread-per-sample int64 kernel
  WinParams wp; SrcParams src[2];
  if (resolve_window(&d, &wp, src) != BHW_OK) return 1.0;
  if (group_eligible(d, wp, src) && family_source(d, BHW_MAX_PHI_WIDTH, key) == BHW_OK) {
    *is_group = true;
    *res = key->pw;
    *stages = key->n_xy;
    const int top = pw < *res ? pw : *res;                   // a lower bound of the family's top level
    const int mode = group_tab_mode(*key, (uint32_t)(*res < 26 ? *res : 26), 192 * 1024);
    double ps;
    if (mode == G_HALF32) ps = 0.70;
    else if (mode == G_Q16) ps = 0.62 + 0.09 * (m - 1);
    else {
      const double base = m <= 3 ? 0.9 : m == 4 ? 1.0 : m == 5 ? 1.22 : 2.56;
      ps = base * (1.0 + 0.1 * (double)(26 - (pw > 26 ? 26 : pw)));
      (void)top;
    }
    if (!whole) ps *= mode == G_GLOBAL ? 2.7 : 1.6;          // unpaired tiles, no spread walk (measured: 7-term 2^26 half 6.8 ps)
    return ps;
  }
  // windows outside the groups (TAYLOR, input-quadrant CORDICs, 64-bit tails, short windows): round-1 fit
  double a = m <= 3 ? 1.0 : m == 4 ? 1.05 : m == 5 ? 1.2 : 1.75;
  if (d.model == BHW_MODEL_RTL && dw >= 31 && fast_tail_mode(wp, src) != TAILMODE_FAST32) a *= 1.55;
  const bool taylor = d.model == BHW_MODEL_RTL && d.sin_type == BHW_SIN_TAYLOR;
  const bool inq = d.model == BHW_MODEL_RTL && (d.sin_type == BHW_SIN_CORDIC48 || d.sin_type == BHW_SIN_CORDIC_SCALED);
  const bool pair = !inq && !(taylor && dw < 19);
  if (!pair) a *= 1.7;
  if (pw < 9) return a * 4.0 * 0.6;                          // general kernel
  const int idx_bits = (inq || taylor) ? pw : (pw < dw ? pw : dw);
  const double table_bytes = 4.0 * (double)(1ull << idx_bits);
  if (table_bytes / (pair ? 2.0 : 1.0) > 192.0 * 1024.0) {
    const double per = table_bytes > 100e6 ? 0.42 : 0.2;
    const double g = (double)(m * (m - 1) / 2) * per * (pair ? 1.0 : 2.0);
    if (g > a) a = g;
  }
  return a * 0.6;
}

// add samples [lo, hi) of window d (N samples) to the slice
static void slice_add(SliceCost& sc, const bhw_desc& d, uint64_t lo, uint64_t hi) {
  const uint64_t N = 1ull << d.phi_width;
  bool grp; SrcParams key; int res = 0, stages = 0;
  const double ps = sample_cost_ps(d, lo == 0 && hi == N, &grp, &key, &res, &stages);
  sc.us += ps * 1e-6 * (double)(hi - lo);
  if (!grp) { sc.us += d.phi_width < 17 ? 0.05 : 4.0; return; }          // a launch (or a share of one) of its own
  const int top = d.phi_width < res ? d.phi_width : res;
  int f = -1;
  for (int i = 0; i < sc.nfam; i++) if (!memcmp(&sc.fam[i].key, &key, sizeof(key))) f = i;
  if (f < 0 && sc.nfam < 16) {
    f = sc.nfam++;
    sc.fam[f].key = key; sc.fam[f].top = 0; sc.fam[f].stages = stages; sc.fam[f].used = true;
    sc.us += 6.0;                                                        // the family's group launches
  }
  if (f >= 0 && top > sc.fam[f].top) {
    if (sc.fam[f].top) sc.us -= build_us(sc.fam[f].top, stages);
    sc.fam[f].top = top;
    sc.us += build_us(top, stages);
  }
}

// Greedy cut for a target slice time T: ranks take samples in order while their estimated time stays within T.
// Cuts fall on window boundaries, or inside a window of >= 2^20 samples at multiples of 2^14 samples (never
// closer than an eighth of the window to one of its ends).  A window that fits no slice and yields no acceptable
// piece - the 2^26-point 7-term window, whose table build alone exceeds the target at 8 ranks - becomes a slice of
// its own ("atom"), exempt from T: the other ranks still balance among themselves.
// -> the cuts (nranks + 1 flat positions); returns false when nranks slices do not suffice.
static bool cut_for_target(const bhw_desc* descs, int nwin, int nranks, double T, double atom_cap, uint64_t total,
                           uint64_t* cuts) {
  int w = 0;
  uint64_t in_w = 0, flat = 0;       // next sample to hand out: sample in_w of window w
  cuts[0] = 0;
  for (int r = 0; r < nranks; r++) {
    SliceCost sc;
    sc.reset();
    while (w < nwin) {
      const uint64_t N = 1ull << descs[w].phi_width;
      SliceCost trial = sc;
      slice_add(trial, descs[w], in_w, N);
      if (trial.us <= T) {                                   // the rest of this window fits
        sc = trial;
        flat += N - in_w;
        in_w = 0;
        w++;
        continue;
      }
      bool took_piece = false;
      if (N >= (1ull << 20)) {                               // take a piece of it: bisect on the piece length
        const uint64_t step = 1ull << 14;
        uint64_t lo = 0, hi = (N - in_w) / step;             // pieces of `step` samples that fit
        while (lo < hi) {
          const uint64_t mid = (lo + hi + 1) / 2;
          SliceCost t2 = sc;
          slice_add(t2, descs[w], in_w, in_w + mid * step);
          if (t2.us <= T) lo = mid; else hi = mid - 1;
        }
        // a piece shorter than an eighth of the window (or leaving less than that) is not worth the pyramid build
        if (lo * step >= N / 8 && N - in_w - lo * step >= N / 8) {
          slice_add(sc, descs[w], in_w, in_w + lo * step);
          flat += lo * step;
          in_w += lo * step;
          took_piece = true;
        }
      }
      if (!took_piece && flat == cuts[r]) {                  // an empty slice and nothing fits: the window is an atom,
        if (in_w != 0 || trial.us > atom_cap) return false;    // if whole and within the cap - else raise T
        flat += N - in_w;
        in_w = 0;
        w++;
      }
      break;
    }
    cuts[r + 1] = flat;
  }
  return flat == total;
}

int bhw_shard_range_cost(const bhw_desc* descs, int nwin, int rank, int nranks, uint64_t* begin,
                                    uint64_t* count) {
  if (!descs || !begin || !count) return BHW_E_NULL;
  if (nwin <= 0 || nranks < 1 || nranks > 1024 || rank < 0 || rank >= nranks) return BHW_E_ARG;
  uint64_t total = 0;
  SliceCost all;
  all.reset();
  for (int w = 0; w < nwin; w++) {
    if (descs[w].phi_width < BHW_MIN_PHI_WIDTH || descs[w].phi_width > BHW_MAX_PHI_WIDTH) return BHW_E_PHI_WIDTH;
    const uint64_t N = 1ull << descs[w].phi_width;
    slice_add(all, descs[w], 0, N);
    total += N;
  }
  std::vector<uint64_t> cuts((size_t)nranks + 1, 0), best((size_t)nranks + 1, 0);
  // smallest target time for which the greedy cut needs no more than nranks slices
  double lo = all.us / nranks, hi = all.us + 1.0;
  bool have = false;
  for (int it = 0; it < 60; it++) {
    const double T = 0.5 * (lo + hi);
    if (cut_for_target(descs, nwin, nranks, T, 0.0, total, cuts.data())) { hi = T; best = cuts; have = true; }
    else lo = T;
    if (hi - lo < 1e-5 * all.us) break;
  }
  // Second pass: hi is the makespan.  When a single costly window sets it, the greedy above has packed the other
  // ranks up to it and left ranks idle; with windows of up to that cost allowed as slices of their own, find the
  // smallest target for everybody else - same makespan by the model, but the other ranks finish early and errors
  // of the model no longer add to the critical path.
  if (have && nranks > 1) {
    const double cap = hi * 1.0001;
    double lo2 = all.us / nranks * 0.5, hi2 = hi;
    for (int it = 0; it < 60; it++) {
      const double T = 0.5 * (lo2 + hi2);
      if (cut_for_target(descs, nwin, nranks, T, cap, total, cuts.data())) { hi2 = T; best = cuts; }
      else lo2 = T;
      if (hi2 - lo2 < 1e-5 * all.us) break;
    }
  }
  if (!have) {                                               // cannot happen (one rank can take everything within all.us)
    for (int r = 0; r <= nranks; r++) best[(size_t)r] = r == 0 ? 0 : total;
  }
  *begin = best[(size_t)rank];
  *count = best[(size_t)rank + 1] - best[(size_t)rank];
  return BHW_OK;
}

}  // extern "C"
