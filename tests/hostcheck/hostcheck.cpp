// hostcheck.cpp - TEST INFRASTRUCTURE ONLY.
//
// Compiles the kernels' per-thread bodies (blackman_harris_win_b200/csrc/bhw_device.cuh) and the
// CUDA-free planning code with g++ and drives them with plain loops that mimic the kernels'
// index mapping.  This lets the CPU-only test tier check the *kernel arithmetic* against the
// oracle where no GPU exists.  It is never linked into the product library; the product has no
// CPU path.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../blackman_harris_win_b200/csrc/bhw_device.cuh"
#include "../../blackman_harris_win_b200/csrc/bhw_plan.h"
#include "../../blackman_harris_win_b200/csrc/bhw_group.cuh"

using namespace bhw;

namespace {

struct HostTable {
  SrcParams canon;
  uint32_t drop;
  std::vector<int32_t> data;
};

// mimic k_table_build for one job
void build_table(const SrcParams& sp, const std::vector<I2>& rom, HostTable& t, int force_generic) {
  t.canon = canonical_source(sp, &t.drop);
  const uint32_t entries = 1u << t.canon.pw;
  t.data.assign(entries, 0x7FFFFFFF);
  TabJob j;
  init_tab_job(t.canon, t.data.data(), &j);
  if (force_generic) j.fast = TABCORE_GENERIC;
  for (uint32_t e = 0; e < j.work; e++) table_build_item(j, rom.data(), e);
  if (!force_generic && table_build_unrolled_ok(j)) {   // the stage-unrolled kernel body must give the same table
    std::vector<int32_t> again(entries, 0x7FFFFFFF);
    TabJob ju = j;
    ju.tab = again.data();
    for (uint32_t e = 0; e < ju.work; e++) {
      if (ju.sp.kind == SRC_INQ) {   // the kernel takes two entries per item (quadrants sharing their z)
        if (e >= ju.work / 2) break;
        if (ju.sp.n_xy == 16) table_build_item_inq_u2<16>(ju, e);
        else if (ju.sp.n_xy == 17) table_build_item_inq_u2<17>(ju, e);
        else if (ju.sp.n_xy == 24) table_build_item_inq_u2<24>(ju, e);
        else table_build_item_inq_u2<32>(ju, e);
      }
      else if (ju.fast == TABCORE_32BIAS) table_build_item_u<31, true>(ju, e);
      else if (ju.sp.n_xy == 15) table_build_item_u<15, false>(ju, e);
      else if (ju.sp.n_xy == 16) table_build_item_u<16, false>(ju, e);
      else table_build_item_u<23, false>(ju, e);
    }
    if (again != t.data) t.data.assign(entries, 0x7FFFFFFE);   // poison: the caller's comparison fails
  }
}

}  // namespace

extern "C" {

// BHW_ALGO_DIRECT body: out[j] = window sample n0 + j (+ stream offset)
int hc_direct(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out) {
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  std::vector<I2> rom;
  if (src[0].kind == SRC_TAYLOR) build_taylor_rom(src[0].dw, src[0].lut, rom);
  const uint64_t nmask = (1ull << wp.pw) - 1;
  SrcCore sc[2];
  for (int u = 0; u < wp.nsrc; u++) init_src_core(src[u], &sc[u]);
  for (uint64_t j = 0; j < count; j++) {
    const uint64_t n = (n0 + j + (uint64_t)wp.stream_offset) & nmask;
    out[j] = direct_sample_core(wp, src, sc, rom.data(), n);          // what k_direct_window runs
    if (direct_sample_generic(wp, src, rom.data(), n) != out[j]) return -100;  // fast cores == generic body
  }
  // the paired body (whole windows in the kernel) must give the same two samples
  uint32_t flip = 0;
  const uint64_t N = 1ull << wp.pw;
  if (N >= 8 && direct_pair_flip(wp, src, &flip)) {
    for (uint64_t j = 0; j < count; j++) {
      const uint64_t pos = n0 + j, partner = (pos + N / 2) & (N - 1);
      if (partner < n0 || partner >= n0 + count) continue;
      int64_t wa, wb;
      direct_sample_core_pair(wp, src, sc, rom.data(), (pos + (uint64_t)wp.stream_offset) & nmask, flip, wa, wb);
      if (wa != out[j] || wb != out[partner - n0]) return -103;
    }
  }
  // ... and the four-samples-per-evaluation body
  uint32_t adv = 0;
  if (direct_quad_adv(wp, src, &adv)) {
    for (uint64_t j = 0; j < count; j++) {
      const uint64_t pos = n0 + j;
      bool all_in = true;
      for (int r = 1; r < 4; r++) {
        const uint64_t pr = (pos + r * (N / 4)) & (N - 1);
        if (pr < n0 || pr >= n0 + count) all_in = false;
      }
      if (!all_in) continue;
      int64_t w4[4];
      direct_sample_core_quad(wp, src, sc, rom.data(), (pos + (uint64_t)wp.stream_offset) & nmask, adv, w4);
      for (int r = 0; r < 4; r++)
        if (w4[r] != out[((pos + r * (N / 4)) & (N - 1)) - n0]) return -107;
    }
  }
  return 0;
}

// the 32-bit TAYLOR direct body (k_direct_taylor); returns 1 when the window is not eligible
int hc_direct_taylor(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out) {
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  DirectTayParams p;
  if (wp.elem64 || !direct_taylor_params(wp, src, &p)) return 1;
  std::vector<I2> rom;
  build_taylor_rom(src[0].dw, src[0].lut, rom);
  for (uint64_t j = 0; j < count; j++) {
    const uint32_t n = (uint32_t)(n0 + j) + p.n_first;
    out[j] = p.tmode == TMODE_ROM ? direct_taylor_sample<TMODE_ROM>(p, rom.data(), n)
           : p.tmode == TMODE_DSP ? direct_taylor_sample<TMODE_DSP>(p, rom.data(), n)
                                  : direct_taylor_sample<TMODE_WIDE>(p, rom.data(), n);
  }
  // the paired body (whole windows in the kernel) must give the same two samples
  const uint64_t N = 1ull << d->phi_width;
  if (p.unit[0].pw == d->phi_width && (p.m == 2 || p.unit[1].pw == d->phi_width - 1)) {
    for (uint64_t j = 0; j < count; j++) {
      const uint64_t pos = n0 + j, partner = (pos + N / 2) & (N - 1);
      if (partner < n0 || partner >= n0 + count) continue;
      const uint32_t n = (uint32_t)pos + p.n_first;
      int32_t wa, wb;
      if (p.tmode == TMODE_ROM) direct_taylor_pair<TMODE_ROM>(p, rom.data(), n, wa, wb);
      else if (p.tmode == TMODE_DSP) direct_taylor_pair<TMODE_DSP>(p, rom.data(), n, wa, wb);
      else direct_taylor_pair<TMODE_WIDE>(p, rom.data(), n, wa, wb);
      if (wa != out[j] || wb != out[partner - n0]) return -101;
    }
  }
  // the quarter-window body (long whole TAY_WIDE windows in the kernel): 16 samples per call
  if (n0 == 0 && count == N && direct_taylor_quad_ok(p, d->phi_width)) {
    for (uint64_t j = 0; j < N / 4; j += 4) {
      int32_t w[4][4];
      direct_taylor_quad4(p, rom.data(), (uint32_t)j, w);
      for (int r = 0; r < 4; r++)
        for (int e = 0; e < 4; e++)
          if (w[r][e] != out[r * (N / 4) + j + e]) return -102;
    }
  }
  return 0;
}

// 1 when hc_direct_taylor(d, 0, N) also runs the quarter-window body (k_direct_taylor's pair == 2 branch)
int hc_taylor_quad_ok(const bhw_desc* d) {
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  DirectTayParams p;
  if (wp.elem64 || !direct_taylor_params(wp, src, &p)) return 0;
  return direct_taylor_quad_ok(p, d->phi_width) ? 1 : 0;
}

// the 32-bit register-resident direct body (k_direct32); returns 1 when the window is not eligible
int hc_direct32(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out) {
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  Direct32Params p;
  if (wp.elem64 || !direct32_params(wp, src, &p)) return 1;
  for (uint64_t j = 0; j < count; j++) {
    const uint32_t n = (uint32_t)(n0 + j) + p.n_first;
    switch (p.n_xy) {   // the unrolled instantiations the launcher uses, plus the run-time loop
      case 7: out[j] = direct32_sample<7>(p, n); break;
      case 11: out[j] = direct32_sample<11>(p, n); break;
      case 15: out[j] = direct32_sample<15>(p, n); break;
      case 16: out[j] = direct32_sample<16>(p, n); break;
      case 19: out[j] = direct32_sample<19>(p, n); break;
      case 23: out[j] = direct32_sample<23>(p, n); break;
      case 24: out[j] = direct32_sample<24>(p, n); break;
      case 29: out[j] = direct32_sample<29>(p, n); break;
      case 30: out[j] = direct32_sample<30>(p, n); break;
      default: out[j] = direct32_sample<0>(p, n); break;
    }
    if (direct32_sample<0>(p, n) != out[j]) return -100;   // unrolled and looped forms agree
  }
  // the paired body (whole windows in the kernel) must give the same two samples
  const uint64_t N = 1ull << d->phi_width;
  if (p.pw == d->phi_width) {
    for (uint64_t j = 0; j < count; j++) {
      const uint64_t pos = n0 + j, partner = (pos + N / 2) & (N - 1);
      if (partner < n0 || partner >= n0 + count) continue;
      int32_t wa, wb;
      const uint32_t n = (uint32_t)pos + p.n_first;
      if (p.n_xy == 15) direct32_pair<15>(p, n, wa, wb);
      else if (p.n_xy == 16) direct32_pair<16>(p, n, wa, wb);
      else direct32_pair<0>(p, n, wa, wb);
      if (wa != out[j] || wb != out[partner - n0]) return -102;
    }
    // ... and the four-samples-per-evaluation body
    for (uint64_t j = 0; j < count && N >= 16; j++) {
      const uint64_t pos = n0 + j;
      bool all_in = true;
      for (int r = 1; r < 4; r++) {
        const uint64_t pr = (pos + r * (N / 4)) & (N - 1);
        if (pr < n0 || pr >= n0 + count) all_in = false;
      }
      if (!all_in) continue;
      int32_t w4[4];
      const uint32_t n = (uint32_t)pos + p.n_first;
      if (p.n_xy == 15) direct32_quad<15>(p, n, w4);
      else if (p.n_xy == 16) direct32_quad<16>(p, n, w4);
      else direct32_quad<0>(p, n, w4);
      for (int r = 0; r < 4; r++)
        if (w4[r] != out[((pos + r * (N / 4)) & (N - 1)) - n0]) return -106;
    }
  }
  return 0;
}

// BHW_ALGO_TABLE bodies: stage 1 (tables) then stage 2 (synthesis), as bhw_api.cu plans them.
// Returns 1 if the descriptor is not eligible for the fast tail (caller should expect the
// generic body instead), negative on error.
int hc_table(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out, int force_generic_core) {
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  if (fast_tail_mode(wp, src) == TAILMODE_GENERIC) return 1;
  std::vector<I2> rom;
  if (src[0].kind == SRC_TAYLOR) build_taylor_rom(src[0].dw, src[0].lut, rom);
  HostTable tabs[2];
  for (int u = 0; u < wp.nsrc; u++) build_table(src[u], rom, tabs[u], force_generic_core);
  WinRec r;
  memset(&r, 0, sizeof(r));
  fill_fast_rec(wp, src, r);
  r.n_first = (uint32_t)wp.stream_offset;
  for (int k = 1; k < wp.m; k++) {
    const TermParams& t = wp.term[k - 1];
    const SrcParams& sp = src[t.src];
    r.tabp[k] = tabs[t.src].data.data();
    r.kstep[k] = t.kmul << (32 - sp.pw);
    r.idx_rsh[k] = (uint32_t)(32 - (sp.pw - (int)tabs[t.src].drop));
  }
  for (uint64_t j = 0; j < count; j++) out[j] = synth_sample(r, (uint32_t)(n0 + j) + r.n_first);
  return 0;
}

// Bank kernel body (k_synth_bank) for one whole window, with the kernel's tile/lane mapping.
// force_mode: -1 = the planner's choice for `smem_limit` bytes, else TAB_* ; force_pair: -1/0/1.
// Returns 1 when the window is not bank-eligible (or the forced combination is not legal).
static uint64_t lin_tiles = 0;   // lane-tiles that took the linear path since the last reset
static uint64_t inq_exceptions_seen = 0;
uint64_t hc_inq_exceptions(int reset) { const uint64_t v = inq_exceptions_seen; if (reset) inq_exceptions_seen = 0; return v; }
uint64_t hc_lin_tiles(int reset) { const uint64_t v = lin_tiles; if (reset) lin_tiles = 0; return v; }

int hc_bank(const bhw_desc* d, int64_t* out, uint64_t smem_limit, int force_mode, int force_pair) {
  const bool no_lin = force_pair >= 16;  // force_pair + 16: keep every tile on the general path
  if (no_lin) force_pair -= 16;
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  if (fast_tail_mode(wp, src) == TAILMODE_GENERIC) return 1;
  std::vector<I2> rom;
  if (src[0].kind == SRC_TAYLOR) build_taylor_rom(src[0].dw, src[0].lut, rom);
  HostTable tabs[2];
  for (int u = 0; u < wp.nsrc; u++) build_table(src[u], rom, tabs[u], 0);
  WinRec r;
  memset(&r, 0, sizeof(r));
  fill_fast_rec(wp, src, r);
  r.n_first = (uint32_t)wp.stream_offset;
  BankTableInfo tk[BHW_MAX_TERMS];
  memset(tk, 0, sizeof(tk));
  for (int k = 1; k < wp.m; k++) {
    const TermParams& t = wp.term[k - 1];
    const SrcParams& sp = src[t.src];
    r.tabp[k] = tabs[t.src].data.data();
    r.kstep[k] = t.kmul << (32 - sp.pw);
    r.idx_rsh[k] = (uint32_t)(32 - (sp.pw - (int)tabs[t.src].drop));
    tk[k].ptr = r.tabp[k];
    tk[k].entries = (uint32_t)tabs[t.src].data.size();
    tk[k].antisym = source_antisymmetric(tabs[t.src].canon);
    tk[k].inq_comp = source_inq_complement(tabs[t.src].canon);
  }
  BankShape sh;
  int mode; bool pair;
  if (!bank_shape(r, tk, smem_limit, &sh, &mode, &pair)) return 1;
  if (force_pair == 1 && !pair) return 1;           // pairing needs an antisymmetric table
  if (force_pair == 0) pair = false;
  if (force_mode >= 0 && force_mode != mode) {
    if (force_mode == TAB_SMEM_HALF) {
      // legal only if the planner itself could have chosen it: re-plan with a limit that forces it
      size_t words = 0;
      for (uint32_t u = 0; u < sh.ntab; u++) words += sh.tentries[u];
      if (!bank_shape(r, tk, (words / 2) * 4, &sh, &mode, &pair) || mode != TAB_SMEM_HALF) return 1;
    } else if (force_mode == TAB_GLOBAL) {
      if (!bank_shape(r, tk, 0, &sh, &mode, &pair) || mode != TAB_GLOBAL) return 1;
      if (force_pair == 0) pair = false;
    } else return 1;
  }
  if (mode == TAB_SMEM_HALF && !pair) return 1;
  // "stage" the tables as the kernel does
  std::vector<int32_t> staged(sh.smem_words ? sh.smem_words : 1);
  const int32_t* tp[2] = {sh.tab[0], sh.tab[1]};
  if (mode != TAB_GLOBAL) {
    for (uint32_t u = 0; u < sh.ntab; u++) {
      const uint32_t words = sh.tentries[u] >> (mode == TAB_SMEM_HALF ? 1 : 0);
      memcpy(staged.data() + sh.toff[u], sh.tab[u], words * sizeof(int32_t));
    }
    tp[0] = staged.data() + sh.toff[0];
    tp[1] = staged.data() + sh.toff[1];
  }
  const uint32_t pw = sh.pw, log_tpw = pw - kBankTileLog2 - (pair ? 1 : 0), half = 1u << (pw - 1);
  for (uint32_t t = 0; t < (1u << log_tpw); t++) {
    const uint32_t nbase = t * kBankTile + r.n_first;
    for (uint32_t lane = 0; lane < 32; lane++) {
      int32_t va[kBankJ], vb[kBankJ];
      const uint32_t n = nbase + lane;
#define HC_TILE2(M, TAB, PAIR, W64)                                                                \
      do {                                                                                         \
        uint32_t lbase[M], lneg;                                                                   \
        if (sh.lin && !no_lin && bank_tile_linear<M, TAB>(sh, nbase, lbase, &lneg)) {              \
          bank_lane_tile_lin<M, TAB, PAIR, W64>(sh, r.A, r.S0, tp, lane, lbase, lneg, va, vb);     \
          lin_tiles++;                                                                             \
        } else if (TAB == TAB_SMEM_HALF)                                                           \
          bank_lane_tile<M, TAB, PAIR, true, W64>(sh, r.A, r.S0, tp, n, nbase, va, vb);            \
        else                                                                                       \
          bank_lane_tile<M, TAB, PAIR, false, W64>(sh, r.A, r.S0, tp, n, nbase, va, vb);           \
      } while (0)
#define HC_TILE(M, TAB, PAIR)                                                                      \
      do { if (sh.acc64) HC_TILE2(M, TAB, PAIR, true); else HC_TILE2(M, TAB, PAIR, false); } while (0)
#define HC_MODE(M)                                                                                 \
      do {                                                                                         \
        if (pair && sh.pair_adj) {      /* ones'-complement pairing: 32-bit tail, never the half-period placement */ \
          if (mode == TAB_SMEM_FULL) HC_TILE2(M, TAB_SMEM_FULL, 2, false); else HC_TILE2(M, TAB_GLOBAL, 2, false);   \
        }                                                                                          \
        else if (mode == TAB_SMEM_FULL) { if (pair) HC_TILE(M, TAB_SMEM_FULL, 1); else HC_TILE(M, TAB_SMEM_FULL, 0); } \
        else if (mode == TAB_SMEM_HALF) HC_TILE(M, TAB_SMEM_HALF, 1);                              \
        else { if (pair) HC_TILE(M, TAB_GLOBAL, 1); else HC_TILE(M, TAB_GLOBAL, 0); }              \
      } while (0)
      switch (sh.m) {
        case 2: HC_MODE(2); break;
        case 3: HC_MODE(3); break;
        case 4: HC_MODE(4); break;
        case 5: HC_MODE(5); break;
        default: HC_MODE(7); break;
      }
      for (int j = 0; j < kBankJ; j++) {
        out[t * kBankTile + lane + 32 * j] = va[j];
        if (pair) out[half + t * kBankTile + lane + 32 * j] = vb[j];
      }
    }
  }
  if (pair && sh.pair_adj) {
    // what k_inq_exceptions / k_inq_patch do: list the entries that break T[i + E/2] == -T[i] - adj, recompute the
    // sample pairs that read them
    if (mode == TAB_SMEM_HALF || sh.acc64) return -300;
    const uint32_t E = sh.tentries[0], nmask = (1u << pw) - 1u;
    const int32_t* T = sh.tab[0];
    uint64_t nexc = 0;
    for (uint32_t i = 0; i < E / 2; i++) {
      if (T[i + E / 2] == -T[i] - (int32_t)sh.pair_adj) continue;
      nexc++;
      for (uint32_t k = 1; k < sh.m; k += 2) {
        uint32_t inv = k;
        inv *= 2u - k * inv; inv *= 2u - k * inv; inv *= 2u - k * inv; inv *= 2u - k * inv;
        const uint32_t n = (i * inv) & nmask;
        for (int h = 0; h < 2; h++) {
          const uint32_t p = (n + (h ? half : 0u)) & nmask;
          out[(p - r.n_first) & nmask] = synth_sample(r, p);
        }
      }
    }
    inq_exceptions_seen += nexc;
  }
  return 0;
}


// Group kernel body (k_synth_group) for a list of windows of ONE family and entity, any PHI_WIDTHs: builds the
// family's half-period pyramid with the table-builder bodies (as bhw_api.cu plans it), then runs every
// window through group_lane_tile with the kernel's tile/lane mapping.  out = the windows' samples concatenated.
// force_tab: -1 = the planner's choice for 192 KB, else G_*; unpaired: 1 = the unpaired instantiation
// (what a window cut by the requested range gets).  Also checks the pyramid against the level rule
// and runs the general-kernel body (WR_HALFTAB records) on the first samples of every window.
// Returns 1 when the windows are not group material (or the forced placement does not fit).
int hc_group(const bhw_desc* descs, int nwin, int64_t* out, int force_tab, int unpaired) {
  if (nwin <= 0) return -1;
  SrcParams key0;
  uint32_t max_pw = 0, min_pw = 99;
  std::vector<WinRec> recs((size_t)nwin);
  for (int w = 0; w < nwin; w++) {
    WinParams wp; SrcParams src[2];
    int st = resolve_window(&descs[w], &wp, src);
    if (st) return st;
    if (!group_eligible(descs[w], wp, src)) return 1;
    SrcParams key;
    if ((st = family_source(descs[w], BHW_MAX_PHI_WIDTH, &key))) return st;
    if (w == 0) key0 = key;
    else if (memcmp(&key, &key0, sizeof(key))) return 1;
    memset(&recs[w], 0, sizeof(WinRec));
    fill_fast_rec(wp, src, recs[w]);
    recs[w].n_first = (uint32_t)wp.stream_offset;
    if (recs[w].m != recs[0].m || recs[w].rc != recs[0].rc || recs[w].lsh != recs[0].lsh || recs[w].rsh != recs[0].rsh) return 1;
    if ((uint32_t)wp.pw > max_pw) max_pw = (uint32_t)wp.pw;
    if ((uint32_t)wp.pw < min_pw) min_pw = (uint32_t)wp.pw;
  }
  const uint32_t res = (uint32_t)key0.pw;
  const uint32_t top = max_pw < res ? max_pw : res;
  const uint32_t low = min_pw < top ? min_pw : top;
  const uint32_t lmin = low > 4 ? low - 2 : 2;
  SrcParams canon;
  int st = family_source(descs[0], (int)top, &canon);
  if (st) return st;
  int tab = group_tab_mode(canon, top, 192 * 1024);
  if (force_tab >= 0) {
    if (force_tab == G_HALF32 && top > 22) return 1;
    if (force_tab == G_Q16 && (canon.dw > 17 || top > 22)) return 1;
    tab = force_tab;
  }
  std::vector<int32_t> pyr((size_t)1 << top, 0x7FFFFFFF);
  std::vector<uint16_t> q16((size_t)1 << (top - 1), 0xFFFF);
  TabJob j;
  init_pyramid_job(canon, lmin, pyr.data(), tab == G_Q16 ? q16.data() : nullptr, &j);
  std::vector<I2> norom;
  for (uint32_t e = 0; e < j.work; e++) table_build_item(j, norom.data(), e);
  if (table_build_unrolled_ok(j)) {      // the stage-unrolled kernel body must give the same pyramid
    std::vector<int32_t> again((size_t)1 << top, 0x7FFFFFFF);
    TabJob ju = j;
    ju.tab = again.data();
    ju.q16 = nullptr;
    for (uint32_t e = 0; e < ju.work; e++) {
      if (ju.fast == TABCORE_32BIAS) table_build_item_u<31, true>(ju, e);
      else if (ju.sp.n_xy == 15) table_build_item_u<15, false>(ju, e);
      else if (ju.sp.n_xy == 16) table_build_item_u<16, false>(ju, e);
      else table_build_item_u<23, false>(ju, e);
    }
    if (again != pyr) return -201;
  }
  // level rule: level L-1 is every second entry of level L
  for (uint32_t L = top; L > lmin; L--)
    for (uint32_t i = 0; i < (1u << (L - 2)); i++)
      if (pyr[(1u << (L - 2)) + i] != pyr[(1u << (L - 1)) + 2 * i]) return -202;
  GroupShape sh;
  group_shape(recs[0], top, lmin, &sh);
  sh.pyr = pyr.data();
  sh.q16 = q16.data();
  const void* img = tab == G_HALF32 ? (const void*)(pyr.data() + ((size_t)1 << (top - 1)))
                  : tab == G_Q16 ? (const void*)q16.data() : (const void*)pyr.data();
  uint64_t off = 0;
  for (int w = 0; w < nwin; w++) {
    const WinRec& r = recs[w];
    const uint32_t pw = r.pw;
    const uint64_t N = 1ull << pw;
    const uint32_t tiles = unpaired ? (uint32_t)(N >> kBankTileLog2) : (uint32_t)(N >> (kBankTileLog2 + 1));
    for (uint32_t t = 0; t < tiles; t++) {
      for (uint32_t lane = 0; lane < 32; lane++) {
        int32_t va[kBankJ], vb[kBankJ];
        const uint32_t nbase = t * kBankTile + r.n_first;
#define HC_G3(M, TAB) do { if (unpaired) group_lane_tile<M, TAB, false>(sh, pw, r.A, r.S0, img, nbase, lane, va, vb); \
                           else group_lane_tile<M, TAB, true>(sh, pw, r.A, r.S0, img, nbase, lane, va, vb); } while (0)
#define HC_G2(M) do { if (tab == G_HALF32) HC_G3(M, G_HALF32); else if (tab == G_Q16) HC_G3(M, G_Q16); else HC_G3(M, G_GLOBAL); } while (0)
        switch (r.m) {
          case 2: HC_G2(2); break;
          case 3: HC_G2(3); break;
          case 4: HC_G2(4); break;
          case 5: HC_G2(5); break;
          default: HC_G2(7); break;
        }
        for (int jj = 0; jj < kBankJ; jj++) {
          out[off + (uint64_t)t * kBankTile + lane + 32 * jj] = va[jj];
          if (!unpaired) out[off + N / 2 + (uint64_t)t * kBankTile + lane + 32 * jj] = vb[jj];
        }
      }
    }
    // the general kernel's body on a record that reads the pyramid (WR_HALFTAB), a few samples
    WinRec rr = r;
    const uint32_t L = pw < top ? pw : top;
    rr.flags |= WR_HALFTAB;
    for (uint32_t k = 1; k < rr.m; k++) {
      rr.kstep[k] = k << (32 - pw);
      rr.idx_rsh[k] = 32 - L;
      rr.tabp[k] = pyr.data() + ((size_t)1 << (L - 1));
    }
    for (uint64_t n = 0; n < N; n += (N > 4096 ? 61 : 1))
      if ((int64_t)synth_sample(rr, (uint32_t)n + rr.n_first) != out[off + n]) return -203;
    off += N;
  }
  return 0;
}

// the sin/cos entry body
int hc_sincos(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out_sin, int64_t* out_cos) {
  int st = validate_desc(d, false);
  if (st) return st;
  SrcParams sp;
  if ((st = resolve_source(d, 0, &sp))) return st;
  std::vector<I2> rom;
  if (sp.kind == SRC_TAYLOR) build_taylor_rom(sp.dw, sp.lut, rom);
  SrcCore sc;
  init_src_core(sp, &sc);
  for (uint64_t j = 0; j < count; j++) {
    int64_t s, c, s2, c2;
    eval_source_core(sp, sc, rom.data(), n0 + j, s, c);               // what k_sincos runs
    eval_source_generic(sp, rom.data(), n0 + j, s2, c2);
    if (s != s2 || c != c2) return -100;
    if (out_sin) out_sin[j] = s;
    if (out_cos) out_cos[j] = c;
    if (sp.kind != SRC_INQ && sp.pw >= 3) {   // the four-phases-per-evaluation body of whole-table requests
      int64_t s4[4], c4[4];
      eval_source_core_quad(sp, sc, rom.data(), n0 + j, s4, c4);
      for (int r = 0; r < 4; r++) {
        eval_source_generic(sp, rom.data(), n0 + j + ((uint64_t)r << (sp.pw - 2)), s2, c2);
        if (s4[r] != s2 || c4[r] != c2) return -104;
      }
    }
  }
  return 0;
}

// the cordic_atan2 body (k_atan2)
int hc_atan2(const bhw_atan2_desc* d, const int32_t* x, const int32_t* y, int32_t* phi, uint64_t count) {
  Atan2Params p;
  int st = resolve_atan2(d, &p);
  if (st) return st;
  for (uint64_t j = 0; j < count; j++) {
    // the quadrant source the kernel picks: the pair itself, or (stream_quadrant) the next pair / zeros after the last
    const int32_t qx = !p.skew ? x[j] : (j + 1 < count ? x[j + 1] : 0), qy = !p.skew ? y[j] : (j + 1 < count ? y[j + 1] : 0);
    phi[j] = atan2_sample(p, x[j], y[j], qx, qy);
    if (p.fast32) {   // the stage-unrolled instantiations k_atan2_u uses
      int32_t u = phi[j];
      if (p.aw == 12) u = atan2_sample32_t<12>(p, x[j], y[j], qx, qy);
      else if (p.aw == 16) u = atan2_sample32_t<16>(p, x[j], y[j], qx, qy);
      else if (p.aw == 20) u = atan2_sample32_t<20>(p, x[j], y[j], qx, qy);
      else if (p.aw == 24) u = atan2_sample32_t<24>(p, x[j], y[j], qx, qy);
      if (u != phi[j]) return -100;
    }
  }
  return 0;
}

// 1 if the planner treats the window's first source as antisymmetric over half a period
int hc_source_antisymmetric(const bhw_desc* d) {
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  uint32_t drop;
  return source_antisymmetric(canonical_source(src[0], &drop)) ? 1 : 0;
}

// The bank kernel's tile walks, replayed with the kernel's own index helpers: every (window, tile) of
// the launch must be visited exactly once.  mode 0: spread walk of one window of U tiles over G warps
// and `grid` CTAs; mode 1: window-minor walk of nwin = G windows of U tiles each over `grid` CTAs of
// 32 warps.  Returns 0, or the number of units visited a wrong number of times.
int hc_walk(int mode, uint32_t U, uint32_t G, uint32_t grid) {
  const uint32_t warps = 32;
  if (mode == 0) {
    std::vector<uint8_t> seen(U, 0);
    const uint32_t L = spread_steps(U, G);
    for (uint32_t cta = 0; cta < grid; cta++) {
      const uint32_t i0 = (uint32_t)((uint64_t)L * cta / grid), i1 = (uint32_t)((uint64_t)L * (cta + 1) / grid);
      for (uint32_t warp = 0; warp < warps; warp++)
        for (uint32_t i = i0; i < i1; i++) {
          uint32_t t;
          if (spread_tile(U, G, warp, i, &t)) { if (t >= U) return -1; seen[t]++; }
        }
    }
    int bad = 0;
    for (uint32_t t = 0; t < U; t++) bad += seen[t] != 1;
    return bad;
  }
  const uint32_t nwin = G;
  const uint64_t units = (uint64_t)nwin * U;
  std::vector<uint8_t> seen(units, 0);
  for (uint32_t cta = 0; cta < grid; cta++) {
    const uint64_t u0 = units * cta / grid, u1 = units * (cta + 1) / grid;
    for (uint32_t warp = 0; warp < warps; warp++)
      for (uint64_t u = u0 + warp; u < u1; u += warps) {
        uint32_t w, t;
        win_minor_unit((uint32_t)u, nwin, &w, &t);
        if (w >= nwin || t >= U) return -1;
        seen[(uint64_t)w * U + t]++;
      }
  }
  int bad = 0;
  for (uint64_t i = 0; i < units; i++) bad += seen[i] != 1;
  return bad;
}

// which synthesis tail the planner picks: 0 = 32-bit, 1 = 64-bit, 2 = generic body
int hc_tail_mode(const bhw_desc* d) {
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  return (int)fast_tail_mode(wp, src);
}

// the cosine table of the window's first source, expanded to one value per phase
int hc_table_cos(const bhw_desc* d, int64_t* out_cos, int force_generic_core) {
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  std::vector<I2> rom;
  if (src[0].kind == SRC_TAYLOR) build_taylor_rom(src[0].dw, src[0].lut, rom);
  HostTable t;
  build_table(src[0], rom, t, force_generic_core);
  const uint64_t N = 1ull << src[0].pw;
  const int ts = table_tshift(t.canon);
  for (uint64_t ph = 0; ph < N; ph++) out_cos[ph] = t.data[ph >> t.drop] >> ts;
  return 0;
}

}  // extern "C"
