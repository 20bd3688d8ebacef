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
  memset(&j, 0, sizeof(j));
  j.sp = t.canon;
  j.tab = t.data.data();
  j.entries = entries;
  j.fast = (!force_generic && fast32_ok(t.canon)) ? 1u : 0u;
  j.work = t.canon.kind == SRC_INQ ? entries : entries / 4;
  for (uint32_t e = 0; e < j.work; e++) table_build_item(j, rom.data(), e);
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
  for (uint64_t j = 0; j < count; j++)
    out[j] = direct_sample_generic(wp, src, rom.data(), (n0 + j + (uint64_t)wp.stream_offset) & nmask);
  return 0;
}

// BHW_ALGO_TABLE bodies: stage 1 (tables) then stage 2 (synthesis), as bhw_api.cu plans them.
// Returns 1 if the descriptor is not eligible for the fast tail (caller should expect the
// generic body instead), negative on error.
int hc_table(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out, int force_generic_core) {
  WinParams wp; SrcParams src[2];
  int st = resolve_window(d, &wp, src);
  if (st) return st;
  if (!fast_tail_exact(wp)) return 1;
  std::vector<I2> rom;
  if (src[0].kind == SRC_TAYLOR) build_taylor_rom(src[0].dw, src[0].lut, rom);
  HostTable tabs[2];
  for (int u = 0; u < wp.nsrc; u++) build_table(src[u], rom, tabs[u], force_generic_core);
  WinRec r;
  memset(&r, 0, sizeof(r));
  fill_fast_rec(wp, r);
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

// the sin/cos entry body
int hc_sincos(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out_sin, int64_t* out_cos) {
  int st = validate_desc(d, false);
  if (st) return st;
  SrcParams sp;
  if ((st = resolve_source(d, 0, &sp))) return st;
  std::vector<I2> rom;
  if (sp.kind == SRC_TAYLOR) build_taylor_rom(sp.dw, sp.lut, rom);
  for (uint64_t j = 0; j < count; j++) {
    int64_t s, c;
    eval_source_generic(sp, rom.data(), n0 + j, s, c);
    if (out_sin) out_sin[j] = s;
    if (out_cos) out_cos[j] = c;
  }
  return 0;
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
  for (uint64_t ph = 0; ph < N; ph++) out_cos[ph] = t.data[ph >> t.drop];
  return 0;
}

}  // extern "C"
