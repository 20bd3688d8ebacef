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

// Fill the 32-bit fast-tail record of a window (see the derivation above WinRec).
void fill_fast_rec(const WinParams& wp, WinRec& r) {
  const int dw = wp.dw, m = wp.m;
  r.m = (uint32_t)m; r.dw = (uint32_t)dw; r.pw = (uint32_t)wp.pw;
  r.flags = 0;
  if (dw > 16) r.flags |= WR_WIDE;
  for (int k = 0; k < m; k++) r.aa[k] = (int32_t)wp.aa[k];
  if (wp.tail == TAIL_HLS) {
    const int sh = 32 - dw;
    r.flags |= WR_HLS;
    r.bshift = dw - 2; r.rnd = 0;
    r.acc0 = (int32_t)((uint32_t)r.aa[0] << sh);
    r.fin_shift = sh;
    for (int k = 1; k < m; k++) r.mul[k] = (int32_t)((k & 1) ? (0u - (1u << sh)) : (1u << sh));
    return;
  }
  r.bshift = dw - 1; r.rnd = 1 << (dw - 2);
  if (wp.tail == TAIL_RTL2) {
    r.flags |= WR_RTL2;
    if (dw > 31) { r.flags |= WR_ACC64; return; }
    const int sh = 31 - dw;  // dsp_pp is DW+1 bits; +1 = the round-half-up increment
    r.acc0 = (int32_t)(((uint32_t)r.aa[0] << sh) + (1u << sh));
    r.fin_shift = sh + 1;
    r.mul[1] = (int32_t)(0u - (1u << sh));
    return;
  }
  if (dw > 30) { r.flags |= WR_ACC64; return; }
  const int sh = 30 - dw;    // dsp_pp is DW+2 bits; +2 = the increment of the bit-1 rounding
  r.acc0 = (int32_t)(((uint32_t)r.aa[0] << sh) + (2u << sh));
  r.fin_shift = sh + 2;
  for (int k = 1; k < m; k++) r.mul[k] = (int32_t)((k & 1) ? (0u - (1u << sh)) : (1u << sh));
}

// Does the fast tail reproduce the entity for these coefficients?  The only case it does not is
// AAk = -2^(DW-1) (k >= 1), where b_k can wrap to DW bits; such windows take the generic body.
bool fast_tail_exact(const WinParams& wp) {
  if (wp.dw > 32) return false;
  if (wp.tail == TAIL_HLS) return true;
  const int64_t lo = -((int64_t)1 << (wp.dw - 1));
  for (int k = 1; k < wp.m; k++) if (wp.aa[k] == lo) return false;
  return true;
}


}  // namespace bhw
