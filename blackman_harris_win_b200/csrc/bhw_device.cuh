// bhw_device.cuh - per-thread arithmetic of the kernels.
//
// Every kernel in bhw_kernels.cu is a thin index-mapping wrapper around a "body" function in
// this header.  Under nvcc the bodies are __device__ only; tests/hostcheck compiles the very same
// text with g++ (BHW_HD = static inline) to compare it with the oracle on a machine without a
// GPU.  That host build is test infrastructure only: the shipped library contains no CPU path.
//
// "Generic" bodies cover every legal descriptor in 64-bit registers with explicit wraps to the
// reference's signal widths (BHW_ALGO_DIRECT, DAT_WIDTH > 32, bhw_sincos, exotic widths).
// "Fast" bodies are the 32-bit specialisations used by the table path.
#pragma once
#include <stdint.h>

#include "bhw_internal.h"

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define BHW_HD __device__ __forceinline__
#define BHW_CONSTANT __constant__
#else
#define BHW_HD static inline
#define BHW_CONSTANT static const
#endif

// Debug build (csrc/build.sh debug -> libbhw_debug.so, -DBHW_BOUNDS_CHECK): every table look-up and every store
// of the synthesis kernels checks its index with a device-side assert.  compute-sanitizer is closed on the
// B200 pool this was developed on; tools/sanitize_smoke.py run against the debug library is the stand-in.
#if defined(BHW_BOUNDS_CHECK) && defined(__CUDA_ARCH__)
#include <assert.h>
#define BHW_CHECK(cond) assert(cond)
#else
#define BHW_CHECK(cond) ((void)0)
#endif

namespace bhw {

struct I2 { int32_t x, y; };  // (cos, sin) ROM word of taylor_sincos

// Two 48-entry x 48-bit atan tables, independently rounded in the reference:
// [0] pi/4 -> 2^46 (src/cordic_dds.vhd:104-117; also hls/cordic/cordic.cpp:57-70),
// [1] pi/4 -> 2^45 (src/cordic_dds48.vhd:115-128; also cpp/cordic_sincos.cpp:97-110).
BHW_CONSTANT int64_t c_atan[2][48] = {
    {0x400000000000ll, 0x25C80A3B3BE6ll, 0x13F670B6BDC7ll, 0x0A2223A83BBBll, 0x05161A861CB1ll,
     0x028BAFC2B209ll, 0x0145EC3CB850ll, 0x00A2F8AA23A9ll, 0x00517CA68DA2ll, 0x0028BE5D7661ll,
     0x00145F300123ll, 0x000A2F982950ll, 0x000517CC19C0ll, 0x00028BE60D83ll, 0x000145F306D6ll,
     0x0000A2F9836Dll, 0x0000517CC1B7ll, 0x000028BE60DCll, 0x0000145F306Ell, 0x00000A2F9837ll,
     0x00000517CC1Bll, 0x0000028BE60Ell, 0x00000145F307ll, 0x000000A2F983ll, 0x000000517CC2ll,
     0x00000028BE61ll, 0x000000145F30ll, 0x0000000A2F98ll, 0x0000000517CCll, 0x000000028BE6ll,
     0x0000000145F3ll, 0x00000000A2FAll, 0x00000000517Dll, 0x0000000028BEll, 0x00000000145Fll,
     0x000000000A30ll, 0x000000000518ll, 0x00000000028Cll, 0x000000000146ll, 0x0000000000A3ll,
     0x000000000051ll, 0x000000000029ll, 0x000000000014ll, 0x00000000000All, 0x000000000005ll,
     0x000000000003ll, 0x000000000001ll, 0x000000000000ll},
    {0x200000000000ll, 0x12E4051D9DF3ll, 0x09FB385B5EE4ll, 0x051111D41DDEll, 0x028B0D430E59ll,
     0x0145D7E15904ll, 0x00A2F61E5C28ll, 0x00517C5511D4ll, 0x0028BE5346D1ll, 0x00145F2EBB31ll,
     0x000A2F980092ll, 0x000517CC14A8ll, 0x00028BE60CE0ll, 0x000145F306C1ll, 0x0000A2F9836Bll,
     0x0000517CC1B7ll, 0x000028BE60DCll, 0x0000145F306Ell, 0x00000A2F9837ll, 0x00000517CC1Bll,
     0x0000028BE60Ell, 0x00000145F307ll, 0x000000A2F983ll, 0x000000517CC2ll, 0x00000028BE61ll,
     0x000000145F30ll, 0x0000000A2F98ll, 0x0000000517CCll, 0x000000028BE6ll, 0x0000000145F3ll,
     0x00000000A2FAll, 0x00000000517Dll, 0x0000000028BEll, 0x00000000145Fll, 0x000000000A30ll,
     0x000000000518ll, 0x00000000028Cll, 0x000000000146ll, 0x0000000000A3ll, 0x000000000051ll,
     0x000000000029ll, 0x000000000014ll, 0x00000000000All, 0x000000000005ll, 0x000000000003ll,
     0x000000000001ll, 0x000000000001ll, 0x000000000000ll}};

// wrap to b-bit two's complement, 1 <= b <= 64 (resp. 32)
BHW_HD int64_t wrapb(int64_t v, int b) {
  const int sh = 64 - b;
  return (int64_t)((uint64_t)v << sh) >> sh;
}
BHW_HD int32_t wrapb32(int32_t v, int b) {
  const int sh = 32 - b;
  return (int32_t)((uint32_t)v << sh) >> sh;
}

// ============================================================================================
// Generic (64-bit) sources
// ============================================================================================

// The shift-add core: returns the un-fixed (sin, cos) pair.  q/low are the quadrant bits and the
// remaining phase bits of a pw-bit phase.
BHW_HD void cordic_core_generic(const SrcParams& p, int q, uint64_t low, int64_t& vs, int64_t& vc) {
  const int pw = p.pw;
  const bool inq = p.kind == SRC_INQ;
  int64_t x = p.gain, y = 0, z;
  if (inq) {
    // quadrant folded into the start vector and the phase sign (src/cordic_dds48.vhd:170-216)
    uint64_t t = low | ((uint64_t)q << (pw - 2));
    if (q == 1) { t = low; x = 0; y = wrapb(-p.gain, p.w); }
    else if (q == 2) { t = low | (3ull << (pw - 2)); x = 0; y = p.gain; }
    z = wrapb((int64_t)(t << p.z_lshift), p.zw);
  } else {
    z = wrapb((int64_t)((low >> p.z_rshift) << p.z_lshift), p.zw);
  }
  for (int i = 0; i < p.n_xy; ++i) {
    const bool zneg = z < 0;
    const bool cw = inq ? !zneg : zneg;  // cw: x += y>>i, y -= x>>i; opposite sense in dds48 (:234-242)
    const int64_t xs = x >> i, ys = y >> i;
    const int64_t xn = cw ? x + ys : x - ys;
    const int64_t yn = cw ? y - xs : y + xs;
    if (i < p.n_z) {
      const int64_t r = (c_atan[p.rom_sel][i] >> p.rom_shift) & p.rom_mask;
      z = wrapb(zneg ? z + r : z - r, p.zw);
    }
    x = wrapb(xn, p.w);
    y = wrapb(yn, p.w);
  }
  vs = y >> p.out_shift;
  vc = x >> p.out_shift;
}

// taylor_sincos ROM look-up + tay1_order first-order correction, before the quadrant fix.
// rom[i] = (cos_i, sin_i), quarter wave, built on the host (src/taylor_sincos.vhd:91-111).
BHW_HD void taylor_core_generic(const SrcParams& p, const I2* rom, uint32_t t, int64_t& vs, int64_t& vc) {
  const int dw = p.dw;
  if (p.tay_mode == TAY_LESS || p.tay_mode == TAY_EQ) {
    const I2 e = rom[p.tay_mode == TAY_LESS ? (t << p.tay_ashift) : t];
    vc = e.x; vs = e.y;
    return;
  }
  const I2 e = rom[t >> p.tay_ashift];
  const int64_t c0 = e.x, s0 = e.y;
  const int64_t acnt = t & ((1u << p.tay_cbits) - 1);
  const int64_t mpi = (p.tay_pi * acnt) & 0xFFFFFF;  // 24-bit ROM word (src/tay1_order.vhd:136-147)
  const int xs = p.tay_xs;
  if (p.tay_mode == TAY_DSP) {
    // P = C -/+ A*B, result = P[xs+dw-1 : xs] (src/tay1_order.vhd:245,316,501-502)
    vc = wrapb(((c0 << xs) - mpi * s0) >> xs, dw);
    vs = wrapb(((s0 << xs) + mpi * c0) >> xs, dw);
  } else {
    const int64_t m1 = wrapb((s0 * mpi) >> xs, dw);                       // :585
    const int64_t m2 = wrapb((c0 * mpi) >> xs, dw);                       // :586
    const int64_t cp = wrapb(c0 - m1, dw), sp = wrapb(s0 + m2, dw);        // :595-596
    const int64_t sat = ((int64_t)1 << (dw - 1)) - 1;
    vc = cp < 0 ? sat : cp;                                               // negative -> max positive (:602-617)
    vs = sp < 0 ? sat : sp;
  }
}

// Output-side quadrant fix.  negw > 0: not(v)+1 in negw bits (src/cordic_dds.vhd:232-246,
// src/taylor_sincos.vhd:237-255, hls/windows/win_function.cpp:135-154); negw == 0: ~v
// (cpp/cordic_sincos.cpp:70-86).
BHW_HD void quadrant_fix(int q, int negw, int64_t vs, int64_t vc, int64_t& so, int64_t& co) {
  const int64_t ns = negw ? wrapb(-vs, negw) : ~vs;
  const int64_t nc = negw ? wrapb(-vc, negw) : ~vc;
  switch (q) {
    case 0: so = vs; co = vc; break;
    case 1: so = vc; co = ns; break;
    case 2: so = ns; co = nc; break;
    default: so = nc; co = vs; break;
  }
}

BHW_HD void eval_source_generic(const SrcParams& p, const I2* rom, uint64_t ph, int64_t& s, int64_t& c) {
  const int pw = p.pw;
  ph &= (1ull << pw) - 1;
  const int q = (int)(ph >> (pw - 2));
  const uint64_t low = ph & ((1ull << (pw - 2)) - 1);
  int64_t vs, vc;
  if (p.kind == SRC_TAYLOR) taylor_core_generic(p, rom, (uint32_t)low, vs, vc);
  else cordic_core_generic(p, q, low, vs, vc);
  if (p.kind != SRC_INQ) quadrant_fix(q, p.negw, vs, vc, vs, vc);
  s = wrapb(vs, p.outw);
  c = wrapb(vc, p.outw);
}

// Window tails.  cosv[k] is the DW-bit cosine of harmonic k (index 0 unused).
BHW_HD int64_t tail_generic(const WinParams& wp, const int64_t* cosv) {
  const int dw = wp.dw, m = wp.m;
  if (wp.tail == TAIL_HLS) {
    // m_k = (a_k*c_k) >> (NW-2); out = (win_t)(a0 - m1 + m2 - ...) (win_function.cpp:182,225,275,332,375)
    int64_t acc = wp.aa[0];
    for (int k = 1; k < m; ++k) {
      const int64_t mk = (wp.aa[k] * cosv[k]) >> (dw - 2);  // |a|,|c| <= 2^31: fits 64 bits
      acc += (k & 1) ? -mk : mk;
    }
    return wrapb(acc, dw);
  }
  // RTL: p = AAk*cos_k (src/int_multNxN_dsp48.vhd:105); r = p[2DW-2:DW-2]; b = (r>>1)+(r&1) in DW bits
  int64_t sum = wp.aa[0];
  for (int k = 1; k < m; ++k) {
    int64_t r;
    if (dw <= 32) r = wrapb((wp.aa[k] * cosv[k]) >> (dw - 2), dw + 1);
    else r = wrapb((int64_t)(((__int128)wp.aa[k] * (__int128)cosv[k]) >> (dw - 2)), dw + 1);
    const int64_t b = wrapb((r >> 1) + (r & 1), dw);
    sum += (k & 1) ? -b : b;
  }
  if (wp.tail == TAIL_RTL2) {
    const int64_t pp = wrapb(sum, dw + 1);         // src/hamming_win.vhd:214
    return wrapb((pp >> 1) + (pp & 1), dw);        // :220-228
  }
  const int64_t pp = wrapb(sum, dw + 2);           // dsp_pp (src/bh_win_3term.vhd:286-288 ...)
  return wrapb((pp >> 2) + ((pp >> 1) & 1), dw);   // rounds on bit 1 (:295-306)
}

// One output sample of BHW_ALGO_DIRECT: every k*phi term evaluated in registers.
BHW_HD int64_t direct_sample_generic(const WinParams& wp, const SrcParams* src, const I2* rom, uint64_t n) {
  int64_t cosv[BHW_MAX_TERMS];
  cosv[0] = 0;
  for (int k = 1; k < wp.m; ++k) {
    const TermParams& t = wp.term[k - 1];
    const uint64_t ph = ((uint64_t)t.kmul * n) & t.ph_mask;
    int64_t s;
    eval_source_generic(src[t.src], rom, ph, s, cosv[k]);
  }
  return tail_generic(wp, cosv);
}

// ============================================================================================
// Trig-table builder bodies (BHW_ALGO_TABLE, stage 1)
// ============================================================================================
// A table holds one full period of the source's cosine output, one int32 per *distinct*
// phase: value(ph) = T[ph >> drop] where `drop` counts the low phase bits the source ignores
// (cordic_dds with PHASE_WIDTH > DATA_WIDTH only looks at the top DATA_WIDTH phase bits,
// src/cordic_dds.vhd:159-162).  The job's source is the canonical one with those bits removed.
// For the output-quadrant sources one core evaluation yields the four entries
// e, e+Q, e+2Q, e+3Q (Q = entries/4): cos = c, -s, -c, s (negations wrap in negw bits).

struct TabJob {
  SrcParams sp;         // canonical source: phase width reduced so that no phase bit is ignored
  int32_t* tab;         // the table (device memory), `entries` int32 words
  uint32_t entries;     // 2^sp.pw
  uint32_t fast;        // TABCORE_*: which shift-add core evaluates this source exactly
  uint32_t work_begin;  // prefix sum of work items (threads) over the jobs of a launch
  uint32_t work;        // work items of this job: entries/4, or entries for SRC_INQ
  uint32_t rom_off;     // Taylor ROM offset (I2 units) in the rom buffer
  uint32_t tshift;      // entries are stored left-shifted by this much (see WinRec)
  uint32_t pyr_lmin;    // > 0: `tab` is a half-period pyramid (bhw_group.cuh) with top level sp.pw and lowest
                        // level pyr_lmin instead of a plain full-period table (output-quadrant sources only)
  uint32_t pad0;
  uint16_t* q16;        // pyramid jobs: also write the uint16 quarter-wave image (cos16[Q], sin16[Q]) here, or NULL
  int32_t rom32[32];    // TABCORE_32 / _32BIAS: atan word of stage i sliced for this register width (0 past n_z);
                        // TABCORE_A64: high half of rom64[i] plus 1 when its low half reads negative as an int32
  int64_t rom64[48];    // TABCORE_A64: atan word of stage i, sliced and left-aligned to bit 63 (0 past n_z)
};
enum : uint32_t { TABCORE_GENERIC = 0, TABCORE_32 = 1, TABCORE_32BIAS = 2, TABCORE_A64 = 3 };

// 32-bit CORDIC for output-quadrant sources whose registers fit int32 and never wrap:
// SRC_DDS with 8 <= DW, DW+PRECISION <= 32 and SRC_HLS with 8 <= NW <= 30.  Amplitude is a
// quarter of the register range, so x, y, z stay inside W-1 bits and the reference's wraps are
// no-ops (DESIGN.md "no-wrap argument").  z0 >= 0, so stage 0 always takes the z >= 0 branch.
//
// BIAS: cordic_dds with DW+PRECISION == 33 (DAT_WIDTH 32).  x and y live in (-2^29, 2^31 + eps):
// one bit too many for int32, so the registers hold X = x - 2^30, Y = y - 2^30.  For i <= 30,
// (X + 2^30) >> i == (X >> i) + 2^(30-i) exactly, so every stage and the output shift stay exact.
// z0 < 2^31 is formed as uint32; after stage 0, |z| <= 2^30.
template <bool BIAS>
BHW_HD void cordic_core_fast32(const SrcParams& p, const int32_t* __restrict__ rom32, uint32_t low, int32_t& vs,
                               int32_t& vc) {
  const uint32_t z0 = (low >> p.z_rshift) << p.z_lshift;
  const int32_t g = (int32_t)p.gain;
  const int32_t K = BIAS ? (1 << 30) : 0;
  int32_t x = g - K, y = g - K;  // stage 0: x - (0>>0), 0 + (x>>0)
  int32_t z = (int32_t)(z0 - (uint32_t)rom32[0]);
  const int n_xy = p.n_xy;
#pragma unroll 4
  for (int i = 1; i < n_xy; ++i) {
    const int32_t d = (z >> 31) | 1;  // -1 when z < 0, else +1
    const int32_t ki = BIAS ? (K >> i) : 0;
    const int32_t xs = (x >> i) + ki, ys = (y >> i) + ki;
    x -= d * ys;                      // z<0: x + (y>>i)   (src/cordic_dds.vhd:199-205)
    y += d * xs;                      // z<0: y - (x>>i)
    z -= d * rom32[i];                // the word is 0 for stages past n_z
  }
  vs = (y >> p.out_shift) + (BIAS ? (K >> p.out_shift) : 0);
  vc = (x >> p.out_shift) + (BIAS ? (K >> p.out_shift) : 0);
}

// The same core with the stage count as a template parameter: every shift is an immediate, the
// bias words are constants, the atan words come from the kernel parameter block (constant bank).
// Uses multiply-ADD forms only (x = ys*nd + x ...) so that each update is one IMAD.
template <int NXY, bool BIAS>
BHW_HD void cordic_core_fast32_u(const SrcParams& p, const int32_t* __restrict__ rom32, uint32_t low, int32_t& vs,
                                 int32_t& vc) {
  const uint32_t z0 = (low >> p.z_rshift) << p.z_lshift;
  const int32_t g = (int32_t)p.gain;
  const int32_t K = BIAS ? (1 << 30) : 0;
  int32_t x = g - K, y = g - K;
  int32_t z = (int32_t)(z0 - (uint32_t)rom32[0]);
#pragma unroll
  for (int i = 1; i < NXY; ++i) {
    const int32_t d = (z >> 31) | 1, nd = -d;
    const int32_t ki = BIAS ? (K >> i) : 0;
    const int32_t xs = (x >> i) + ki, ys = (y >> i) + ki;
    x = ys * nd + x;
    y = xs * d + y;
    z = rom32[i] * nd + z;
  }
  vs = (y >> p.out_shift) + (BIAS ? (K >> p.out_shift) : 0);
  vc = (x >> p.out_shift) + (BIAS ? (K >> p.out_shift) : 0);
}

// 64-bit core with the registers left-aligned to bit 63: X = x << (64-w), Z = z << (64-zw).  The
// reference's wrap of every sum to w (zw) bits is then the natural overflow of the 64-bit add, and
// (x >> i) << (64-w) == (X >> i) with the low 64-w bits cleared.  Covers the input-quadrant
// CORDICs (src/cordic_dds48.vhd:170-246, src/cordic_dds_scaled.vhd) and any output-quadrant one
// too wide for the 32-bit cores; same results as cordic_core_generic.
BHW_HD void cordic_core_aligned64(const SrcParams& p, const int64_t* __restrict__ rom64, int q, uint64_t low,
                                  int64_t& vs, int64_t& vc) {
  const int pw = p.pw;
  const bool inq = p.kind == SRC_INQ;
  const int ax = 64 - p.w, az = 64 - p.zw;
  const int64_t mx = (int64_t)(~0ull << ax);
  const int64_t G = (int64_t)((uint64_t)p.gain << ax);
  int64_t X = G, Y = 0, Z;
  if (inq) {
    uint64_t t = low | ((uint64_t)q << (pw - 2));
    if (q == 1) { t = low; X = 0; Y = (int64_t)(0ull - (uint64_t)G); }
    else if (q == 2) { t = low | (3ull << (pw - 2)); X = 0; Y = G; }
    Z = (int64_t)(t << (p.z_lshift + az));
  } else {
    Z = (int64_t)(((low >> p.z_rshift) << p.z_lshift) << az);
  }
  const int n_xy = p.n_xy;
#pragma unroll 2
  for (int i = 0; i < n_xy; ++i) {
    const bool zneg = Z < 0;
    const bool cw = inq ? !zneg : zneg;
    const uint64_t Xs = (uint64_t)((X >> i) & mx), Ys = (uint64_t)((Y >> i) & mx);
    const uint64_t r = (uint64_t)rom64[i];
    X = (int64_t)(cw ? (uint64_t)X + Ys : (uint64_t)X - Ys);
    Y = (int64_t)(cw ? (uint64_t)Y - Xs : (uint64_t)Y + Xs);
    Z = (int64_t)(zneg ? (uint64_t)Z + r : (uint64_t)Z - r);
  }
  vs = Y >> (ax + p.out_shift);
  vc = X >> (ax + p.out_shift);
}

// Half-period pyramid: see bhw_group.cuh.  Work item e (a quarter-wave phase of the top level) writes
// `c` = cos of quadrant 0 and `ns` = cos of quadrant 1 (= -sin), already wrapped and shifted as table
// entries, into every level that contains phase e: level L starts at word 2^(L-1) and holds the first half
// period at L-bit phase resolution.  (vs, vc): the raw pair, for the uint16 quarter-wave image.
constexpr uint32_t kQ16Bias = 1024;   // uint16 image: stored value = cos + bias (the CORDIC's error is << bias)
BHW_HD void pyramid_store(int32_t* H, uint32_t top, uint32_t lmin, uint16_t* q16, uint32_t e, int32_t c, int32_t ns,
                          int32_t vs, int32_t vc) {
  for (uint32_t L = top; L >= lmin; --L) {
    const uint32_t d = top - L;
    if (e & ((1u << d) - 1u)) break;
    const uint32_t i = e >> d;
    H[(1u << (L - 1)) + i] = c;
    H[(1u << (L - 1)) + (1u << (L - 2)) + i] = ns;
  }
  if (q16) {
    const uint32_t Q = 1u << (top - 2);
    q16[e] = (uint16_t)((uint32_t)vc + kQ16Bias);
    q16[Q + e] = (uint16_t)((uint32_t)vs + kQ16Bias);
  }
}

// the four entries one core evaluation yields: cos = c, -s, -c, s in quadrants 0..3
BHW_HD void table_store_quadrants(const TabJob& job, uint32_t e, int64_t vs, int64_t vc) {
  const SrcParams& p = job.sp;
  int32_t* T = job.tab;
  const uint32_t Q = job.entries >> 2;
  const int64_t t = (int64_t)1 << job.tshift;
  const int64_t ns = wrapb(-vs, p.negw), nc = wrapb(-vc, p.negw);
  if (job.pyr_lmin) {
    pyramid_store(T, (uint32_t)p.pw, job.pyr_lmin, job.q16, e, (int32_t)(wrapb(vc, p.outw) * t),
                  (int32_t)(wrapb(ns, p.outw) * t), (int32_t)vs, (int32_t)vc);
    return;
  }
  T[e] = (int32_t)(wrapb(vc, p.outw) * t);          // quadrant 0: cos =  c
  T[e + Q] = (int32_t)(wrapb(ns, p.outw) * t);      // quadrant 1: cos = -s
  T[e + 2 * Q] = (int32_t)(wrapb(nc, p.outw) * t);  // quadrant 2: cos = -c
  T[e + 3 * Q] = (int32_t)(wrapb(vs, p.outw) * t);  // quadrant 3: cos =  s
}

// Work item `e` of a table job -> table entries.
BHW_HD void table_build_item(const TabJob& job, const I2* rom, uint32_t e) {
  const SrcParams& p = job.sp;
  int32_t* T = job.tab;
  if (p.kind == SRC_INQ) {  // one phase per item, no output symmetry
    int64_t s, c;
    if (job.fast == TABCORE_A64) {
      const int q = (int)(e >> (p.pw - 2));
      cordic_core_aligned64(p, job.rom64, q, (uint64_t)(e & ((1u << (p.pw - 2)) - 1u)), s, c);
      c = wrapb(c, p.outw);
    } else {
      eval_source_generic(p, rom, (uint64_t)e, s, c);
    }
    T[e] = (int32_t)(c * ((int64_t)1 << job.tshift));
    return;
  }
  const uint32_t low = e;
  int64_t vs, vc;
  if (p.kind == SRC_TAYLOR) taylor_core_generic(p, rom + job.rom_off, low, vs, vc);
  else if (job.fast == TABCORE_32) { int32_t s32, c32; cordic_core_fast32<false>(p, job.rom32, low, s32, c32); vs = s32; vc = c32; }
  else if (job.fast == TABCORE_32BIAS) { int32_t s32, c32; cordic_core_fast32<true>(p, job.rom32, low, s32, c32); vs = s32; vc = c32; }
  else if (job.fast == TABCORE_A64) cordic_core_aligned64(p, job.rom64, 0, low, vs, vc);
  else cordic_core_generic(p, 0, low, vs, vc);
  table_store_quadrants(job, e, vs, vc);
}

// acc += a*b, signed 32x32 -> 64 with a 64-bit addend: one IMAD.WIDE
BHW_HD void madw(int64_t& acc, int32_t a, int32_t b) {
#if defined(__CUDA_ARCH__)
  asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(acc) : "r"(a), "r"(b));
#else
  acc = (int64_t)((uint64_t)acc + (uint64_t)((int64_t)a * (int64_t)b));
#endif
}
// high word of acc += a*b (mod 2^32): one IMAD
BHW_HD void madhi(int64_t& acc, int32_t a, int32_t b) {
#if defined(__CUDA_ARCH__)
  asm("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmad.lo.s32 hi, %1, %2, hi;\n\tmov.b64 %0, {lo, hi};\n\t}"
      : "+l"(acc) : "r"(a), "r"(b));
#else
  acc = (int64_t)((uint64_t)acc + ((uint64_t)((uint32_t)a * (uint32_t)b) << 32));
#endif
}
// 32x32 multiply-add kept as one IMAD
BHW_HD int32_t mad32(int32_t a, int32_t b, int32_t c) {
#if defined(__CUDA_ARCH__)
  int32_t d;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
#else
  return (int32_t)((uint32_t)a * (uint32_t)b + (uint32_t)c);
#endif
}
// acc += s * v for a 64-bit v and s = +-1: v = vt*2^32 + (int32)vl with vt = vh + (vl >> 31), so the
// product is one signed IMAD.WIDE on the low half plus one IMAD into the high word - no sign-
// dependent select and no carry chain on the alu pipe
BHW_HD void mad64_pm1(int64_t& acc, int64_t v, int32_t s) {
  const uint32_t vl = (uint32_t)(uint64_t)v;
  const int32_t vt = (int32_t)(uint32_t)((uint64_t)v >> 32) + (int32_t)(vl >> 31);
  madw(acc, (int32_t)vl, s);
  madhi(acc, vt, s);
}

// The input-quadrant core with the stage count as a template parameter (cordic_dds48,
// cordic_dds_scaled at DAT_WIDTH 16, 17, 24, 32; src/cordic_dds48.vhd:170-258,
// src/cordic_dds_scaled.vhd:196-283).  X and Y are held as plain (right-aligned) 64-bit integers:
// the gain puts the rotating vector at a quarter of the w-bit register range (|x|, |y| <=
// 2^(w-2) * (1 + eps) at every stage, the CORDIC magnitude only grows towards its final value), so
// the reference's wrap of x and y to w bits never acts and (x >> i) is the arithmetic shift of the
// value itself.  Z stays left-aligned to bit 63 (its wrap is then the natural overflow).  With
// m = z >> 63 and s = 2m + 1 (+1: z >= 0, -1: z < 0; src/cordic_dds48.vhd:234-250):
//   x += s * (y >> i)      y -= s * (x >> i)      z -= s * atan_i
// each as IMAD.WIDE + IMAD (mad64_pm1; the atan word arrives split the same way, `romh`).  Per
// stage: 8 alu instructions (sign mask, s, 2 x (funnel shift, high shift, LEA.HI)) and 7 fma ones,
// against 22 + 6 for the select-based form.
// Two phases per evaluation of z: quadrants 0 and 1 run on t = "00" & low, quadrants 3 and 2 on
// t = "11" & low (src/cordic_dds48.vhd:170-216) - the same z, hence the same rotation directions,
// for two different start vectors.  `hi` = 0: entries of quadrants 0 (a) and 1 (b); 1: quadrants 3
// (a) and 2 (b).  Saves the z update (a third of the fma work of a stage) for every second entry.
template <int NXY>
BHW_HD void cordic_core_inq_u2(const SrcParams& p, const int64_t* __restrict__ rom64, const int32_t* __restrict__ romh,
                               int hi, uint64_t low, int64_t& ca, int64_t& cb) {
  const int pw = p.pw;
  const int az = 64 - p.zw;
  const int64_t G = (int64_t)p.gain;
  int64_t Xa = G, Ya = 0;                    // quadrants 0 / 3
  int64_t Xb = 0, Yb = hi ? G : -G;          // quadrant 2: (0, +G); quadrant 1: (0, -G)
  const uint64_t t = low | (hi ? (3ull << (pw - 2)) : 0ull);
  int64_t Z = (int64_t)(t << (p.z_lshift + az));
#pragma unroll
  for (int i = 0; i < NXY; ++i) {
    const int32_t m = (int32_t)(Z >> 63);
    const int32_t s = 2 * m + 1, ns = mad32(m, -2, -1);
    const int64_t Xsa = Xa >> i, Ysa = Ya >> i, Xsb = Xb >> i, Ysb = Yb >> i;
    mad64_pm1(Xa, Ysa, s);
    mad64_pm1(Ya, Xsa, ns);
    mad64_pm1(Xb, Ysb, s);
    mad64_pm1(Yb, Xsb, ns);
    madw(Z, (int32_t)(uint32_t)(uint64_t)rom64[i], ns);
    madhi(Z, romh[i], ns);
  }
  ca = Xa >> p.out_shift;
  cb = Xb >> p.out_shift;
}

// Work item `e` of the dedicated input-quadrant kernel when it takes two entries per item
// (e < entries / 2): see cordic_core_inq_u2.
template <int NXY>
BHW_HD void table_build_item_inq_u2(const TabJob& job, uint32_t e) {
  const SrcParams& p = job.sp;
  const uint32_t Q = 1u << (p.pw - 2);
  const int hi = (int)(e >> (p.pw - 2));
  const uint32_t low = e & (Q - 1u);
  int64_t ca, cb;
  cordic_core_inq_u2<NXY>(p, job.rom64, job.rom32, hi, (uint64_t)low, ca, cb);
  const int64_t t = (int64_t)1 << job.tshift;
  job.tab[(hi ? 3u * Q : 0u) + low] = (int32_t)(wrapb(ca, p.outw) * t);
  job.tab[(hi ? 2u * Q : Q) + low] = (int32_t)(wrapb(cb, p.outw) * t);
}

// Work item `e` of a job whose core is the 32-bit one with NXY stages (the dedicated kernel for
// large tables); same entries as table_build_item.
template <int NXY, bool BIAS>
BHW_HD void table_build_item_u(const TabJob& job, uint32_t e) {
  int32_t s32, c32;
  cordic_core_fast32_u<NXY, BIAS>(job.sp, job.rom32, e, s32, c32);
  table_store_quadrants(job, e, (int64_t)s32, (int64_t)c32);
}

// ---- direct evaluation through the fast cores ------------------------------------------------
// Which core evaluates a (non-canonical) source in the one-thread-per-sample kernels, with the
// atan words sliced for it; lives in the kernel parameter block (constant bank).
struct SrcCore {
  uint32_t core;        // TABCORE_*
  uint32_t pad;
  int32_t rom32[32];
  int64_t rom64[48];
};

BHW_HD void eval_source_core(const SrcParams& p, const SrcCore& sc, const I2* rom, uint64_t ph, int64_t& s,
                             int64_t& c) {
  const int pw = p.pw;
  ph &= (1ull << pw) - 1;
  const int q = (int)(ph >> (pw - 2));
  const uint64_t low = ph & ((1ull << (pw - 2)) - 1);
  int64_t vs, vc;
  if (p.kind == SRC_TAYLOR) taylor_core_generic(p, rom, (uint32_t)low, vs, vc);
  else if (sc.core == TABCORE_32) { int32_t s32, c32; cordic_core_fast32<false>(p, sc.rom32, (uint32_t)low, s32, c32); vs = s32; vc = c32; }
  else if (sc.core == TABCORE_32BIAS) { int32_t s32, c32; cordic_core_fast32<true>(p, sc.rom32, (uint32_t)low, s32, c32); vs = s32; vc = c32; }
  else if (sc.core == TABCORE_A64) cordic_core_aligned64(p, sc.rom64, q, low, vs, vc);
  else cordic_core_generic(p, q, low, vs, vc);
  if (p.kind != SRC_INQ) quadrant_fix(q, p.negw, vs, vc, vs, vc);
  s = wrapb(vs, p.outw);
  c = wrapb(vc, p.outw);
}

// The four phases low, low + N/4, low + N/2, low + 3N/4 of an output-quadrant source from one
// shift-add evaluation: they share the quarter-wave phase `low`, and the entity's quadrant mux
// (quadrant_fix) turns the one (sin, cos) pair into the four outputs (src/cordic_dds.vhd:232-246,
// src/taylor_sincos.vhd:237-255).  s[r], c[r] belong to phase ph + r*N/4.  Not for SRC_INQ.
BHW_HD void eval_source_core_quad(const SrcParams& p, const SrcCore& sc, const I2* rom, uint64_t ph, int64_t* s,
                                  int64_t* c) {
  const int pw = p.pw;
  ph &= (1ull << pw) - 1;
  const int q = (int)(ph >> (pw - 2));
  const uint64_t low = ph & ((1ull << (pw - 2)) - 1);
  int64_t vs, vc;
  if (p.kind == SRC_TAYLOR) taylor_core_generic(p, rom, (uint32_t)low, vs, vc);
  else if (sc.core == TABCORE_32) { int32_t s32, c32; cordic_core_fast32<false>(p, sc.rom32, (uint32_t)low, s32, c32); vs = s32; vc = c32; }
  else if (sc.core == TABCORE_32BIAS) { int32_t s32, c32; cordic_core_fast32<true>(p, sc.rom32, (uint32_t)low, s32, c32); vs = s32; vc = c32; }
  else if (sc.core == TABCORE_A64) cordic_core_aligned64(p, sc.rom64, q, low, vs, vc);
  else cordic_core_generic(p, q, low, vs, vc);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int64_t so, co;
    quadrant_fix((q + r) & 3, p.negw, vs, vc, so, co);
    s[r] = wrapb(so, p.outw);
    c[r] = wrapb(co, p.outw);
  }
}

// One output sample of BHW_ALGO_DIRECT with per-source cores.
BHW_HD int64_t direct_sample_core(const WinParams& wp, const SrcParams* src, const SrcCore* sc, const I2* rom,
                                  uint64_t n) {
  int64_t cosv[BHW_MAX_TERMS];
  cosv[0] = 0;
  for (int k = 1; k < wp.m; ++k) {
    const TermParams& t = wp.term[k - 1];
    const uint64_t ph = ((uint64_t)t.kmul * n) & t.ph_mask;
    int64_t s;
    eval_source_core(src[t.src], sc[t.src], rom, ph, s, cosv[k]);
  }
  return tail_generic(wp, cosv);
}

// Samples n and n + N/2 of a window from one evaluation per harmonic (whole-window requests of
// k_direct_window).  Half a window later a harmonic's phase has either not moved (even harmonics; the
// second, PHI_WIDTH-1 bit unit of bh_win_3term) or advanced by half the source's period - then the
// quadrant has advanced by two on the same low phase bits and the output mux picks the other value
// of each (v, negated v) pair, which quadrant_fix gives without a second shift-add evaluation.  Bit k
// of `flip` marks the harmonics of the second kind (direct_pair_flip() decides, and excludes the
// input-quadrant CORDICs, which have no such mux).
BHW_HD void direct_sample_core_pair(const WinParams& wp, const SrcParams* src, const SrcCore* sc, const I2* rom,
                                    uint64_t n, uint32_t flip, int64_t& wa, int64_t& wb) {
  int64_t ca[BHW_MAX_TERMS], cb[BHW_MAX_TERMS];
  ca[0] = cb[0] = 0;
  for (int k = 1; k < wp.m; ++k) {
    const TermParams& t = wp.term[k - 1];
    const SrcParams& p = src[t.src];
    const SrcCore& c = sc[t.src];
    const uint64_t ph = ((uint64_t)t.kmul * n) & t.ph_mask;
    const int pw = p.pw;
    const int q = (int)(ph >> (pw - 2));
    const uint64_t low = ph & ((1ull << (pw - 2)) - 1);
    int64_t vs, vc;
    if (p.kind == SRC_TAYLOR) taylor_core_generic(p, rom, (uint32_t)low, vs, vc);
    else if (c.core == TABCORE_32) { int32_t s32, c32; cordic_core_fast32<false>(p, c.rom32, (uint32_t)low, s32, c32); vs = s32; vc = c32; }
    else if (c.core == TABCORE_32BIAS) { int32_t s32, c32; cordic_core_fast32<true>(p, c.rom32, (uint32_t)low, s32, c32); vs = s32; vc = c32; }
    else if (c.core == TABCORE_A64) cordic_core_aligned64(p, c.rom64, q, low, vs, vc);
    else cordic_core_generic(p, q, low, vs, vc);
    int64_t s0, c0, s1, c1;
    if (p.kind != SRC_INQ) {
      quadrant_fix(q, p.negw, vs, vc, s0, c0);
      quadrant_fix((q + 2) & 3, p.negw, vs, vc, s1, c1);
    } else {
      c0 = c1 = vc;            // flip is never set for these
    }
    ca[k] = wrapb(c0, p.outw);
    cb[k] = ((flip >> k) & 1u) ? wrapb(c1, p.outw) : ca[k];
  }
  wa = tail_generic(wp, ca);
  wb = tail_generic(wp, cb);
}

// The four samples n + r*N/4 (r = 0..3) of a window from one evaluation per harmonic: a quarter
// window later harmonic k's phase has advanced by adv_k of its source's own quarter periods on the
// same low phase bits (adv_k = 2 bits at position 2k of `adv`; direct_quad_adv() decides), so
// quadrant_fix at quadrant q + adv_k*r gives its cosine.  Sources with an output quadrant mux only.
BHW_HD void direct_sample_core_quad(const WinParams& wp, const SrcParams* src, const SrcCore* sc, const I2* rom,
                                    uint64_t n, uint32_t adv, int64_t* w) {
  int64_t cq[4][BHW_MAX_TERMS];
#pragma unroll
  for (int r = 0; r < 4; ++r) cq[r][0] = 0;
  for (int k = 1; k < wp.m; ++k) {
    const TermParams& t = wp.term[k - 1];
    const SrcParams& p = src[t.src];
    const SrcCore& c = sc[t.src];
    const uint64_t ph = ((uint64_t)t.kmul * n) & t.ph_mask;
    const int pw = p.pw;
    const int q = (int)(ph >> (pw - 2));
    const uint64_t low = ph & ((1ull << (pw - 2)) - 1);
    int64_t vs, vc;
    if (p.kind == SRC_TAYLOR) taylor_core_generic(p, rom, (uint32_t)low, vs, vc);
    else if (c.core == TABCORE_32) { int32_t s32, c32; cordic_core_fast32<false>(p, c.rom32, (uint32_t)low, s32, c32); vs = s32; vc = c32; }
    else if (c.core == TABCORE_32BIAS) { int32_t s32, c32; cordic_core_fast32<true>(p, c.rom32, (uint32_t)low, s32, c32); vs = s32; vc = c32; }
    else if (c.core == TABCORE_A64) cordic_core_aligned64(p, c.rom64, q, low, vs, vc);
    else cordic_core_generic(p, q, low, vs, vc);
    const int a = (int)((adv >> (2 * k)) & 3u);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int64_t so, co;
      quadrant_fix((q + a * r) & 3, p.negw, vs, vc, so, co);
      cq[r][k] = wrapb(co, p.outw);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) w[r] = tail_generic(wp, cq[r]);
}

// ============================================================================================
// Window synthesis bodies (BHW_ALGO_TABLE, stage 2)
// ============================================================================================
// Per window, resolved for the 32-bit fast tail:
//   phase32_k(n) = n * kstep[k]  (mod 2^32)  - the harmonic's phase left-aligned in 32 bits, so
//                                              the modulo 2^PHI_WIDTH of the RTL counter is free
//   C2_k        = tabp[k][phase32_k >> idx_rsh[k]]   - the table stores cos << tshift
//   RTL  : b_k  = (r>>1)+(r&1), r = (AAk*cos_k)[2DW-2:DW-2]  ==  floor((AAk*cos_k + 2^(DW-2)) / 2^(DW-1))
//               = hi32(A_k*C2_k + 2^31)            with A_k = AAk << (32-DW), tshift = 1
//          S    = AA0 + add + sum_k (-1)^k b_k     (mod 2^32), add = 2 (1 for the 2-term entity):
//                 (pp>>2)+((pp>>1)&1) == (pp+2)>>2, (pp>>1)+(pp&1) == (pp+1)>>1
//          out  = (S << lsh) >> rsh (arithmetic)   lsh = 30-DW (31-DW), rsh = 32-DW: the left shift
//                 wraps S to the DW+2 (DW+1) bits of dsp_pp, the right shift drops the rounded-off
//                 bits and sign-extends the DW-bit DT_WIN in one go.
//   HLS  : m_k  = (a_k*cos_k) >> (NW-2) = hi32(A_k*C2_k), A_k = a_k << (32-NW), tshift = 2;
//          out  = (S << (32-NW)) >> (32-NW).
// `rc` is the low word of the 64-bit addend of the product (RTL 0x80000000, HLS 0).
struct WinRec {
  uint32_t flags;        // WR_*
  uint32_t m;            // terms
  uint32_t dw;
  uint32_t pw;
  int32_t S0;            // AA0 + add
  int32_t lsh, rsh;      // final shifts
  uint32_t rc;           // product rounding addend
  int32_t A[BHW_MAX_TERMS];        // A[k], k = 1..m-1 (A[0] unused)
  int32_t aa[BHW_MAX_TERMS];       // raw AA0.. (64-bit tail)
  uint32_t kstep[BHW_MAX_TERMS];   // kstep[k], k = 1..m-1
  uint32_t idx_rsh[BHW_MAX_TERMS];
  uint32_t n_first;      // n of the window's first flat sample (stream offset folded in)
  uint32_t gen_idx;      // index into the generic-parameter array when WR_GENERIC
  uint32_t tshift;       // left shift of the table entries
  uint32_t pad;
  const int32_t* tabp[BHW_MAX_TERMS];  // tabp[k], k = 1..m-1: trig table of harmonic k
  uint32_t pad2[2];
};
static_assert(sizeof(WinRec) == 320 && sizeof(WinRec) % 16 == 0, "WinRec is copied to shared memory word by word");
enum : uint32_t {
  WR_ACC64 = 2u,    // accumulator / shifts need more than 32 bits: 64-bit tail on table values
  WR_HLS = 4u,      // HLS tail
  WR_RTL2 = 8u,     // 2-term entity
  WR_GENERIC = 16u, // fall back to the generic 64-bit body (direct evaluation)
  WR_HALFTAB = 32u  // tabp[k] is one level of a half-period pyramid (bhw_group.cuh): it holds the first half
                    // period only, the second half is its negation; idx_rsh[k] = 32 - level
};

// hi32(a*b + rc): one IMAD.WIDE with a 64-bit addend, high word taken
BHW_HD int32_t mulhi_rc(int32_t a, int32_t b, uint32_t rc) {
  return (int32_t)(((int64_t)a * (int64_t)b + (int64_t)(uint64_t)rc) >> 32);
}

// fast tail, one sample, 32-bit accumulator (the wraps are intended)
template <int M>
BHW_HD int32_t synth_sample32(const WinRec& r, uint32_t n) {
  uint32_t S = (uint32_t)r.S0;
  const bool halftab = (r.flags & WR_HALFTAB) != 0;
#pragma unroll
  for (int k = 1; k < M; ++k) {
    const uint32_t ph = n * r.kstep[k];
    int32_t c2;
    if (halftab) {
      const int32_t t = r.tabp[k][(ph & 0x7FFFFFFFu) >> r.idx_rsh[k]];
      c2 = (ph >> 31) ? -t : t;     // exact: the table is antisymmetric over half a period
    } else {
      c2 = r.tabp[k][ph >> r.idx_rsh[k]];
    }
    const uint32_t b = (uint32_t)mulhi_rc(r.A[k], c2, r.rc);
    S = (k & 1) ? S - b : S + b;
  }
  return (int32_t)(S << r.lsh) >> r.rsh;
}

// tail for RTL DW 31..32 (and TAYLOR DW 32): 64-bit sum, then the entity's rounding on bit 0 / bit 1.
template <int M>
BHW_HD int32_t synth_sample64(const WinRec& r, uint32_t n) {
  const int dw = (int)r.dw;
  int64_t S = r.aa[0];
#pragma unroll
  for (int k = 1; k < M; ++k) {
    const uint32_t ph = n * r.kstep[k];
    const int32_t c = r.tabp[k][ph >> r.idx_rsh[k]] >> r.tshift;
    const int32_t b = (int32_t)wrapb(((int64_t)r.aa[k] * c + ((int64_t)1 << (dw - 2))) >> (dw - 1), dw);
    S += (k & 1) ? -(int64_t)b : (int64_t)b;
  }
  if (r.flags & WR_RTL2) {
    const int64_t pp = wrapb(S, dw + 1);
    return (int32_t)wrapb((pp >> 1) + (pp & 1), dw);
  }
  const int64_t pp = wrapb(S, dw + 2);
  return (int32_t)wrapb((pp >> 2) + ((pp >> 1) & 1), dw);
}

template <int M>
BHW_HD int32_t synth_sample_m(const WinRec& r, uint32_t n) {
  if (r.flags & WR_ACC64) return synth_sample64<M>(r, n);
  return synth_sample32<M>(r, n);
}

BHW_HD int32_t synth_sample(const WinRec& r, uint32_t n) {
  switch (r.m) {
    case 2: return synth_sample_m<2>(r, n);
    case 3: return synth_sample_m<3>(r, n);
    case 4: return synth_sample_m<4>(r, n);
    case 5: return synth_sample_m<5>(r, n);
    case 6: return synth_sample_m<6>(r, n);     // 6 and 8..11 terms: BHW_WIN_MTERM_* (general kernel only)
    case 7: return synth_sample_m<7>(r, n);
    case 8: return synth_sample_m<8>(r, n);
    case 9: return synth_sample_m<9>(r, n);
    case 10: return synth_sample_m<10>(r, n);
    default: return synth_sample_m<11>(r, n);
  }
}

// ============================================================================================
// Register-resident 32-bit direct evaluation (BHW_ALGO_DIRECT fast path)
// ============================================================================================
// One thread per output sample, every k*phi term evaluated by fully unrolled shift-add stages in
// int32 registers; the atan words arrive pre-sliced for this register width in the kernel
// parameter block (constant bank).  Valid for the output-quadrant CORDICs whose registers fit
// 32 bits without ever wrapping (fast32_ok(): cordic_dds with DW+PRECISION <= 32, the HLS cordic
// with NW <= 30, both DW >= 8) combined with a 32-bit tail (TAILMODE_FAST32).
struct Direct32Params {
  int32_t m, pw;
  int32_t n_xy, n_z;             // x/y stages, z stages
  int32_t z_rshift, z_lshift;    // z0 = ((phase without quadrant bits) >> z_rshift) << z_lshift
  int32_t out_shift;             // cos = x >> out_shift
  int32_t tshift;                // the tail wants cos << tshift (WinRec comment)
  int32_t gain;                  // x0
  int32_t S0, lsh, rsh;          // tail constants (WinRec)
  uint32_t rc;
  uint32_t n_first;              // sample index of output element 0 (stream offset folded in)
  int32_t A[BHW_MAX_TERMS];      // A[k], k = 1..m-1
  uint32_t kmul[BHW_MAX_TERMS];  // phase step of harmonic k
  int32_t rom[32];               // atan word of stage i, already shifted/masked for this width
};

// cos of a pw-bit phase by the unrolled CORDIC (src/cordic_dds.vhd:170-246 with the quadrant fix;
// identical structure in hls/windows/win_function.cpp:86-154).  NXY = number of x/y stages, a
// compile-time constant so that every shift is an immediate and the stages are straight-line code;
// NXY == 0 selects the run-time loop (any stage count).  The z update of a stage past n_z uses a
// zero atan word (HLS: the last stage reads no table entry) - z is dead after the last stage.
template <int NXY>
BHW_HD void direct32_sc(const Direct32Params& p, uint32_t ph, int32_t& vs, int32_t& vc) {
  const int pw = p.pw;
  const uint32_t low = ph & ((1u << (pw - 2)) - 1u);
  int32_t z = (int32_t)((low >> p.z_rshift) << p.z_lshift);
  int32_t x = p.gain, y = p.gain;  // stage 0 with z0 >= 0: x - (0>>0), 0 + (x>>0)
  z -= p.rom[0];
  if (NXY > 0) {
#pragma unroll
    for (int i = 1; i < (NXY > 0 ? NXY : 1); ++i) {
      // d = -1 when z < 0, else +1; only multiply-ADDs so that each update is one IMAD
      const int32_t d = (z >> 31) | 1, nd = -d;
      const int32_t xs = x >> i, ys = y >> i;
      x = ys * nd + x;               // z<0: x + (y>>i)   (src/cordic_dds.vhd:199-205)
      y = xs * d + y;                // z<0: y - (x>>i)
      z = p.rom[i] * nd + z;
    }
  } else {
    const int n_xy = p.n_xy;
#pragma unroll 4
    for (int i = 1; i < n_xy; ++i) {
      const int32_t d = (z >> 31) | 1;
      const int32_t xs = x >> i, ys = y >> i;
      x -= d * ys;
      y += d * xs;
      z -= d * p.rom[i];
    }
  }
  vc = x >> p.out_shift;
  vs = y >> p.out_shift;
}
// quadrant fix for the cosine output: c, -s, -c, s.  |vc|,|vs| <= 2^(DW-2)+eps, so the
// reference's wrapped negation is the plain one
BHW_HD int32_t direct32_fix(uint32_t q, int32_t vs, int32_t vc) {
  const int32_t v = (q & 1u) ? vs : vc;
  return ((q + 1u) & 2u) ? -v : v;
}
template <int NXY>
BHW_HD int32_t direct32_cos(const Direct32Params& p, uint32_t ph) {
  int32_t vs, vc;
  direct32_sc<NXY>(p, ph, vs, vc);
  return direct32_fix(ph >> (p.pw - 2), vs, vc);
}

// one output sample; the harmonics are a run-time loop (m is a kernel parameter)
template <int NXY>
BHW_HD int32_t direct32_sample(const Direct32Params& p, uint32_t n) {
  const uint32_t pmask = (1u << p.pw) - 1u;
  uint32_t S = (uint32_t)p.S0;
  for (int k = 1; k < p.m; ++k) {
    const int32_t c = direct32_cos<NXY>(p, (p.kmul[k] * n) & pmask);
    const uint32_t b = (uint32_t)mulhi_rc(p.A[k], c << p.tshift, p.rc);
    S = (k & 1) ? S - b : S + b;
  }
  return (int32_t)(S << p.lsh) >> p.rsh;
}

// Samples n and n + N/2 from one set of CORDIC evaluations: half a window later the phase of an
// odd harmonic has its quadrant advanced by two on the same low bits, so the output mux above picks
// the negated value (direct32_cos: -v instead of v); the phase of an even harmonic is unchanged.
template <int NXY>
BHW_HD void direct32_pair(const Direct32Params& p, uint32_t n, int32_t& wa, int32_t& wb) {
  const uint32_t pmask = (1u << p.pw) - 1u;
  uint32_t Sa = (uint32_t)p.S0, Sb = Sa;
  for (int k = 1; k < p.m; ++k) {
    const uint32_t km = p.kmul[k];
    const int32_t c = direct32_cos<NXY>(p, (km * n) & pmask);
    const uint32_t ba = (uint32_t)mulhi_rc(p.A[k], c << p.tshift, p.rc);
    const uint32_t bb = (km & 1u) ? (uint32_t)mulhi_rc(p.A[k], (-c) << p.tshift, p.rc) : ba;
    Sa = (k & 1) ? Sa - ba : Sa + ba;
    Sb = (k & 1) ? Sb - bb : Sb + bb;
  }
  wa = (int32_t)(Sa << p.lsh) >> p.rsh;
  wb = (int32_t)(Sb << p.lsh) >> p.rsh;
}

// Samples n + r*N/4, r = 0..3, from one set of CORDIC evaluations: a quarter window later the phase
// of harmonic k has advanced by k quarter periods on the same low bits, so its cosine is the
// quadrant fix of the same (sin, cos) pair at quadrant q + k*r.
template <int NXY>
BHW_HD void direct32_quad(const Direct32Params& p, uint32_t n, int32_t* w) {
  const uint32_t pmask = (1u << p.pw) - 1u;
  uint32_t S[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) S[r] = (uint32_t)p.S0;
  for (int k = 1; k < p.m; ++k) {
    const uint32_t km = p.kmul[k];
    const uint32_t ph = (km * n) & pmask;
    const uint32_t q = ph >> (p.pw - 2);
    int32_t vs, vc;
    direct32_sc<NXY>(p, ph, vs, vc);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int32_t c = direct32_fix((q + km * (uint32_t)r) & 3u, vs, vc);
      const uint32_t b = (uint32_t)mulhi_rc(p.A[k], c << p.tshift, p.rc);
      S[r] = (k & 1) ? S[r] - b : S[r] + b;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) w[r] = (int32_t)(S[r] << p.lsh) >> p.rsh;
}

// ============================================================================================
// Register-resident 32-bit direct evaluation, TAYLOR source (k_direct_taylor)
// ============================================================================================
// taylor_sincos + tay1_order in int32 registers with 32x32->64 products (IMAD.WIDE): valid for every
// legal TAYLOR window (DAT_WIDTH <= 32); same integers as taylor_core_generic.  A Taylor evaluation
// is a ROM look-up and two multiplies, so for long windows evaluating it per sample is cheaper than
// building, storing and re-reading a table as large as the window itself.
struct TayUnit {
  int32_t pw;       // phase width of the unit (PHI_WIDTH, or PHI_WIDTH-1 for the 2nd harmonic of bh_win_3term)
  int32_t mode;     // TayMode
  int32_t ashift;   // ROM address shift
  int32_t cbits;    // bits of acnt
  int32_t xs;       // XSHIFT = 19 + LUT_SIZE
  int32_t pi;       // ramb_pi = round(pi * 2^(17-STAGE))
};
struct DirectTayParams {
  int32_t m, dw;
  int32_t tshift;                // the tail wants cos << tshift (WinRec comment)
  int32_t S0, lsh, rsh;          // tail constants (WinRec)
  uint32_t rc;
  uint32_t n_first;              // sample index of output element 0 (stream offset folded in)
  int32_t A[4];                  // A[k], k = 1..m-1
  TayUnit unit[2];               // unit[k-1] feeds harmonic k
  uint32_t rom_entries;          // 2^LUT_SIZE (cos, sin) words; the kernel keeps them in shared memory
  uint32_t tmode;                // TMODE_*
};

// TMODE: the datapath both units of the window sit in (they share DAT_WIDTH, and bh_win_3term
// rejects the one LUT_SIZE that would split them): 0 ROM only (TAY_LESS / TAY_EQ per unit),
// 1 TAY_DSP, 2 TAY_WIDE.
enum : int { TMODE_ROM = 0, TMODE_DSP = 1, TMODE_WIDE = 2 };

template <int TMODE>
BHW_HD void taylor_core_fast32(const TayUnit& u, int dw, const I2* __restrict__ rom, uint32_t t, int32_t& vs,
                               int32_t& vc) {
  if (TMODE == TMODE_ROM) {  // ROM only (src/taylor_sincos.vhd:157-167)
    const I2 e = rom[u.mode == TAY_LESS ? (t << u.ashift) : t];
    vc = e.x; vs = e.y;
    return;
  }
  const I2 e = rom[t >> u.ashift];
  const int32_t c0 = e.x, s0 = e.y;
  const uint32_t acnt = t & ((1u << u.cbits) - 1u);
  const int32_t mpi = (int32_t)(((uint32_t)u.pi * acnt) & 0xFFFFFFu);  // 24-bit ROM word (src/tay1_order.vhd:136-147)
  const int xs = u.xs;
  if (TMODE == TMODE_DSP) {
    // P = C -/+ A*B, result = P[xs+dw-1 : xs] (src/tay1_order.vhd:245,316,501-502)
    const int64_t pc = (int64_t)((uint64_t)(int64_t)c0 << xs) - (int64_t)mpi * (int64_t)s0;
    const int64_t ps = (int64_t)((uint64_t)(int64_t)s0 << xs) + (int64_t)mpi * (int64_t)c0;
    vc = wrapb32((int32_t)(pc >> xs), dw);
    vs = wrapb32((int32_t)(ps >> xs), dw);
  } else {
    // 0 <= s0, c0 < 2^(dw-1) and 0 <= mpi < 2^20 <= 2^xs, so 0 <= m1, m2 < 2^(dw-1): the DW-bit
    // wraps of :585-586 and of c0 - m1 are no-ops, and s0 + m2 < 2^dw goes negative in DW bits
    // exactly when it exceeds 2^(dw-1) - 1 (unsigned compare; also right for dw = 32)
    const int32_t m1 = (int32_t)(((int64_t)s0 * (int64_t)mpi) >> xs);               // :585
    const int32_t m2 = (int32_t)(((int64_t)c0 * (int64_t)mpi) >> xs);               // :586
    const int32_t cp = c0 - m1;                                                     // :595
    const uint32_t sp = (uint32_t)s0 + (uint32_t)m2;                                // :596
    const int32_t sat = (int32_t)((1u << (dw - 1)) - 1u);
    vc = cp < 0 ? sat : cp;                                                         // negative -> max positive (:602-617)
    vs = sp > (uint32_t)sat ? sat : (int32_t)sp;
  }
}

// one output sample: every harmonic through its own taylor_sincos unit (each unit owns a +1
// phase counter, src/bh_win_3term.vhd:221-233), output-side quadrant fix (src/taylor_sincos.vhd:237-255)
template <int TMODE>
BHW_HD int32_t direct_taylor_sample(const DirectTayParams& p, const I2* __restrict__ rom, uint32_t n) {
  uint32_t S = (uint32_t)p.S0;
#pragma unroll
  for (int k = 1; k < 3; ++k) {
    if (k < p.m) {
      const TayUnit& u = p.unit[k - 1];
      const uint32_t ph = n & ((1u << u.pw) - 1u);
      const uint32_t q = ph >> (u.pw - 2);
      int32_t vs, vc;
      taylor_core_fast32<TMODE>(u, p.dw, rom, ph & ((1u << (u.pw - 2)) - 1u), vs, vc);
      const int32_t v = (q & 1u) ? vs : vc;                                       // cos: c, -s, -c, s
      // ROM-only and TAY_WIDE values lie in [0, 2^(dw-1)-1]: the DW-bit negation cannot wrap
      const int32_t nv = TMODE == TMODE_DSP ? wrapb32((int32_t)(0u - (uint32_t)v), p.dw) : -v;
      const int32_t c = ((q + 1u) & 2u) ? nv : v;
      const uint32_t b = (uint32_t)mulhi_rc(p.A[k], (int32_t)((uint32_t)c << p.tshift), p.rc);
      S = (k & 1) ? S - b : S + b;
    }
  }
  return (int32_t)(S << p.lsh) >> p.rsh;
}

// Samples n and n + N/2 from one evaluation of the units.  Half a window later the first unit's
// phase has its quadrant advanced by two and the same low bits: the entity's quadrant mux then picks
// the other one of (v, -v) (src/taylor_sincos.vhd:237-255) - nothing is assumed about v, so this holds
// for the wrapping DSP48 branch too.  The second unit of bh_win_3term counts PHI_WIDTH-1 bits
// (src/bh_win_3term.vhd:221-233): its phase, and with it b_2, is the same for both samples.
template <int TMODE>
BHW_HD void direct_taylor_pair(const DirectTayParams& p, const I2* __restrict__ rom, uint32_t n, int32_t& wa, int32_t& wb) {
  uint32_t Sa = (uint32_t)p.S0, Sb = Sa;
  {
    const TayUnit& u = p.unit[0];
    const uint32_t ph = n & ((1u << u.pw) - 1u);
    const uint32_t q = ph >> (u.pw - 2);
    int32_t vs, vc;
    taylor_core_fast32<TMODE>(u, p.dw, rom, ph & ((1u << (u.pw - 2)) - 1u), vs, vc);
    const int32_t v = (q & 1u) ? vs : vc;
    const int32_t nv = TMODE == TMODE_DSP ? wrapb32((int32_t)(0u - (uint32_t)v), p.dw) : -v;
    const bool neg = ((q + 1u) & 2u) != 0;
    const int32_t ca = neg ? nv : v, cb = neg ? v : nv;
    Sa -= (uint32_t)mulhi_rc(p.A[1], (int32_t)((uint32_t)ca << p.tshift), p.rc);
    Sb -= (uint32_t)mulhi_rc(p.A[1], (int32_t)((uint32_t)cb << p.tshift), p.rc);
  }
  if (p.m > 2) {
    const TayUnit& u = p.unit[1];
    const uint32_t ph = n & ((1u << u.pw) - 1u);
    const uint32_t q = ph >> (u.pw - 2);
    int32_t vs, vc;
    taylor_core_fast32<TMODE>(u, p.dw, rom, ph & ((1u << (u.pw - 2)) - 1u), vs, vc);
    const int32_t v = (q & 1u) ? vs : vc;
    const int32_t nv = TMODE == TMODE_DSP ? wrapb32((int32_t)(0u - (uint32_t)v), p.dw) : -v;
    const int32_t c = ((q + 1u) & 2u) ? nv : v;
    const uint32_t b = (uint32_t)mulhi_rc(p.A[2], (int32_t)((uint32_t)c << p.tshift), p.rc);
    Sa += b;
    Sb += b;
  }
  wa = (int32_t)(Sa << p.lsh) >> p.rsh;
  wb = (int32_t)(Sb << p.lsh) >> p.rsh;
}

// Sixteen samples from two ROM words: four consecutive samples n0 .. n0+3 of the first quarter window
// (n0 a multiple of 4, stream offset 0) and their partners a quarter, a half and three quarters of a window
// later, TAY_WIDE datapath.  What is shared:
//   * the four samples address one ROM word per unit (its counter bits below the ROM address are >= 2 wide)
//     and sit in one quadrant, so the word, the quadrant mux and the product rounding constants are loaded
//     once and pi*acnt advances by `pi` per sample (0 <= pi*acnt < 2^20: the 24-bit ROM word never wraps);
//   * a quarter window later the first unit's quadrant advances by one: cos = vc, -vs, -vc, vs
//     (src/taylor_sincos.vhd:237-255); TAY_WIDE values lie in [0, 2^(dw-1)-1] so the DW-bit negation is plain;
//   * the product of a negated value is taken from the same 64-bit product with the other rounding addend:
//     b(-P) = -hi32(P + rcn), rcn = 2^32 - 1 - rc (BankShape comment);
//   * the second unit of bh_win_3term counts PHI_WIDTH-1 bits (src/bh_win_3term.vhd:221-233): a quarter
//     window is half its period, so its value alternates v, -v, v, -v over the four partners.
// w[r][e] = sample n0 + e + r*N/4.  Same integers as direct_taylor_sample (tests/hostcheck).
BHW_HD void taylor_wide4(const TayUnit& u, int dw, const I2* __restrict__ rom, uint32_t t0, int32_t* vs, int32_t* vc) {
  const I2 w = rom[t0 >> u.ashift];
  const int32_t c0 = w.x, s0 = w.y;
  const int32_t sat = (int32_t)((1u << (dw - 1)) - 1u);
  const int xs = u.xs;
  int32_t mpi = u.pi * (int32_t)(t0 & ((1u << u.cbits) - 1u));
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int32_t m1 = (int32_t)(((int64_t)s0 * (int64_t)mpi) >> xs);               // src/tay1_order.vhd:585
    const int32_t m2 = (int32_t)(((int64_t)c0 * (int64_t)mpi) >> xs);               // :586
    const uint32_t cp = (uint32_t)(c0 - m1), sp = (uint32_t)s0 + (uint32_t)m2;      // :595-596
    vc[e] = (int32_t)(cp < (uint32_t)sat ? cp : (uint32_t)sat);                     // negative (huge unsigned) -> max positive
    vs[e] = (int32_t)(sp < (uint32_t)sat ? sp : (uint32_t)sat);
    mpi += u.pi;
  }
}

BHW_HD void direct_taylor_quad4(const DirectTayParams& p, const I2* __restrict__ rom, uint32_t n0, int32_t (*w)[4]) {
  const int64_t rc = (int64_t)(uint64_t)p.rc, rcn = (int64_t)(uint64_t)(0xFFFFFFFFu - p.rc);
  uint32_t S[4][4];
  {
    int32_t vs[4], vc[4];
    taylor_wide4(p.unit[0], p.dw, rom, n0, vs, vc);                                // n0 < N/4: quadrant 0, low bits = n0
    const int64_t A1 = p.A[1];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t Pc = A1 * (int64_t)(int32_t)((uint32_t)vc[e] << p.tshift);
      const int64_t Ps = A1 * (int64_t)(int32_t)((uint32_t)vs[e] << p.tshift);
      S[0][e] = (uint32_t)p.S0 - (uint32_t)((Pc + rc) >> 32);                      // - b( vc)
      S[1][e] = (uint32_t)p.S0 + (uint32_t)((Ps + rcn) >> 32);                     // - b(-vs)
      S[2][e] = (uint32_t)p.S0 + (uint32_t)((Pc + rcn) >> 32);                     // - b(-vc)
      S[3][e] = (uint32_t)p.S0 - (uint32_t)((Ps + rc) >> 32);                      // - b( vs)
    }
  }
  if (p.m > 2) {
    const TayUnit& u = p.unit[1];
    const uint32_t q1 = n0 >> (u.pw - 2);                                          // 0 or 1: n0 < N/4 = half its period
    int32_t vs[4], vc[4];
    taylor_wide4(u, p.dw, rom, n0 & ((1u << (u.pw - 2)) - 1u), vs, vc);
    const int64_t A2 = p.A[2];
    const int64_t r_even = q1 ? rcn : rc, r_odd = q1 ? rc : rcn;                   // quadrant 1: cos = -vs
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t P = A2 * (int64_t)(int32_t)((uint32_t)(q1 ? vs[e] : vc[e]) << p.tshift);
      const uint32_t he = (uint32_t)((P + r_even) >> 32), ho = (uint32_t)((P + r_odd) >> 32);
      const uint32_t be = q1 ? 0u - he : he, bo = q1 ? ho : 0u - ho;              // b(cos) at even / odd partners
      S[0][e] += be; S[1][e] += bo; S[2][e] += be; S[3][e] += bo;
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int e = 0; e < 4; ++e) w[r][e] = (int32_t)(S[r][e] << p.lsh) >> p.rsh;
}

// ============================================================================================
// cordic_atan2 (src/cordic_atan2.vhd:80-220)
// ============================================================================================
// Vectoring CORDIC: W = ANGLE_WIDTH + PRECISION bit registers that DO wrap (a full-scale input
// grows past W bits at PRECISION 1), so the registers are kept left-aligned to bit 63 like
// cordic_core_aligned64: the wrap is the natural overflow of the 64-bit add.
struct Atan2Params {
  int32_t iw, aw, w;     // INPUT_WIDTH, ANGLE_WIDTH, ANGLE_WIDTH + PRECISION
  int32_t fast32;        // 1: W <= 32, the 32-bit body applies
  int32_t skew;          // 1: PHI_DT of pair t takes the quadrant of pair t+1 (bhw_atan2_desc::stream_quadrant)
  int32_t pad;
  int64_t rom64[48];     // ROM_TABLE(ii) (:97-108) left-aligned to bit 63; aw-1 entries used
  uint32_t rom32[32];    // the same words left-aligned to bit 31 (fast32)
};

// W <= 32: the same algorithm with the registers left-aligned in 32 bits.  d = +1/-1 from the sign
// of Y turns the three conditional add/subtracts into multiply-adds (mod 2^32 = the W-bit wrap).
// AW > 0: ANGLE_WIDTH as a compile-time constant (stage count, shifts and masks become immediates);
// AW == 0: run-time loop.
// (qx, qy): the pair whose sign bits select the output quadrant - the pair itself, or (Atan2Params::skew) the NEXT
// pair of the stream, which is what the entity as written does: its quadrant shift registers are one stage shorter
// than the data path (src/cordic_atan2.vhd:127-129 vs :136-184; found by executing the VHDL, oracle/vhdl_sim.py).
template <int AW>
BHW_HD int32_t atan2_sample32_t(const Atan2Params& p, int32_t xin, int32_t yin, int32_t qx, int32_t qy) {
  const int aw = AW > 0 ? AW : p.aw, ax = 32 - p.w;
  const uint32_t sxb = ((uint32_t)xin >> (p.iw - 1)) & 1u, syb = ((uint32_t)yin >> (p.iw - 1)) & 1u;
  const uint32_t qxb = ((uint32_t)qx >> (p.iw - 1)) & 1u, qyb = ((uint32_t)qy >> (p.iw - 1)) & 1u;
  const uint32_t mag_mask = (1u << (aw - 1)) - 1u;
  const int32_t mx = (int32_t)(~0u << ax);
  uint32_t X = (((uint32_t)xin ^ (0u - sxb)) & mag_mask) << ax;
  uint32_t Y = (((uint32_t)yin ^ (0u - syb)) & mag_mask) << ax;
  uint32_t Z = 0;
#pragma unroll
  for (int i = 0; i <= (AW > 0 ? AW - 2 : -1); ++i) {
    const uint32_t d = (uint32_t)(((int32_t)Y >> 31) | 1), nd = 0u - d;   // -1 when y < 0, else +1
    const uint32_t Xs = (uint32_t)(((int32_t)X >> i) & mx), Ys = (uint32_t)(((int32_t)Y >> i) & mx);
    X = Ys * d + X;                                           // y >= 0: x + (y >> i)   (:169-175)
    Y = Xs * nd + Y;                                          //         y - (x >> i)
    Z = p.rom32[i] * nd + Z;                                  //         z - ROM_TABLE(i)
  }
  if (AW == 0) {
#pragma unroll 4
    for (int i = 0; i <= aw - 2; ++i) {
      const uint32_t d = (uint32_t)(((int32_t)Y >> 31) | 1);
      const uint32_t Xs = (uint32_t)(((int32_t)X >> i) & mx), Ys = (uint32_t)(((int32_t)Y >> i) & mx);
      X += d * Ys;
      Y -= d * Xs;
      Z -= d * p.rom32[i];
    }
  }
  const int32_t phi = (int32_t)Z >> (32 - aw);                // top ANGLE_WIDTH bits, sign-extended
  const int32_t pi = 1 << (aw - 2);
  int32_t o;
  switch ((qxb << 1) | qyb) {
    case 0: o = phi; break;
    case 1: o = (int32_t)((uint32_t)phi + (uint32_t)pi); break;
    case 2: o = (int32_t)(0u - (uint32_t)phi); break;
    default: o = (int32_t)((uint32_t)phi - (uint32_t)pi); break;
  }
  return wrapb32(o, aw);
}
template <int AW>
BHW_HD int32_t atan2_sample32_t(const Atan2Params& p, int32_t xin, int32_t yin) { return atan2_sample32_t<AW>(p, xin, yin, xin, yin); }
BHW_HD int32_t atan2_sample32(const Atan2Params& p, int32_t xin, int32_t yin, int32_t qx, int32_t qy) {
  return atan2_sample32_t<0>(p, xin, yin, qx, qy);
}

BHW_HD int32_t atan2_sample(const Atan2Params& p, int32_t xin, int32_t yin, int32_t qx, int32_t qy) {
  if (p.fast32) return atan2_sample32(p, xin, yin, qx, qy);
  const int aw = p.aw, ax = 64 - p.w;
  const uint32_t sxb = ((uint32_t)xin >> (p.iw - 1)) & 1u, syb = ((uint32_t)yin >> (p.iw - 1)) & 1u;
  const uint32_t qxb = ((uint32_t)qx >> (p.iw - 1)) & 1u, qyb = ((uint32_t)qy >> (p.iw - 1)) & 1u;
  // init_x(ii) = VEC_DX(ii) xor VEC_DX(INPUT_WIDTH-1), ii = 0..ANGLE_WIDTH-2; upper bits zero (:136-146)
  const uint32_t mag_mask = aw - 1 >= 32 ? 0xFFFFFFFFu : ((1u << (aw - 1)) - 1u);
  const uint64_t ix = ((uint32_t)xin ^ (0u - sxb)) & mag_mask, iy = ((uint32_t)yin ^ (0u - syb)) & mag_mask;
  const int64_t mx = (int64_t)(~0ull << ax);
  uint64_t X = ix << ax, Y = iy << ax, Z = 0;
  for (int i = 0; i <= aw - 2; ++i) {                     // lpXY / lpZ (:166-184)
    const bool yneg = (int64_t)Y < 0;
    const uint64_t Xs = (uint64_t)(((int64_t)X >> i) & mx), Ys = (uint64_t)(((int64_t)Y >> i) & mx);
    const uint64_t r = (uint64_t)p.rom64[i];
    X = yneg ? X - Ys : X + Ys;
    Y = yneg ? Y + Xs : Y - Xs;
    Z = yneg ? Z + r : Z - r;
  }
  // dat_phi = sigZ(W-1 downto PRECISION): the top ANGLE_WIDTH bits (:188)
  const int64_t phi = (int64_t)Z >> (64 - aw);
  const int64_t pi = (int64_t)1 << (aw - 2);              // PHI_PI: only bit ANGLE_WIDTH-2 set (:121)
  int64_t o;
  switch ((qxb << 1) | qyb) {                             // quadrant = sign(X) & sign(Y) (:129-131,203-208)
    case 0: o = phi; break;
    case 1: o = phi + pi; break;
    case 2: o = -phi; break;
    default: o = phi - pi; break;
  }
  return (int32_t)wrapb(o, aw);
}
BHW_HD int32_t atan2_sample(const Atan2Params& p, int32_t xin, int32_t yin) { return atan2_sample(p, xin, yin, xin, yin); }

// ============================================================================================
// Bank synthesis body (k_synth_bank): whole windows of one shape
// ============================================================================================
// A "bank" is a run of windows that differ only in their AAk ports (and stream offset): same
// entity, PHI_WIDTH, DAT_WIDTH and sin/cos source, hence the same trig tables and the same
// constants below.  The kernel stages the tables in shared memory when they fit and gives every
// lane *pairs* of samples (n, n + N/2): for the output-quadrant sources the table is exactly
// antisymmetric over half a period, so cos_k(n + N/2) = (-1)^k cos_k(n) and one look-up serves
// both samples - odd harmonics with the product negated, even harmonics unchanged:
//   b(-P) = floor((-P + rc) / 2^32) = -hi32(P + rcn),  rcn = 2^32 - 1 - rc.
// When even half a period does not fit, only the first half-period is staged (TAB_SMEM_HALF) and
// the sign of a look-up in the second half moves into the coefficient: (-A) * C2 == A * (-C2).
struct BankShape {
  uint32_t m, pw;
  uint32_t rc, rcn;
  int32_t lsh, rsh;
  // 64-bit tail (RTL DAT_WIDTH 31..32): b = (int32)((AAk*C2 + prnd) >> psh), S in 64 bits,
  // out = sx(((S + fadd) >> fin), DW); A[k] are then the raw AAk and S0 the raw AA0
  uint32_t acc64, psh, prnd, prndn, fin, fadd, flsh;
  uint32_t pair_adj;                    // pairing through the ones'-complement relation of the input-quadrant CORDICs:
                                        // T[i + E/2] == -T[i] - pair_adj (pair_adj = 1 << tshift; 0: exact antisymmetry)
  uint32_t ntab;                        // distinct tables (1, or 2 for 3-term TAYLOR)
  uint32_t smem_words;                  // staged words in total
  uint32_t kstep[BHW_MAX_TERMS];
  uint32_t idx_rsh[BHW_MAX_TERMS];      // full-period index = phase32 >> idx_rsh
  uint32_t tsel[BHW_MAX_TERMS];         // which distinct table harmonic k reads
  uint32_t toff[2];                     // word offset of distinct table u in the staged copy
  uint32_t tentries[2];                 // entries of distinct table u (full period)
  const int32_t* tab[2];                // distinct table u in global memory
  // linear indexing: when no phase bit is dropped, index_k(n) = (n * lin_step[k]) mod entries, so
  // inside a tile that crosses no boundary of the staged domain the index is affine in (lane, j)
  uint32_t lin;                         // 1: every harmonic has an integer index step
  uint32_t lin_step[BHW_MAX_TERMS];     // k for the CORDIC entities, 1 for a TAYLOR unit
  uint32_t lin_dmask[BHW_MAX_TERMS];    // staged domain size - 1 (half a period with TAB_SMEM_HALF)
  uint32_t lin_dbit[BHW_MAX_TERMS];     // TAB_SMEM_HALF: the index bit that selects the negated half
};
enum : int { TAB_SMEM_FULL = 0, TAB_SMEM_HALF = 1, TAB_GLOBAL = 2 };
constexpr int kBankJ = 8;               // samples per lane and tile (per half when paired)
constexpr int kBankTile = 32 * kBankJ;  // samples per warp tile (per half when paired)
constexpr int kBankTileLog2 = 8;

// Tile walks of the bank kernel over tables that stay in L2 (bhw_kernels.cu, k_synth_bank).
// Spread walk of one window of U tiles: warp j of G owns tiles [U*j/G, U*(j+1)/G) and takes the i-th
// of them at step i; the steps 0 .. spread_steps()-1 are split over the CTAs of the grid.
BHW_HD uint32_t spread_steps(uint32_t U, uint32_t G) { return (U + G - 1) / G; }
BHW_HD bool spread_tile(uint32_t U, uint32_t G, uint32_t warp, uint32_t step, uint32_t* tile) {
  const uint32_t b0 = (uint32_t)((uint64_t)U * warp / G), b1 = (uint32_t)((uint64_t)U * (warp + 1) / G);
  *tile = b0 + step;
  return warp < G && b0 + step < b1;
}
// Window-minor walk of a bank of nwin windows: unit u is tile u / nwin of window u % nwin.
BHW_HD void win_minor_unit(uint32_t u, uint32_t nwin, uint32_t* w, uint32_t* tile) {
  *tile = u / nwin;
  *w = u - *tile * nwin;
}

// Do all samples of the tile starting at sample nbase see, for every harmonic, a phase in one
// and the same half-period?  (TAB_SMEM_HALF only; needs 127*kstep < 2^31, guaranteed by the host.)
template <int M>
BHW_HD bool bank_tile_sign_uniform(const BankShape& sh, uint32_t nbase) {
  uint32_t diff = 0;
#pragma unroll
  for (int k = 1; k < M; ++k) {
    const uint32_t first = nbase * sh.kstep[k];
    diff |= first ^ (first + (uint32_t)(kBankTile - 1) * sh.kstep[k]);
  }
  return (diff >> 31) == 0;
}

template <bool W64> struct BankAcc { typedef uint32_t type; };
template <> struct BankAcc<true> { typedef uint64_t type; };

// b_k from the 64-bit product P = A*C2; `negated`: the value v with b(-P) == -v
template <bool W64>
BHW_HD typename BankAcc<W64>::type bank_term(const BankShape& sh, int64_t P, bool negated) {
  if (W64) return (typename BankAcc<W64>::type)(int64_t)(int32_t)((P + (int64_t)(negated ? sh.prndn : sh.prnd)) >> sh.psh);
  return (typename BankAcc<W64>::type)(uint32_t)((P + (int64_t)(uint64_t)(negated ? sh.rcn : sh.rc)) >> 32);
}
template <bool W64>
BHW_HD typename BankAcc<W64>::type bank_init(const BankShape& sh, int32_t S0) {
  if (W64) return (typename BankAcc<W64>::type)((int64_t)S0 + (int64_t)sh.fadd);
  return (typename BankAcc<W64>::type)(uint32_t)S0;
}
template <bool W64>
BHW_HD int32_t bank_finish(const BankShape& sh, typename BankAcc<W64>::type S) {
  if (W64) {
    const int32_t t = (int32_t)((int64_t)S >> sh.fin);
    return (int32_t)((uint32_t)t << sh.flsh) >> sh.flsh;
  }
  return (int32_t)((uint32_t)S << sh.lsh) >> sh.rsh;
}

// Linear tiles.  Returns true when, for every harmonic, the 128 samples starting at nbase stay
// inside one staged domain (no wrap of the table index, no sign change); base[k] is then the
// index of sample nbase and bit k of *neg says whether the harmonic sits in the negated half.
template <int M, int TAB>
BHW_HD bool bank_tile_linear(const BankShape& sh, uint32_t nbase, uint32_t* base, uint32_t* neg) {
  bool ok = true;
  uint32_t ng = 0;
#pragma unroll
  for (int k = 1; k < M; ++k) {
    const uint32_t pos = nbase * sh.lin_step[k];
    const uint32_t p = pos & sh.lin_dmask[k];
    ok = ok && (p + (uint32_t)(kBankTile - 1) * sh.lin_step[k] <= sh.lin_dmask[k]);
    base[k] = p;
    if (TAB == TAB_SMEM_HALF) ng |= (pos & sh.lin_dbit[k]) ? (1u << k) : 0u;
  }
  *neg = ng;
  return ok;
}

// One lane's share of a linear tile: the look-ups of harmonic k are T[base + step*(lane + 32*j)].
// PAIR: 0 = single samples; 1 = pairs (n, n + N/2) over an exactly antisymmetric table; 2 = pairs over an
// input-quadrant CORDIC's table, where T[i + E/2] == ~T[i] (the ones' complement: the entity negates a wide
// register and then floors, src/cordic_dds48.vhd:196-216,257-258) except at a few data-dependent entries that a
// patch pass recomputes (k_inq_patch): the partner's product is A * -(C2 + pair_adj).
template <int M, int TAB, int PAIR, bool W64>
BHW_HD void bank_lane_tile_lin(const BankShape& sh, const int32_t* A, int32_t S0, const int32_t* const* tabs,
                               uint32_t lane, const uint32_t* base, uint32_t neg, int32_t* va, int32_t* vb) {
  typedef typename BankAcc<W64>::type acc_t;
  acc_t Sa[kBankJ], Sb[kBankJ];
#pragma unroll
  for (int j = 0; j < kBankJ; ++j) { Sa[j] = bank_init<W64>(sh, S0); Sb[j] = Sa[j]; }
#pragma unroll
  for (int k = 1; k < M; ++k) {
    const uint32_t step = sh.lin_step[k];
    const int32_t* T = (sh.tsel[k] ? tabs[1] : tabs[0]) + base[k] + step * lane;
    const int32_t Ak = (TAB == TAB_SMEM_HALF && ((neg >> k) & 1u)) ? -A[k] : A[k];
#pragma unroll
    for (int j = 0; j < kBankJ; ++j) {
      BHW_CHECK(base[k] + step * (lane + (uint32_t)(32 * j)) <= sh.lin_dmask[k]);
      const int32_t c2 = T[(uint32_t)(32 * j) * step];
      const int64_t P = (int64_t)Ak * (int64_t)c2;
      const acc_t ba = bank_term<W64>(sh, P, false);
      Sa[j] = (k & 1) ? Sa[j] - ba : Sa[j] + ba;
      if (PAIR) {
        if (k & 1) Sb[j] += bank_term<W64>(sh, PAIR == 2 ? (int64_t)Ak * (int64_t)(c2 + (int32_t)sh.pair_adj) : P, true);
        else Sb[j] += ba;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kBankJ; ++j) {
    va[j] = bank_finish<W64>(sh, Sa[j]);
    if (PAIR) vb[j] = bank_finish<W64>(sh, Sb[j]);
  }
}

// One lane's share of a tile: samples n + 32*j (j = 0..kBankJ-1) -> va[j], and their partners half a
// window later -> vb[j] when PAIR.  `tabs[u]` is distinct table u as the kernel sees it (staged
// or global).  A[k] are the window's pre-shifted coefficients, S0 its initial accumulator.
template <int M, int TAB, int PAIR, bool LANE_SIGN, bool W64>
BHW_HD void bank_lane_tile(const BankShape& sh, const int32_t* A, int32_t S0, const int32_t* const* tabs,
                           uint32_t n, uint32_t nbase, int32_t* va, int32_t* vb) {
  typedef typename BankAcc<W64>::type acc_t;
  acc_t Sa[kBankJ], Sb[kBankJ];
#pragma unroll
  for (int j = 0; j < kBankJ; ++j) { Sa[j] = bank_init<W64>(sh, S0); Sb[j] = Sa[j]; }
#pragma unroll
  for (int k = 1; k < M; ++k) {
    const uint32_t ks = sh.kstep[k];
    const int32_t* T = sh.tsel[k] ? tabs[1] : tabs[0];
    const uint32_t ph0 = n * ks;
    int32_t Ak = A[k];
    if (TAB == TAB_SMEM_HALF && !LANE_SIGN) Ak = ((int32_t)(nbase * ks) < 0) ? -Ak : Ak;
#pragma unroll
    for (int j = 0; j < kBankJ; ++j) {
      const uint32_t ph = ph0 + (uint32_t)(32 * j) * ks;
      const uint32_t idx = TAB == TAB_SMEM_HALF ? ((ph << 1) >> (sh.idx_rsh[k] + 1)) : (ph >> sh.idx_rsh[k]);
      BHW_CHECK(idx < (sh.tentries[sh.tsel[k]] >> (TAB == TAB_SMEM_HALF ? 1 : 0)));
      const int32_t c2 = T[idx];
      int32_t Ae = Ak;
      if (TAB == TAB_SMEM_HALF && LANE_SIGN) Ae = ((int32_t)ph < 0) ? -Ak : Ak;
      const int64_t P = (int64_t)Ae * (int64_t)c2;
      const acc_t ba = bank_term<W64>(sh, P, false);
      Sa[j] = (k & 1) ? Sa[j] - ba : Sa[j] + ba;
      if (PAIR) {
        // b(-P) = -bank_term(P, negated)
        if (k & 1) Sb[j] += bank_term<W64>(sh, PAIR == 2 ? (int64_t)Ae * (int64_t)(c2 + (int32_t)sh.pair_adj) : P, true);
        else Sb[j] += ba;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kBankJ; ++j) {
    va[j] = bank_finish<W64>(sh, Sa[j]);
    if (PAIR) vb[j] = bank_finish<W64>(sh, Sb[j]);
  }
}

// Parameters of windows that take the generic body inside a batch launch.
struct GenRec {
  WinParams wp;
  SrcParams src[2];
  uint32_t rom_off;  // Taylor ROM offset (I2 units); both units of a 3-term window share one ROM
  uint32_t pad[3];
};

}  // namespace bhw
