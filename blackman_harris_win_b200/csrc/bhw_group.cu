// bhw_group.cu - k_synth_group: one launch generates every window of a group (bhw_group.cuh): windows of one
// family (sin/cos source + DAT_WIDTH) and one entity (number of terms), of ANY mix of PHI_WIDTHs and ports.
// Persistent, one CTA of 1024 threads per SM; the family's table is staged in shared memory once per CTA
// (G_HALF32 / G_Q16) or gathered from the half-period pyramid through L1/L2 (G_GLOBAL); every warp takes
// tiles of 256 samples (sample pairs (n, n + N/2) in the paired instantiations: one look-up serves both),
// lane-interleaved so that every store instruction writes 128 contiguous bytes, streaming.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "bhw_group.cuh"
#include "bhw_launch.h"

namespace bhw {

constexpr int kGroupThreads = 1024;
constexpr int kGroupWarps = kGroupThreads / 32;

// Epilogue of one lane-tile.  APPLY 0: store the window samples (streaming, 128 B per warp instruction).
// APPLY 1 / 2: the window is never written - every frame of the caller's signal is multiplied by it on the fly,
// y = x * w as int_multNxN_dsp48 does (DAT_Q <= SIGNED(sig_a) * SIGNED(sig_b), 2*DTW bits,
// src/int_multNxN_dsp48.vhd:82,105): 1 = the exact product (int64), 2 = the window entities' own use of DAT_Q,
// r = DAT_Q[2DW-2 : DW-2], y = (r >> 1) + (r & 1) in DW bits (src/hamming_win.vhd:195-208) (int32).
template <bool PAIR, int APPLY>
__device__ __forceinline__ void group_epilogue(const GroupArgs& a, int32_t* o, size_t idx, size_t half, const int32_t* va,
                                               const int32_t* vb) {
  BHW_CHECK(idx + 32 * (kBankJ - 1) < 2 * half && (!PAIR || idx + 32 * (kBankJ - 1) < half));
  if (APPLY == 0) {
    int32_t* ot = o + idx;
#pragma unroll
    for (int j = 0; j < kBankJ; ++j) {
      __stcs(ot + 32 * j, va[j]);
      if (PAIR) __stcs(ot + half + 32 * j, vb[j]);
    }
    return;
  }
  if (APPLY == 3) {
    // packed output (BHW_OUT_INT16, DAT_WIDTH <= 16): `o` already points at the window's first int16 element
    short* ot = reinterpret_cast<short*>(o) + idx;
#pragma unroll
    for (int j = 0; j < kBankJ; ++j) {
      __stcs(ot + 32 * j, (short)va[j]);
      if (PAIR) __stcs(ot + half + 32 * j, (short)vb[j]);
    }
    return;
  }
  const int dw = (int)a.apply_dw;
  const int xsh = 32 - dw;
  // the frames are split over gridDim.y (launch_synth_group): a short window has too few tiles to fill the GPU,
  // so several CTAs generate the same tile and each multiplies its share of the frames
  const uint64_t f_begin = a.frames * blockIdx.y / gridDim.y, f_end = a.frames * (blockIdx.y + 1) / gridDim.y;
  for (uint64_t f = f_begin; f < f_end; ++f) {
    const size_t base = (size_t)f * (half * 2) + idx;
    const int32_t* xf = a.x + base;
#pragma unroll
    for (int h = 0; h < (PAIR ? 2 : 1); ++h) {
      int32_t xv[kBankJ];
#pragma unroll
      for (int j = 0; j < kBankJ; ++j) xv[j] = __ldcs(xf + h * half + 32 * j);
#pragma unroll
      for (int j = 0; j < kBankJ; ++j) {
        const int64_t p = (int64_t)((int32_t)((uint32_t)xv[j] << xsh) >> xsh) * (int64_t)(h ? vb[j] : va[j]);
        if (APPLY == 1) {
          __stcs(reinterpret_cast<long long*>(a.y) + base + h * half + 32 * j, (long long)p);
        } else {
          const int64_t r = wrapb(p >> (dw - 2), dw + 1);
          __stcs(reinterpret_cast<int32_t*>(a.y) + base + h * half + 32 * j, (int32_t)wrapb((r >> 1) + (r & 1), dw));
        }
      }
    }
  }
}

// G_GLOBAL: pull the pyramid lines the tile at sample `nbase` is going to gather into L1 ahead of time (one
// prefetch instruction per harmonic for up to 32 lines of 128 B), so that the gathers of the next tile find them
// there instead of waiting ~300 cycles for L2: the kernel is otherwise bound by exactly that wait (ncu: long
// scoreboard stalls, issue slots 54 % busy on a 2^26-point 5-term window).  A tile whose phase crosses a half
// period is skipped (rare).  `max_lines`: L1 is finite - harmonics whose span needs more are left alone.
template <int M>
__device__ __forceinline__ void group_prefetch_tile(const GroupShape& sh, uint32_t pw, uint32_t nbase, uint32_t lane,
                                                    uint32_t max_lines) {
  const uint32_t sl = 32u - pw;
#pragma unroll
  for (int k = 1; k < M; ++k) {
    const uint32_t ks = (uint32_t)k << sl;
    const uint32_t ph = (nbase * ks) & 0x7FFFFFFFu;
    const uint32_t span = (uint32_t)(kBankTile - 1) * ks;
    if ((((uint32_t)k * (uint32_t)(kBankTile - 1)) >> (pw - 1)) != 0u || ph + span >= 0x80000000u) continue;
    const uint32_t want = pw - harmonic_log2(k);
    const uint32_t L = want < sh.top ? want : sh.top;
    const uint32_t rsh = 32u - L;
    const uint32_t i0 = (ph | 0x80000000u) >> rsh, i1 = ((ph + span) | 0x80000000u) >> rsh;   // heap element indices
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(sh.pyr + i0) & ~(uintptr_t)127;
    const uintptr_t a1 = reinterpret_cast<uintptr_t>(sh.pyr + i1);
    const uint32_t nlines = (uint32_t)((a1 - a0) >> 7) + 1u;
    if (nlines > max_lines) continue;
    if (lane < nlines) asm volatile("prefetch.global.L1 [%0];" ::"l"(a0 + ((uintptr_t)lane << 7)));
  }
}

template <int M, int TAB, bool PAIR, int APPLY>
__global__ void __launch_bounds__(kGroupThreads, 1)
k_synth_group(const __grid_constant__ GroupArgs a) {
  extern __shared__ __align__(16) int32_t s_img[];
  const GroupShape& sh = a.sh;
  // launched with programmatic stream serialization behind the table build (a no-op otherwise)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const void* tab = sh.pyr;
  if (TAB != G_GLOBAL) {
    // G_HALF32: level `top` of the pyramid (2^(top-1) words from word 2^(top-1)); G_Q16: 2 * 2^(top-2) uint16
    const uint32_t words = TAB == G_HALF32 ? (1u << (sh.top - 1)) : (1u << (sh.top - 2));
    const int4* src = TAB == G_HALF32 ? reinterpret_cast<const int4*>(sh.pyr + (1u << (sh.top - 1)))
                                      : reinterpret_cast<const int4*>(sh.q16);
    int4* dst = reinterpret_cast<int4*>(s_img);
    for (uint32_t i = threadIdx.x; i < words / 4; i += kGroupThreads) dst[i] = __ldg(src + i);
    __syncthreads();
    tab = s_img;
  }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t U = a.nunits;
  if (TAB == G_GLOBAL && a.spread) {
    // One long window over a pyramid that lives in L2 / HBM.  Odd harmonic k of tile n reads its level around
    // k*n at stride k: k sectors per 8 look-ups, of which it uses every k-th word; the other words belong to
    // the tiles a k-th of a period away.  So warp j of G takes the j-th G-th of the window (tiles U*j/G + i,
    // all warps at the same i): with G = 30 the warps j, j + 10, j + 20 read the same sectors for k = 3 (and
    // j + 6m for k = 5) at the same time, through L1 - every sector crosses L2 -> SM once and every level is
    // swept once per harmonic (r1: N = 2^26 7-term 340 -> 226 us; with the pyramid's even-harmonic levels less).
    const uint32_t G = a.spread;
    if (warp >= G) return;
    const GroupWin* gw = a.wins ? a.wins : a.iw;
    const uint32_t pw = gw->pw;
    const WinRec* r = a.recs + gw->rec;
    int32_t A[M];
    A[0] = 0;
#pragma unroll
    for (int k = 1; k < M; ++k) A[k] = __ldg(&r->A[k]);
    const int32_t S0 = __ldg(&r->S0);
    const uint32_t n_first = __ldg(&r->n_first);
    int32_t* o = APPLY == 3 ? reinterpret_cast<int32_t*>(reinterpret_cast<short*>(a.out) + gw->out_off) : a.out + gw->out_off;
    const size_t half = (size_t)1 << (pw - 1);
    const uint32_t L = spread_steps(U, G);
    const uint32_t i0 = (uint32_t)((uint64_t)L * blockIdx.x / gridDim.x);
    const uint32_t i1 = (uint32_t)((uint64_t)L * (blockIdx.x + 1) / gridDim.x);
    for (uint32_t i = i0; i < i1; ++i) {
      uint32_t t;
      if (!spread_tile(U, G, warp, i, &t)) continue;
      if (a.prefetch_lines) group_prefetch_tile<M>(sh, pw, (t + 1) * kBankTile + n_first, lane, a.prefetch_lines);
      int32_t va[kBankJ], vb[kBankJ];
      group_lane_tile<M, TAB, PAIR>(sh, pw, A, S0, tab, t * kBankTile + n_first, lane, va, vb);
      group_epilogue<PAIR, APPLY>(a, o, (size_t)t * kBankTile + lane, half, va, vb);
    }
    return;
  }
  uint32_t u, u_end, u_step;
  if (sh.interleave) {
    u = blockIdx.x * kGroupWarps + warp; u_end = U; u_step = gridDim.x * kGroupWarps;
  } else {
    const uint32_t u0 = (uint32_t)((uint64_t)U * blockIdx.x / gridDim.x);
    u_end = (uint32_t)((uint64_t)U * (blockIdx.x + 1) / gridDim.x);
    u = u0 + warp; u_step = kGroupWarps;
  }
  if (u >= u_end) return;
  const GroupWin* wins = a.wins ? a.wins : a.iw;
  u += a.unit_base;
  u_end += a.unit_base;
  uint32_t w = group_find_window(wins, a.nwin, u);
  uint32_t w_begin = 0, w_next = 0, pw = 0, n_first = 0, tile_first = 0;
  int32_t A[M], S0 = 0;
  int32_t* o = nullptr;
  bool loaded = false;
  for (; u < u_end; u += u_step) {
    if (!loaded || u >= w_next) {
      if (loaded) {
        // the next window is usually the neighbour; an interleaved walk over a list of short windows jumps far
        // (65,536 windows of 1024 samples: hundreds of windows per step) - then search instead of walking
        uint32_t hops = 0;
        do { ++w; } while (u >= wins[w + 1].unit_begin && ++hops < 16u);
        if (u >= wins[w + 1].unit_begin) w = group_find_window(wins, a.nwin, u);
      }
      const GroupWin* gw = wins + w;
      w_begin = gw->unit_begin;
      w_next = gw[1].unit_begin;
      pw = gw->pw;
      tile_first = gw->tile_first;
      o = APPLY == 3 ? reinterpret_cast<int32_t*>(reinterpret_cast<short*>(a.out) + gw->out_off) : a.out + gw->out_off;
      const WinRec* r = a.recs + gw->rec;
      A[0] = 0;
#pragma unroll
      for (int k = 1; k < M; ++k) A[k] = __ldg(&r->A[k]);
      S0 = __ldg(&r->S0);
      n_first = __ldg(&r->n_first);
      loaded = true;
    }
    const uint32_t t = u - w_begin + tile_first;
    int32_t va[kBankJ], vb[kBankJ];
    group_lane_tile<M, TAB, PAIR>(sh, pw, A, S0, tab, t * kBankTile + n_first, lane, va, vb);
    group_epilogue<PAIR, APPLY>(a, o, (size_t)t * kBankTile + lane, (size_t)1 << (pw - 1), va, vb);
  }
}

size_t group_smem_limit() { return 192u * 1024u; }

template <int M, int TAB, bool PAIR, int APPLY>
static cudaError_t launch_group_t(const GroupArgs& a, unsigned grid, size_t smem, cudaStream_t stream, bool pdl) {
  // fused apply step: fewer CTAs than two per SM -> split the frames over gridDim.y
  unsigned fy = 1;
  if ((APPLY == 1 || APPLY == 2) && a.frames > 1) {
    const unsigned want = 2u * (unsigned)device_sm_count();
    if (grid < want) fy = (want + grid - 1) / grid;
    if ((uint64_t)fy > a.frames) fy = (unsigned)a.frames;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_synth_group<M, TAB, PAIR, APPLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)group_smem_limit());
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid, fy);
  cfg.blockDim = dim3(kGroupThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, k_synth_group<M, TAB, PAIR, APPLY>, a);
}

template <int M, int TAB>
static cudaError_t launch_group_tab(const GroupArgs& a, bool pair, int apply, unsigned grid, size_t smem, cudaStream_t stream,
                                    bool pdl) {
  if (apply == 1) return pair ? launch_group_t<M, TAB, true, 1>(a, grid, smem, stream, pdl) : cudaErrorInvalidValue;
  if (apply == 2) return pair ? launch_group_t<M, TAB, true, 2>(a, grid, smem, stream, pdl) : cudaErrorInvalidValue;
  if (apply == 3) return pair ? launch_group_t<M, TAB, true, 3>(a, grid, smem, stream, pdl)
                              : launch_group_t<M, TAB, false, 3>(a, grid, smem, stream, pdl);
  return pair ? launch_group_t<M, TAB, true, 0>(a, grid, smem, stream, pdl)
              : launch_group_t<M, TAB, false, 0>(a, grid, smem, stream, pdl);
}

template <int M>
static cudaError_t launch_group_m(const GroupArgs& a, int tab, bool pair, int apply, unsigned grid, size_t smem,
                                  cudaStream_t stream, bool pdl) {
  if (tab == G_HALF32) return launch_group_tab<M, G_HALF32>(a, pair, apply, grid, smem, stream, pdl);
  if (tab == G_Q16) return launch_group_tab<M, G_Q16>(a, pair, apply, grid, smem, stream, pdl);
  return launch_group_tab<M, G_GLOBAL>(a, pair, apply, grid, 0, stream, pdl);
}

cudaError_t launch_synth_group(const GroupArgs& a, int tab, bool pair, cudaStream_t stream, bool pdl) {
  const int apply = a.x ? (int)a.apply_mode : a.pack16 ? 3 : 0;
  if (a.x && (!a.y || !a.frames || apply > 2 || a.pack16)) return cudaErrorInvalidValue;
  if (!a.nunits || !a.nwin) return cudaSuccess;
  if (a.spread && (tab != G_GLOBAL || a.nwin != 1 || a.unit_base || a.spread > (uint32_t)kGroupWarps)) return cudaErrorInvalidValue;
  const uint64_t ctas = a.spread ? ((uint64_t)a.nunits + a.spread - 1) / a.spread
                                 : ((uint64_t)a.nunits + kGroupWarps - 1) / kGroupWarps;
  const unsigned sms = (unsigned)device_sm_count();
  const unsigned grid = (unsigned)(ctas < sms ? ctas : sms);
  size_t smem = 0;
  if (tab == G_HALF32) smem = (size_t)4 << (a.sh.top - 1);
  else if (tab == G_Q16) smem = (size_t)4 << (a.sh.top - 2);
  if (smem > group_smem_limit()) return cudaErrorInvalidValue;
  switch (a.sh.m) {
    case 2: return launch_group_m<2>(a, tab, pair, apply, grid, smem, stream, pdl);
    case 3: return launch_group_m<3>(a, tab, pair, apply, grid, smem, stream, pdl);
    case 4: return launch_group_m<4>(a, tab, pair, apply, grid, smem, stream, pdl);
    case 5: return launch_group_m<5>(a, tab, pair, apply, grid, smem, stream, pdl);
    case 7: return launch_group_m<7>(a, tab, pair, apply, grid, smem, stream, pdl);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace bhw
