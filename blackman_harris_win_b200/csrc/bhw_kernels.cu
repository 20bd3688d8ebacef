// bhw_kernels.cu - sm_100a kernels of the window generator and their launchers.
//
// Nothing here is a dense contraction, so there are no tensor-core instructions: the work is
// integer shift-add (CORDIC), table gathers and coalesced stores.  Three kernel families:
//   k_table_build  : stage 1 of BHW_ALGO_TABLE - one thread per distinct source phase, the
//                    shift-add stages run in registers with the atan words in constant memory;
//                    writes the full-period cosine table of each sin/cos source of a launch.
//   k_synth        : stage 2 - one warp per 128 consecutive output samples (lane-interleaved so
//                    that both the gathers of harmonic k and the stores are coalesced), gathers
//                    cos(k*phi) from the tables and applies the entity's multiply / round / sum /
//                    round tail bit-exactly.  Handles a whole batch of windows in one launch.
//   k_direct_*     : BHW_ALGO_DIRECT - one thread per output sample, every k*phi term evaluated
//                    in registers (also the only path for DAT_WIDTH > 32), and the sin/cos entry.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "bhw_device.cuh"
#include "bhw_launch.h"

namespace bhw {

// -------------------------------------------------------------------------------------------
// stage 1: trig tables
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_table_build(const TabJob* __restrict__ jobs, int njobs, uint32_t total_work,
              const I2* __restrict__ rom) {
  // programmatic dependent launch: the synthesis kernel queued behind this one may be scheduled
  // now; it blocks in griddepcontrol.wait until this grid has completed and flushed its tables
  asm volatile("griddepcontrol.launch_dependents;");
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total_work;
       i += gridDim.x * blockDim.x) {
    // job of work item i: last job with work_begin <= i (few jobs; the search is noise next
    // to hundreds of shift-add operations)
    int lo = 0, hi = njobs - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].work_begin <= i) lo = mid; else hi = mid - 1;
    }
    const TabJob& job = jobs[lo];
    table_build_item(job, rom, i - job.work_begin);
  }
}

// One large table whose source runs on the 32-bit core: the job travels in the parameter block,
// the stages are unrolled (NXY), one thread per quarter-wave phase.
template <int NXY, bool BIAS>
__global__ void __launch_bounds__(256)
k_table_build_u(const __grid_constant__ TabJob job) {
  asm volatile("griddepcontrol.launch_dependents;");
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < job.work; i += gridDim.x * blockDim.x)
    table_build_item_u<NXY, BIAS>(job, i);
}

// the same for an input-quadrant source (one thread per phase, left-aligned 64-bit registers)
template <int NXY>
__global__ void __launch_bounds__(256)
k_table_build_inq_u(const __grid_constant__ TabJob job) {
  asm volatile("griddepcontrol.launch_dependents;");
  // two entries per thread: quadrants 0/1 and 3/2 share their z (cordic_core_inq_u2)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < job.work / 2; i += gridDim.x * blockDim.x)
    table_build_item_inq_u2<NXY>(job, i);
}

// -------------------------------------------------------------------------------------------
// stage 2: synthesis
// -------------------------------------------------------------------------------------------
constexpr int kSynthThreads = 256;
constexpr int kSynthWarps = kSynthThreads / 32;
constexpr int kTile = 128;  // samples per warp tile: 4 per lane, lane-interleaved

// OutT = int32_t, or short for the packed output (BHW_OUT_INT16)
template <int M, typename OutT>
__device__ __forceinline__ void synth_tile(const WinRec& r, uint32_t n, OutT* __restrict__ out,
                                           uint32_t valid) {
  // lane owns samples n, n+32, n+64, n+96 of the tile; `valid` = samples left from this lane's first
  if (r.flags & WR_ACC64) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((uint32_t)(32 * j) < valid) out[32 * j] = (OutT)synth_sample64<M>(r, n + 32 * j);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((uint32_t)(32 * j) < valid) out[32 * j] = (OutT)synth_sample32<M>(r, n + 32 * j);
  }
}

__device__ __forceinline__ int find_window(const uint64_t* __restrict__ off, int nwin, uint64_t f) {
  int lo = 0, hi = nwin - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= f) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <typename OutT>
__global__ void __launch_bounds__(kSynthThreads)
k_synth(const __grid_constant__ SynthArgs a) {
  __shared__ WinRec s_rec[kSynthWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WinRec& rec = s_rec[warp];
  int cur_w = -1, cur_r = -1;
  const uint64_t ntiles = a.npieces ? (uint64_t)a.piece_tile0[a.npieces] : (a.flat_count + kTile - 1) / kTile;
  OutT* const out = reinterpret_cast<OutT*>(a.out);
  for (uint64_t tile = (uint64_t)blockIdx.x * kSynthWarps + warp; tile < ntiles;
       tile += (uint64_t)gridDim.x * kSynthWarps) {
    uint64_t o0, f0, left;
    if (a.npieces) {
      // several disjoint ranges in one launch: piece of this tile (few pieces: linear search)
      uint32_t p = 0;
      while (p + 1 < a.npieces && (uint32_t)tile >= a.piece_tile0[p + 1]) ++p;
      f0 = a.piece_begin[p] + (tile - a.piece_tile0[p]) * kTile;
      left = a.piece_end[p] - f0;
      o0 = f0 - a.out_flat0;
    } else {
      o0 = tile * kTile;                            // first output index of the tile
      f0 = a.flat_begin + o0;                       // its flat sample index
      left = a.flat_count - o0;                     // samples left in the request
    }
    const uint32_t tile_n = left < kTile ? (uint32_t)left : kTile;
    int w;
    if (a.uniform_pw >= 0) w = (int)(f0 >> a.uniform_pw);
    else if (cur_w >= 0 && f0 >= a.flat_off[cur_w] && f0 < a.flat_off[cur_w + 1]) w = cur_w;
    else w = find_window(a.flat_off, a.nwin, f0);
    const uint64_t wbeg = a.uniform_pw >= 0 ? ((uint64_t)w << a.uniform_pw) : a.flat_off[w];
    const uint64_t wend = a.uniform_pw >= 0 ? ((uint64_t)(w + 1) << a.uniform_pw) : a.flat_off[w + 1];
    if (f0 + tile_n <= wend) {
      // whole tile inside window w (always the case for windows of >= 128 samples cut at
      // multiples of 128): warp-uniform record, staged once in shared memory
      if (w != cur_w) {
        const int ri = a.win_rec ? (int)a.win_rec[w] : 0;
        if (ri != cur_r) {
          __syncwarp();
          const uint32_t* src = reinterpret_cast<const uint32_t*>(a.recs + ri);
          uint32_t* dst = reinterpret_cast<uint32_t*>(&rec);
          for (int i = lane; i < (int)(sizeof(WinRec) / 4); i += 32) dst[i] = src[i];
          __syncwarp();
          cur_r = ri;
        }
        cur_w = w;
      }
      const uint32_t n = (uint32_t)(f0 - wbeg) + rec.n_first + lane;
      OutT* o = out + o0 + lane;
      const uint32_t valid = tile_n > (uint32_t)lane ? tile_n - lane : 0;
      if (rec.flags & WR_GENERIC) {
        const GenRec& g = a.gens[rec.gen_idx];
        const uint64_t nmask = (1ull << g.wp.pw) - 1;
        for (int j = 0; j < 4; ++j)
          if ((uint32_t)(32 * j) < valid)
            o[32 * j] = (OutT)direct_sample_generic(g.wp, g.src, a.rom + g.rom_off,
                                                    (uint64_t)(n + 32 * j) & nmask);
      } else {
        switch (rec.m) {
          case 2: synth_tile<2, OutT>(rec, n, o, valid); break;
          case 3: synth_tile<3, OutT>(rec, n, o, valid); break;
          case 4: synth_tile<4, OutT>(rec, n, o, valid); break;
          case 5: synth_tile<5, OutT>(rec, n, o, valid); break;
          case 7: synth_tile<7, OutT>(rec, n, o, valid); break;
          default:   // 6 and 8..11 terms (BHW_WIN_MTERM_*): the run-time form
            for (int j = 0; j < 4; ++j)
              if ((uint32_t)(32 * j) < valid) o[32 * j] = (OutT)synth_sample(rec, n + 32 * j);
            break;
        }
      }
    } else {
      // tile straddles windows (windows shorter than 128 samples, or ragged request ends):
      // per-sample look-up straight from global memory
      cur_w = -1;
      for (int j = 0; j < 4; ++j) {
        const uint32_t i = lane + 32 * j;
        if (i >= tile_n) break;
        const uint64_t f = f0 + i;
        int ww = w;
        while (f >= (a.uniform_pw >= 0 ? ((uint64_t)(ww + 1) << a.uniform_pw) : a.flat_off[ww + 1])) ++ww;
        const WinRec& r = a.recs[a.win_rec ? a.win_rec[ww] : 0];
        const uint64_t wb = a.uniform_pw >= 0 ? ((uint64_t)ww << a.uniform_pw) : a.flat_off[ww];
        const uint32_t n = (uint32_t)(f - wb) + r.n_first;
        int32_t v;
        if (r.flags & WR_GENERIC) {
          const GenRec& g = a.gens[r.gen_idx];
          v = (int32_t)direct_sample_generic(g.wp, g.src, a.rom + g.rom_off,
                                             (uint64_t)n & ((1ull << g.wp.pw) - 1));
        } else {
          v = synth_sample(r, n);
        }
        out[o0 + i] = (OutT)v;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// stage 2, bank form: whole windows of one shape (see BankShape in bhw_device.cuh)
// -------------------------------------------------------------------------------------------
// Persistent: one CTA per SM, each CTA owns a contiguous range of warp tiles, its warps take
// consecutive tiles so that a CTA sweeps its output range front to back.  The trig tables are
// staged in shared memory once per CTA (TAB_SMEM_*), gathers then cost one LDS each; with
// TAB_GLOBAL they stay in L2/L1.  Stores are 128 B per warp instruction, streaming.
constexpr int kBankThreads = 1024;
constexpr int kBankWarps = kBankThreads / 32;

// OutT: int32_t, or short for the packed output of DAT_WIDTH <= 16 windows (BHW_OUT_INT16)
template <int M, int TAB, int PAIR, bool W64, typename OutT = int32_t>
__global__ void __launch_bounds__(kBankThreads, 1)
k_synth_bank(const __grid_constant__ BankArgs a) {
  extern __shared__ __align__(16) int32_t s_tab[];
  const BankShape& sh = a.sh;
  const int32_t* tab0 = sh.tab[0];
  const int32_t* tab1 = sh.tab[1];
  // launched with programmatic stream serialization right behind k_table_build: wait here for
  // its tables (a no-op for an ordinary launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (TAB != TAB_GLOBAL) {
    for (uint32_t u = 0; u < sh.ntab; ++u) {
      const uint32_t words = sh.tentries[u] >> (TAB == TAB_SMEM_HALF ? 1 : 0);
      const int4* src = reinterpret_cast<const int4*>(sh.tab[u]);
      int4* dst = reinterpret_cast<int4*>(s_tab + sh.toff[u]);
      for (uint32_t i = threadIdx.x; i < words / 4; i += kBankThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    tab0 = s_tab + sh.toff[0];
    tab1 = s_tab + sh.toff[1];
  }
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t pw = sh.pw;
  const uint32_t log_tpw = pw - kBankTileLog2 - (PAIR ? 1 : 0);  // log2(tiles per window)
  const uint32_t half = 1u << (pw - 1);
  // whole windows, or a tile range inside one window (a shard or a requested range that cuts a long
  // window): unit u is tile (u + tile_off) of the launch's windows
  // (only the unpaired global-table instantiations take tile ranges - launch_synth_bank() - so the
  // staged, paired kernels of whole-window banks carry none of this)
  constexpr bool kRange = TAB == TAB_GLOBAL && PAIR == 0;
  const uint64_t U = kRange && a.ntiles ? (uint64_t)a.ntiles : (uint64_t)a.nwin << log_tpw;
  const uint64_t toff = kRange && a.ntiles ? (uint64_t)a.tile_off : 0;
  const uint64_t u0 = U * blockIdx.x / gridDim.x, u1 = U * (blockIdx.x + 1) / gridDim.x;
  // ports of a window (warp-uniform)
  struct Ports { int32_t A[M]; int32_t S0; uint32_t n_first; };
  auto load_ports = [&](uint32_t w, Ports& p) {
    const WinRec* r = a.recs + (a.win_rec ? __ldg(a.win_rec + a.w_first + w) : 0u);
    p.A[0] = 0;
#pragma unroll
    for (int k = 1; k < M; ++k) p.A[k] = __ldg(&r->A[k]);
    p.S0 = __ldg(&r->S0);
    p.n_first = __ldg(&r->n_first);
  };
  auto do_tile = [&](const Ports& p, uint32_t w, uint32_t t) {
    const uint32_t nbase = t * kBankTile + p.n_first;
    const uint32_t n = nbase + lane;
    const int32_t* tabs[2] = {tab0, tab1};
    int32_t va[kBankJ], vb[kBankJ];
    uint32_t lbase[M], lneg;
    if (sh.lin && bank_tile_linear<M, TAB>(sh, nbase, lbase, &lneg))   // warp-uniform
      bank_lane_tile_lin<M, TAB, PAIR, W64>(sh, p.A, p.S0, tabs, lane, lbase, lneg, va, vb);
    else if (TAB == TAB_SMEM_HALF)
      bank_lane_tile<M, TAB, PAIR, true, W64>(sh, p.A, p.S0, tabs, n, nbase, va, vb);
    else
      bank_lane_tile<M, TAB, PAIR, false, W64>(sh, p.A, p.S0, tabs, n, nbase, va, vb);
    OutT* o = reinterpret_cast<OutT*>(a.out) + ((uint64_t)w << pw) + t * kBankTile + lane;
    if (kRange) o -= toff * kBankTile;
#pragma unroll
    for (int j = 0; j < kBankJ; ++j) {
      __stcs(o + 32 * j, (OutT)va[j]);
      if (PAIR) __stcs(o + half + 32 * j, (OutT)vb[j]);
    }
  };
  Ports cur;
  uint32_t cur_w = 0xFFFFFFFFu;
  if (TAB == TAB_GLOBAL && a.win_minor) {
    // Tables read from L2: a bank of windows over one table walks it window-minor - unit u is tile
    // u / nwin of window u % nwin - so that the warps of a CTA work on the same tile of consecutive
    // windows at the same time and all but the first find the table sectors in L1 (a stride-k
    // gather pulls 4k sectors per warp instruction; window-major, every tile pulled them from L2
    // again: 2.7 GB per 256 MB of 7-term output, L2 82 % busy).
    const uint32_t nwin = a.nwin;
    for (uint64_t u = u0 + warp; u < u1; u += kBankWarps) {
      uint32_t t, w;
      win_minor_unit((uint32_t)u, nwin, &w, &t);
      load_ports(w, cur);
      do_tile(cur, w, t);
    }
    return;
  }
  if (TAB == TAB_GLOBAL && a.spread) {
    // One long window over an L2-resident (or larger) table.  Harmonic k of tile n reads the table
    // around k*n at stride k: 4k sectors per warp instruction of which it uses every k-th word; the
    // other words belong to the tiles a k-th of a period away.  So warp j of G takes the j-th G-th
    // of the window (tiles U*j/G + i, all warps at the same i): harmonic k then reads G/gcd(k,G)
    // distinct regions instead of G, their sectors shared through L1 by the warps whose offsets
    // differ by a multiple of period/k.  G = 30 for 7 terms: 7 sector fetches per tile row instead of
    // 21 at best (measured, 7-term DAT_WIDTH 32: N=2^25 135 -> 92 us, N=2^26 340 -> 226 us, cordic_dds48
    // N=2^24 80 -> 48 us; a CTA barrier per step to keep the warps together measured slower).
    const uint32_t G = a.spread;
    if (warp < G) {
      const uint32_t Ut = (uint32_t)U;
      const uint32_t L = spread_steps(Ut, G);
      const uint32_t i0 = (uint32_t)((uint64_t)L * blockIdx.x / gridDim.x);
      const uint32_t i1 = (uint32_t)((uint64_t)L * (blockIdx.x + 1) / gridDim.x);
      load_ports(0, cur);
      for (uint32_t i = i0; i < i1; ++i) {
        uint32_t t;
        if (spread_tile(Ut, G, warp, i, &t)) do_tile(cur, 0, t);
      }
    }
    return;
  }
  if (M <= 3) {
    // 2- and 3-term windows have registers to spare: fetch the next window's ports one tile ahead,
    // so that the two dependent loads (window -> record -> ports) never sit in front of a tile.
    // Short windows start a new window with nearly every tile (N = 1024: two tile pairs per window).
    // These light kernels are paced by the store path, not by instruction issue, and the store
    // bandwidth an SM gets is uneven (tools/store_pattern_probe.cu): tiles are interleaved over the
    // grid (the whole GPU sweeps the output front to back) instead of one contiguous share per CTA.
    // (Unpaired tiles - sources without the half-period antisymmetry - measured slower interleaved
    // and keep the contiguous share.)
    Ports nxt;
    const uint64_t first = PAIR ? (uint64_t)blockIdx.x * kBankWarps + warp : u0 + warp;
    const uint64_t stride = PAIR ? (uint64_t)gridDim.x * kBankWarps : (uint64_t)kBankWarps;
    const uint64_t last = PAIR ? U : u1;
    if (first < last) {
      cur_w = (uint32_t)((first + toff) >> log_tpw);
      load_ports(cur_w, cur);
    }
    nxt = cur;
    for (uint64_t u = first; u < last; u += stride) {
      const uint32_t w = cur_w;
      const uint64_t u2 = u + stride;
      const uint32_t w2 = u2 < last ? (uint32_t)((u2 + toff) >> log_tpw) : w;
      if (w2 != w) load_ports(w2, nxt);  // consumed after this tile
      do_tile(cur, w, (uint32_t)(u + toff) & ((1u << log_tpw) - 1));
      if (w2 != w) { cur = nxt; cur_w = w2; }
    }
  } else {
    // 4 and more terms: paced by instruction issue; one contiguous share per CTA measured faster
    // than any interleaving (bh4: 171 vs 177 us, bh5: 197 vs 207 us per GiB)
    for (uint64_t u = u0 + warp; u < u1; u += kBankWarps) {
      const uint32_t w = (uint32_t)((u + toff) >> log_tpw);
      if (w != cur_w) {
        load_ports(w, cur);
        cur_w = w;
      }
      do_tile(cur, w, (uint32_t)(u + toff) & ((1u << log_tpw) - 1));
    }
  }
}

// -------------------------------------------------------------------------------------------
// direct evaluation (one thread per sample) and the sin/cos entry
// -------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256)
k_direct_window(const __grid_constant__ DirectArgs a, OutT* __restrict__ out) {
  // taylor_sincos keeps its quarter-wave ROM in shared memory when it fits
  extern __shared__ I2 s_rom[];
  const I2* rom = a.rom;
  if (a.rom_smem_entries) {
    for (uint32_t i = threadIdx.x; i < a.rom_smem_entries; i += blockDim.x) s_rom[i] = a.rom[i];
    __syncthreads();
    rom = s_rom;
  }
  const uint64_t nmask = (1ull << a.wp.pw) - 1;
  if (a.quad_adv >> 31) {    // whole window: one shift-add evaluation per harmonic and four samples
    const uint64_t quarter = a.count / 4;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < quarter;
         j += (uint64_t)gridDim.x * blockDim.x) {
      int64_t w[4];
      direct_sample_core_quad(a.wp, a.src, a.sc, rom, (a.n_first + j) & nmask, a.quad_adv & 0x7FFFFFFFu, w);
#pragma unroll
      for (int r = 0; r < 4; ++r) out[r * quarter + j] = (OutT)w[r];
    }
    return;
  }
  if (a.pair_flip >> 31) {   // whole window: one shift-add evaluation per harmonic and sample pair
    const uint64_t half = a.count / 2;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half;
         j += (uint64_t)gridDim.x * blockDim.x) {
      int64_t wa, wb;
      direct_sample_core_pair(a.wp, a.src, a.sc, rom, (a.n_first + j) & nmask, a.pair_flip & 0x7FFFFFFFu, wa, wb);
      out[j] = (OutT)wa;
      out[half + j] = (OutT)wb;
    }
    return;
  }
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.count;
       j += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t n = (a.n_first + j) & nmask;
    out[j] = (OutT)direct_sample_core(a.wp, a.src, a.sc, rom, n);
  }
}

// Register-resident 32-bit form: 4 consecutive samples per thread (four independent shift-add
// chains in flight), one 128-bit store.  NXY = compile-time stage count (0: run-time loop).
template <int NXY>
__global__ void __launch_bounds__(256)
k_direct32(const __grid_constant__ Direct32Args a, int32_t* __restrict__ out) {
  const Direct32Params& p = a.p;
  if (a.narrow) {
    // a short request: one sample (or sample pair) per thread.  The CORDIC stages of one evaluation
    // are a dependent chain, so a thread that owns 4 samples x (M-1) harmonics runs ~4 x (M-1) x NXY
    // stages back to back while most of the GPU idles; spreading them shortens the kernel 4-fold
    // (N = 65536 4-term: 6.2 -> 4.x us per call).
    const uint64_t items = a.pair == 2 ? a.count / 4 : a.pair ? a.count / 2 : a.count;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < items; j += (uint64_t)gridDim.x * blockDim.x) {
      const uint32_t n = (uint32_t)(a.n0 + j) + p.n_first;
      if (a.pair == 2) {
        int32_t w[4];
        direct32_quad<NXY>(p, n, w);
#pragma unroll
        for (int r = 0; r < 4; ++r) out[r * items + j] = w[r];
      } else if (a.pair) {
        int32_t wa, wb;
        direct32_pair<NXY>(p, n, wa, wb);
        out[j] = wa;
        out[items + j] = wb;
      } else {
        out[j] = direct32_sample<NXY>(p, n);
      }
    }
    return;
  }
  const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const uint32_t pmask = (1u << p.pw) - 1u;
  if (a.pair == 2) {
    // whole window: 4 consecutive samples of the first quarter and their partners a quarter, half and
    // three quarters of a window later - one set of CORDIC evaluations for the sixteen of them
    const uint64_t quarter = a.count / 4;
    for (uint64_t qd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; qd < quarter / 4;
         qd += (uint64_t)gridDim.x * blockDim.x) {
      const uint64_t j = qd * 4;
      const uint32_t n = (uint32_t)j + p.n_first;
      int32_t v[4][4];   // [sample e][quarter r]
#pragma unroll
      for (int e = 0; e < 4; ++e) direct32_quad<NXY>(p, n + e, v[e]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (aligned) {
          __stcs(reinterpret_cast<int4*>(out + r * quarter + j), make_int4(v[0][r], v[1][r], v[2][r], v[3][r]));
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) out[r * quarter + j + e] = v[e][r];
        }
      }
    }
    return;
  }
  if (a.pair) {
    // whole window: 4 consecutive samples of the first half and their partners half a window later
    const uint64_t half = a.count / 2;
    for (uint64_t qd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; qd < half / 4;
         qd += (uint64_t)gridDim.x * blockDim.x) {
      const uint64_t j = qd * 4;
      const uint32_t n = (uint32_t)j + p.n_first;
      uint32_t Sa[4], Sb[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) { Sa[e] = (uint32_t)p.S0; Sb[e] = Sa[e]; }
      for (int k = 1; k < p.m; ++k) {
        const uint32_t km = p.kmul[k];
        const int32_t Ak = p.A[k];
        const bool odd = (km & 1u) != 0;      // odd harmonic: the partner sees the negated cosine
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int32_t c = direct32_cos<NXY>(p, (km * (n + e)) & pmask);
          const uint32_t ba = (uint32_t)mulhi_rc(Ak, c << p.tshift, p.rc);
          const uint32_t bb = odd ? (uint32_t)mulhi_rc(Ak, (-c) << p.tshift, p.rc) : ba;
          Sa[e] = (k & 1) ? Sa[e] - ba : Sa[e] + ba;
          Sb[e] = (k & 1) ? Sb[e] - bb : Sb[e] + bb;
        }
      }
      int32_t va[4], vb[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        va[e] = (int32_t)(Sa[e] << p.lsh) >> p.rsh;
        vb[e] = (int32_t)(Sb[e] << p.lsh) >> p.rsh;
      }
      if (aligned) {
        __stcs(reinterpret_cast<int4*>(out + j), make_int4(va[0], va[1], va[2], va[3]));
        __stcs(reinterpret_cast<int4*>(out + half + j), make_int4(vb[0], vb[1], vb[2], vb[3]));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) { out[j + e] = va[e]; out[half + j + e] = vb[e]; }
      }
    }
    return;
  }
  const uint64_t quads = (a.count + 3) / 4;
  for (uint64_t qd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; qd < quads;
       qd += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t j = qd * 4;
    const uint32_t n = (uint32_t)(a.n0 + j) + p.n_first;  // taken modulo 2^pw below
    uint32_t S[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) S[e] = (uint32_t)p.S0;
    for (int k = 1; k < p.m; ++k) {
      const uint32_t km = p.kmul[k];
      const int32_t Ak = p.A[k];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int32_t c = direct32_cos<NXY>(p, (km * (n + e)) & pmask);
        const uint32_t b = (uint32_t)mulhi_rc(Ak, c << p.tshift, p.rc);
        S[e] = (k & 1) ? S[e] - b : S[e] + b;
      }
    }
    int32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (int32_t)(S[e] << p.lsh) >> p.rsh;
    if (aligned && j + 4 <= a.count) {
      __stcs(reinterpret_cast<int4*>(out + j), make_int4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (j + e < a.count) out[j + e] = v[e];
    }
  }
}

// TAYLOR windows, register-resident 32-bit form: 4 consecutive samples per thread, the quarter-wave
// sine ROM in shared memory, one 128-bit store.  TMODE = the tay1_order datapath (compile time).
template <int TMODE>
__global__ void __launch_bounds__(256)
k_direct_taylor(const __grid_constant__ DirectTayArgs a, int32_t* __restrict__ out) {
  extern __shared__ I2 s_rom[];
  const DirectTayParams& p = a.p;
  for (uint32_t i = threadIdx.x; i < p.rom_entries; i += blockDim.x) s_rom[i] = a.rom[i];
  __syncthreads();
  const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (a.pair == 2 && TMODE == TMODE_WIDE && aligned) {
    // whole long window: 4 consecutive samples of the first quarter and their three partners (direct_taylor_quad4)
    const uint64_t quarter = a.count / 4;
    for (uint64_t qd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; qd < quarter / 4;
         qd += (uint64_t)gridDim.x * blockDim.x) {
      const uint64_t j = qd * 4;
      int32_t w[4][4];
      direct_taylor_quad4(p, s_rom, (uint32_t)j, w);
#pragma unroll
      for (int r = 0; r < 4; ++r)
        __stcs(reinterpret_cast<int4*>(out + r * quarter + j), make_int4(w[r][0], w[r][1], w[r][2], w[r][3]));
    }
    return;
  }
  if (a.pair) {
    // whole window: 4 consecutive samples of the first half and their partners half a window later
    const uint64_t half = a.count / 2;
    for (uint64_t qd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; qd < half / 4;
         qd += (uint64_t)gridDim.x * blockDim.x) {
      const uint64_t j = qd * 4;
      const uint32_t n = (uint32_t)j + p.n_first;
      int32_t va[4], vb[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) direct_taylor_pair<TMODE>(p, s_rom, n + e, va[e], vb[e]);
      if (aligned) {
        __stcs(reinterpret_cast<int4*>(out + j), make_int4(va[0], va[1], va[2], va[3]));
        __stcs(reinterpret_cast<int4*>(out + half + j), make_int4(vb[0], vb[1], vb[2], vb[3]));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) { out[j + e] = va[e]; out[half + j + e] = vb[e]; }
      }
    }
    return;
  }
  const uint64_t quads = (a.count + 3) / 4;
  for (uint64_t qd = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; qd < quads;
       qd += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t j = qd * 4;
    const uint32_t n = (uint32_t)(a.n0 + j) + p.n_first;  // each unit masks it to its own phase width
    int32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = direct_taylor_sample<TMODE>(p, s_rom, n + e);
    if (aligned && j + 4 <= a.count) {
      __stcs(reinterpret_cast<int4*>(out + j), make_int4(v[0], v[1], v[2], v[3]));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) if (j + e < a.count) out[j + e] = v[e];
    }
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256)
k_sincos(const __grid_constant__ SinCosArgs a, OutT* __restrict__ out_sin, OutT* __restrict__ out_cos) {
  if (a.quad) {
    const uint64_t Q = a.count / 4;
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < Q; j += (uint64_t)gridDim.x * blockDim.x) {
      int64_t s[4], c[4];
      eval_source_core_quad(a.src, a.sc, a.rom, a.n_first + j, s, c);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (out_sin) out_sin[j + r * Q] = (OutT)s[r];
        if (out_cos) out_cos[j + r * Q] = (OutT)c[r];
      }
    }
    return;
  }
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.count;
       j += (uint64_t)gridDim.x * blockDim.x) {
    int64_t s, c;
    eval_source_core(a.src, a.sc, a.rom, a.n_first + j, s, c);
    if (out_sin) out_sin[j] = (OutT)s;
    if (out_cos) out_cos[j] = (OutT)c;
  }
}

// cordic_atan2: one thread per (VEC_DX, VEC_DY) pair; 8 B read + 4 B written per sample, but
// ~10 instructions per stage make it integer-issue-bound long before HBM.
__global__ void __launch_bounds__(256)
k_atan2(const __grid_constant__ Atan2Params p, const int32_t* __restrict__ x, const int32_t* __restrict__ y,
        int32_t* __restrict__ phi, uint64_t count, uint64_t avail) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count;
       j += (uint64_t)gridDim.x * blockDim.x)
  {
    const int32_t xv = __ldcs(x + j), yv = __ldcs(y + j);
    // stream_quadrant: the quadrant of the next pair of the stream (the inputs after the last pair read 0)
    const int32_t qx = !p.skew ? xv : (j + 1 < avail ? __ldg(x + j + 1) : 0), qy = !p.skew ? yv : (j + 1 < avail ? __ldg(y + j + 1) : 0);
    __stcs(phi + j, atan2_sample(p, xv, yv, qx, qy));
  }
}

// ANGLE_WIDTH known at compile time (W <= 32): unrolled stages, 4 pairs per thread, 128-bit accesses
template <int AW>
__global__ void __launch_bounds__(256)
k_atan2_u(const __grid_constant__ Atan2Params p, const int32_t* __restrict__ x, const int32_t* __restrict__ y,
          int32_t* __restrict__ phi, uint64_t count) {
  const uint64_t quads = count / 4;
  for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += (uint64_t)gridDim.x * blockDim.x) {
    const int4 xv = __ldcs(reinterpret_cast<const int4*>(x) + q), yv = __ldcs(reinterpret_cast<const int4*>(y) + q);
    int4 o;
    o.x = atan2_sample32_t<AW>(p, xv.x, yv.x);
    o.y = atan2_sample32_t<AW>(p, xv.y, yv.y);
    o.z = atan2_sample32_t<AW>(p, xv.z, yv.z);
    o.w = atan2_sample32_t<AW>(p, xv.w, yv.w);
    __stcs(reinterpret_cast<int4*>(phi) + q, o);
  }
  // tail (count not a multiple of 4)
  const uint64_t j = quads * 4 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < count) phi[j] = atan2_sample32_t<AW>(p, x[j], y[j]);
}

// Pairing over an input-quadrant CORDIC's table (BankShape::pair_adj): which entries break T[i + E/2] == -T[i] - adj?
// exc[0] = their number (zeroed by the host before the launch), exc[1 ...] = the indices i < E/2, any order.
__global__ void __launch_bounds__(256)
k_inq_exceptions(const int32_t* __restrict__ tab, uint32_t entries, int32_t adj, uint32_t* __restrict__ exc) {
  const uint32_t half = entries >> 1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < half; i += gridDim.x * blockDim.x)
    if (tab[i + half] != -tab[i] - adj) exc[1 + atomicAdd(exc, 1u)] = i;
}

// ... and the samples that read them: for every window of the launch, exception i and odd harmonic k, the sample
// pair (n, n + N/2) with k*n == i (mod N) is recomputed from the table itself (the general kernel's body).
__global__ void __launch_bounds__(256)
k_inq_patch(const WinRec* __restrict__ recs, const uint32_t* __restrict__ win_rec, uint32_t w_first, uint32_t nwin,
            uint32_t pw, uint32_t m, const uint32_t* __restrict__ exc, int32_t* __restrict__ out) {
  const uint32_t nexc = exc[0];
  const uint32_t nodd = m / 2;                         // odd harmonics 1, 3, 5 below m
  const uint64_t work = (uint64_t)nexc * nodd * nwin;
  const uint32_t nmask = (1u << pw) - 1u, half = 1u << (pw - 1);
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t w = (uint32_t)(t % nwin);
    const uint32_t r2 = (uint32_t)(t / nwin);
    const uint32_t k = 2u * (r2 % nodd) + 1u;
    const uint32_t i = exc[1 + r2 / nodd];
    // inverse of the odd k modulo 2^32 (Newton), then n = i * k^-1 mod N
    uint32_t inv = k;
    inv *= 2u - k * inv; inv *= 2u - k * inv; inv *= 2u - k * inv; inv *= 2u - k * inv;
    const uint32_t n = (i * inv) & nmask;
    const WinRec& r = recs[win_rec ? win_rec[w_first + w] : 0u];
    int32_t* o = out + ((size_t)w << pw);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t p = (n + (h ? half : 0u)) & nmask;   // phase index of the sample
      o[(p - r.n_first) & nmask] = synth_sample(r, p);
    }
  }
}

// The apply step for windows the fused kernel (k_synth_group) does not take (TAYLOR, the input-quadrant
// CORDICs, 64-bit tails ...): y[f*N + n] = x[f*N + n] * w[n] from a window generated into scratch memory.
// mode 1: the exact product DAT_Q (int64); mode 2: the entities' rounded slice of it (int32) - see group_epilogue.
__global__ void __launch_bounds__(256)
k_apply_mul(const int32_t* __restrict__ x, const int32_t* __restrict__ w, void* __restrict__ y, uint64_t n, uint64_t frames,
            int mode, int dw) {
  const int xsh = 32 - dw;
  const uint64_t total = n * frames;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const int32_t xv = (int32_t)((uint32_t)__ldcs(x + i) << xsh) >> xsh;
    const int64_t p = (int64_t)xv * (int64_t)__ldg(w + (i % n));
    if (mode == 1) {
      __stcs(reinterpret_cast<long long*>(y) + i, (long long)p);
    } else {
      const int64_t r = wrapb(p >> (dw - 2), dw + 1);
      __stcs(reinterpret_cast<int32_t*>(y) + i, (int32_t)wrapb((r >> 1) + (r & 1), dw));
    }
  }
}

// -------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------
cudaError_t launch_apply_mul(const int32_t* x, const int32_t* w, void* y, uint64_t n, uint64_t frames, int mode, int dw,
                             cudaStream_t stream);
static int g_sm_count[64] = {0};

static int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!g_sm_count[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_sm_count[dev] = n;
  }
  return g_sm_count[dev];
}

int device_sm_count() { return sm_count(); }

// grid sized as a multiple of the SM count, capped at `per_sm` resident CTAs per SM
static unsigned grid_for(uint64_t ctas_needed, int per_sm) {
  const uint64_t cap = (uint64_t)sm_count() * (uint64_t)per_sm;
  if (ctas_needed == 0) return 1;
  return (unsigned)(ctas_needed < cap ? ctas_needed : cap);
}

cudaError_t launch_table_build(const TabJob* jobs_dev, int njobs, uint32_t total_work, const I2* rom_dev,
                               cudaStream_t stream) {
  if (!total_work) return cudaSuccess;
  const unsigned grid = grid_for(((uint64_t)total_work + 255) / 256, 8);
  k_table_build<<<grid, 256, 0, stream>>>(jobs_dev, njobs, total_work, rom_dev);
  return cudaGetLastError();
}

cudaError_t launch_table_build_unrolled(const TabJob& j, cudaStream_t stream, int ctas_per_sm) {
  if (!j.work) return cudaSuccess;
  const unsigned grid = grid_for(((uint64_t)j.work / (j.sp.kind == SRC_INQ ? 2 : 1) + 255) / 256, ctas_per_sm);
  if (j.sp.kind == SRC_INQ) {
    switch (j.sp.n_xy) {
      case 16: k_table_build_inq_u<16><<<grid, 256, 0, stream>>>(j); break;
      case 17: k_table_build_inq_u<17><<<grid, 256, 0, stream>>>(j); break;
      case 24: k_table_build_inq_u<24><<<grid, 256, 0, stream>>>(j); break;
      case 32: k_table_build_inq_u<32><<<grid, 256, 0, stream>>>(j); break;
      default: return cudaErrorInvalidValue;
    }
  }
  else if (j.fast == TABCORE_32BIAS && j.sp.n_xy == 31) k_table_build_u<31, true><<<grid, 256, 0, stream>>>(j);
  else if (j.fast == TABCORE_32 && j.sp.n_xy == 15) k_table_build_u<15, false><<<grid, 256, 0, stream>>>(j);
  else if (j.fast == TABCORE_32 && j.sp.n_xy == 16) k_table_build_u<16, false><<<grid, 256, 0, stream>>>(j);
  else if (j.fast == TABCORE_32 && j.sp.n_xy == 23) k_table_build_u<23, false><<<grid, 256, 0, stream>>>(j);
  else return cudaErrorInvalidValue;
  return cudaGetLastError();
}

cudaError_t launch_synth(const SynthArgs& a, cudaStream_t stream) {
  if (!a.npieces && !a.flat_count) return cudaSuccess;
  const uint64_t ntiles = a.npieces ? (uint64_t)a.piece_tile0[a.npieces] : (a.flat_count + kTile - 1) / kTile;
  if (!ntiles) return cudaSuccess;
  const unsigned grid = grid_for((ntiles + kSynthWarps - 1) / kSynthWarps, 8);
  if (a.pack16) k_synth<short><<<grid, kSynthThreads, 0, stream>>>(a);
  else k_synth<int32_t><<<grid, kSynthThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

size_t bank_smem_limit() { return 192u * 1024u; }

template <int M, int TAB, int PAIR, bool W64, typename OutT = int32_t>
static cudaError_t launch_bank_t(const BankArgs& a, unsigned grid, size_t smem, cudaStream_t stream, bool pdl) {
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (smem > 48 * 1024 && dev >= 0 && dev < 64 && !attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_synth_bank<M, TAB, PAIR, W64, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)bank_smem_limit());
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  if (pdl) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kBankThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_synth_bank<M, TAB, PAIR, W64, OutT>, a);
  }
  k_synth_bank<M, TAB, PAIR, W64, OutT><<<grid, kBankThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

template <int M, bool W64>
static cudaError_t launch_bank_m(const BankArgs& a, int tab, bool pair, unsigned grid, size_t smem,
                                 cudaStream_t stream, bool pdl) {
  if (pair && a.sh.pair_adj) {       // ones'-complement pairing (input-quadrant CORDICs): 32-bit tail, no half-period staging
    if (W64 || tab == TAB_SMEM_HALF) return cudaErrorInvalidValue;
    return tab == TAB_SMEM_FULL ? launch_bank_t<M, TAB_SMEM_FULL, 2, false>(a, grid, smem, stream, pdl)
                                : launch_bank_t<M, TAB_GLOBAL, 2, false>(a, grid, 0, stream, pdl);
  }
  if (tab == TAB_SMEM_FULL) return pair ? launch_bank_t<M, TAB_SMEM_FULL, 1, W64>(a, grid, smem, stream, pdl)
                                        : launch_bank_t<M, TAB_SMEM_FULL, 0, W64>(a, grid, smem, stream, pdl);
  if (tab == TAB_SMEM_HALF) return launch_bank_t<M, TAB_SMEM_HALF, 1, W64>(a, grid, smem, stream, pdl);
  return pair ? launch_bank_t<M, TAB_GLOBAL, 1, W64>(a, grid, 0, stream, pdl)
              : launch_bank_t<M, TAB_GLOBAL, 0, W64>(a, grid, 0, stream, pdl);
}

// packed output: 2- and 3-term banks only (what DAT_WIDTH <= 16 windows are in practice; the rest of a packed
// plan goes through k_synth_group / k_synth), 32-bit tail, no ones'-complement pairing
template <int M>
static cudaError_t launch_bank_m16(const BankArgs& a, int tab, bool pair, unsigned grid, size_t smem, cudaStream_t stream,
                                   bool pdl) {
  if (tab == TAB_SMEM_FULL) return pair ? launch_bank_t<M, TAB_SMEM_FULL, 1, false, short>(a, grid, smem, stream, pdl)
                                        : launch_bank_t<M, TAB_SMEM_FULL, 0, false, short>(a, grid, smem, stream, pdl);
  if (tab == TAB_SMEM_HALF) return launch_bank_t<M, TAB_SMEM_HALF, 1, false, short>(a, grid, smem, stream, pdl);
  return pair ? launch_bank_t<M, TAB_GLOBAL, 1, false, short>(a, grid, 0, stream, pdl)
              : launch_bank_t<M, TAB_GLOBAL, 0, false, short>(a, grid, 0, stream, pdl);
}

cudaError_t launch_synth_bank(const BankArgs& a, int tab, bool pair, cudaStream_t stream, bool pdl) {
  if (!a.nwin) return cudaSuccess;
  if (tab == TAB_SMEM_HALF && !pair) return cudaErrorInvalidValue;
  if (a.ntiles && (pair || tab != TAB_GLOBAL || a.nwin != 1)) return cudaErrorInvalidValue;
  const uint32_t log_tpw = a.sh.pw - kBankTileLog2 - (pair ? 1 : 0);
  const uint64_t units = a.ntiles ? (uint64_t)a.ntiles : (uint64_t)a.nwin << log_tpw;
  if (a.win_minor && (tab != TAB_GLOBAL || a.ntiles || units >> 32)) return cudaErrorInvalidValue;
  if (a.spread && (tab != TAB_GLOBAL || a.ntiles || a.win_minor || a.nwin != 1 || a.spread > (uint32_t)kBankWarps || units >> 32))
    return cudaErrorInvalidValue;
  const uint64_t ctas = a.spread ? (units + a.spread - 1) / a.spread : (units + kBankWarps - 1) / kBankWarps;
  const unsigned grid = (unsigned)(ctas < (uint64_t)sm_count() ? ctas : (uint64_t)sm_count());
  const size_t smem = tab == TAB_GLOBAL ? 0 : (size_t)a.sh.smem_words * sizeof(int32_t);
  const bool w64 = a.sh.acc64 != 0;
  if (a.pack16) {
    if (w64 || (pair && a.sh.pair_adj)) return cudaErrorInvalidValue;
    if (a.sh.m == 2) return launch_bank_m16<2>(a, tab, pair, grid, smem, stream, pdl);
    if (a.sh.m == 3) return launch_bank_m16<3>(a, tab, pair, grid, smem, stream, pdl);
    return cudaErrorInvalidValue;
  }
  switch (a.sh.m) {
    case 2: return w64 ? launch_bank_m<2, true>(a, tab, pair, grid, smem, stream, pdl) : launch_bank_m<2, false>(a, tab, pair, grid, smem, stream, pdl);
    case 3: return w64 ? launch_bank_m<3, true>(a, tab, pair, grid, smem, stream, pdl) : launch_bank_m<3, false>(a, tab, pair, grid, smem, stream, pdl);
    case 4: return w64 ? launch_bank_m<4, true>(a, tab, pair, grid, smem, stream, pdl) : launch_bank_m<4, false>(a, tab, pair, grid, smem, stream, pdl);
    case 5: return w64 ? launch_bank_m<5, true>(a, tab, pair, grid, smem, stream, pdl) : launch_bank_m<5, false>(a, tab, pair, grid, smem, stream, pdl);
    case 7: return w64 ? launch_bank_m<7, true>(a, tab, pair, grid, smem, stream, pdl) : launch_bank_m<7, false>(a, tab, pair, grid, smem, stream, pdl);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_direct_window(const DirectArgs& a, void* out, cudaStream_t stream) {
  if (!a.count) return cudaSuccess;
  const unsigned grid = grid_for((a.count / ((a.quad_adv >> 31) ? 4 : (a.pair_flip >> 31) ? 2 : 1) + 255) / 256, 8);
  const size_t smem = (size_t)a.rom_smem_entries * sizeof(I2);
  if (a.wp.elem64) k_direct_window<int64_t><<<grid, 256, smem, stream>>>(a, (int64_t*)out);
  else k_direct_window<int32_t><<<grid, 256, smem, stream>>>(a, (int32_t*)out);
  return cudaGetLastError();
}

template <int NXY>
static void launch_direct32_t(const Direct32Args& a, int32_t* out, unsigned grid, cudaStream_t stream) {
  k_direct32<NXY><<<grid, 256, 0, stream>>>(a, out);
}

cudaError_t launch_direct32(const Direct32Args& a_in, int32_t* out, cudaStream_t stream) {
  if (!a_in.count) return cudaSuccess;
  Direct32Args a = a_in;
  // four samples per evaluation set only from about one 256-thread CTA of pairs per SM up: below that
  // the kernel is a latency chain and more, lighter threads finish sooner (7-term DW 24, us per call
  // pairs / fours: N = 2^14 5.1 / 6.2, 2^16 6.2 / 6.2, 2^18 10.1 / 6.7)
  if (a.pair == 2 && a.count / 2 <= (uint64_t)sm_count() * 256u) a.pair = 1;
  const uint64_t items = a.count / (a.pair == 2 ? 4 : a.pair ? 2 : 1);
  a.narrow = items <= (uint64_t)sm_count() * 8u * 256u ? 1u : 0u;   // fewer items than thread slots: one per thread
  const unsigned grid = grid_for(((a.narrow ? items : (items + 3) / 4) + 255) / 256, 8);
  switch (a.p.n_xy) {  // DAT_WIDTH 8..31 (cordic_dds: DW-1 stages; HLS: NW stages)
#define BHW_D32(N) case N: launch_direct32_t<N>(a, out, grid, stream); break;
    BHW_D32(7) BHW_D32(8) BHW_D32(9) BHW_D32(10) BHW_D32(11) BHW_D32(12) BHW_D32(13) BHW_D32(14)
    BHW_D32(15) BHW_D32(16) BHW_D32(17) BHW_D32(18) BHW_D32(19) BHW_D32(20) BHW_D32(21) BHW_D32(22)
    BHW_D32(23) BHW_D32(24) BHW_D32(25) BHW_D32(26) BHW_D32(27) BHW_D32(28) BHW_D32(29) BHW_D32(30)
#undef BHW_D32
    default: launch_direct32_t<0>(a, out, grid, stream); break;
  }
  return cudaGetLastError();
}

cudaError_t launch_direct_taylor(const DirectTayArgs& a, int32_t* out, cudaStream_t stream) {
  if (!a.count) return cudaSuccess;
  const bool quad = a.pair == 2 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  const unsigned grid = grid_for(((a.count / (quad ? 4 : a.pair ? 2 : 1) + 3) / 4 + 255) / 256, 8);
  const size_t smem = (size_t)a.p.rom_entries * sizeof(I2);  // <= 32 KB (LUT_SIZE <= 12)
  if (a.p.tmode == TMODE_ROM) k_direct_taylor<TMODE_ROM><<<grid, 256, smem, stream>>>(a, out);
  else if (a.p.tmode == TMODE_DSP) k_direct_taylor<TMODE_DSP><<<grid, 256, smem, stream>>>(a, out);
  else k_direct_taylor<TMODE_WIDE><<<grid, 256, smem, stream>>>(a, out);
  return cudaGetLastError();
}

cudaError_t launch_atan2(const Atan2Params& p, const int32_t* x, const int32_t* y, int32_t* phi, uint64_t count,
                         cudaStream_t stream, uint64_t avail) {
  if (avail < count) avail = count;
  if (!count) return cudaSuccess;
  const bool vec = p.fast32 && !p.skew && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                                 reinterpret_cast<uintptr_t>(phi)) & 15) == 0;
  const unsigned gridv = grid_for((count / 4 + 255) / 256 + 1, 8);
  if (vec && p.aw == 16) k_atan2_u<16><<<gridv, 256, 0, stream>>>(p, x, y, phi, count);
  else if (vec && p.aw == 24) k_atan2_u<24><<<gridv, 256, 0, stream>>>(p, x, y, phi, count);
  else if (vec && p.aw == 20) k_atan2_u<20><<<gridv, 256, 0, stream>>>(p, x, y, phi, count);
  else if (vec && p.aw == 12) k_atan2_u<12><<<gridv, 256, 0, stream>>>(p, x, y, phi, count);
  else {
    const unsigned grid = grid_for((count + 255) / 256, 8);
    k_atan2<<<grid, 256, 0, stream>>>(p, x, y, phi, count, avail);
  }
  return cudaGetLastError();
}

cudaError_t launch_inq_exceptions(const int32_t* tab, uint32_t entries, int32_t adj, uint32_t* exc, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(exc, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return e;
  const unsigned grid = grid_for(((uint64_t)entries / 2 + 255) / 256, 8);
  k_inq_exceptions<<<grid, 256, 0, stream>>>(tab, entries, adj, exc);
  return cudaGetLastError();
}

cudaError_t launch_inq_patch(const BankArgs& a, const uint32_t* exc, cudaStream_t stream) {
  // the exception count lives on the device: a small fixed grid strides over whatever there is
  k_inq_patch<<<(unsigned)sm_count(), 256, 0, stream>>>(a.recs, a.win_rec, a.w_first, a.nwin, a.sh.pw, a.sh.m, exc, a.out);
  return cudaGetLastError();
}

cudaError_t launch_apply_mul(const int32_t* x, const int32_t* w, void* y, uint64_t n, uint64_t frames, int mode, int dw,
                             cudaStream_t stream) {
  if (!n || !frames) return cudaSuccess;
  const unsigned grid = grid_for((n * frames + 255) / 256, 8);
  k_apply_mul<<<grid, 256, 0, stream>>>(x, w, y, n, frames, mode, dw);
  return cudaGetLastError();
}

cudaError_t launch_sincos(const SinCosArgs& a, void* out_sin, void* out_cos, bool elem64, cudaStream_t stream) {
  if (!a.count) return cudaSuccess;
  const unsigned grid = grid_for((a.count / (a.quad ? 4 : 1) + 255) / 256, 8);
  if (elem64) k_sincos<int64_t><<<grid, 256, 0, stream>>>(a, (int64_t*)out_sin, (int64_t*)out_cos);
  else k_sincos<int32_t><<<grid, 256, 0, stream>>>(a, (int32_t*)out_sin, (int32_t*)out_cos);
  return cudaGetLastError();
}

}  // namespace bhw
