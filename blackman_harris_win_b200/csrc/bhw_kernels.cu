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

#include "bhw_device.cuh"
#include "bhw_launch.h"

namespace bhw {

// -------------------------------------------------------------------------------------------
// stage 1: trig tables
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_table_build(const TabJob* __restrict__ jobs, int njobs, uint32_t total_work,
              const I2* __restrict__ rom) {
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

// -------------------------------------------------------------------------------------------
// stage 2: synthesis
// -------------------------------------------------------------------------------------------
constexpr int kSynthThreads = 256;
constexpr int kSynthWarps = kSynthThreads / 32;
constexpr int kTile = 128;  // samples per warp tile: 4 per lane, lane-interleaved

template <int M>
__device__ __forceinline__ void synth_tile(const WinRec& r, uint32_t n, int32_t* __restrict__ out,
                                           uint32_t valid) {
  // lane owns samples n, n+32, n+64, n+96 of the tile; `valid` = samples left from this lane's first
  if (r.flags & WR_ACC64) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((uint32_t)(32 * j) < valid) out[32 * j] = synth_sample64<M>(r, n + 32 * j);
  } else if (r.flags & WR_WIDE) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((uint32_t)(32 * j) < valid) out[32 * j] = synth_sample32<M, true>(r, n + 32 * j);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if ((uint32_t)(32 * j) < valid) out[32 * j] = synth_sample32<M, false>(r, n + 32 * j);
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

__global__ void __launch_bounds__(kSynthThreads)
k_synth(SynthArgs a) {
  __shared__ WinRec s_rec[kSynthWarps];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WinRec& rec = s_rec[warp];
  int cur_w = -1, cur_r = -1;
  const uint64_t ntiles = (a.flat_count + kTile - 1) / kTile;
  int32_t* const out = reinterpret_cast<int32_t*>(a.out);
  for (uint64_t tile = (uint64_t)blockIdx.x * kSynthWarps + warp; tile < ntiles;
       tile += (uint64_t)gridDim.x * kSynthWarps) {
    const uint64_t o0 = tile * kTile;               // first output index of the tile
    const uint64_t f0 = a.flat_begin + o0;          // its flat sample index
    const uint64_t left = a.flat_count - o0;        // samples left in the request
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
      int32_t* o = out + o0 + lane;
      const uint32_t valid = tile_n > (uint32_t)lane ? tile_n - lane : 0;
      if (rec.flags & WR_GENERIC) {
        const GenRec& g = a.gens[rec.gen_idx];
        const uint64_t nmask = (1ull << g.wp.pw) - 1;
        for (int j = 0; j < 4; ++j)
          if ((uint32_t)(32 * j) < valid)
            o[32 * j] = (int32_t)direct_sample_generic(g.wp, g.src, a.rom + g.rom_off,
                                                       (uint64_t)(n + 32 * j) & nmask);
      } else {
        switch (rec.m) {
          case 2: synth_tile<2>(rec, n, o, valid); break;
          case 3: synth_tile<3>(rec, n, o, valid); break;
          case 4: synth_tile<4>(rec, n, o, valid); break;
          case 5: synth_tile<5>(rec, n, o, valid); break;
          default: synth_tile<7>(rec, n, o, valid); break;
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
        out[o0 + i] = v;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// direct evaluation (one thread per sample) and the sin/cos entry
// -------------------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256)
k_direct_window(DirectArgs a, OutT* __restrict__ out) {
  // taylor_sincos keeps its quarter-wave ROM in shared memory when it fits
  extern __shared__ I2 s_rom[];
  const I2* rom = a.rom;
  if (a.rom_smem_entries) {
    for (uint32_t i = threadIdx.x; i < a.rom_smem_entries; i += blockDim.x) s_rom[i] = a.rom[i];
    __syncthreads();
    rom = s_rom;
  }
  const uint64_t nmask = (1ull << a.wp.pw) - 1;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.count;
       j += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t n = (a.n_first + j) & nmask;
    out[j] = (OutT)direct_sample_generic(a.wp, a.src, rom, n);
  }
}

template <typename OutT>
__global__ void __launch_bounds__(256)
k_sincos(SinCosArgs a, OutT* __restrict__ out_sin, OutT* __restrict__ out_cos) {
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < a.count;
       j += (uint64_t)gridDim.x * blockDim.x) {
    int64_t s, c;
    eval_source_generic(a.src, a.rom, a.n_first + j, s, c);
    if (out_sin) out_sin[j] = (OutT)s;
    if (out_cos) out_cos[j] = (OutT)c;
  }
}

// -------------------------------------------------------------------------------------------
// launchers
// -------------------------------------------------------------------------------------------
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

cudaError_t launch_synth(const SynthArgs& a, cudaStream_t stream) {
  if (!a.flat_count) return cudaSuccess;
  const uint64_t ntiles = (a.flat_count + kTile - 1) / kTile;
  const unsigned grid = grid_for((ntiles + kSynthWarps - 1) / kSynthWarps, 8);
  k_synth<<<grid, kSynthThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_direct_window(const DirectArgs& a, void* out, cudaStream_t stream) {
  if (!a.count) return cudaSuccess;
  const unsigned grid = grid_for((a.count + 255) / 256, 8);
  const size_t smem = (size_t)a.rom_smem_entries * sizeof(I2);
  if (a.wp.elem64) k_direct_window<int64_t><<<grid, 256, smem, stream>>>(a, (int64_t*)out);
  else k_direct_window<int32_t><<<grid, 256, smem, stream>>>(a, (int32_t*)out);
  return cudaGetLastError();
}

cudaError_t launch_sincos(const SinCosArgs& a, void* out_sin, void* out_cos, bool elem64, cudaStream_t stream) {
  if (!a.count) return cudaSuccess;
  const unsigned grid = grid_for((a.count + 255) / 256, 8);
  if (elem64) k_sincos<int64_t><<<grid, 256, 0, stream>>>(a, (int64_t*)out_sin, (int64_t*)out_cos);
  else k_sincos<int32_t><<<grid, 256, 0, stream>>>(a, (int32_t*)out_sin, (int32_t*)out_cos);
  return cudaGetLastError();
}

}  // namespace bhw
