// store_pattern_probe.cu - what store patterns reach on this GPU with no arithmetic in front of them.
// Writes 1 GiB per launch with
//   * the bank kernel's exact pattern (148 persistent CTAs x 32 warps, 256-sample tiles, (n, n+N/2)
//     pairs, equal contiguous shares per CTA), 32-bit and 128-bit stores, four cache operators;
//   * the same tiles interleaved over the grid, launched one chunk per CTA (not persistent), and
//     handed out dynamically (per warp and per CTA) from a global counter;
//   * plain fills (persistent grid-stride and one thread per int4).
// Round-1 findings on B200 (GB/s): equal static shares 5.9 k, interleaved 6.1 k, one-shot CTAs or
// CTA-granular dynamic chunks 7.3 k, one thread per int4 7.37 k; the cache operator and the store
// width make no difference.  A store-only persistent kernel is limited by unequal per-SM store
// bandwidth (static shares finish up to 59 us apart).  k_synth_bank itself does NOT show that
// imbalance (its CTAs finish within 4 % of each other: it is paced by instruction issue, not by the
// store path), so dynamic scheduling was tried there and dropped - see DESIGN.md.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/store_pattern_probe tools/store_pattern_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kPw = 16, kTile = 256, kJ = 8;

// store flavour: 0 st.global.cs (streaming), 1 plain st.global, 2 st.global.cg, 3 st.global.wt
template <int OP, typename T>
__device__ __forceinline__ void st(T* p, T v) {
  if (OP == 0) __stcs(p, v);
  else if (OP == 1) *p = v;
  else if (OP == 2) __stcg(p, v);
  else __stwt(p, v);
}

template <int VEC, int OP>
__global__ void __launch_bounds__(1024, 1) k_bank_pattern(int32_t* out, uint32_t nwin, int32_t v) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t log_tpw = kPw - 8 - 1, half = 1u << (kPw - 1);
  const uint64_t U = (uint64_t)nwin << log_tpw;
  const uint64_t u0 = U * blockIdx.x / gridDim.x, u1 = U * (blockIdx.x + 1) / gridDim.x;
  for (uint64_t u = u0 + warp; u < u1; u += 32) {
    const uint32_t w = (uint32_t)(u >> log_tpw), t = (uint32_t)u & ((1u << log_tpw) - 1);
    int32_t* o = out + ((uint64_t)w << kPw) + t * kTile;
    if (VEC == 1) {
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        st<OP>(o + lane + 32 * j, v + j);
        st<OP>(o + half + lane + 32 * j, v - j);
      }
    } else {
#pragma unroll
      for (int j = 0; j < kJ / 4; ++j) {
        st<OP>(reinterpret_cast<int4*>(o + 4 * lane + 128 * j), make_int4(v, v + 1, v + 2, v + j));
        st<OP>(reinterpret_cast<int4*>(o + half + 4 * lane + 128 * j), make_int4(v, v - 1, v - 2, v - j));
      }
    }
  }
}

// MAP 0: persistent, tiles interleaved over the grid (tile u = i*grid*32 + cta*32 + warp): the whole GPU
//        sweeps the output front to back;  MAP 1: one tile per warp, CTAs launched in order (not persistent)
// PAIR: the lane also writes the partner tile half a window later (as k_synth_bank does)
template <int MAP, bool PAIR, int OP>
__global__ void __launch_bounds__(1024, 1) k_tiles(int32_t* out, uint32_t nwin, int32_t v) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t log_tpw = kPw - 8 - (PAIR ? 1 : 0), half = 1u << (kPw - 1);
  const uint64_t U = (uint64_t)nwin << log_tpw;
  const uint64_t first = (uint64_t)blockIdx.x * 32 + warp;
  const uint64_t stride = MAP == 0 ? (uint64_t)gridDim.x * 32 : U;
  for (uint64_t u = first; u < U; u += stride) {
    const uint32_t w = (uint32_t)(u >> log_tpw), t = (uint32_t)u & ((1u << log_tpw) - 1);
    int32_t* o = out + ((uint64_t)w << kPw) + t * kTile;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      st<OP>(o + lane + 32 * j, v + j);
      if (PAIR) st<OP>(o + half + lane + 32 * j, v - j);
    }
  }
}

// persistent CTAs, every warp takes groups of G tile pairs from one global counter (first group static);
// the last warp to run dry resets the counters for the next launch
template <int G>
__global__ void __launch_bounds__(1024, 1) k_dynamic(int32_t* out, uint32_t nwin, int32_t v, unsigned* ctr) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t log_tpw = kPw - 8 - 1, half = 1u << (kPw - 1);
  const uint32_t groups = (uint32_t)(((uint64_t)nwin << log_tpw) / G);
  const uint32_t nwarps = gridDim.x * 32;
  uint32_t g = blockIdx.x * 32 + warp;
  while (g < groups) {
    uint32_t nxt = 0;
    if (lane == 0) nxt = atomicAdd(ctr, 1u) + nwarps;   // prefetch the next group
#pragma unroll 1
    for (int i = 0; i < G; ++i) {
      const uint32_t u = g * G + i;
      const uint32_t w = u >> log_tpw, t = u & ((1u << log_tpw) - 1);
      int32_t* o = out + ((uint64_t)w << kPw) + t * kTile;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        __stcs(o + lane + 32 * j, v + j);
        __stcs(o + half + lane + 32 * j, v - j);
      }
    }
    g = __shfl_sync(0xffffffffu, nxt, 0);
  }
  if (lane == 0 && atomicAdd(ctr + 1, 1u) == nwarps - 1) { ctr[0] = 0; ctr[1] = 0; }
}

// persistent CTAs, CTA-granular dynamic scheduling: one thread takes the next chunk (32*C tile pairs) from
// a global counter while the CTA works on the current one; one __syncthreads per chunk
template <int C>
__global__ void __launch_bounds__(1024, 1) k_dynamic_cta(int32_t* out, uint32_t nwin, int32_t v, unsigned* ctr) {
  __shared__ uint32_t s_next[2];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t log_tpw = kPw - 8 - 1, half = 1u << (kPw - 1);
  const uint32_t chunks = (uint32_t)(((uint64_t)nwin << log_tpw) / (32 * C));
  uint32_t c = blockIdx.x, it = 0;
  while (c < chunks) {
    if (threadIdx.x == 0) s_next[it & 1] = atomicAdd(ctr, 1u) + gridDim.x;
#pragma unroll 1
    for (int i = 0; i < C; ++i) {
      const uint32_t u = (c * C + i) * 32 + warp;
      const uint32_t w = u >> log_tpw, t = u & ((1u << log_tpw) - 1);
      int32_t* o = out + ((uint64_t)w << kPw) + t * kTile;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        __stcs(o + lane + 32 * j, v + j);
        __stcs(o + half + lane + 32 * j, v - j);
      }
    }
    __syncthreads();
    c = s_next[it & 1];
    ++it;
  }
  if (threadIdx.x == 0 && atomicAdd(ctr + 1, 1u) == gridDim.x - 1) { ctr[0] = 0; ctr[1] = 0; }
}

// static contiguous ranges (the bank kernel's mapping); records each CTA's finish time
__global__ void __launch_bounds__(1024, 1) k_static_timed(int32_t* out, uint32_t nwin, int32_t v, long long* t_end) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t log_tpw = kPw - 8 - 1, half = 1u << (kPw - 1);
  const uint64_t U = (uint64_t)nwin << log_tpw;
  const uint64_t u0 = U * blockIdx.x / gridDim.x, u1 = U * (blockIdx.x + 1) / gridDim.x;
  for (uint64_t u = u0 + warp; u < u1; u += 32) {
    const uint32_t w = (uint32_t)(u >> log_tpw), t = (uint32_t)u & ((1u << log_tpw) - 1);
    int32_t* o = out + ((uint64_t)w << kPw) + t * kTile;
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      __stcs(o + lane + 32 * j, v + j);
      __stcs(o + half + lane + 32 * j, v - j);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); t_end[blockIdx.x] = t; }
}

template <int OP>
__global__ void __launch_bounds__(256) k_fill(int4* out, uint64_t n4, int32_t v) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x)
    st<OP>(out + i, make_int4(v, v, v, v));
}

template <typename F>
static double best_gbs(F launch, size_t bytes) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0;
  for (int rep = 0; rep < 12; ++rep) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double g = bytes / (ms * 1e-3) / 1e9;
    if (rep >= 2 && g > best) best = g;
  }
  return best;
}

int main() {
  const uint32_t nwin = 4096;
  const size_t bytes = (size_t)nwin << (kPw + 2);
  int32_t* out = nullptr;
  if (cudaMalloc(&out, bytes) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double r[4][4];
#define ROW(OP)                                                                                              \
  r[OP][0] = best_gbs([&] { k_bank_pattern<1, OP><<<sms, 1024>>>(out, nwin, 3); }, bytes);                      \
  r[OP][1] = best_gbs([&] { k_bank_pattern<4, OP><<<sms, 1024>>>(out, nwin, 3); }, bytes);                      \
  r[OP][2] = best_gbs([&] { k_fill<OP><<<sms * 8, 256>>>((int4*)out, bytes / 16, 3); }, bytes);                 \
  r[OP][3] = best_gbs([&] { k_fill<OP><<<(unsigned)(bytes / 16 / 256), 256>>>((int4*)out, bytes / 16, 3); }, bytes);
  ROW(0) ROW(1) ROW(2) ROW(3)
  if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 1; }
  const uint32_t Up = nwin << (kPw - 9), Uu = nwin << (kPw - 8);
  const double t0 = best_gbs([&] { k_tiles<0, true, 0><<<sms, 1024>>>(out, nwin, 3); }, bytes);
  const double t1 = best_gbs([&] { k_tiles<0, false, 0><<<sms, 1024>>>(out, nwin, 3); }, bytes);
  const double t2 = best_gbs([&] { k_tiles<1, true, 0><<<Up / 32, 1024>>>(out, nwin, 3); }, bytes);
  const double t3 = best_gbs([&] { k_tiles<1, false, 0><<<Uu / 32, 1024>>>(out, nwin, 3); }, bytes);
  const double t4 = best_gbs([&] { k_tiles<0, true, 0><<<sms * 2, 1024>>>(out, nwin, 3); }, bytes);
  printf("{\"interleaved_persistent_pair\": %.1f, \"interleaved_persistent_nopair\": %.1f, \"oneshot_pair\": %.1f, "
         "\"oneshot_nopair\": %.1f, \"interleaved_2cta_per_sm_pair\": %.1f}\n", t0, t1, t2, t3, t4);
  unsigned* ctr = nullptr;
  cudaMalloc(&ctr, 8);
  cudaMemset(ctr, 0, 8);
  const double d1 = best_gbs([&] { k_dynamic<1><<<sms, 1024>>>(out, nwin, 3, ctr); }, bytes);
  const double d2 = best_gbs([&] { k_dynamic<2><<<sms, 1024>>>(out, nwin, 3, ctr); }, bytes);
  const double d4 = best_gbs([&] { k_dynamic<4><<<sms, 1024>>>(out, nwin, 3, ctr); }, bytes);
  const double d8 = best_gbs([&] { k_dynamic<8><<<sms, 1024>>>(out, nwin, 3, ctr); }, bytes);
  const double d16 = best_gbs([&] { k_dynamic<16><<<sms, 1024>>>(out, nwin, 3, ctr); }, bytes);
  printf("{\"dynamic_warp_groups\": {\"G1\": %.1f, \"G2\": %.1f, \"G4\": %.1f, \"G8\": %.1f, \"G16\": %.1f}}\n", d1, d2, d4, d8, d16);
  const double c1 = best_gbs([&] { k_dynamic_cta<1><<<sms, 1024>>>(out, nwin, 3, ctr); }, bytes);
  const double c2 = best_gbs([&] { k_dynamic_cta<2><<<sms, 1024>>>(out, nwin, 3, ctr); }, bytes);
  const double c4 = best_gbs([&] { k_dynamic_cta<4><<<sms, 1024>>>(out, nwin, 3, ctr); }, bytes);
  printf("{\"dynamic_cta_chunks\": {\"C1\": %.1f, \"C2\": %.1f, \"C4\": %.1f}}\n", c1, c2, c4);
  {
    long long* t_end = nullptr;
    cudaMalloc(&t_end, sizeof(long long) * sms);
    for (int rep = 0; rep < 3; ++rep) k_static_timed<<<sms, 1024>>>(out, nwin, 3, t_end);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, t_end, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long lo = h[0], hi = h[0];
    for (int i = 1; i < sms; ++i) { if (h[i] < lo) lo = h[i]; if (h[i] > hi) hi = h[i]; }
    int early = 0;
    for (int i = 0; i < sms; ++i) if (hi - h[i] > 20000) early++;
    printf("{\"static_cta_finish_spread_us\": %.1f, \"ctas_done_20us_before_last\": %d}\n", (hi - lo) / 1e3, early);
  }
  const char* names[4] = {"cs", "plain", "cg", "wt"};
  printf("{");
  for (int op = 0; op < 4; ++op)
    printf("\"%s\": {\"bank_pattern_st32_gbs\": %.1f, \"bank_pattern_st128_gbs\": %.1f, \"persistent_fill_st128_gbs\": %.1f, "
           "\"one_thread_per_int4_fill_gbs\": %.1f}, ", names[op], r[op][0], r[op][1], r[op][2], r[op][3]);
  printf("\"bytes\": %zu, \"how\": \"best of 10, CUDA events, 1 GiB per launch\"}\n", bytes);
  cudaFree(out);
  return 0;
}
