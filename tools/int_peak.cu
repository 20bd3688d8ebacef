// int_peak.cu - measured INT32 issue peaks of the GPU, the roofline denominator of the
// one-thread-per-sample direct kernels ("INT32 peak measured by a
// dependent-free IADD3/SHF micro-benchmark on the same box").
//
// Three kernels, each thread runs ILP independent register chains of the same instruction mix
// the CORDIC stages use:
//   alu : SHF + LOP3 only                  (alu pipe; an add would be fused into LEA or moved to IMAD)
//   fma : IMAD only                        (fma pipe)
//   mix : 1 SHF/LOP3 : 1 IMAD interleaved  (both pipes; the shape of k_direct32's stage)
// Prints one JSON line with lane-ops/s (32 x warp instructions per second).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/int_peak tools/int_peak.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int kIlp = 8;
constexpr int kInner = 64;  // instructions per chain and outer iteration

template <int MODE>
__global__ void __launch_bounds__(256) k_peak(int outer, uint32_t seed, uint32_t* sink) {
  uint32_t r[kIlp];
#pragma unroll
  for (int i = 0; i < kIlp; ++i) r[i] = seed + threadIdx.x * 977u + i * 131u;
  const uint32_t c = seed | 1u;
  for (int o = 0; o < outer; ++o) {
#pragma unroll
    for (int j = 0; j < kInner; ++j) {
#pragma unroll
      for (int i = 0; i < kIlp; ++i) {
        if (MODE == 0) {        // alu pipe: alternate funnel shift and logic op
          if (j & 1) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(c));
          else asm volatile("xor.b32 %0, %0, %1;" : "+r"(r[i]) : "r"(c));
        } else if (MODE == 1) { // fma pipe: integer multiply-add
          asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(c));
        } else if (MODE == 2) { // both pipes, 1:1
          if (j & 1) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(c));
          else asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(c));
        } else if (MODE == 3) { // multiply-high with addend (IMAD.HI): the table path's b_k
          asm volatile("mad.hi.s32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(c));
        } else if (MODE == 4) { // 32x32+64 (IMAD.WIDE), result feeds the next multiplicand
          asm volatile("{ .reg .b64 t; .reg .b32 lo, hi; mul.wide.s32 t, %0, %1; "
                       "mov.b64 {lo, hi}, t; xor.b32 %0, lo, hi; }" : "+r"(r[i]) : "r"(c));
        } else {                // IMAD.HI : SHF 1:1
          if (j & 1) asm volatile("mad.hi.s32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(c));
          else asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(c));
        }
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < kIlp; ++i) acc ^= r[i];
  if (acc == 0x12345u) sink[0] = acc;  // keeps the chains alive; practically never true
}

template <int MODE>
static double run(int sms, uint32_t* sink) {
  const int outer = 2000, grid = sms * 8;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    k_peak<MODE><<<grid, 256>>>(outer, 12345u + rep, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double ops = (double)grid * 256.0 * outer * kInner * kIlp;
    const double rate = ops / (ms * 1e-3);
    if (rep && rate > best) best = rate;  // rep 0 is the warm-up
  }
  return best;
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  uint32_t* sink = nullptr;
  cudaMalloc(&sink, 4);
  const double alu = run<0>(sms, sink), fma = run<1>(sms, sink), mix = run<2>(sms, sink);
  const double hi = run<3>(sms, sink), wide = run<4>(sms, sink), himix = run<5>(sms, sink);
  if (cudaDeviceSynchronize() != cudaSuccess) { fprintf(stderr, "kernel failed\n"); return 1; }
  const double per_clk = 1.0 / ((double)sms * khz * 1e3);
  printf("{\"int32_alu_tops\": %.2f, \"int32_fma_tops\": %.2f, \"int32_mix_tops\": %.2f, "
         "\"alu_lanes_per_clk_sm\": %.1f, \"fma_lanes_per_clk_sm\": %.1f, \"mix_lanes_per_clk_sm\": %.1f, "
         "\"imad_hi_tops\": %.2f, \"imad_wide_plus_lop_tops\": %.2f, \"imad_hi_shf_mix_tops\": %.2f, \"sms\": %d, \"sm_mhz_max\": %d, \"how\": \"tools/int_peak.cu: 8 independent chains per thread, "
         "8 CTAs x 256 threads per SM, best of 4 after warm-up, CUDA events\"}\n",
         alu / 1e12, fma / 1e12, mix / 1e12, alu * per_clk, fma * per_clk, mix * per_clk, hi / 1e12,
         wide / 1e12, himix / 1e12, sms, khz / 1000);
  cudaFree(sink);
  return 0;
}
