// call_latency.cpp - microseconds per call of the one-shot entry point, straight through the C ABI
// (no Python): BASELINE config 1 is launch-bound (1024 samples = 4 KiB), so the figure that matters
// there is how long one bhw_generate takes, back to back on one stream.  Prints one JSON line.
//
//   g++ -std=c++17 -O2 -I include tools/call_latency.cpp -L blackman_harris_win_b200 -lbhw \
//       -Wl,-rpath,$PWD/blackman_harris_win_b200 -L/usr/local/cuda/lib64 -lcudart -o tools/call_latency
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <cuda_runtime_api.h>

#include "bhw.h"

static double time_calls(const bhw_desc& d, void* out, uint64_t n, int reps) {
  for (int i = 0; i < 20; i++) bhw_generate(&d, out, 0, n, nullptr);
  cudaDeviceSynchronize();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; i++) {
    if (bhw_generate(&d, out, 0, n, nullptr) != BHW_OK) return -1.0;
  }
  cudaDeviceSynchronize();
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}

static double time_host_calls(const bhw_desc& d, void* out_pinned, uint64_t n, int reps) {
  for (int i = 0; i < 10; i++) bhw_generate_host(&d, out_pinned, 0, n);
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < reps; i++) {
    if (bhw_generate_host(&d, out_pinned, 0, n) != BHW_OK) return -1.0;
  }
  return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
}

static bhw_desc make(int win_type, int pw, int dw, int variant, int sin_type) {
  bhw_desc d;
  memset(&d, 0, sizeof(d));
  d.win_type = win_type; d.phi_width = pw; d.dat_width = dw; d.sin_type = sin_type;
  int32_t wt = 0;
  bhw_quantize(variant, BHW_RULE_TB, dw, d.aa, &wt);
  return d;
}

int main(int argc, char** argv) {
  void* out = nullptr;
  if (cudaMalloc(&out, 64u << 20) != cudaSuccess) { fprintf(stderr, "no CUDA device\n"); return 1; }
  if (argc > 1 && !strcmp(argv[1], "--route")) {
    // where BHW_ALGO_AUTO should switch from the direct kernel to the table path: whole windows of
    // PHI_WIDTH 14..22 through each strategy, one JSON line per entity
    const struct { const char* name; int m, dw, variant; } ent[] = {{"hamming_dw16", 2, 16, 1}, {"bh3_dw16", 3, 16, 4},
      {"bh4_dw17", 4, 17, 6}, {"bh5_dw24", 5, 24, 9}, {"bh7_dw24", 7, 24, 10}};
    for (const auto& e : ent) {
      printf("{\"entity\": \"%s\", \"us_per_call [auto, table, direct]\": {", e.name);
      for (int pw = 14; pw <= 22; pw++) {
        bhw_desc d = make(e.m, pw, e.dw, e.variant, BHW_SIN_CORDIC);
        double t[3];
        const int algos[3] = {BHW_ALGO_AUTO, BHW_ALGO_TABLE, BHW_ALGO_DIRECT};
        for (int a = 0; a < 3; a++) { d.algo = algos[a]; t[a] = time_calls(d, out, 1ull << pw, 300); }
        printf("%s\"%d\": [%.2f, %.2f, %.2f]", pw > 14 ? ", " : "", pw, t[0], t[1], t[2]);
      }
      printf("}}\n");
    }
    cudaFree(out);
    return 0;
  }
  const bhw_desc c1 = make(2, 10, 16, 1, BHW_SIN_CORDIC);      // config 1: Hamming N=1024 DW=16
  const bhw_desc c2 = make(4, 16, 17, 6, BHW_SIN_CORDIC);      // config 2: BH4 N=65536 DW=17
  const bhw_desc c3 = make(7, 20, 32, 10, BHW_SIN_CORDIC48);   // config 3: BH7 N=1M DW=32 cordic_dds48
  const bhw_desc c3b = make(7, 20, 32, 10, BHW_SIN_CORDIC);    // config 3 with the entity's own cordic_dds
  const bhw_desc c4 = make(3, 24, 24, 3, BHW_SIN_TAYLOR);      // config 4: Blackman TAYLOR N=16M DW=24
  void* pinned = nullptr;
  cudaMallocHost(&pinned, 64u << 20);
  const double h1 = time_host_calls(c1, pinned, 1u << 10, 500), h2 = time_host_calls(c2, pinned, 1u << 16, 500);
  const double h3 = time_host_calls(c3b, pinned, 1u << 20, 100), h4 = time_host_calls(c4, pinned, 1u << 24, 50);
  printf("{\"us_per_host_call (bhw_generate_host, pinned destination, blocking)\": {\"cfg1_hamming_n1024\": %.2f, "
         "\"cfg2_bh4_n65536\": %.2f, \"cfg3_bh7_n1m_dds\": %.2f, \"cfg4_blackman_taylor_n16m\": %.2f}}\n", h1, h2, h3, h4);
  printf("{\"us_per_call\": {\"cfg1_hamming_n1024\": %.2f, \"cfg2_bh4_n65536\": %.2f, \"cfg3_bh7_n1m_dds48\": %.2f, "
         "\"cfg3_bh7_n1m_dds\": %.2f, \"cfg4_blackman_taylor_n16m\": %.2f}, "
         "\"how\": \"bhw_generate back to back on the default stream, wall clock over 2000/200 calls incl. the final "
         "synchronize; C ABI, no Python\"}\n",
         time_calls(c1, out, 1u << 10, 2000), time_calls(c2, out, 1u << 16, 2000), time_calls(c3, out, 1u << 20, 200),
         time_calls(c3b, out, 1u << 20, 200), time_calls(c4, out, 1u << 24, 200));
  cudaFree(out);
  return 0;
}
