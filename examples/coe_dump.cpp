// coe_dump.cpp - the reference's table-dump flow (cpp/cordic_sincos.cpp:131-139 writes "s c" lines to
// coe.dat; hls/windows/window_test.cpp:93-99 writes one coefficient per line) on top of bhw.hpp.
//
//   g++ -std=c++17 -I include examples/coe_dump.cpp -L blackman_harris_win_b200 -lbhw -o coe_dump
//   LD_LIBRARY_PATH=blackman_harris_win_b200 ./coe_dump BH4TERM 16 17 6 > bh4.dat
//
// Arguments: WIN_TYPE PHI_WIDTH DAT_WIDTH variant [SIN_TYPE [LUT_SIZE]]; with "--check" it only
// validates the generics/ports (works without a GPU).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bhw.hpp"

int main(int argc, char** argv) {
  bool check_only = false;
  if (argc > 1 && !strcmp(argv[1], "--check")) { check_only = true; argv++; argc--; }
  if (argc < 5) {
    fprintf(stderr, "usage: coe_dump [--check] WIN_TYPE PHI_WIDTH DAT_WIDTH variant [SIN_TYPE [LUT_SIZE]]\n");
    return 2;
  }
  try {
    bhw::win_selector w(atoi(argv[2]), atoi(argv[3]), argv[1], argc > 5 ? argv[5] : "CORDIC",
                        argc > 6 ? atoi(argv[6]) : 9);
    const bhw_desc d = w.desc_variant(atoi(argv[4]));
    if (check_only) {
      printf("ok: %llu samples of %d bytes, AA0=%lld\n", (unsigned long long)w.length(), bhw_elem_bytes(&d),
             (long long)d.aa[0]);
      return 0;
    }
    if (bhw_elem_bytes(&d) == 4) for (int32_t v : w.stream(d)) printf("%d\n", v);
    else for (int64_t v : w.stream64(d)) printf("%lld\n", (long long)v);
  } catch (const bhw::error& e) {
    fprintf(stderr, "coe_dump: %s\n", e.what());
    return 1;
  }
  return 0;
}
