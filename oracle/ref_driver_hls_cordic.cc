// Driver appended to the UNMODIFIED reference file hls/cordic/cordic.cpp
// (oracle/_ref/hls_cordic_np*_nw*.so).  TEST INFRASTRUCTURE ONLY.
// Loop of hls/cordic/cordic_test.cpp:66-68 without file I/O.
extern "C" {
int ref_nphase(void) { return NPHASE; }
int ref_nwidth(void) { return NWIDTH; }
void ref_hls_cordic(long long n0, long long count, long long* out_sin, long long* out_cos) {
  for (long long j = 0; j < count; j++) {
    out_t s, c;
    cordic((phi_t)(int)(n0 + j), &c, &s);
    out_sin[j] = s.to_int64();
    out_cos[j] = c.to_int64();
  }
}
}
