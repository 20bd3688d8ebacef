// Driver appended (same translation unit) to the UNMODIFIED reference file
// hls/windows/win_function.cpp when oracle/build_ref.sh builds oracle/_ref/hls_win_np*_nw*.so.
// TEST INFRASTRUCTURE ONLY.  It is the per-sample loop of hls/windows/window_test.cpp:93,193
// without the file I/O and the 2*N-object stack arrays (:79-80).
extern "C" {
int ref_nphase(void) { return NPHASE; }
int ref_nwidth(void) { return NWIDTH; }
// out[j] = win_function(win_type, n0+j)
void ref_hls_window(int win_type, long long n0, long long count, long long* out) {
  for (long long j = 0; j < count; j++) {
    win_t r;
    win_function((char)win_type, (phi_t)(int)(n0 + j), &r);
    out[j] = r.to_int64();
  }
}
void ref_hls_window_i32(int win_type, long long n0, long long count, int* out) {
  for (long long j = 0; j < count; j++) {
    win_t r;
    win_function((char)win_type, (phi_t)(int)(n0 + j), &r);
    out[j] = (int)r.to_int64();
  }
}
// the CORDIC inside win_function.cpp (:47-156)
void ref_hls_cordic(long long n0, long long count, long long* out_sin, long long* out_cos) {
  for (long long j = 0; j < count; j++) {
    win_t s, c;
    cordic((phi_t)(int)(n0 + j), &c, &s);
    out_sin[j] = s.to_int64();
    out_cos[j] = c.to_int64();
  }
}
}
