// Driver appended to the UNMODIFIED reference file cpp/cordic_sincos.cpp
// (oracle/_ref/cpp_cordic_pw*_dw*.so; its main() is renamed ref_cpp_main by -Dmain=...).
// TEST INFRASTRUCTURE ONLY.  Loop of cpp/cordic_sincos.cpp:135-138 without fprintf.
// The reference keeps its atan table local to main() (:97-110), so the driver carries the
// same 48 constants to hand to cordic(theta, lut, &s, &c).
static long long ref_lut_table[48] = {
    0x200000000000, 0x12E4051D9DF3, 0x09FB385B5EE4, 0x051111D41DDE, 0x028B0D430E59, 0x0145D7E15904,
    0x00A2F61E5C28, 0x00517C5511D4, 0x0028BE5346D1, 0x00145F2EBB31, 0x000A2F980092, 0x000517CC14A8,
    0x00028BE60CE0, 0x000145F306C1, 0x0000A2F9836B, 0x0000517CC1B7, 0x000028BE60DC, 0x0000145F306E,
    0x00000A2F9837, 0x00000517CC1B, 0x0000028BE60E, 0x00000145F307, 0x000000A2F983, 0x000000517CC2,
    0x00000028BE61, 0x000000145F30, 0x0000000A2F98, 0x0000000517CC, 0x000000028BE6, 0x0000000145F3,
    0x00000000A2FA, 0x00000000517D, 0x0000000028BE, 0x00000000145F, 0x000000000A30, 0x000000000518,
    0x00000000028C, 0x000000000146, 0x0000000000A3, 0x000000000051, 0x000000000029, 0x000000000014,
    0x00000000000A, 0x000000000005, 0x000000000003, 0x000000000001, 0x000000000001, 0x000000000000};
extern "C" {
int ref_phase_width(void) { return PHASE_WIDTH; }
int ref_data_width(void) { return DATA_WIDTH; }
void ref_cpp_cordic(long long n0, long long count, int* out_sin, int* out_cos) {
  for (long long j = 0; j < count; j++) cordic((int)(n0 + j), ref_lut_table, &out_sin[j], &out_cos[j]);
}
int ref_cpp_main(void) { return ref_cpp_main_impl(0, 0); }
}
