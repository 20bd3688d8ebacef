// bhw.hpp - header-only C++ mirror of the reference's interfaces over the C ABI (bhw.h).
//
// The reference's host-side language is C++ (cpp/, hls/); its "operator interface" for this path
// is the entity generic/port list and three free functions.  This header keeps those names:
//   bhw::win_selector   generics as constructor arguments, AA0..AA6 as call arguments
//                       (src/win_selector.vhd:60-87)
//   bhw::win_function   void win_function(char win_type, phi_t i, win_t* out)   vector form
//                       (hls/windows/win_function.h:65-69, hls/windows/win_function.cpp:380-422)
//   bhw::cordic         void cordic(phi_t, out_t* cos, out_t* sin)               vector form
//                       (hls/cordic/cordic.h:58-62; model = BHW_MODEL_CPP: cpp/cordic_sincos.cpp:10)
// Errors become bhw::error (the C ABI itself never throws).  Host-buffer calls generate on the
// current CUDA device and copy back; the *_device calls are stream-ordered.
#ifndef BHW_HPP_
#define BHW_HPP_
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "bhw.h"

namespace bhw {

struct error : std::runtime_error {
  int status;
  error(int st, const char* where)
      : std::runtime_error(std::string(where) + ": " + bhw_strerror(st) +
                           (st == BHW_E_CUDA ? std::string(" (") + bhw_last_cuda_error() + ")" : std::string())),
        status(st) {}
};
inline void check(int st, const char* where) { if (st != BHW_OK) throw error(st, where); }

class win_selector {
 public:
  // generic map of win_selector; XSERIES is accepted and ignored (no numeric effect)
  win_selector(int PHI_WIDTH = 10, int DAT_WIDTH = 16, const std::string& WIN_TYPE = "HAMMING",
               const std::string& SIN_TYPE = "CORDIC", int LUT_SIZE = 9, const std::string& XSERIES = "ULTRA") {
    (void)XSERIES;
    d_ = bhw_desc();
    d_.phi_width = PHI_WIDTH;
    d_.dat_width = DAT_WIDTH;
    if (WIN_TYPE == "HAMMING") d_.win_type = BHW_WIN_HAMMING;
    else if (WIN_TYPE == "BH3TERM") d_.win_type = BHW_WIN_BH3TERM;
    else if (WIN_TYPE == "BH4TERM") d_.win_type = BHW_WIN_BH4TERM;
    else if (WIN_TYPE == "BH5TERM") d_.win_type = BHW_WIN_BH5TERM;
    else if (WIN_TYPE == "BH7TERM") d_.win_type = BHW_WIN_BH7TERM;
    else throw error(BHW_E_WIN_TYPE, "win_selector");
    if (SIN_TYPE == "CORDIC") d_.sin_type = BHW_SIN_CORDIC;
    else if (SIN_TYPE == "TAYLOR") { d_.sin_type = BHW_SIN_TAYLOR; d_.lut_size = LUT_SIZE; }
    else if (SIN_TYPE == "CORDIC48") d_.sin_type = BHW_SIN_CORDIC48;          // entity swap: cordic_dds48
    else if (SIN_TYPE == "CORDIC_SCALED") d_.sin_type = BHW_SIN_CORDIC_SCALED;  // cordic_dds_scaled
    else throw error(BHW_E_SIN_TYPE, "win_selector");
  }
  win_selector& dt_vld_order(bool on) { d_.stream_offset = on ? 1 : 0; return *this; }
  win_selector& algo(int a) { d_.algo = a; return *this; }

  // port map: raw DAT_WIDTH-bit AA0..AA6
  bhw_desc desc(int64_t AA0, int64_t AA1 = 0, int64_t AA2 = 0, int64_t AA3 = 0, int64_t AA4 = 0, int64_t AA5 = 0,
                int64_t AA6 = 0) const {
    bhw_desc d = d_;
    const int64_t aa[BHW_MAX_TERMS] = {AA0, AA1, AA2, AA3, AA4, AA5, AA6};
    for (int k = 0; k < BHW_MAX_TERMS; k++) d.aa[k] = aa[k];
    check(bhw_validate(&d), "win_selector");
    return d;
  }
  // coefficients by README variant (1..10) through the testbench's quantisation rule
  bhw_desc desc_variant(int variant) const {
    bhw_desc d = d_;
    int32_t wt = 0;
    check(bhw_quantize(variant, BHW_RULE_TB, d.dat_width, d.aa, &wt), "bhw_quantize");
    if (wt != d.win_type) throw error(BHW_E_VARIANT, "win_selector: variant does not belong to this entity");
    check(bhw_validate(&d), "win_selector");
    return d;
  }

  uint64_t length() const { return 1ull << d_.phi_width; }

  // DT_WIN for `count` enabled clocks (DAT_WIDTH <= 32), host memory
  std::vector<int32_t> stream(const bhw_desc& d, uint64_t count = 0) const {
    if (bhw_elem_bytes(&d) != 4) throw error(BHW_E_ELEM, "stream: use stream64 for DAT_WIDTH > 32");
    if (!count) count = length();
    std::vector<int32_t> out(count);
    check(bhw_generate_host(&d, out.data(), 0, count), "bhw_generate_host");
    return out;
  }
  std::vector<int64_t> stream64(const bhw_desc& d, uint64_t count = 0) const {
    if (bhw_elem_bytes(&d) != 8) throw error(BHW_E_ELEM, "stream64: use stream for DAT_WIDTH <= 32");
    if (!count) count = length();
    std::vector<int64_t> out(count);
    check(bhw_generate_host(&d, out.data(), 0, count), "bhw_generate_host");
    return out;
  }
  // the same into device memory, stream-ordered (stream: cudaStream_t as void*)
  void stream_device(const bhw_desc& d, void* out_dev, uint64_t n0, uint64_t count, void* stream = nullptr) const {
    check(bhw_generate(&d, out_dev, n0, count, stream), "bhw_generate");
  }

 private:
  bhw_desc d_;
};

// Extension (no reference entity): a window of `terms` = 6 or 8..11 cosine terms built like bh_win_7term
// (BHW_WIN_MTERM_*), coefficients by quantize variant 14..18 (doc/blackman-harris coef.jpg) or raw ports.
inline bhw_desc mterm_desc(int terms, int PHI_WIDTH, int DAT_WIDTH, const std::vector<int64_t>& aa = {},
                           int sin_type = BHW_SIN_CORDIC) {
  bhw_desc d = bhw_desc();
  d.win_type = terms; d.phi_width = PHI_WIDTH; d.dat_width = DAT_WIDTH; d.sin_type = sin_type;
  if (aa.empty()) {
    const int variant = terms == 6 ? 14 : terms + 7;   // 8 -> 15 ... 11 -> 18
    int32_t wt = 0;
    check(bhw_quantize(variant, BHW_RULE_TB, DAT_WIDTH, d.aa, &wt), "bhw_quantize");
    if (wt != terms) throw error(BHW_E_WIN_TYPE, "mterm_desc");
  } else {
    if (aa.size() > (size_t)BHW_MAX_TERMS) throw error(BHW_E_ARG, "mterm_desc: more than AA0..AA10");
    for (size_t k = 0; k < aa.size(); k++) d.aa[k] = aa[k];
  }
  check(bhw_validate(&d), "mterm_desc");
  return d;
}

// HLS win_function for i = i0 .. i0+count-1; win_type codes as in the reference (1 Hamming, 2 Hann,
// 3 Blackman(-Harris 3), 4 BH4, 5 BH5, 7 BH7); any other code yields zeros like win_empty().
inline std::vector<int32_t> win_function(char win_type, int nphase, int nwidth, uint64_t i0 = 0, uint64_t count = 0) {
  if (!count) count = (1ull << nphase) - i0;
  std::vector<int32_t> out(count, 0);
  int variant;
  switch (win_type) {
    case 1: variant = 1; break;
    case 2: variant = 2; break;
    case 3: variant = 3; break;
    case 4: variant = 6; break;
    case 5: variant = 9; break;
    case 7: variant = 10; break;
    default: return out;
  }
  bhw_desc d = bhw_desc();
  d.model = BHW_MODEL_HLS;
  d.phi_width = nphase;
  d.dat_width = nwidth;
  check(bhw_quantize(variant, BHW_RULE_HLS, nwidth, d.aa, &d.win_type), "bhw_quantize");
  check(bhw_generate_host(&d, out.data(), i0, count), "bhw_generate_host");
  return out;
}

// entity cordic_atan2 (src/cordic_atan2.vhd:64-78): PHI_DT for every (VEC_DX, VEC_DY) pair, in order.
// stream_quadrant = true: as the entity streams it (the quadrant of pair t+1 corrects pair t, src/cordic_atan2.vhd:110-205)
inline std::vector<int32_t> cordic_atan2(int input_width, int angle_width, const std::vector<int32_t>& vec_dx,
                                         const std::vector<int32_t>& vec_dy, int precision = 1,
                                         bool stream_quadrant = false) {
  if (vec_dx.size() != vec_dy.size()) throw error(BHW_E_ARG, "cordic_atan2: VEC_DX and VEC_DY differ in length");
  bhw_atan2_desc d = bhw_atan2_desc();
  d.input_width = input_width;
  d.angle_width = angle_width;
  d.precision = precision;
  d.stream_quadrant = stream_quadrant ? 1 : 0;
  std::vector<int32_t> phi(vec_dx.size());
  check(bhw_atan2_host(&d, vec_dx.data(), vec_dy.data(), phi.data(), vec_dx.size()), "bhw_atan2_host");
  return phi;
}

}  // namespace bhw
#endif
