// bhw_internal.h - host/device shared definitions of the window generator (not part of the ABI).
//
// A bhw_desc (include/bhw.h) is resolved on the host into small POD "parameter blocks" that the
// kernels take by value.  Nothing here depends on torch; nothing here touches oracle/.
#pragma once
#include <stdint.h>
#include "../../include/bhw.h"

namespace bhw {

// ---- sin/cos source ------------------------------------------------------------------------
// Which reference algorithm a source block describes (reference file:line in each comment).
enum SrcKind : int32_t {
  SRC_DDS = 0,    // cordic_dds, output-side quadrant fix        (src/cordic_dds.vhd:97-249)
  SRC_INQ = 1,    // cordic_dds48 / cordic_dds_scaled, input-side quadrant
                  //                                            (src/cordic_dds48.vhd:110-259,
                  //                                             src/cordic_dds_scaled.vhd:100-285)
  SRC_HLS = 2,    // HLS cordic()                               (hls/windows/win_function.cpp:47-156)
  SRC_CPP = 3,    // plain C++ cordic()                         (cpp/cordic_sincos.cpp:10-92)
  SRC_TAYLOR = 4  // taylor_sincos + tay1_order                 (src/taylor_sincos.vhd:86-255,
                  //                                             src/tay1_order.vhd:112-640)
};

enum TayMode : int32_t {
  TAY_LESS = 0,  // PHASE_WIDTH-LUT_SIZE < 2 : ROM only, address left-shifted (taylor_sincos.vhd:157-161)
  TAY_EQ = 1,    // = 2                      : ROM only                        (:164-167)
  TAY_DSP = 2,   // > 2, DATA_WIDTH < 19     : DSP48 MACC datapath             (tay1_order.vhd:180-504)
  TAY_WIDE = 3   // > 2, DATA_WIDTH > 18     : 35x27 multiply + saturation     (tay1_order.vhd:506-637)
};

struct SrcParams {
  int32_t kind;       // SrcKind
  int32_t pw;         // phase width of this unit
  int32_t dw;         // output width
  int32_t w;          // x/y register width (wrap)
  int32_t zw;         // z register width (wrap)
  int32_t n_xy;       // x/y iterations
  int32_t n_z;        // z iterations (= atan entries used)
  int32_t rom_sel;    // 0: table with pi/4 -> 2^46, 1: table with pi/4 -> 2^45
  int32_t rom_shift;  // right shift applied to the 48-bit atan words
  int32_t z_rshift;   // z0 = ((t >> z_rshift) << z_lshift), t = phase without the quadrant bits
  int32_t z_lshift;
  int32_t out_shift;  // result = x >> out_shift
  int32_t negw;       // width in which the quadrant negation wraps; 0 = ones' complement (CPP)
  int32_t outw;       // final truncation width
  int64_t rom_mask;   // AND mask applied after the shift (HLS 40 bits, CPP 48 bits)
  int64_t gain;       // x0
  // Taylor only
  int32_t lut;        // LUT_SIZE
  int32_t tay_mode;   // TayMode
  int32_t tay_ashift; // ROM address: TAY_LESS t << ashift ; TAY_DSP/WIDE t >> ashift
  int32_t tay_xs;     // XSHIFT = 19 + LUT_SIZE
  int32_t tay_cbits;  // bits of acnt
  int32_t tay_pad;
  int64_t tay_pi;     // ramb_pi = round(pi * 2^(17-STAGE))
};

// ---- trig table (memoised source) -----------------------------------------------------------
// value(ph) = T[(ph & idx_mask) >> idx_shift], two's-complement negated in dw bits when
// (ph & neg_bit) != 0.  For the output-quadrant sources the table covers half a period (the
// second half is the negation of the first, exactly); for SRC_INQ it covers the full period.
struct TabLookup {
  uint32_t idx_mask;
  uint32_t idx_shift;
  uint32_t neg_bit;
  uint32_t entries;
};

// ---- window tail ----------------------------------------------------------------------------
enum TailKind : int32_t {
  TAIL_RTL2 = 0,  // hamming_win: (AA0 - b1) rounded on bit 0      (src/hamming_win.vhd:192-231)
  TAIL_RTLM = 1,  // 3/4/5/7-term: signed sum rounded on bit 1     (src/bh_win_3term.vhd:258-306 ...)
  TAIL_HLS = 2    // a0 - m1 + m2 ..., floor shifts, truncation    (hls/windows/win_function.cpp:168-377)
};

struct TermParams {
  uint32_t kmul;     // phase step of this harmonic (k for CORDIC; 1 for a Taylor unit)
  uint32_t ph_mask;  // 2^(unit phase width) - 1
  int32_t src;       // index into the plan's sources / tables
  int32_t pad;
};

struct WinParams {
  int32_t m;          // number of terms (2,3,4,5,7)
  int32_t dw;
  int32_t pw;
  int32_t tail;       // TailKind
  int32_t stream_offset;
  int32_t elem64;     // output element is int64
  int32_t nsrc;       // distinct sources (1, or 2 for 3-term TAYLOR)
  int32_t pad;
  int64_t aa[BHW_MAX_TERMS];  // sign-wrapped to dw bits
  TermParams term[BHW_MAX_TERMS - 1];
};

// Host-side resolution (bhw_resolve.cpp).  All return a bhw_status.
int validate_desc(const bhw_desc* d, bool for_window);
int batch_elem_bytes(const bhw_desc* descs, int nwin, size_t* esz);   // one container per batch, else BHW_E_ELEM
int resolve_source(const bhw_desc* d, int unit, SrcParams* out);     // unit: 0, or 1 = 2nd Taylor unit
int resolve_window(const bhw_desc* d, WinParams* wp, SrcParams src[2]);
TabLookup table_lookup_for(const SrcParams& sp);

}  // namespace bhw
