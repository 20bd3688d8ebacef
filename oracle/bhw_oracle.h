/*
 * bhw_oracle.h - CPU restatement of the reference's window-generation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (blackman_harris_win_b200/,
 * include/) links, imports or calls this; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may.
 *
 * Parity status: the HLS-model and CPP-model functions are pinned against the
 * reference's own C++ compiled here (oracle/_ref, see oracle/Makefile).  The
 * RTL-model functions are pinned by golden vectors recorded while executing the
 * reference's own VHDL in oracle/vhdl_sim.py (a cycle simulator written for
 * this repo: no ghdl / nvc / vendor simulator exists in the image) - see
 * tests/golden/make_rtl_golden.py and tests/test_rtl_vhdl_sim.py.  The TAYLOR
 * entities sit on Xilinx DSP48E1/E2 primitives and ieee.math_real, which the
 * reference does not carry: those two are modelled (vhdl_sim.Dsp48, libm), so
 * TAYLOR is pinned "up to the primitive model".  An independent bit-vector
 * restatement (oracle/rtl_bitvec.py, tests/test_rtl_bitvec.py) and the
 * known-answer hashes in tests/golden/ remain as second anchors.
 */
#ifndef BHW_ORACLE_H_
#define BHW_ORACLE_H_
#include <stdint.h>
#include "../include/bhw.h"

#ifdef __cplusplus
extern "C" {
#endif

/* sin/cos sources: phase in, (sin, cos) out as sign-extended DATA_WIDTH-bit ints */
void orc_cordic_dds(int pw, int dw, int prec, uint64_t ph, int64_t* s, int64_t* c);
void orc_cordic_dds48(int pw, int dw, uint64_t ph, int64_t* s, int64_t* c);
void orc_cordic_dds_scaled(int pw, int dw, uint64_t ph, int64_t* s, int64_t* c);
void orc_taylor_sincos(int pw, int dw, int lut, uint64_t cnt, int64_t* s, int64_t* c);
void orc_taylor_rom(int dw, int lut, int64_t* rom_cos, int64_t* rom_sin); /* 2^lut entries each */
void orc_hls_cordic(int np, int nw, uint64_t phi, int64_t* s, int64_t* c);
void orc_cpp_cordic(int pw, int dw, int theta, int* s, int* c);

/* whole-descriptor entry points (same descriptor as the product ABI) */
int orc_validate(const bhw_desc* d);
int orc_window(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out);
int orc_sincos(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out_sin, int64_t* out_cos);
int orc_quantize(int variant, int rule, int dat_width, int64_t aa_out[BHW_MAX_TERMS], int32_t* win_type);

/* cordic_atan2 (src/cordic_atan2.vhd): PHI_DT of one (VEC_DX, VEC_DY) pair, sign-extended from
 * ANGLE_WIDTH bits.  RTL-only entity: pinned by the executed VHDL (see header comment); the entity's own
 * stream pairs PHI_DT of pair t with the quadrant of pair t+1 - orc_atan2_stream restates that. */
int orc_atan2_validate(int input_width, int angle_width, int precision);
int64_t orc_cordic_atan2(int input_width, int angle_width, int precision, int64_t vec_dx, int64_t vec_dy);
int orc_atan2(int input_width, int angle_width, int precision, const int32_t* x, const int32_t* y, int32_t* phi,
              uint64_t count);

/* multi-threaded fill used by bench.py's CPU legs (pthreads, one contiguous slice per thread) */
int orc_window_mt(const bhw_desc* d, uint64_t n0, uint64_t count, int64_t* out, int nthreads);
int orc_window_i32(const bhw_desc* d, uint64_t n0, uint64_t count, int32_t* out);
int orc_atan2_stream(int iw, int aw, int prec, const int32_t* x, const int32_t* y, int32_t* phi, uint64_t count);
int orc_apply(const bhw_desc* d, int mode, const int32_t* x, uint64_t frames, int64_t* y);

#ifdef __cplusplus
}
#endif
#endif
