// bhw_plan.h - CUDA-free planning helpers (see bhw_plan.cpp).
#pragma once
#include <vector>

#include "bhw_device.cuh"

namespace bhw {

SrcParams canonical_source(const SrcParams& sp, uint32_t* drop);
bool fast32_ok(const SrcParams& sp);
int table_core32(const SrcParams& sp);
// Does the job's source have a stage-unrolled instantiation of the 32-bit core (k_table_build_u)?
// cordic_dds at DAT_WIDTH 16, 17, 24 (15, 16, 23 stages) and 32 (31 stages, biased);
// cordic_dds48 / cordic_dds_scaled at DAT_WIDTH 16, 17, 24, 32 (k_table_build_inq_u).
bool table_build_unrolled_ok(const TabJob& j);
void init_src_core(const SrcParams& sp, SrcCore* sc);
void init_tab_job(const SrcParams& canon, int32_t* tab, TabJob* j);
void build_taylor_rom(int dw, int lut, std::vector<I2>& rom);
enum TailMode { TAILMODE_FAST32 = 0, TAILMODE_ACC64 = 1, TAILMODE_GENERIC = 2 };
int table_tshift(const SrcParams& sp);
TailMode fast_tail_mode(const WinParams& wp, const SrcParams* src);
void fill_fast_rec(const WinParams& wp, const SrcParams* src, WinRec& r);

// Parameters of the register-resident 32-bit direct kernel; false when the window needs the
// generic 64-bit body (wide registers, input-quadrant CORDICs, TAYLOR, 64-bit tails).
bool direct32_params(const WinParams& wp, const SrcParams* src, Direct32Params* out);

// Parameters of the register-resident TAYLOR direct kernel; false when the window is not a
// 2-/3-term TAYLOR window with a 32-bit tail.
bool direct_taylor_params(const WinParams& wp, const SrcParams* src, DirectTayParams* out);

// cordic_atan2: validation (+ kernel parameters when p != NULL); returns a bhw_status.
int resolve_atan2(const bhw_atan2_desc* d, Atan2Params* p);

// Trig table of one harmonic as the bank planner sees it.
struct BankTableInfo { const int32_t* ptr; uint32_t entries; bool antisym; };
// Is the source's cosine table provably antisymmetric over half a period, T[i + E/2] == -T[i]
// as plain integers?  (It is whenever no entry can be the most negative DW-bit number, whose
// negation wraps onto itself.)
bool source_antisymmetric(const SrcParams& sp);
// k_direct_window on a whole window: can samples (n, n + N/2) share one evaluation per harmonic, and
// which harmonics (bit k) see the flipped quadrant half a window later (direct_sample_core_pair)?
bool direct_pair_flip(const WinParams& wp, const SrcParams* src, uint32_t* flip);
// ... and the four samples n + r*N/4 (direct_sample_core_quad)?  adv: 2 bits per harmonic.
bool direct_quad_adv(const WinParams& wp, const SrcParams* src, uint32_t* adv);
// Shape of a record for the bank kernel (tables_k[k] = table of harmonic k), the table placement
// (TAB_*) and whether lanes take (n, n + N/2) pairs; false when the record cannot go there.
// allow_pair = false: lanes own single samples even where pairing would be valid (tile ranges inside
// one window, where the partner half is somebody else's).
bool bank_shape(const WinRec& r, const BankTableInfo* tables_k, size_t smem_limit_bytes, BankShape* sh,
                int* tab_mode, bool* pair, bool allow_pair = true);

}  // namespace bhw
