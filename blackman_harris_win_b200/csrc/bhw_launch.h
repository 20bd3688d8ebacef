// bhw_launch.h - kernel argument blocks and launcher prototypes (bhw_kernels.cu <-> bhw_api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bhw_device.cuh"
#include "bhw_group.cuh"

namespace bhw {

constexpr int kSynthMaxPieces = 24;

struct SynthArgs {
  const WinRec* recs;        // [nrec] distinct window records
  const uint32_t* win_rec;   // [nwin] record of each window; NULL: record 0 for every window
  const uint64_t* flat_off;  // [nwin+1] first flat sample of each window; unused when uniform_pw >= 0
  const GenRec* gens;        // parameters of WR_GENERIC windows
  const I2* rom;             // Taylor ROM words (generic body only)
  void* out;                 // int32 (pack16: int16) output, element 0 = flat sample flat_begin
  uint64_t flat_begin;
  uint64_t flat_count;
  int32_t nwin;
  int32_t uniform_pw;        // >= 0: every window has 2^uniform_pw samples (no search needed)
  // npieces > 0: the launch covers several disjoint flat ranges ("pieces", ascending) instead of
  // [flat_begin, flat_begin + flat_count): piece p = [piece_begin[p], piece_end[p]), its 128-sample tiles are
  // numbered from piece_tile0[p]; `out` is then element flat index `out_flat0` of the batch
  uint32_t npieces;
  uint32_t pack16;           // 1: `out` is an int16 array (BHW_OUT_INT16, every window DAT_WIDTH <= 16)
  uint64_t out_flat0;
  uint64_t piece_begin[kSynthMaxPieces];
  uint64_t piece_end[kSynthMaxPieces];
  uint32_t piece_tile0[kSynthMaxPieces + 1];
};

struct BankArgs {
  BankShape sh;
  const WinRec* recs;        // per-window records (A[k], S0, n_first are read)
  const uint32_t* win_rec;   // record of each window; NULL: record 0 for every window
  int32_t* out;              // sample 0 of window w_first (an int16 array when pack16)
  uint32_t w_first;          // first window of the launch (index into win_rec)
  uint32_t nwin;             // whole windows to generate
  uint32_t tile_off;         // ntiles > 0: tiles [tile_off, tile_off + ntiles) of window w_first only
  uint32_t ntiles;           //             (unpaired shape); `out` is then the first of those tiles
  uint32_t win_minor;        // TAB_GLOBAL, whole windows: walk the bank tile by tile across its windows
  uint32_t spread;           // TAB_GLOBAL, one whole window: G > 0 = warp j of G takes the j-th G-th of the window
  uint32_t pack16;           // 1: `out` is an int16 array (BHW_OUT_INT16); 2- and 3-term shapes, 32-bit tail only
  uint32_t pad3;
};

struct GroupArgs {
  GroupShape sh;
  const WinRec* recs;        // per-window records (A[k], S0, n_first are read)
  const GroupWin* wins;      // [nwin + 1] windows of the launch in unit order (the last one is a sentinel whose
                             // unit_begin closes the unit range); NULL: the list is `iw` below
  GroupWin iw[3];            // inline list for the one or two windows a requested range cuts
  int32_t* out;              // GroupWin::out_off is relative to this
  uint32_t nwin;
  uint32_t unit_base;        // the launch covers units [unit_base, unit_base + nunits) of the list's numbering
  uint32_t nunits;           // tiles of 256 samples (sample pairs) in the launch
  uint32_t spread;           // G_GLOBAL, one whole window: G > 0 = warp j of G takes the j-th G-th of the window
  // fused apply step (bhw_apply): x != NULL -> multiply every frame of x by the window instead of storing it
  const int32_t* x;          // frames x N samples, frame-major; the low DAT_WIDTH bits are the sample
  void* y;                   // frames x N products: int64 (mode 1) or int32 (mode 2)
  uint64_t frames;
  uint32_t apply_mode;       // BHW_APPLY_EXACT + 1 / BHW_APPLY_ROUNDED + 1
  uint32_t apply_dw;         // DAT_WIDTH
  uint32_t prefetch_lines;   // spread walk: prefetch the next tile's pyramid lines into L1, harmonics of up to this many lines
  uint32_t pack16;           // 1: `out` is an int16 array (BHW_OUT_INT16); not together with x
};

struct DirectArgs {
  WinParams wp;
  SrcParams src[2];
  SrcCore sc[2];              // shift-add core of each source + its sliced atan words
  const I2* rom;              // Taylor ROM (global)
  uint32_t rom_smem_entries;  // > 0: copy that many ROM words to shared memory first
  uint32_t pair_flip;         // bit 31: whole window, samples (n, n + N/2) from one evaluation per harmonic;
                              // bits 1..6: the harmonics whose quadrant is flipped half a window later
  uint32_t quad_adv;          // bit 31: whole window, the four samples n + r*N/4 from one evaluation per
                              // harmonic (wins over pair_flip); bits 2k, 2k+1: quadrants harmonic k advances
                              // per quarter window
  uint32_t pad2;
  uint64_t n_first;           // n of output element 0 (stream offset folded in)
  uint64_t count;
};

struct Direct32Args {
  Direct32Params p;
  uint64_t n0;      // first sample (the stream offset lives in p.n_first)
  uint64_t count;
  uint32_t pair;    // whole window (n0 = 0, count = N): 1 = one set of CORDIC evaluations serves samples n and
                    // n + N/2 (direct32_pair, N >= 8); 2 = and n + N/4, n + 3N/4 as well (direct32_quad, N >= 32)
  uint32_t narrow;  // short request: one sample (pair) per thread instead of four
};

struct DirectTayArgs {
  DirectTayParams p;
  const I2* rom;    // Taylor ROM (global)
  uint64_t n0;      // first sample (the stream offset lives in p.n_first)
  uint64_t count;
  uint32_t pair;    // whole window (n0 = 0, count = N >= 8, bh_win_3term's second unit one bit narrower): 1 = one
                    // evaluation serves samples n and n + N/2 (direct_taylor_pair); 2 = long TAY_WIDE window,
                    // stream offset 0: two ROM words serve 4 consecutive samples and their three quarter-window
                    // partners (direct_taylor_quad4, direct_taylor_quad_ok).  (Round 1 tried the four partners
                    // with the per-sample body and measured no gain, 32.8 vs 30.8 us on config 4: the cost was the
                    // per-sample ROM look-up, mode and quadrant logic, which quad4 hoists out.)
};

struct SinCosArgs {
  SrcParams src;
  SrcCore sc;
  const I2* rom;
  uint64_t n_first;
  uint64_t count;
  uint32_t quad;   // whole table of an output-quadrant source: phases j, j + N/4, j + N/2, j + 3N/4 from one evaluation
};

cudaError_t launch_table_build(const TabJob* jobs_dev, int njobs, uint32_t total_work, const I2* rom_dev,
                               cudaStream_t stream);
// a single large job of a 32-bit-core source with a stage count that has an unrolled instantiation
// (DAT_WIDTH 16, 17, 24 and 32 of cordic_dds): its own launch, the job in the parameter block
// ctas_per_sm: resident 256-thread CTAs per SM the grid is sized for (8 fills an SM; 2 leaves room for a
// 1024-thread synthesis CTA of another stream next to it)
cudaError_t launch_table_build_unrolled(const TabJob& j, cudaStream_t stream, int ctas_per_sm = 8);
cudaError_t launch_synth(const SynthArgs& a, cudaStream_t stream);
// tab: TAB_SMEM_FULL / TAB_SMEM_HALF / TAB_GLOBAL; pair: lanes own (n, n + N/2) sample pairs
// pdl: launch with programmatic stream serialization (the kernel directly ahead in the stream is
// k_table_build, which releases its dependents early; k_synth_bank waits for it before its first
// table read)
cudaError_t launch_synth_bank(const BankArgs& a, int tab, bool pair, cudaStream_t stream, bool pdl = false);
// tab: G_HALF32 / G_Q16 / G_GLOBAL; pair: units are tiles of sample pairs (whole windows only)
cudaError_t launch_synth_group(const GroupArgs& a, int tab, bool pair, cudaStream_t stream, bool pdl = false);
cudaError_t launch_apply_mul(const int32_t* x, const int32_t* w, void* y, uint64_t n, uint64_t frames, int mode, int dw,
                             cudaStream_t stream);
size_t group_smem_limit();  // bytes of shared memory a group launch may use for the staged table image
// pairing over input-quadrant CORDIC tables: find the entries that break the ones'-complement relation (exc[0] =
// count, exc[1..] = indices; room for entries/2 + 1 words), and recompute the sample pairs of a paired bank launch
// that read them
cudaError_t launch_inq_exceptions(const int32_t* tab, uint32_t entries, int32_t adj, uint32_t* exc, cudaStream_t stream);
cudaError_t launch_inq_patch(const BankArgs& a, const uint32_t* exc, cudaStream_t stream);
size_t bank_smem_limit();  // bytes of shared memory a bank launch may use for staged tables
int device_sm_count();     // SMs of the current device (148 on B200)
cudaError_t launch_direct_window(const DirectArgs& a, void* out, cudaStream_t stream);
cudaError_t launch_direct32(const Direct32Args& a, int32_t* out, cudaStream_t stream);
cudaError_t launch_direct_taylor(const DirectTayArgs& a, int32_t* out, cudaStream_t stream);
// avail: input pairs present in x / y (>= count): with Atan2Params::skew pair count-1 reads the quadrant of pair `count`
cudaError_t launch_atan2(const Atan2Params& p, const int32_t* x, const int32_t* y, int32_t* phi, uint64_t count,
                         cudaStream_t stream, uint64_t avail = 0);
cudaError_t launch_sincos(const SinCosArgs& a, void* out_sin, void* out_cos, bool elem64, cudaStream_t stream);

}  // namespace bhw
