/*
 * bhw.h - C ABI of the B200-native window-coefficient generator.
 *
 * Drop-in boundary for the table-generation path of hukenovs/blackman_harris_win.
 * The reference has no FFI; its "interface" is (i) the VHDL entity generics and
 * ports and (ii) three small C/C++ model functions.  Every entry point below
 * cites the reference interface it replaces (paths relative to the reference
 * checkout).  Plain C types only: no torch, no C++ types, no exceptions.
 *
 * Conventions
 *   - All functions return 0 (BHW_OK) or a negative bhw_status code; nothing
 *     aborts or throws.  bhw_strerror() names a code.
 *   - "device" pointers are CUDA device pointers on the *current* device of
 *     the calling thread; "stream" is a cudaStream_t passed as void* (NULL =
 *     legacy default stream).  Calls are stream-ordered and do not synchronise
 *     unless the name ends in _host.
 *   - Output elements are sign-extended two's complement: int32_t when
 *     dat_width <= 32, int64_t otherwise (bhw_elem_bytes()).  Index order is
 *     the entity's phase order: out[j] = DT_WIN for phase (n0 + j +
 *     stream_offset) mod 2^phi_width.
 *   - The library is re-entrant; its only persistent state is a per-device
 *     cache of TAYLOR sine ROMs and of the *_host staging buffers
 *     (bhw_cache_clear() drops it).  Trig tables belong to plans.
 */
#ifndef BHW_H_
#define BHW_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* every entry point is exported; the library is built with -fvisibility=hidden */
#if defined(__GNUC__)
#define BHW_API __attribute__((visibility("default")))
#else
#define BHW_API
#endif

#define BHW_VERSION 0x000100 /* 0.1.0 */
#define BHW_MAX_TERMS 11  /* bh_win_7term is the reference's widest entity; 6 and 8..11 terms: see BHW_WIN_MTERM_* */
#define BHW_MIN_PHI_WIDTH 4   /* 16 points ...                                  */
#define BHW_MAX_PHI_WIDTH 26  /* ... 64M points, the reference's range (README.md:2) */

/* ---- status codes ------------------------------------------------------ */
typedef enum bhw_status {
  BHW_OK = 0,
  BHW_E_NULL = -1,        /* a required pointer is NULL                       */
  BHW_E_WIN_TYPE = -2,    /* win_type is not 2..11, or 6 / 8..11 outside the
                             RTL model with a CORDIC source                   */
  BHW_E_SIN_TYPE = -3,    /* unknown sin_type, or not available for the entity
                             (4/5/7-term have no TAYLOR: src/bh_win_4term.vhd:57-61) */
  BHW_E_MODEL = -4,       /* unknown model, or model/sin_type/op mismatch     */
  BHW_E_PHI_WIDTH = -5,   /* phi_width outside the supported range            */
  BHW_E_DAT_WIDTH = -6,   /* dat_width outside the range of the sin source    */
  BHW_E_PRECISION = -7,   /* cordic_dds PRECISION outside 1..7 or W > 49      */
  BHW_E_LUT_SIZE = -8,    /* TAYLOR LUT_SIZE invalid (STAGE > 15, width
                             overflow of the DSP slice, or the mis-aligned
                             3-term case PHI_WIDTH-LUT_SIZE == 3)             */
  BHW_E_COEFF = -9,       /* an AAk does not fit dat_width bits               */
  BHW_E_RANGE = -10,      /* n0/count/flat range outside the window / batch   */
  BHW_E_ELEM = -11,       /* mixed or wrong output element size in a batch    */
  BHW_E_CUDA = -12,       /* a CUDA runtime call failed (bhw_last_cuda_error) */
  BHW_E_NO_DEVICE = -13,  /* no CUDA device / device index out of range       */
  BHW_E_ALLOC = -14,      /* workspace allocation failed                      */
  BHW_E_VARIANT = -15,    /* bhw_quantize: unknown variant or rule            */
  BHW_E_ARG = -16,        /* any other invalid argument                       */
  BHW_E_CAPTURE = -17     /* a one-shot batch call on a stream that is being
                             captured into a CUDA graph (capture
                             bhw_plan_execute instead)                        */
} bhw_status;

/* ---- enumerations mirroring the generics ------------------------------- */
/* WIN_TYPE generic of win_selector (src/win_selector.vhd:64,93-199).  The
 * value is the number of terms, i.e. the entity that is instantiated. */
enum {
  BHW_WIN_HAMMING = 2, /* "HAMMING" -> hamming_win  (src/hamming_win.vhd:60-82)  */
  BHW_WIN_BH3TERM = 3, /* "BH3TERM" -> bh_win_3term (src/bh_win_3term.vhd:68-90) */
  BHW_WIN_BH4TERM = 4, /* "BH4TERM" -> bh_win_4term (src/bh_win_4term.vhd:56-75) */
  BHW_WIN_BH5TERM = 5, /* "BH5TERM" -> bh_win_5term (src/bh_win_5term.vhd:70-90) */
  BHW_WIN_BH7TERM = 7, /* "BH7TERM" -> bh_win_7term (src/bh_win_7term.vhd:58-80) */
  /* Extension, not a reference entity: the minimum-sidelobe sets with 6 and 8..11 terms that the reference
   * only tabulates (doc/blackman-harris coef.jpg).  The window is what the structure shared by
   * bh_win_3term .. bh_win_7term gives for M terms: one cordic_dds (or pin-compatible CORDIC) per harmonic
   * k = 1..M-1 at phase k*n, each product sliced and rounded as in every entity
   * (src/bh_win_7term.vhd:350-400), the alternating sum AA0 - b1 + b2 - ... in DAT_WIDTH+2 bits and the
   * rounding on bit 1 (src/bh_win_3term.vhd:295-306).  BHW_MODEL_RTL with a CORDIC source only. */
  BHW_WIN_MTERM_6 = 6,
  BHW_WIN_MTERM_8 = 8,
  BHW_WIN_MTERM_9 = 9,
  BHW_WIN_MTERM_10 = 10,
  BHW_WIN_MTERM_11 = 11
};

/* SIN_TYPE generic (src/win_selector.vhd:65) plus the two pin-compatible
 * CORDIC entities that can be swapped in for cordic_dds. */
enum {
  BHW_SIN_CORDIC = 0,        /* cordic_dds        (src/cordic_dds.vhd:77-92)        */
  BHW_SIN_TAYLOR = 1,        /* taylor_sincos     (src/taylor_sincos.vhd:65-81)     */
  BHW_SIN_CORDIC48 = 2,      /* cordic_dds48      (src/cordic_dds48.vhd:91-105)     */
  BHW_SIN_CORDIC_SCALED = 3  /* cordic_dds_scaled (src/cordic_dds_scaled.vhd:81-95) */
};

/* Which of the reference's three (mutually non-bit-identical) models the
 * integers must equal. */
enum {
  BHW_MODEL_RTL = 0, /* the VHDL entities (src/ tree)                               */
  BHW_MODEL_HLS = 1, /* hls/windows/win_function.cpp + hls/cordic/cordic.cpp      */
  BHW_MODEL_CPP = 2  /* cpp/cordic_sincos.cpp (sin/cos table only: bhw_sincos)    */
};

/* Output container (bhw_desc.out_format). */
enum {
  BHW_OUT_DEFAULT = 0, /* int32, or int64 when DAT_WIDTH > 32 */
  BHW_OUT_INT16 = 1    /* int16; DAT_WIDTH <= 16 only (BHW_E_DAT_WIDTH otherwise) */
};

/* Evaluation strategy (results are identical; this is a performance knob). */
enum {
  BHW_ALGO_AUTO = 0,
  BHW_ALGO_DIRECT = 1, /* one thread per sample, every k*phi term evaluated in registers */
  BHW_ALGO_TABLE = 2   /* sin/cos source evaluated once per distinct phase, then gathered */
};

/* ---- the descriptor: one field per generic / port ---------------------- */
/* Mirrors win_selector's generic list and AA0..AA6 ports (AA7..AA10: BHW_WIN_MTERM_*)
 * (src/win_selector.vhd:60-87).  For BHW_MODEL_HLS, phi_width/dat_width are
 * NPHASE/NWIDTH (hls/windows/win_function.h:51-52) and aa[] are the a_k
 * integers the HLS functions derive from their double constants
 * (hls/windows/win_function.cpp:176-177 etc.; see bhw_quantize). */
typedef struct bhw_desc {
  int32_t win_type;      /* BHW_WIN_*     (WIN_TYPE)                               */
  int32_t sin_type;      /* BHW_SIN_*     (SIN_TYPE / entity swap)                 */
  int32_t model;         /* BHW_MODEL_*                                            */
  int32_t phi_width;     /* PHI_WIDTH: window length N = 2^phi_width, 4..26        */
  int32_t dat_width;     /* DAT_WIDTH: bits of AAk and DT_WIN                      */
  int32_t precision;     /* cordic_dds PRECISION generic; 0 means the entity
                            default 1 (src/cordic_dds.vhd:79). Ignored elsewhere.  */
  int32_t lut_size;      /* LUT_SIZE (TAYLOR only); 0 means the entity default 9
                            (src/hamming_win.vhd:67)                               */
  int32_t stream_offset; /* 0: out[j] = w[n0+j]; 1: the DT_VLD-gated order
                            w[1], w[2], ..., w[N-1], w[0] of the entities as written
                            (cordic_dds / TAYLOR; confirmed by executing the VHDL,
                            DESIGN.md section 2).  With cordic_dds48 / _scaled swapped
                            in, the same rotation is applied as a convention only */
  int32_t algo;          /* BHW_ALGO_*                                             */
  int32_t out_format;    /* BHW_OUT_*: container of one output sample.  0: int32
                            (int64 for DAT_WIDTH > 32).  BHW_OUT_INT16: int16, for
                            DAT_WIDTH <= 16 (DT_WIN is DAT_WIDTH bits wide,
                            src/hamming_win.vhd:60-82) - half the bytes in HBM and
                            over the host link; batch / plan entry points and
                            bhw_generate[_host] only, one format per batch.
                            (XSERIES needs no field: it has no numeric effect.)      */
  int64_t aa[BHW_MAX_TERMS]; /* raw two's-complement AA0..AA10 port values; terms
                                beyond win_type are ignored                        */
} bhw_desc;

/* ---- helpers ----------------------------------------------------------- */
BHW_API const char* bhw_strerror(int status);
BHW_API int bhw_version(void);

/* Check a descriptor exactly as the generators do.  Rejections correspond to
 * configurations the reference cannot elaborate or documents as invalid. */
BHW_API int bhw_validate(const bhw_desc* d);

/* 4 or 8: bytes per output element for this descriptor. */
BHW_API int bhw_elem_bytes(const bhw_desc* d);

/* Coefficient front end.  The reference leaves quantisation to the caller; the
 * rules it uses itself are in src/tb/tb_windows.vhd:75-127 (rule BHW_RULE_TB)
 * and hls/windows/win_function.cpp:176-355 (rule BHW_RULE_HLS).
 * variant: 1 Hamming, 2 Hann, 3 Blackman, 4 Blackman-Harris-3, 5 Nuttall,
 * 6 Blackman-Harris-4, 7 Blackman-Nuttall, 8 Flat-top, 9 Blackman-Harris-5,
 * 10 Blackman-Harris-7 (README.md:30-41); and the alternative sets the reference
 * prints beside them: 11 Blackman-Harris-7 as in README.md:45-51, 12 Hamming
 * 0.5383554 / 0.4616446 (src/hamming_win.vhd:21-23), 13 Flat-top normalised
 * (src/bh_win_5term.vhd:28-33); 14..18: the 6-, 8-, 9-, 10- and 11-term
 * minimum-sidelobe sets of doc/blackman-harris coef.jpg (rule TB only, scaled like
 * the 7-term entity's: 2^(DAT_WIDTH-1)-1) for the BHW_WIN_MTERM_* extension.
 * Writes aa_out[0..10] (unused = 0) and *win_type (the number of terms). */
enum { BHW_RULE_TB = 0, BHW_RULE_HLS = 1 };
BHW_API int bhw_quantize(int variant, int rule, int dat_width, int64_t aa_out[BHW_MAX_TERMS],
                 int32_t* win_type);
/* Real-valued coefficients of a variant as the reference spells them
 * (for rule TB; rule HLS halves variant 3: win_function.cpp:206-208). */
BHW_API int bhw_variant_coeffs(int variant, int rule, double a_out[BHW_MAX_TERMS], int32_t* nterms);

/* ---- generation: one window ------------------------------------------- */
/* Replaces: one win_selector instance streaming `count` DT_WIN words
 * (src/win_selector.vhd:60-87) / the per-sample loop around
 * win_function(win_type, i, &out) (hls/windows/window_test.cpp:93,193,
 * hls/windows/win_function.h:65-69).  out_dev: device buffer of `count`
 * elements. Range: n0 + count <= 2^phi_width. */
BHW_API int bhw_generate(const bhw_desc* d, void* out_dev, uint64_t n0, uint64_t count, void* stream);

/* `reps` back-to-back bhw_generate calls from one host call (a per-call latency measurement without a
 * scripting language in the loop): call i writes to out_dev + (i % out_slots) * out_stride elements
 * (out_slots = 0: always out_dev). */
BHW_API int bhw_generate_repeat(const bhw_desc* d, void* out_dev, uint64_t n0, uint64_t count, int reps,
                        uint64_t out_stride, uint64_t out_slots, void* stream);

/* Same with a HOST output buffer: generates on the current device and copies
 * back (pinned staging, chunked, copy overlapped with generation).
 * Synchronises before returning. */
BHW_API int bhw_generate_host(const bhw_desc* d, void* out_host, uint64_t n0, uint64_t count);

/* ---- generation: a batch of windows, sharded by flat sample range ------ */
/* The batch is the concatenation of the nwin full windows in order; "flat"
 * indices address that concatenation.  Writes flat samples
 * [flat_begin, flat_begin+flat_count) to out_dev[0 .. flat_count).  All
 * windows in a batch must share one element size.  This is the sharding
 * primitive: rank r of R calls it with its contiguous slice
 * (bhw_shard_range) - no collective is involved. */
BHW_API int bhw_batch_total(const bhw_desc* descs, int nwin, uint64_t* total_samples);
BHW_API int bhw_shard_range(uint64_t total_samples, int rank, int nranks, uint64_t* begin,
                    uint64_t* count);
/* Cost-balanced variant for batches of unlike windows (the win_selector sweep): still one contiguous
 * flat slice per rank, but the cuts equalise an estimate of the generation time instead of the
 * sample count - a sample of a 7-term 32-bit window whose table lives in L2 costs several times a
 * sample of a Hann window.  Cuts fall on multiples of 4 samples; the slices of ranks 0..nranks-1
 * tile [0, total) in order.  The estimate is a heuristic of this implementation, not of the
 * reference (which has no notion of sharding). */
BHW_API int bhw_shard_range_cost(const bhw_desc* descs, int nwin, int rank, int nranks, uint64_t* begin,
                         uint64_t* count);
BHW_API int bhw_generate_batch(const bhw_desc* descs, int nwin, uint64_t flat_begin,
                       uint64_t flat_count, void* out_dev, void* stream);
BHW_API int bhw_generate_batch_host(const bhw_desc* descs, int nwin, uint64_t flat_begin,
                            uint64_t flat_count, void* out_host);
/* Single-process multi-GPU form: device g (0..ngpus-1) receives shard g of the
 * flat range in outs_dev[g] (allocated by the caller on device g, at least
 * the bhw_shard_range count).  Synchronises all devices before returning. */
BHW_API int bhw_generate_batch_multi(const bhw_desc* descs, int nwin, int ngpus, void* const* outs_dev);

/* Which windows of a batch a flat range touches: descs[*first_win .. *first_win + *nwin_touched)
 * and the range's offset inside that sub-batch.  A rank that owns flat slice [begin, begin+count)
 * can plan only those windows. */
BHW_API int bhw_shard_windows(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count,
                      int* first_win, int* nwin_touched, uint64_t* local_begin);

/* ---- the apply step: what the coefficient stream is for ------------------------------------ */
/* Replaces: a win_selector instance feeding int_multNxN_dsp48 together with the signal, the way the
 * window entities themselves use that multiplier (src/int_multNxN_dsp48.vhd:75-110:
 * DAT_Q <= SIGNED(DAT_A) * SIGNED(DAT_B), DTW = DAT_WIDTH bits in, 2*DTW bits out; instantiated e.g.
 * src/hamming_win.vhd:183-191).  x_dev holds `frames` frames of N = 2^phi_width samples, frame-major, one
 * int32 per sample whose low DAT_WIDTH bits are the DAT_A port bits (sign bit = bit DAT_WIDTH-1).
 * y[f*N + n] = x[f*N + n] * w[n] with w the stream bhw_generate(d) would write (stream_offset applies):
 *   BHW_APPLY_EXACT   : DAT_Q itself, one int64 per sample
 *   BHW_APPLY_ROUNDED : the entities' own slice of DAT_Q - r = DAT_Q[2DW-2 : DW-2],
 *                       y = (r >> 1) + (r & 1) wrapped to DW bits (src/hamming_win.vhd:195-208) - one
 *                       int32 per sample, sign-extended
 * For the windows k_synth_group generates (cordic_dds / HLS families, 32-bit tail, N >= 512) the window
 * is never written: it is produced in registers and multiplied into every frame on the fly (4 B read
 * + 4 or 8 B written per sample); any other window is generated into scratch memory first.
 * dat_width <= 32.  Stream-ordered; not capturable (BHW_E_CAPTURE). */
enum { BHW_APPLY_EXACT = 0, BHW_APPLY_ROUNDED = 1 };
BHW_API int bhw_apply(const bhw_desc* d, int mode, const int32_t* x_dev, void* y_dev, uint64_t frames, void* stream);

/* ---- plans: resolve once, execute many times ------------------------------ */
/* A plan is a batch resolved and resident on the current device: per-window records, trig-table
 * storage and (for TAYLOR) its own copy of every sine ROM the batch needs (any mix of DAT_WIDTH /
 * LUT_SIZE).  It corresponds to the elaborated entity instances of the reference (generics fixed at
 * elaboration, src/win_selector.vhd:61-70); executing it is the ENABLE burst.  bhw_plan_execute
 * writes flat samples [flat_begin, flat_begin+flat_count) of the batch to out_dev, stream-ordered,
 * without host-side planning work.  Trig tables are built by the first eager execute and kept
 * (an execute on another stream waits for that build through an event), unless the table cache is
 * off (bhw_set_table_cache(0)), in which case every execute rebuilds them and concurrent executes
 * of one plan must share one stream.  bhw_plan_execute may be captured into a CUDA graph: a build
 * that is only captured does not count as done, the graph then rebuilds the tables on every
 * replay.  A plan must be executed and destroyed on the device it was created on. */
typedef struct bhw_plan bhw_plan;
BHW_API int bhw_plan_create(const bhw_desc* descs, int nwin, bhw_plan** plan_out);
BHW_API int bhw_plan_execute(bhw_plan* plan, uint64_t flat_begin, uint64_t flat_count, void* out_dev,
                     void* stream);
BHW_API int bhw_plan_total(const bhw_plan* plan, uint64_t* total_samples);
BHW_API int bhw_plan_destroy(bhw_plan* plan);

/* ---- sin/cos tables (the DDS entities on their own) -------------------- */
/* Replaces: cordic_dds / cordic_dds48 / cordic_dds_scaled / taylor_sincos
 * driven by a free-running phase counter (ports PH_IN -> DT_SIN, DT_COS:
 * src/cordic_dds.vhd:83-91; taylor: src/taylor_sincos.vhd:73-80), the HLS
 * cordic(phi, &cos, &sin) (hls/cordic/cordic.h:58-62) and the C++
 * cordic(theta, lut, &s, &c) (cpp/cordic_sincos.cpp:10).  win_type/aa are
 * ignored.  Either output may be NULL.  Elements as for bhw_generate. */
BHW_API int bhw_sincos(const bhw_desc* d, void* out_sin_dev, void* out_cos_dev, uint64_t n0,
               uint64_t count, void* stream);

/* ---- cordic_atan2: the vectoring sibling of the DDS cores ------------------ */
/* Replaces entity cordic_atan2 (src/cordic_atan2.vhd:64-78: generics PRECISION, INPUT_WIDTH,
 * ANGLE_WIDTH; ports VEC_DX, VEC_DY -> PHI_DT).  No window entity instantiates it; it is here
 * because it shares the DDS cores' ROM and stage structure.  One PHI_DT per
 * (VEC_DX, VEC_DY) pair, in input order (the entity's ANGLE_WIDTH+1 clocks of latency and PHI_VL
 * are timing, not data).  x_dev/y_dev: int32 per element, the low INPUT_WIDTH bits are the port
 * bits (bit INPUT_WIDTH-1 is the sign); phi_dev: int32, PHI_DT sign-extended from ANGLE_WIDTH bits.
 * Valid: 4 <= angle_width <= 32, angle_width - 1 <= input_width <= 32 (the entity indexes
 * VEC_DX(ANGLE_WIDTH-2), src/cordic_atan2.vhd:140-141), 1 <= precision <= 7. */
typedef struct bhw_atan2_desc {
  int32_t input_width;   /* INPUT_WIDTH */
  int32_t angle_width;   /* ANGLE_WIDTH */
  int32_t precision;     /* PRECISION, 0 = the entity default 1 */
  int32_t stream_quadrant; /* 0: every pair is corrected with its own quadrant (the evident intent);
                              1: as the entity really streams - its quadrant shift registers are one stage
                              shorter than the data path (src/cordic_atan2.vhd:127-129 vs :136-184), so on a
                              stream of one pair per clock PHI_DT of pair t takes the quadrant of pair t+1
                              (the inputs after the last pair read 0).  Found by executing the VHDL. */
} bhw_atan2_desc;
BHW_API int bhw_atan2_validate(const bhw_atan2_desc* d);
BHW_API int bhw_atan2(const bhw_atan2_desc* d, const int32_t* x_dev, const int32_t* y_dev, int32_t* phi_dev,
              uint64_t count, void* stream);
/* The same with host buffers (pinned or pageable): chunks are staged through the library's device
 * buffers; blocking. */
BHW_API int bhw_atan2_host(const bhw_atan2_desc* d, const int32_t* x_host, const int32_t* y_host, int32_t* phi_host,
                   uint64_t count);

/* ---- cache / introspection --------------------------------------------- */
BHW_API int bhw_cache_clear(void);             /* free the per-device sine ROMs of the one-shot direct
                                          kernels, the host-pipeline staging, the one-shot memory
                                          pool and the side streams.  Synchronises every device
                                          first; no other library call may be in flight.  Plans
                                          own their ROMs and tables and are not affected.     */
BHW_API int bhw_set_table_cache(int enabled);  /* 1 (default): a plan builds its trig tables on its
                                          first execute and keeps them; 0: every execute
                                          rebuilds them (one-shot calls always build)        */
BHW_API int bhw_set_side_streams(int n);        /* 0..8 (default 4): an execute with many independent
                                          launches (a batch of differently shaped windows) fans
                                          them out over n internal side streams, forked from and
                                          joined back into the caller's stream; 0 keeps every
                                          launch on the caller's stream                       */
BHW_API uint64_t bhw_launch_count(void);       /* kernels launched by this library so far */
BHW_API const char* bhw_last_cuda_error(void); /* text of the last CUDA failure           */
BHW_API int bhw_device_count(void);

/* ---- per-kernel device timing (for bench.py's roofline line) ------------- */
/* When enabled, every kernel launch of the library is bracketed by CUDA events recorded on the
 * launching stream.  bhw_timing_read() synchronises those events and returns, for one kernel
 * class, the number of launches and their summed device time since the last reset.  Off by
 * default; costs two cudaEventRecord per launch when on. */
enum {
  BHW_KERNEL_TABLE_BUILD = 0, /* k_table_build: sin/cos source evaluated once per distinct phase */
  BHW_KERNEL_SYNTH = 1,       /* k_synth: any flat range of any batch (gather + tail + store)   */
  BHW_KERNEL_DIRECT = 2,      /* k_direct*: one thread per sample, sources evaluated in registers */
  BHW_KERNEL_SINCOS = 3,      /* k_sincos                                                        */
  BHW_KERNEL_SYNTH_BANK = 4,  /* k_synth_bank: whole windows of one shape, tables in shared memory */
  BHW_KERNEL_ATAN2 = 5,       /* k_atan2                                                          */
  BHW_KERNEL_SYNTH_GROUP = 6, /* k_synth_group: all windows of one family and entity, any PHI_WIDTHs  */
  BHW_KERNEL_APPLY = 7,       /* k_apply_mul: the unfused half of bhw_apply                           */
  BHW_KERNEL_CLASSES = 8
};
/* One finished launch of a timed region: its class, a kernel-specific shape word (k_synth_group /
 * k_synth_bank: terms | table placement << 8 | paired << 16 | spread walk << 17 | top level or PHI_WIDTH
 * << 24), the algorithmic bytes it wrote and its device time. */
typedef struct bhw_launch_record {
  int32_t kernel_class;
  uint32_t tag;
  uint64_t bytes;
  double ms;
} bhw_launch_record;
BHW_API int bhw_timing_enable(int enabled);
BHW_API int bhw_timing_reset(void);
BHW_API int bhw_timing_read(int kernel_class, double* total_ms, uint64_t* launches);
/* The launches recorded since the last reset, oldest first: writes up to max_records of them to `out`
 * (may be NULL) and their total number to *n_records. */
BHW_API int bhw_timing_launches(bhw_launch_record* out, uint64_t max_records, uint64_t* n_records);

#ifdef __cplusplus
}
#endif
#endif /* BHW_H_ */
