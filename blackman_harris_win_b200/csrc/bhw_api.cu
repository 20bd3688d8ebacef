// bhw_api.cu - the GPU half of the C ABI (include/bhw.h): plans, trig tables, sine ROM cache,
// launches, the host-buffer pipeline and the single-process multi-GPU form.
//
// There is deliberately no CPU evaluation path in this file: every generate/sincos entry point
// ends in a kernel launch or returns an error.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <new>
#include <unordered_map>
#include <mutex>
#include <string>
#include <vector>

#include "bhw_device.cuh"
#include "bhw_launch.h"
#include "bhw_plan.h"

namespace bhw {

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_cache_enabled{1};
static std::atomic<int> g_side_streams{4};  // bhw_set_side_streams
static std::atomic<int> g_spread_min_terms{5}; // long windows over a pyramid in L2/HBM get a spread launch of their own from this many terms
static std::atomic<int> g_prefetch_lines{0};    // spread launches: L1 prefetch of the next tile (0 = off)
static std::atomic<int> g_defer_ctas{0};    // resident CTAs per SM of a table build that shares the GPU with synthesis
static thread_local std::string t_cuda_err;

static int cuda_fail(cudaError_t e, const char* where) {
  t_cuda_err = std::string(where) + ": " + cudaGetErrorString(e);
  return BHW_E_CUDA;
}
#define BHW_CUDA(call)                                  \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// ---- optional per-kernel timing (bhw_timing_*) -------------------------------------------------
struct TimedSpan { int cls; int dev; cudaEvent_t a, b; uint32_t tag; uint64_t bytes; };
static std::atomic<int> g_timing{0};
static std::mutex g_timing_mu;
static std::vector<TimedSpan> g_spans;                 // recorded, not yet read
static std::vector<cudaEvent_t> g_event_pool[64];      // reusable events per device
static std::vector<bhw_launch_record> g_launch_log;   // finished spans since the last reset (bhw_timing_launches)
static double g_time_ms[BHW_KERNEL_CLASSES] = {0};
static uint64_t g_time_n[BHW_KERNEL_CLASSES] = {0};

static cudaEvent_t pool_event(int dev) {
  if (!g_event_pool[dev].empty()) { cudaEvent_t e = g_event_pool[dev].back(); g_event_pool[dev].pop_back(); return e; }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  return e;
}

// Brackets one launch with events on its stream when timing is on.
struct LaunchTimer {
  TimedSpan sp{0, 0, nullptr, nullptr, 0, 0};
  cudaStream_t stream;
  bool on;
  // tag: kernel-specific shape word (bhw_launch_record::tag); bytes: algorithmic bytes the launch writes
  LaunchTimer(int cls, cudaStream_t s, uint32_t tag = 0, uint64_t bytes = 0) : stream(s), on(g_timing.load() != 0) {
    sp.tag = tag; sp.bytes = bytes;
    if (!on) return;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { on = false; return; }
    std::lock_guard<std::mutex> lk(g_timing_mu);
    sp.cls = cls; sp.dev = dev; sp.a = pool_event(dev); sp.b = pool_event(dev);
    if (!sp.a || !sp.b) { on = false; return; }
    cudaEventRecord(sp.a, stream);
  }
  ~LaunchTimer() {
    if (!on) return;
    cudaEventRecord(sp.b, stream);
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_spans.push_back(sp);
  }
};

// ---- per-device persistent state -------------------------------------------------------------
struct CachedRom {
  I2* ptr = nullptr;
  uint32_t entries = 0;
};
static const int kMaxSideStreams = 8;
struct HostPipe {  // staging of the *_host entry points: two device chunks, two streams
  cudaStream_t s_gen = nullptr, s_copy = nullptr;
  cudaEvent_t ev_gen[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
  char* buf[2] = {nullptr, nullptr};
  size_t buf_bytes = 0;
};
struct DeviceState {
  std::mutex mu;
  std::map<uint32_t, CachedRom> roms;  // key: dw << 8 | lut
  cudaMemPool_t pool = nullptr;        // stream-ordered pool of the one-shot plans (kept warm between calls)
  std::mutex pipe_mu;                  // one host-buffer call at a time per device
  HostPipe pipe;
  // fork/join of the independent launches of one execute (see LaunchFan)
  std::mutex fan_mu;
  cudaStream_t side[kMaxSideStreams] = {nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxSideStreams] = {nullptr};
};
static DeviceState g_dev[64];

static int get_rom(int dev, int dw, int lut, const I2** out) {
  DeviceState& ds = g_dev[dev];
  std::lock_guard<std::mutex> lk(ds.mu);
  const uint32_t key = ((uint32_t)dw << 8) | (uint32_t)lut;
  auto it = ds.roms.find(key);
  if (it == ds.roms.end()) {
    std::vector<I2> rom;
    build_taylor_rom(dw, lut, rom);
    CachedRom cr;
    cr.entries = (uint32_t)rom.size();
    BHW_CUDA(cudaMalloc((void**)&cr.ptr, rom.size() * sizeof(I2)));
    // blocking copy: ROMs are tiny and uploaded once per (device, DW, LUT_SIZE)
    cudaError_t e = cudaMemcpy(cr.ptr, rom.data(), rom.size() * sizeof(I2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(cr.ptr); return cuda_fail(e, "cudaMemcpy(rom)"); }
    it = ds.roms.emplace(key, cr).first;
  }
  *out = it->second.ptr;
  return BHW_OK;
}

// The library's own stream-ordered pool: one-shot plans take their tables and records from it.
// Freed blocks stay in the pool (up to kPoolKeepBytes) instead of going back to the driver at
// the next synchronisation, so a repeated call does not pay cudaMalloc again.
static const uint64_t kPoolKeepBytes = 1ull << 30;

static cudaError_t device_pool(int dev, cudaMemPool_t* out) {
  DeviceState& ds = g_dev[dev];
  std::lock_guard<std::mutex> lk(ds.mu);
  if (!ds.pool) {
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    cudaError_t e = cudaMemPoolCreate(&ds.pool, &props);
    if (e != cudaSuccess) { ds.pool = nullptr; return e; }
    uint64_t keep = kPoolKeepBytes;
    cudaMemPoolSetAttribute(ds.pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  *out = ds.pool;
  return cudaSuccess;
}

static int current_device(int* dev) {
  cudaError_t e = cudaGetDevice(dev);
  if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return BHW_E_NO_DEVICE; }
  if (*dev < 0 || *dev >= 64) return BHW_E_NO_DEVICE;
  return BHW_OK;
}

}  // namespace bhw

// ---- the plan object ---------------------------------------------------------------------------
struct PlanTable {
  bhw::SrcParams canon;
  uint32_t drop;
  uint32_t entries;
  int32_t* ptr;
  uint32_t* exc = nullptr;   // input-quadrant CORDIC tables some bank run pairs: entries that break T[i+E/2] == ~T[i]
  bool need_exc = false;
};

struct bhw_plan {
  int dev = 0;
  int nwin = 0;
  bool elem64 = false;
  bool pack16 = false;   // BHW_OUT_INT16: int16 output - k_synth<short>, k_synth_group<.., 3>, k_synth_bank<.., short> (2/3 terms)
  bool transient = false;        // one-shot plan: device memory comes from / returns to the stream pool
  uint64_t total = 0;            // flat samples of the whole batch
  // DAT_WIDTH <= 32: table + synthesis path
  std::vector<bhw::WinRec> recs;
  std::vector<bhw::GenRec> gens;
  std::vector<uint32_t> win_rec;
  std::vector<uint64_t> flat_off;
  std::vector<PlanTable> tables;
  std::vector<bhw::TabJob> jobs;       // small / other-core jobs: one combined k_table_build launch
  std::vector<bhw::TabJob> big_jobs;   // large 32-bit-core jobs: one k_table_build_u launch each
  // Taylor sine ROMs of the plan, one per distinct (DAT_WIDTH, LUT_SIZE), concatenated; they live in
  // the plan's own metadata blob (the per-device ROM cache only serves the one-shot direct kernels)
  std::vector<bhw::I2> rom_host;
  std::map<uint32_t, uint32_t> rom_offs;   // (dw << 8 | lut) -> offset in rom_host (I2 units)
  const bhw::I2* rom = nullptr;            // device copy (inside blob_dev)
  int uniform_pw = -1;
  bool all_same = false;
  char* blob_dev = nullptr;      // [recs][gens][jobs][flat_off][win_rec]
  size_t o_recs = 0, o_gens = 0, o_jobs = 0, o_off = 0, o_wr = 0, o_rom = 0;
  uint32_t table_work = 0;
  // Tables kept by the plan (table cache on): built by the first eager execute; `ev_built` is recorded
  // behind that build so that executes on other streams can order themselves after it.
  bool tables_built = false;
  cudaEvent_t ev_built = nullptr;
  cudaStream_t built_stream = nullptr;
  // runs of consecutive same-shape windows that the bank kernel can take whole
  struct BankRun {
    int w_begin, w_end;      // windows [w_begin, w_end)
    uint64_t flat_off;       // flat sample of window w_begin
    bhw::BankShape sh;
    int tab_mode;
    bool pair;
    // shape for a tile range inside one window of the run (unpaired, table read from global memory)
    bhw::BankShape sh_part;
    bool part_ok;
    int exc_table = -1;      // paired over an input-quadrant CORDIC table: that table (its exception list feeds the patch pass)
  };
  std::vector<BankRun> runs;
  // families (one half-period pyramid each) and groups (bhw_group.cuh): windows that k_synth_group takes
  struct Family {
    bhw::SrcParams key;          // canonical source at the maximum PHI_WIDTH: the family's identity
    bhw::SrcParams canon;        // canonical source at PHI_WIDTH = top: what the pyramid job evaluates
    bhw_desc proto;              // a descriptor of the family
    uint32_t top = 0, lmin = 0;  // pyramid levels
    uint32_t max_pw = 0, min_pw = 99;
    int tab_mode = 0;            // G_*
    int32_t* pyr = nullptr;      // 2^top words
    uint16_t* q16 = nullptr;     // G_Q16: 2^(top-1) uint16
    int big_job = -1;            // index in big_jobs when the pyramid has a launch of its own
  };
  struct Group {
    int family = 0;
    bhw::GroupShape sh;
    std::vector<uint32_t> wins;          // member windows, ascending
    std::vector<bhw::GroupWin> list;     // whole-window (paired) launch list, one entry per member + sentinel
    size_t o_list = 0;                   // its place in blob_dev
    // long windows over a pyramid in L2 / HBM get a launch of their own (spread walk, k_synth_group):
    // entry + sentinel per window
    std::vector<bhw::GroupWin> singles;
    size_t o_singles = 0;
  };
  std::vector<Family> families;
  std::vector<Group> groups;
  // what an execute asks of each group: the piece [i0, i1) of its launch list, up to two windows the range cuts,
  // the windows launched singly; `covered` = the flat intervals those launches write (ascending, merged)
  struct Covered { uint64_t b, e; };
  struct GroupWork { uint32_t i0 = 0, i1 = 0; bhw::GroupWin part[2]; uint32_t nparts = 0; std::vector<uint32_t> singles; };
  // ... of an execute of the whole batch, kept from the first one (a bank of 65,536 short windows: the walk over
  // the windows costs more host time than the kernels take)
  bool whole_cached = false;
  std::vector<GroupWork> whole_gwork;
  std::vector<Covered> whole_covered;
  size_t whole_group_launches = 0;
  std::vector<int32_t> win_group;        // [nwin] group of each window, -1: none
  std::vector<uint32_t> win_gidx;        // [nwin] index of the window in its group's list (or in `singles`, bit 31 set)
  // DAT_WIDTH > 32: one direct launch per window
  std::vector<bhw::DirectArgs> wins64;
  std::mutex mu;
};

namespace bhw {

// work items from which a table job gets its own stage-unrolled launch
static const uint32_t kBigTableWork = 1u << 13;

static int find_or_add_table(bhw_plan& plan, const SrcParams& sp, int* index) {
  uint32_t drop;
  const SrcParams canon = canonical_source(sp, &drop);
  for (size_t i = 0; i < plan.tables.size(); i++)
    if (!memcmp(&plan.tables[i].canon, &canon, sizeof(canon))) { *index = (int)i; return BHW_OK; }
  PlanTable pt;
  pt.canon = canon;
  pt.drop = drop;
  pt.entries = 1u << canon.pw;
  pt.ptr = nullptr;
  plan.tables.push_back(pt);
  *index = (int)plan.tables.size() - 1;
  return BHW_OK;
}

// offset (I2 units) of the (dw, lut) sine ROM inside the plan's ROM block, appending it on first use
static uint32_t plan_rom_off(bhw_plan& plan, int dw, int lut) {
  const uint32_t key = ((uint32_t)dw << 8) | (uint32_t)lut;
  auto it = plan.rom_offs.find(key);
  if (it != plan.rom_offs.end()) return it->second;
  std::vector<I2> rom;
  build_taylor_rom(dw, lut, rom);
  const uint32_t off = (uint32_t)plan.rom_host.size();
  plan.rom_host.insert(plan.rom_host.end(), rom.begin(), rom.end());
  plan.rom_offs.emplace(key, off);
  return off;
}

static void plan_free_device(bhw_plan& plan, cudaStream_t stream) {
  if (plan.ev_built) { cudaEventDestroy(plan.ev_built); plan.ev_built = nullptr; }
  for (auto& pt : plan.tables)
    if (pt.ptr) { if (plan.transient) cudaFreeAsync(pt.ptr, stream); else cudaFree(pt.ptr); pt.ptr = nullptr; }
  for (auto& pt : plan.tables)
    if (pt.exc) { if (plan.transient) cudaFreeAsync(pt.exc, stream); else cudaFree(pt.exc); pt.exc = nullptr; }
  for (auto& f : plan.families) {
    if (f.pyr) { if (plan.transient) cudaFreeAsync(f.pyr, stream); else cudaFree(f.pyr); f.pyr = nullptr; }
    if (f.q16) { if (plan.transient) cudaFreeAsync(f.q16, stream); else cudaFree(f.q16); f.q16 = nullptr; }
  }
  if (plan.blob_dev) {
    if (plan.transient) cudaFreeAsync(plan.blob_dev, stream); else cudaFree(plan.blob_dev);
    plan.blob_dev = nullptr;
  }
}

static cudaError_t plan_alloc(bhw_plan& plan, void** p, size_t bytes, cudaStream_t stream) {
  if (!plan.transient) return cudaMalloc(p, bytes);
  cudaMemPool_t pool = nullptr;
  cudaError_t e = device_pool(plan.dev, &pool);
  if (e != cudaSuccess) return e;
  return cudaMallocFromPoolAsync(p, bytes, pool, stream);
}


static void build_bank_runs(bhw_plan& plan, const std::vector<int>& rec_tab) {
  plan.runs.clear();
  uint64_t off = 0;
  for (int w = 0; w < plan.nwin; w++) {
    const uint32_t ri = plan.win_rec[(size_t)w];
    const WinRec& r = plan.recs[ri];
    const uint64_t N = 1ull << r.pw;
    BankShape sh;
    int mode = 0;
    bool pair = false;
    BankTableInfo ti[BHW_MAX_TERMS];
    for (int k = 0; k < BHW_MAX_TERMS; k++) {
      const int t = rec_tab[(size_t)ri * BHW_MAX_TERMS + (size_t)k];
      ti[k].ptr = t < 0 ? nullptr : plan.tables[(size_t)t].ptr;
      ti[k].entries = t < 0 ? 0 : plan.tables[(size_t)t].entries;
      ti[k].antisym = t >= 0 && source_antisymmetric(plan.tables[(size_t)t].canon);
      ti[k].inq_comp = t >= 0 && source_inq_complement(plan.tables[(size_t)t].canon);
    }
    // a packed plan (int16 output) has bank kernels for 2- and 3-term shapes with the 32-bit tail only
    if (bank_shape(r, ti, bank_smem_limit(), &sh, &mode, &pair) && (!plan.pack16 || (sh.m <= 3 && !sh.acc64))) {
      if (!plan.runs.empty()) {
        bhw_plan::BankRun& last = plan.runs.back();
        if (last.w_end == w && last.tab_mode == mode && last.pair == pair && !memcmp(&last.sh, &sh, sizeof(sh))) {
          last.w_end = w + 1;
          off += N;
          continue;
        }
      }
      bhw_plan::BankRun run;
      run.w_begin = w; run.w_end = w + 1; run.flat_off = off; run.sh = sh; run.tab_mode = mode; run.pair = pair;
      int mode_p = 0;
      bool pair_p = false;
      run.part_ok = bank_shape(r, ti, 0, &run.sh_part, &mode_p, &pair_p, false) && mode_p == TAB_GLOBAL && !pair_p;
      if (pair && sh.pair_adj) run.exc_table = rec_tab[(size_t)ri * BHW_MAX_TERMS + 1];
      plan.runs.push_back(run);
    }
    off += N;
  }
  // A run too short to be worth a launch of its own (table staging, one CTA or two) is left to the
  // general kernel, which takes any number of neighbouring short windows in one launch.
  // (measured on the win_selector sweep: 2^17 samples is the best cut - 1.68 -> 1.58 ms per sweep, 200 -> 110
  // launches; all ten variants x PHI_WIDTH 4..14 in one plan: 257 -> 14 us)
  const uint64_t min_run = 1ull << 17;
  size_t keep = 0;
  for (size_t i = 0; i < plan.runs.size(); i++) {
    const bhw_plan::BankRun& run = plan.runs[i];
    if (((uint64_t)(run.w_end - run.w_begin) << run.sh.pw) >= min_run) plan.runs[keep++] = run;
  }
  plan.runs.resize(keep);
  // Pairing over an input-quadrant CORDIC's table costs two more launches (exception scan, patch pass): below
  // 2^23 samples the run keeps single samples (a one-shot N = 2^20 cordic_dds48 window: 40 us unpaired, 51 us paired)
  // (a packed plan has no bank kernel for that pairing: it keeps single samples, or leaves the run to k_synth)
  if (plan.pack16) {
    keep = 0;
    for (size_t i = 0; i < plan.runs.size(); i++)
      if (plan.runs[i].exc_table < 0 || plan.runs[i].part_ok) plan.runs[keep++] = plan.runs[i];
    plan.runs.resize(keep);
  }
  for (bhw_plan::BankRun& run : plan.runs) {
    if (run.exc_table < 0) continue;
    if ((((uint64_t)(run.w_end - run.w_begin) << run.sh.pw) < (1ull << 23) || plan.pack16) && run.part_ok) {
      run.sh = run.sh_part;
      run.tab_mode = TAB_GLOBAL;
      run.pair = false;
      run.exc_table = -1;
    } else {
      plan.tables[(size_t)run.exc_table].need_exc = true;
    }
  }
}

// Resolve a batch into `plan` and make it resident on the current device.  [hint_begin,
// hint_begin+hint_count) is the flat range the caller is going to execute (the whole batch for
// persistent plans): windows outside it get no tables, and BHW_ALGO_AUTO sends a short request
// into a long window through the direct body instead of building a table for it.
static int plan_build(bhw_plan& plan, const bhw_desc* descs, int nwin, uint64_t hint_begin,
                      uint64_t hint_count, cudaStream_t stream) {
  size_t esz_ = 4;
  int st = batch_elem_bytes(descs, nwin, &esz_);   // before anything touches a device: the error is the caller's
  if (st) return st;
  if ((st = current_device(&plan.dev))) return st;
  plan.nwin = nwin;
  plan.elem64 = esz_ == 8;
  plan.pack16 = esz_ == 2;
  plan.flat_off.resize((size_t)nwin + 1);
  uint64_t off = 0;
  for (int w = 0; w < nwin; w++) {
    if (descs[w].phi_width < BHW_MIN_PHI_WIDTH || descs[w].phi_width > BHW_MAX_PHI_WIDTH) return BHW_E_PHI_WIDTH;
    plan.flat_off[w] = off;
    off += 1ull << descs[w].phi_width;
  }
  plan.flat_off[nwin] = off;
  plan.total = off;
  if (hint_begin > off || hint_count > off - hint_begin) return BHW_E_RANGE;
  const uint64_t hint_end = hint_begin + hint_count;

  if (plan.elem64) {
    // DAT_WIDTH > 32: window by window through the direct kernel
    plan.wins64.resize((size_t)nwin);
    for (int w = 0; w < nwin; w++) {
      DirectArgs& a = plan.wins64[(size_t)w];
      memset(&a, 0, sizeof(a));
      if ((st = resolve_window(&descs[w], &a.wp, a.src))) return st;
      for (int u = 0; u < a.wp.nsrc; u++) init_src_core(a.src[u], &a.sc[u]);
    }
    return BHW_OK;
  }

  // pass 1: windows with a byte-identical descriptor share one record.  Runs of two or more consecutive
  // windows of one shape (a bank: everything but the ports equal) of at least 2^17 samples keep the uniform
  // bank kernel and its per-shape tables (`banked`); everything else may join a group (bhw_group.cuh).
  auto same_shape = [](const bhw_desc& a, const bhw_desc& b) {
    return a.win_type == b.win_type && a.sin_type == b.sin_type && a.model == b.model && a.phi_width == b.phi_width &&
           a.dat_width == b.dat_width && a.precision == b.precision && a.lut_size == b.lut_size && a.algo == b.algo;
  };
  std::vector<uint8_t> banked((size_t)nwin, 0);
  for (int w = 0; w < nwin;) {
    int e = w + 1;
    while (e < nwin && same_shape(descs[w], descs[e])) e++;
    if ((!plan.pack16 || descs[w].win_type <= 3) && e - w >= 2 && ((uint64_t)(e - w) << descs[w].phi_width) >= (1ull << 17))
      for (int i = w; i < e; i++) banked[(size_t)i] = 1;
    w = e;
  }
  plan.win_rec.resize((size_t)nwin);
  std::vector<int> rec_desc;       // descriptor index that defines record i
  std::vector<uint64_t> rec_used;  // hinted samples that fall into windows of record i
  std::unordered_map<std::string, int> seen;
  const int pw0 = descs[0].phi_width;
  bool uniform = true;
  for (int w = 0; w < nwin; w++) {
    const bhw_desc& d = descs[w];
    if (d.phi_width != pw0) uniform = false;
    const uint64_t b = plan.flat_off[w], e = plan.flat_off[w + 1];
    const uint64_t lo = b > hint_begin ? b : hint_begin;
    const uint64_t hi = e < hint_end ? e : hint_end;
    const uint64_t used = lo < hi ? hi - lo : 0;
    int ri;
    if (w > 0 && banked[(size_t)w] == banked[(size_t)w - 1] && !memcmp(&d, &descs[w - 1], sizeof(d))) ri = (int)plan.win_rec[w - 1];
    else {
      std::string key((const char*)&d, sizeof(d));
      key.push_back((char)banked[(size_t)w]);
      auto ins = seen.emplace(std::move(key), (int)rec_desc.size());
      ri = ins.first->second;
      if (ins.second) { rec_desc.push_back(w); rec_used.push_back(0); }
    }
    plan.win_rec[w] = (uint32_t)ri;
    rec_used[(size_t)ri] += used;
  }
  plan.uniform_pw = uniform ? pw0 : -1;
  plan.all_same = rec_desc.size() == 1;

  // pass 2: validate + resolve every distinct descriptor, give the used ones their trig tables
  struct TabRef { int rec, k, tab; };
  std::vector<TabRef> refs;
  std::vector<int> rec_family(rec_desc.size(), -1);
  plan.recs.reserve(rec_desc.size());
  for (size_t i = 0; i < rec_desc.size(); i++) {
    const bhw_desc& d = descs[rec_desc[i]];
    WinParams wp; SrcParams src[2];
    if ((st = resolve_window(&d, &wp, src))) return st;
    WinRec r;
    memset(&r, 0, sizeof(r));
    r.n_first = (uint32_t)wp.stream_offset;
    bool generic = d.algo == BHW_ALGO_DIRECT || fast_tail_mode(wp, src) == TAILMODE_GENERIC;
    if (!generic && d.algo == BHW_ALGO_AUTO && rec_used[i]) {
      uint64_t table_work = 0;
      for (int u = 0; u < wp.nsrc; u++) {
        uint32_t drop;
        const SrcParams canon = canonical_source(src[u], &drop);
        table_work += (1ull << canon.pw) / (canon.kind == SRC_INQ ? 1 : 4);
      }
      if (rec_used[i] * (uint64_t)(wp.m - 1) < table_work / 4) generic = true;
    }
    if (generic) {
      GenRec g;
      memset(&g, 0, sizeof(g));
      g.wp = wp; g.src[0] = src[0]; g.src[1] = src[1];
      r.flags = WR_GENERIC;
      r.m = (uint32_t)wp.m; r.dw = (uint32_t)wp.dw; r.pw = (uint32_t)wp.pw;
      r.gen_idx = (uint32_t)plan.gens.size();
      if (src[0].kind == SRC_TAYLOR && rec_used[i]) g.rom_off = plan_rom_off(plan, src[0].dw, src[0].lut);
      plan.gens.push_back(g);
    } else if (!banked[(size_t)rec_desc[i]] && rec_used[i] && group_eligible(d, wp, src)) {
      // member of a family: no table of its own, it reads the family's pyramid (pointers set below)
      fill_fast_rec(wp, src, r);
      SrcParams key;
      if ((st = family_source(d, BHW_MAX_PHI_WIDTH, &key))) return st;
      int f = -1;
      for (size_t j = 0; j < plan.families.size(); j++)
        if (!memcmp(&plan.families[j].key, &key, sizeof(key))) { f = (int)j; break; }
      if (f < 0) {
        bhw_plan::Family fam;
        fam.key = key;
        fam.proto = d;
        plan.families.push_back(fam);
        f = (int)plan.families.size() - 1;
      }
      bhw_plan::Family& fam = plan.families[(size_t)f];
      if ((uint32_t)wp.pw > fam.max_pw) fam.max_pw = (uint32_t)wp.pw;
      if ((uint32_t)wp.pw < fam.min_pw) fam.min_pw = (uint32_t)wp.pw;
      rec_family[i] = f;
    } else {
      fill_fast_rec(wp, src, r);
      if (rec_used[i]) {
        for (int k = 1; k < wp.m; k++) {
          const TermParams& t = wp.term[k - 1];
          const SrcParams& sp = src[t.src];
          int ti;
          if ((st = find_or_add_table(plan, sp, &ti))) return st;
          r.kstep[k] = t.kmul << (32 - sp.pw);
          // full-period index = phase32 >> (32 - log2(entries)): windows whose PHI_WIDTH exceeds
          // the source's input resolution share one table and simply drop more phase bits
          r.idx_rsh[k] = (uint32_t)(32 - plan.tables[(size_t)ti].canon.pw);
          refs.push_back({(int)i, k, ti});
        }
      }
    }
    plan.recs.push_back(r);
  }

  // device memory: tables, then the metadata blob (records carry the table pointers)
  uint32_t work = 0;
  for (auto& pt : plan.tables) {
    cudaError_t e = plan_alloc(plan, (void**)&pt.ptr, (size_t)pt.entries * sizeof(int32_t), stream);
    if (e != cudaSuccess) return cuda_fail(e, "alloc(trig table)");
    TabJob j;
    init_tab_job(pt.canon, pt.ptr, &j);
    if (j.work >= kBigTableWork && table_build_unrolled_ok(j)) {
      j.work_begin = 0;
      plan.big_jobs.push_back(j);
      continue;
    }
    j.work_begin = work;
    if (pt.canon.kind == SRC_TAYLOR) j.rom_off = plan_rom_off(plan, pt.canon.dw, pt.canon.lut);
    work += j.work;
    plan.jobs.push_back(j);
  }
  // families: one half-period pyramid each, built by one more table job
  for (auto& fam : plan.families) {
    const uint32_t res = (uint32_t)fam.key.pw;                   // phase bits the source looks at
    fam.top = fam.max_pw < res ? fam.max_pw : res;
    const uint32_t low = fam.min_pw < fam.top ? fam.min_pw : fam.top;
    fam.lmin = low > 4 ? low - 2 : 2;                            // even harmonics read up to two levels down
    if ((st = family_source(fam.proto, (int)fam.top, &fam.canon))) return st;
    fam.tab_mode = group_tab_mode(fam.canon, fam.top, group_smem_limit());
    cudaError_t e = plan_alloc(plan, (void**)&fam.pyr, sizeof(int32_t) << fam.top, stream);
    if (e != cudaSuccess) return cuda_fail(e, "alloc(table pyramid)");
    if (fam.tab_mode == G_Q16) {
      e = plan_alloc(plan, (void**)&fam.q16, sizeof(uint16_t) << (fam.top - 1), stream);
      if (e != cudaSuccess) return cuda_fail(e, "alloc(quarter-wave image)");
    }
    TabJob j;
    init_pyramid_job(fam.canon, fam.lmin, fam.pyr, fam.q16, &j);
    if (j.work >= kBigTableWork && table_build_unrolled_ok(j)) {
      j.work_begin = 0;
      fam.big_job = (int)plan.big_jobs.size();
      plan.big_jobs.push_back(j);
    } else {
      j.work_begin = work;
      work += j.work;
      plan.jobs.push_back(j);
    }
  }
  // records of family members read the pyramid level of their own PHI_WIDTH (general kernel)
  for (size_t i = 0; i < plan.recs.size(); i++) {
    if (rec_family[i] < 0) continue;
    const bhw_plan::Family& fam = plan.families[(size_t)rec_family[i]];
    WinRec& r = plan.recs[i];
    const uint32_t L = r.pw < fam.top ? r.pw : fam.top;
    r.flags |= WR_HALFTAB;
    for (uint32_t k = 1; k < r.m; k++) {
      r.kstep[k] = k << (32 - r.pw);
      r.idx_rsh[k] = 32 - L;
      r.tabp[k] = fam.pyr + ((size_t)1 << (L - 1));
    }
  }
  // groups: the windows of a family with one entity (number of terms) and tail, in window order
  plan.win_group.assign((size_t)nwin, -1);
  plan.win_gidx.assign((size_t)nwin, 0);
  for (int w = 0; w < nwin && !plan.families.empty(); w++) {
    const uint32_t ri = plan.win_rec[(size_t)w];
    const int f = rec_family[ri];
    if (f < 0) continue;
    const WinRec& r = plan.recs[ri];
    int g = -1;
    for (size_t j = 0; j < plan.groups.size(); j++) {
      const bhw_plan::Group& gr = plan.groups[j];
      if (gr.family == f && gr.sh.m == r.m && gr.sh.rc == r.rc && gr.sh.lsh == r.lsh && gr.sh.rsh == r.rsh) { g = (int)j; break; }
    }
    if (g < 0) {
      bhw_plan::Group gr;
      gr.family = f;
      const bhw_plan::Family& fam = plan.families[(size_t)f];
      group_shape(r, fam.top, fam.lmin, &gr.sh);
      gr.sh.pyr = fam.pyr;
      gr.sh.q16 = fam.q16;
      plan.groups.push_back(gr);
      g = (int)plan.groups.size() - 1;
    }
    bhw_plan::Group& gr = plan.groups[(size_t)g];
    plan.win_group[(size_t)w] = g;
    gr.wins.push_back((uint32_t)w);
  }
  const uint32_t kSpreadG = 30;
  for (auto& gr : plan.groups) {
    const bool can_spread = plan.families[(size_t)gr.family].tab_mode == G_GLOBAL && (int)gr.sh.m >= g_spread_min_terms.load();
    uint32_t units = 0;
    for (uint32_t w : gr.wins) {
      GroupWin gw;
      memset(&gw, 0, sizeof(gw));
      gw.pw = (uint32_t)descs[w].phi_width;
      gw.rec = plan.win_rec[w];
      gw.out_off = (int64_t)plan.flat_off[w];
      const uint32_t wu = 1u << (gw.pw - kBankTileLog2 - 1);   // tiles of 256 sample pairs
      if (can_spread && wu >= kSpreadG * (uint32_t)device_sm_count() * 2u) {
        plan.win_gidx[w] = 0x80000000u | (uint32_t)(gr.singles.size() / 2);
        gr.singles.push_back(gw);                              // unit_begin 0
        GroupWin end;
        memset(&end, 0, sizeof(end));
        end.unit_begin = wu;
        gr.singles.push_back(end);
        continue;
      }
      plan.win_gidx[w] = (uint32_t)gr.list.size();
      gw.unit_begin = units;
      gr.list.push_back(gw);
      units += wu;
    }
    GroupWin end;
    memset(&end, 0, sizeof(end));
    end.unit_begin = units;
    gr.list.push_back(end);
  }
  plan.table_work = work;
  std::vector<int> rec_tab(plan.recs.size() * BHW_MAX_TERMS, -1);
  for (const TabRef& tr : refs) {
    plan.recs[(size_t)tr.rec].tabp[tr.k] = plan.tables[(size_t)tr.tab].ptr;
    rec_tab[(size_t)tr.rec * BHW_MAX_TERMS + (size_t)tr.k] = tr.tab;
  }
  build_bank_runs(plan, rec_tab);
  for (auto& pt : plan.tables) {
    if (!pt.need_exc) continue;
    cudaError_t ee = plan_alloc(plan, (void**)&pt.exc, ((size_t)pt.entries / 2 + 1) * sizeof(uint32_t), stream);
    if (ee != cudaSuccess) return cuda_fail(ee, "alloc(exception list)");
  }

  auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const bool need_off = plan.uniform_pw < 0, need_wr = !plan.all_same;
  plan.o_recs = 0;
  plan.o_gens = align16(plan.o_recs + plan.recs.size() * sizeof(WinRec));
  plan.o_jobs = align16(plan.o_gens + plan.gens.size() * sizeof(GenRec));
  plan.o_off = align16(plan.o_jobs + plan.jobs.size() * sizeof(TabJob));
  plan.o_wr = align16(plan.o_off + (need_off ? plan.flat_off.size() * sizeof(uint64_t) : 0));
  plan.o_rom = align16(plan.o_wr + (need_wr ? plan.win_rec.size() * sizeof(uint32_t) : 0));
  size_t total = align16(plan.o_rom + plan.rom_host.size() * sizeof(I2));
  for (auto& gr : plan.groups) {
    gr.o_list = total; total = align16(total + gr.list.size() * sizeof(GroupWin));
    gr.o_singles = total; total = align16(total + gr.singles.size() * sizeof(GroupWin));
  }
  std::vector<char> blob(total);
  for (auto& gr : plan.groups) {
    memcpy(blob.data() + gr.o_list, gr.list.data(), gr.list.size() * sizeof(GroupWin));
    if (!gr.singles.empty()) memcpy(blob.data() + gr.o_singles, gr.singles.data(), gr.singles.size() * sizeof(GroupWin));
  }
  if (!plan.rom_host.empty()) memcpy(blob.data() + plan.o_rom, plan.rom_host.data(), plan.rom_host.size() * sizeof(I2));
  memcpy(blob.data() + plan.o_recs, plan.recs.data(), plan.recs.size() * sizeof(WinRec));
  if (!plan.gens.empty()) memcpy(blob.data() + plan.o_gens, plan.gens.data(), plan.gens.size() * sizeof(GenRec));
  if (!plan.jobs.empty()) memcpy(blob.data() + plan.o_jobs, plan.jobs.data(), plan.jobs.size() * sizeof(TabJob));
  if (need_off) memcpy(blob.data() + plan.o_off, plan.flat_off.data(), plan.flat_off.size() * sizeof(uint64_t));
  if (need_wr) memcpy(blob.data() + plan.o_wr, plan.win_rec.data(), plan.win_rec.size() * sizeof(uint32_t));
  cudaError_t e = plan_alloc(plan, (void**)&plan.blob_dev, total, stream);
  if (e != cudaSuccess) return cuda_fail(e, "alloc(plan metadata)");
  // pageable source: the runtime stages it before returning, so `blob` may die with this frame
  e = cudaMemcpyAsync(plan.blob_dev, blob.data(), total, cudaMemcpyHostToDevice, stream);
  if (e != cudaSuccess) return cuda_fail(e, "cudaMemcpyAsync(plan metadata)");
  if (!plan.transient) {
    e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaStreamSynchronize(plan)");
  }
  if (!plan.rom_host.empty()) plan.rom = (const I2*)(plan.blob_dev + plan.o_rom);
  return BHW_OK;
}

// The synthesis launches of one execute write disjoint output ranges and only read the trig
// tables, so they are independent of each other.  A batch of many differently shaped windows
// (the win_selector sweep: ~200 launches, most of them far too small to fill 148 SMs) is
// therefore fanned out over a few side streams: an event recorded on the caller's stream after
// the table build forks them, one event per side stream joins them back.  Stream-ordered as a
// whole - the caller sees nothing but its own stream.
struct LaunchFan {
  DeviceState* ds = nullptr;
  cudaStream_t main = nullptr;
  int nside = 0;          // side streams in use for this execute (0: everything on `main`)
  unsigned used = 0;      // side streams that received work
  unsigned rr = 0;
  cudaError_t err = cudaSuccess;
  std::unique_lock<std::mutex> lock;

  // `launches`: how many independent launches are coming
  void begin(int dev, cudaStream_t stream, size_t launches) {
    main = stream;
    int want = g_side_streams.load();
    if (want > kMaxSideStreams) want = kMaxSideStreams;
    if (want <= 0 || launches < 4) return;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) return;
    ds = &g_dev[dev];
    lock = std::unique_lock<std::mutex>(ds->fan_mu);
    if (!ds->ev_fork && (err = cudaEventCreateWithFlags(&ds->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return;
    for (int i = 0; i < want; i++) {
      if (!ds->side[i] && (err = cudaStreamCreateWithFlags(&ds->side[i], cudaStreamNonBlocking)) != cudaSuccess) return;
      if (!ds->ev_join[i] && (err = cudaEventCreateWithFlags(&ds->ev_join[i], cudaEventDisableTiming)) != cudaSuccess) return;
    }
    if ((err = cudaEventRecord(ds->ev_fork, main)) != cudaSuccess) return;
    nside = want;
  }
  unsigned reserved = 0;  // side streams handed out by reserve(): next() leaves them alone
  // a side stream of its own for a chain of dependent launches (a family's table build and its groups);
  // `main` when none is free
  cudaStream_t reserve() {
    for (int i = 0; i < nside; i++) {
      if (reserved & (1u << i)) continue;
      if (!(used & (1u << i))) {
        cudaError_t e = cudaStreamWaitEvent(ds->side[i], ds->ev_fork, 0);
        if (e != cudaSuccess) { err = e; return main; }
        used |= 1u << i;
      }
      reserved |= 1u << i;
      return ds->side[i];
    }
    return main;
  }
  // stream of the next launch
  cudaStream_t next() {
    if (!nside) return main;
    unsigned slot = rr++ % (unsigned)(nside + 1);
    for (int tries = 0; slot != 0 && (reserved & (1u << (slot - 1))) && tries <= nside; tries++) slot = rr++ % (unsigned)(nside + 1);
    if (slot == 0 || (reserved & (1u << (slot - 1)))) return main;
    const unsigned i = slot - 1;
    if (!(used & (1u << i))) {
      cudaError_t e = cudaStreamWaitEvent(ds->side[i], ds->ev_fork, 0);
      if (e != cudaSuccess) { err = e; return main; }
      used |= 1u << i;
    }
    return ds->side[i];
  }
  // make `main` wait for everything that went to a side stream
  cudaError_t join() {
    for (int i = 0; i < nside; i++) {
      if (!(used & (1u << i))) continue;
      cudaError_t e = cudaEventRecord(ds->ev_join[i], ds->side[i]);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(main, ds->ev_join[i], 0);
      if (e != cudaSuccess && err == cudaSuccess) err = e;
    }
    used = 0;
    reserved = 0;
    nside = 0;
    return err;
  }
  ~LaunchFan() { join(); }
};

// fused apply step of bhw_apply: the plan's single whole window multiplies `frames` frames of x instead of
// being stored
struct ApplyInfo { const int32_t* x; void* y; uint64_t frames; int mode; int dw; };

static int plan_execute(bhw_plan& plan, uint64_t flat_begin, uint64_t flat_count, void* out_dev,
                        cudaStream_t stream, const ApplyInfo* ap = nullptr) {
  if (flat_begin > plan.total || flat_count > plan.total - flat_begin) return BHW_E_RANGE;
  if (!flat_count) return BHW_OK;
  if (!out_dev) return BHW_E_NULL;
  if (plan.elem64) {
    uint64_t off = 0;
    for (int w = 0; w < plan.nwin; w++) {
      DirectArgs a = plan.wins64[(size_t)w];
      const uint64_t N = 1ull << a.wp.pw;
      const uint64_t lo = off > flat_begin ? off : flat_begin;
      const uint64_t hi = off + N < flat_begin + flat_count ? off + N : flat_begin + flat_count;
      if (lo < hi) {
        a.n_first = (lo - off) + (uint64_t)a.wp.stream_offset;
        a.count = hi - lo;
        uint32_t flip = 0;   // a whole window: sample pairs (n, n + N/2) from one evaluation per harmonic
        a.pair_flip = (lo == off && hi == off + N && N >= 8 && direct_pair_flip(a.wp, a.src, &flip)) ? (0x80000000u | flip) : 0u;
        uint32_t adv = 0;
        a.quad_adv = (a.pair_flip && N / 2 > (uint64_t)device_sm_count() * 256u && direct_quad_adv(a.wp, a.src, &adv))
                         ? (0x80000000u | adv) : 0u;
        cudaError_t e;
        {
          LaunchTimer tm(BHW_KERNEL_DIRECT, stream);
          e = launch_direct_window(a, (int64_t*)out_dev + (lo - flat_begin), stream);
        }
        if (e != cudaSuccess) return cuda_fail(e, "k_direct_window");
        g_launches++;
      }
      off += N;
    }
    return BHW_OK;
  }
  std::lock_guard<std::mutex> lk(plan.mu);
  cudaError_t e = cudaSuccess;
  bool table_ahead = false;  // k_table_build is the last thing enqueued on `stream`
  bool tm_on = false;
  bool building = false, defer_family_builds = false;
  const bool have_jobs = !plan.jobs.empty() || !plan.big_jobs.empty();
  // a one-shot plan lives for one call: the chunks of a host-buffer request share the tables its first chunk built
  const bool keep = plan.transient || g_cache_enabled.load() != 0;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
  const bool capturing = cap != cudaStreamCaptureStatusNone;
  if (have_jobs && keep && plan.tables_built) {
    // kept tables: an execute on another stream than the one that built them orders itself behind the
    // build (inside a capture the event lies outside the graph, so the host waits for it instead)
    if (plan.ev_built && stream != plan.built_stream) {
      e = capturing ? cudaEventSynchronize(plan.ev_built) : cudaStreamWaitEvent(stream, plan.ev_built, 0);
      if (e != cudaSuccess) return cuda_fail(e, "wait(table build)");
    }
  } else if (have_jobs) {
    if (!plan.jobs.empty()) {
      LaunchTimer tm(BHW_KERNEL_TABLE_BUILD, stream);
      tm_on = tm.on;
      e = launch_table_build((const TabJob*)(plan.blob_dev + plan.o_jobs), (int)plan.jobs.size(),
                             plan.table_work, plan.rom, stream);
      if (e != cudaSuccess) return cuda_fail(e, "k_table_build");
      g_launches++;
    }
    // A family's pyramid is read by that family's groups only: when the execute fans its launches out over
    // side streams, such a build goes to a side stream of its own together with its groups (below), so that
    // an integer-bound build overlaps the store-bound synthesis of the other families.
    defer_family_builds = !capturing && g_side_streams.load() > 0 && g_defer_ctas.load() > 0 && plan.groups.size() >= 3;
    for (size_t ji = 0; ji < plan.big_jobs.size(); ji++) {
      bool deferred = false;
      if (defer_family_builds)
        for (const auto& fam : plan.families) if (fam.big_job == (int)ji) deferred = true;
      if (deferred) continue;
      LaunchTimer tm(BHW_KERNEL_TABLE_BUILD, stream);
      tm_on = tm.on;
      e = launch_table_build_unrolled(plan.big_jobs[ji], stream);
      if (e != cudaSuccess) return cuda_fail(e, "k_table_build_u");
      g_launches++;
    }
    // tables some bank run pairs through the ones'-complement relation: list the entries that break it
    for (const auto& pt : plan.tables) {
      if (!pt.exc) continue;
      LaunchTimer tm(BHW_KERNEL_TABLE_BUILD, stream);
      tm_on = true;                                            // no longer directly behind the build: no PDL
      e = launch_inq_exceptions(pt.ptr, pt.entries, (int32_t)(1u << table_tshift(pt.canon)), pt.exc, stream);
      if (e != cudaSuccess) return cuda_fail(e, "k_inq_exceptions");
      g_launches++;
    }
    building = true;
    table_ahead = !tm_on && !defer_family_builds;
  }
  SynthArgs a;
  memset(&a, 0, sizeof(a));
  a.recs = (const WinRec*)(plan.blob_dev + plan.o_recs);
  a.win_rec = !plan.all_same ? (const uint32_t*)(plan.blob_dev + plan.o_wr) : nullptr;
  a.flat_off = plan.uniform_pw < 0 ? (const uint64_t*)(plan.blob_dev + plan.o_off) : nullptr;
  a.gens = (const GenRec*)(plan.blob_dev + plan.o_gens);
  a.rom = plan.rom;
  a.nwin = plan.nwin;
  a.uniform_pw = plan.uniform_pw;
  a.pack16 = plan.pack16 ? 1u : 0u;
  // whole windows of bank runs go to the bank kernel, everything in between to the general one
  const uint64_t flat_end = flat_begin + flat_count;
  size_t runs_hit = 0;
  for (const bhw_plan::BankRun& run : plan.runs) {
    const uint64_t rb = run.flat_off, re = rb + ((uint64_t)(run.w_end - run.w_begin) << run.sh.pw);
    if (re > flat_begin && rb < flat_end) runs_hit++;
  }
  // Groups (bhw_group.cuh): the member windows that lie wholly inside the range form one contiguous piece of
  // their group's launch list -> one paired launch per group; a member window the range cuts contributes its
  // whole 256-sample tiles to an unpaired launch (inline list); `covered` collects what these launches write.
  using Covered = bhw_plan::Covered;
  using GroupWork = bhw_plan::GroupWork;
  const uint32_t kSpreadG_ = 30;
  std::vector<Covered> covered;
  std::vector<GroupWork> gwork(plan.groups.size());
  size_t group_launches = 0;
  const bool whole = flat_begin == 0 && flat_count == plan.total;
  auto cover = [&](uint64_t b, uint64_t e_) {
    if (!covered.empty() && covered.back().e == b) covered.back().e = e_;
    else covered.push_back({b, e_});
  };
  if (whole && plan.whole_cached) {
    gwork = plan.whole_gwork;
    covered = plan.whole_covered;
    group_launches = plan.whole_group_launches;
  } else if (!plan.groups.empty()) {
    size_t w = (size_t)(std::upper_bound(plan.flat_off.begin(), plan.flat_off.end(), flat_begin) - plan.flat_off.begin());
    w = w ? w - 1 : 0;
    for (; w < (size_t)plan.nwin && plan.flat_off[w] < flat_end; w++) {
      const int g = plan.win_group[w];
      if (g < 0) continue;
      GroupWork& gk = gwork[(size_t)g];
      const uint64_t wb = plan.flat_off[w], we = plan.flat_off[w + 1];
      const uint64_t lo = wb > flat_begin ? wb : flat_begin, hi = we < flat_end ? we : flat_end;
      if (lo >= hi) continue;
      if (lo == wb && hi == we) {
        const uint32_t gi = plan.win_gidx[w];
        if (gi & 0x80000000u) { gk.singles.push_back(gi & 0x7FFFFFFFu); group_launches++; }
        else {
          if (gk.i1 == gk.i0) { gk.i0 = gi; group_launches++; }
          gk.i1 = gi + 1;
        }
        cover(wb, we);
      } else {
        const uint64_t ta = (lo - wb + kBankTile - 1) / kBankTile, tb = (hi - wb) / kBankTile;
        if (tb <= ta || gk.nparts >= 2) continue;
        GroupWin& pw_ = gk.part[gk.nparts];
        memset(&pw_, 0, sizeof(pw_));
        pw_.pw = plan.recs[plan.win_rec[w]].pw;
        pw_.rec = plan.win_rec[w];
        pw_.tile_first = (uint32_t)ta;
        pw_.out_off = (int64_t)wb;
        pw_.unit_begin = (uint32_t)(tb - ta);                   // tile count for now; prefix-summed at launch
        if (!gk.nparts++) group_launches++;
        cover(wb + ta * kBankTile, wb + tb * kBankTile);
      }
    }
    if (whole) {
      plan.whole_gwork = gwork;
      plan.whole_covered = covered;
      plan.whole_group_launches = group_launches;
      plan.whole_cached = true;
    }
  }
  LaunchFan fan;
  fan.begin(plan.dev, stream, runs_hit + group_launches);
  if (fan.err != cudaSuccess) return cuda_fail(fan.err, "side streams");
  // deferred family builds: the largest first, each on a side stream of its own (or on `stream` when none is free)
  std::vector<cudaStream_t> fam_stream(plan.families.size(), nullptr);
  bool any_deferred = false;
  if (building && defer_family_builds) {
    std::vector<size_t> order;
    for (size_t f = 0; f < plan.families.size(); f++) if (plan.families[f].big_job >= 0) order.push_back(f);
    std::sort(order.begin(), order.end(), [&](size_t x, size_t y) {
      return plan.big_jobs[(size_t)plan.families[x].big_job].work > plan.big_jobs[(size_t)plan.families[y].big_job].work; });
    for (size_t f : order) {
      cudaStream_t fs = fan.reserve();
      LaunchTimer tm(BHW_KERNEL_TABLE_BUILD, fs);
      e = launch_table_build_unrolled(plan.big_jobs[(size_t)plan.families[f].big_job], fs, fs == stream ? 8 : g_defer_ctas.load());
      if (e != cudaSuccess) return cuda_fail(e, "k_table_build_u");
      g_launches++;
      fam_stream[f] = fs;
      any_deferred = true;
    }
  }
  // flat sub-ranges for the general kernel are collected and launched together (several disjoint ranges per
  // launch): the win_selector sweep leaves ten little gaps of windows shorter than a tile pair
  std::vector<Covered> pieces;
  auto general_raw = [&](uint64_t b, uint64_t e_) -> int {
    if (b < e_) pieces.push_back({b, e_});
    return BHW_OK;
  };
  auto flush_pieces = [&]() -> int {
    for (size_t i = 0; i < pieces.size();) {
      const size_t n = pieces.size() - i < (size_t)kSynthMaxPieces ? pieces.size() - i : (size_t)kSynthMaxPieces;
      uint64_t bytes = 0;
      if (n == 1) {
        a.npieces = 0;
        a.out = (char*)out_dev + (pieces[i].b - flat_begin) * (plan.pack16 ? 2 : 4);
        a.flat_begin = pieces[i].b;
        a.flat_count = pieces[i].e - pieces[i].b;
        bytes = a.flat_count * (plan.pack16 ? 2 : 4);
      } else {
        a.npieces = (uint32_t)n;
        a.out = out_dev;
        a.out_flat0 = flat_begin;
        a.flat_begin = pieces[i].b;
        a.flat_count = 0;
        uint32_t tiles = 0;
        for (size_t j = 0; j < n; j++) {
          a.piece_begin[j] = pieces[i + j].b;
          a.piece_end[j] = pieces[i + j].e;
          a.piece_tile0[j] = tiles;
          tiles += (uint32_t)((pieces[i + j].e - pieces[i + j].b + 127) / 128);
          bytes += (pieces[i + j].e - pieces[i + j].b) * (plan.pack16 ? 2 : 4);
        }
        a.piece_tile0[n] = tiles;
      }
      cudaStream_t ls = fan.next();
      table_ahead = false;
      cudaError_t ce;
      {
        LaunchTimer tm(BHW_KERNEL_SYNTH, ls, (uint32_t)n, bytes);
        ce = launch_synth(a, ls);
      }
      if (ce != cudaSuccess) return cuda_fail(ce, "k_synth");
      g_launches++;
      i += n;
    }
    pieces.clear();
    return BHW_OK;
  };
  // ... minus what the group launches cover
  auto general = [&](uint64_t b, uint64_t e_) -> int {
    if (b >= e_) return BHW_OK;
    auto it = std::lower_bound(covered.begin(), covered.end(), b, [](const Covered& c, uint64_t v) { return c.e <= v; });
    uint64_t cur = b;
    for (; it != covered.end() && it->b < e_; ++it) {
      if (it->b > cur) { int st_ = general_raw(cur, it->b); if (st_) return st_; }
      if (it->e > cur) cur = it->e;
    }
    return cur < e_ ? general_raw(cur, e_) : BHW_OK;
  };
  uint64_t cursor = flat_begin;
  // one bank launch: whole windows [wa, wb) of the run, or (ntiles > 0) tiles [tile_off, +ntiles) of window wa
  auto bank = [&](const bhw_plan::BankRun& run, uint64_t out_flat, uint32_t wa, uint32_t wb, uint32_t tile_off,
                  uint32_t ntiles) -> int {
    BankArgs ba;
    memset(&ba, 0, sizeof(ba));
    ba.sh = ntiles ? run.sh_part : run.sh;
    ba.recs = a.recs;
    ba.win_rec = a.win_rec;
    ba.out = (int32_t*)((char*)out_dev + (out_flat - flat_begin) * (plan.pack16 ? 2 : 4));
    ba.pack16 = plan.pack16 ? 1u : 0u;
    ba.w_first = (uint32_t)run.w_begin + wa;
    ba.nwin = wb - wa;
    ba.tile_off = tile_off;
    ba.ntiles = ntiles;
    {
      // window-minor walk (BankArgs::win_minor): pays where the stride-k gathers dominate - no phase
      // bit dropped and 5 or more terms (measured on 256 MB banks: 7-term DAT_WIDTH 32 N=2^20 123 -> 99 us,
      // cordic_dds48 239 -> 145 us, 5-term DAT_WIDTH 24 79 -> 71 us; 2-term 48 -> 54 us, so not there)
      const uint32_t log_tpw = run.sh.pw - kBankTileLog2 - (run.pair ? 1 : 0);
      ba.win_minor = (!ntiles && run.tab_mode == TAB_GLOBAL && ba.nwin > 1 && run.sh.lin && run.sh.m >= 5 &&
                      !(((uint64_t)ba.nwin << log_tpw) >> 32)) ? 1u : 0u;
      // one long 7-term window: spread the warps of a CTA over the window (BankArgs::spread); 5 and
      // fewer terms measured no different (they are not bound by the L2 -> L1 gather traffic)
      const uint32_t G = 30u;
      ba.spread = (!ntiles && run.tab_mode == TAB_GLOBAL && ba.nwin == 1 && run.sh.lin && run.sh.m >= 7 &&
                   log_tpw < 32 && (1ull << log_tpw) >= (uint64_t)G * (uint64_t)device_sm_count()) ? G : 0u;
    }
    cudaStream_t ls = fan.next();
    cudaError_t ce;
    {
      const uint64_t bank_bytes = (ntiles ? (uint64_t)ntiles * kBankTile : ((uint64_t)ba.nwin << run.sh.pw)) * (plan.pack16 ? 2 : 4);
      LaunchTimer tm(BHW_KERNEL_SYNTH_BANK, ls, run.sh.m | ((uint32_t)(ntiles ? TAB_GLOBAL : run.tab_mode) << 8) |
                     ((uint32_t)(ntiles ? 0 : run.pair) << 16) | (run.sh.pw << 24), bank_bytes);
      const bool pdl = table_ahead && !fan.nside && !tm.on && ls == stream;
      ce = ntiles ? launch_synth_bank(ba, TAB_GLOBAL, false, ls, pdl) : launch_synth_bank(ba, run.tab_mode, run.pair, ls, pdl);
      table_ahead = false;
    }
    if (ce != cudaSuccess) return cuda_fail(ce, "k_synth_bank");
    g_launches++;
    if (!ntiles && run.exc_table >= 0) {                       // recompute the sample pairs that read an exception entry
      LaunchTimer tm(BHW_KERNEL_SYNTH, ls, 0, 0);
      ce = launch_inq_patch(ba, plan.tables[(size_t)run.exc_table].exc, ls);
      if (ce != cudaSuccess) return cuda_fail(ce, "k_inq_patch");
      g_launches++;
    }
    return BHW_OK;
  };
  // samples [pa, pb) inside window widx of the run: whole tiles through the bank kernel (unpaired
  // shape), the ragged ends through the general kernel
  auto partial = [&](const bhw_plan::BankRun& run, uint32_t widx, uint64_t pa, uint64_t pb) -> int {
    const uint64_t ws = run.flat_off + ((uint64_t)widx << run.sh.pw);
    const uint64_t ta = (pa - ws + kBankTile - 1) / kBankTile, tb = (pb - ws) / kBankTile;
    if (!run.part_ok || tb < ta + 32) return general(pa, pb);       // short: not worth a launch of its own
    int st = general(pa, ws + ta * kBankTile);
    if (!st) st = bank(run, ws + ta * kBankTile, widx, widx + 1, (uint32_t)ta, (uint32_t)(tb - ta));
    if (!st) st = general(ws + tb * kBankTile, pb);
    return st;
  };
  for (const bhw_plan::BankRun& run : plan.runs) {
    const uint32_t pw = run.sh.pw;
    const uint64_t N = 1ull << pw;
    const uint64_t rb = run.flat_off, re = rb + ((uint64_t)(run.w_end - run.w_begin) << pw);
    if (re <= cursor) continue;
    if (rb >= flat_end) break;
    uint64_t lo = rb > cursor ? rb : cursor;
    const uint64_t hi = re < flat_end ? re : flat_end;
    int st = general(cursor, lo);                                     // windows between the runs
    if (st) return st;
    if ((lo - rb) & (N - 1)) {                                        // the range enters the run inside a window
      const uint32_t widx = (uint32_t)((lo - rb) >> pw);
      const uint64_t wend = rb + ((uint64_t)(widx + 1) << pw);
      const uint64_t pe = wend < hi ? wend : hi;
      if ((st = partial(run, widx, lo, pe))) return st;
      lo = pe;
    }
    const uint64_t wlo = (lo - rb) >> pw, whi = (hi - rb) >> pw;      // whole windows [wlo, whi) of the run
    if (whi > wlo) {
      if ((st = bank(run, rb + (wlo << pw), (uint32_t)wlo, (uint32_t)whi, 0, 0))) return st;
      lo = rb + (whi << pw);
    }
    if (lo < hi && (st = partial(run, (uint32_t)whi, lo, hi))) return st;   // ... and leaves it inside one
    cursor = hi;
  }
  int st = general(cursor, flat_end);
  // with family builds on side streams, pieces that read a pyramid (the ragged ends of cut group windows) run
  // after the join; the others now
  std::vector<Covered> late;
  if (!st && any_deferred) {
    std::vector<Covered> now;
    for (const Covered& pc : pieces) {
      bool reads_pyramid = false;
      size_t w = (size_t)(std::upper_bound(plan.flat_off.begin(), plan.flat_off.end(), pc.b) - plan.flat_off.begin());
      for (w = w ? w - 1 : 0; w < (size_t)plan.nwin && plan.flat_off[w] < pc.e; w++)
        if (plan.win_group[w] >= 0 && fam_stream[(size_t)plan.groups[(size_t)plan.win_group[w]].family]) reads_pyramid = true;
      (reads_pyramid ? late : now).push_back(pc);
    }
    pieces.swap(now);
  }
  if (!st) st = flush_pieces();
  if (st) return st;
  // the group launches (after the small general pieces: these kernels fill the GPU)
  for (size_t g = 0; g < plan.groups.size(); g++) {
    const bhw_plan::Group& gr = plan.groups[g];
    GroupWork& gk = gwork[g];
    const int tabm = plan.families[(size_t)gr.family].tab_mode;
    const int npass = 2 + (int)gk.singles.size();
    for (int pass = 0; pass < npass; pass++) {
      GroupArgs ga;
      memset(&ga, 0, sizeof(ga));
      ga.sh = gr.sh;
      ga.recs = a.recs;
      ga.out = plan.pack16 ? (int32_t*)((int16_t*)out_dev - flat_begin)   // GroupWin::out_off is a flat index of the batch
                           : (int32_t*)out_dev - flat_begin;
      ga.pack16 = plan.pack16 ? 1u : 0u;
      if (pass == 0) {
        if (gk.i1 == gk.i0) continue;
        ga.wins = (const GroupWin*)(plan.blob_dev + gr.o_list) + gk.i0;
        ga.nwin = gk.i1 - gk.i0;
        ga.unit_base = gr.list[gk.i0].unit_begin;
        ga.nunits = gr.list[gk.i1].unit_begin - ga.unit_base;
        // the interleaved walk jumps gridDim * 8 units per step: over a list of short windows that is a window search
        // per unit (65,536 windows of 1024 samples: 0.11 ms instead of 0.05) - such lists are walked contiguously
        if (ga.nunits / ga.nwin < 64u) ga.sh.interleave = 0;
      } else if (pass >= 2) {
        const uint32_t si = gk.singles[(size_t)pass - 2];
        ga.wins = (const GroupWin*)(plan.blob_dev + gr.o_singles) + 2 * si;
        ga.nwin = 1;
        ga.nunits = gr.singles[2 * (size_t)si + 1].unit_begin;
        ga.spread = kSpreadG_;
        ga.prefetch_lines = (uint32_t)g_prefetch_lines.load();
      } else {
        if (!gk.nparts) continue;
        uint32_t units = 0;
        for (uint32_t i = 0; i < gk.nparts; i++) {
          ga.iw[i] = gk.part[i];
          const uint32_t cnt = gk.part[i].unit_begin;
          ga.iw[i].unit_begin = units;
          units += cnt;
        }
        ga.iw[gk.nparts].unit_begin = units;
        ga.wins = nullptr;
        ga.nwin = gk.nparts;
        ga.nunits = units;
        ga.sh.interleave = 0;
      }
      if (ap) {
        ga.x = ap->x; ga.y = ap->y; ga.frames = ap->frames;
        ga.apply_mode = (uint32_t)ap->mode + 1u; ga.apply_dw = (uint32_t)ap->dw;
      }
      cudaStream_t ls = fam_stream[(size_t)gr.family] ? fam_stream[(size_t)gr.family] : fan.next();
      cudaError_t ce;
      {
        LaunchTimer tm(BHW_KERNEL_SYNTH_GROUP, ls, gr.sh.m | ((uint32_t)tabm << 8) | ((uint32_t)(pass != 1) << 16) |
                       ((uint32_t)(ga.spread ? 1 : 0) << 17) | (gr.sh.top << 24),
                       (uint64_t)ga.nunits * kBankTile * (pass != 1 ? 8 : 4));
        const bool pdl = table_ahead && !fan.nside && !tm.on && ls == stream;
        ce = launch_synth_group(ga, tabm, pass != 1, ls, pdl);
        table_ahead = false;
      }
      if (ce != cudaSuccess) return cuda_fail(ce, "k_synth_group");
      g_launches++;
    }
  }
  e = fan.join();
  if (e != cudaSuccess) return cuda_fail(e, "side streams (join)");
  if (!late.empty()) {
    pieces.swap(late);
    if ((st = flush_pieces())) return st;
  }
  // A build that is only captured has not happened: the flag stays clear, so an eager execute (or the next
  // capture) builds again; replays of the graph rebuild the tables every time.
  if (building && keep && !capturing) {
    if (!plan.ev_built && (e = cudaEventCreateWithFlags(&plan.ev_built, cudaEventDisableTiming)) != cudaSuccess)
      return cuda_fail(e, "cudaEventCreate(table build)");
    if ((e = cudaEventRecord(plan.ev_built, stream)) != cudaSuccess) return cuda_fail(e, "cudaEventRecord(table build)");
    plan.built_stream = stream;
    plan.tables_built = true;
  }
  return BHW_OK;
}

// windows touched by a flat range
static int shard_windows(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count,
                         int* first_win, int* nwin_touched, uint64_t* local_begin) {
  uint64_t off = 0;
  int first = -1, last = -1;
  uint64_t first_off = 0;
  const uint64_t end = flat_begin + flat_count;
  for (int w = 0; w < nwin; w++) {
    if (descs[w].phi_width < BHW_MIN_PHI_WIDTH || descs[w].phi_width > BHW_MAX_PHI_WIDTH) return BHW_E_PHI_WIDTH;
    const uint64_t N = 1ull << descs[w].phi_width;
    if (flat_count && off < end && off + N > flat_begin) {
      if (first < 0) { first = w; first_off = off; }
      last = w;
    }
    off += N;
  }
  if (flat_begin > off || flat_count > off - flat_begin) return BHW_E_RANGE;
  if (first < 0) { *first_win = 0; *nwin_touched = 0; *local_begin = 0; return BHW_OK; }
  *first_win = first;
  *nwin_touched = last - first + 1;
  *local_begin = flat_begin - first_off;
  return BHW_OK;
}

// One-shot: plan the touched windows from the stream's memory pool, execute, release.
static int run_batch(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count, void* out_dev,
                     cudaStream_t stream) {
  if (!descs) return BHW_E_NULL;
  if (nwin <= 0) return BHW_E_ARG;
  if (!out_dev && flat_count) return BHW_E_NULL;
  // validate every descriptor of the batch, touched or not: errors do not depend on the range
  for (int w = 0; w < nwin; w++) {
    int st = validate_desc(&descs[w], true);
    if (st) return st;
  }
  size_t esz = 4;
  int st = batch_elem_bytes(descs, nwin, &esz);
  if (st) return st;
  int first = 0, touched = 0;
  uint64_t local = 0;
  st = shard_windows(descs, nwin, flat_begin, flat_count, &first, &touched, &local);
  if (st) return st;
  if (!touched) return BHW_OK;
  // a one-shot call uploads its records from a stack-lifetime buffer: it cannot be recorded into a
  // CUDA graph (capture a bhw_plan_execute instead)
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) return BHW_E_CAPTURE;
  bhw_plan plan;
  plan.transient = true;
  st = plan_build(plan, descs + first, touched, local, flat_count, stream);
  if (!st) st = plan_execute(plan, local, flat_count, out_dev, stream);
  plan_free_device(plan, stream);
  return st;
}

// ---- host-buffer pipeline ----------------------------------------------------------------------
static const uint64_t kHostChunkBytes = 32ull << 20;

static cudaError_t pipe_ensure(HostPipe& p) {
  cudaError_t e = cudaSuccess;
  if (!p.s_gen && (e = cudaStreamCreateWithFlags(&p.s_gen, cudaStreamNonBlocking)) != cudaSuccess) return e;
  if (!p.s_copy && (e = cudaStreamCreateWithFlags(&p.s_copy, cudaStreamNonBlocking)) != cudaSuccess) return e;
  for (int i = 0; i < 2; i++) {
    if (!p.ev_gen[i] && (e = cudaEventCreateWithFlags(&p.ev_gen[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    if (!p.ev_copy[i] && (e = cudaEventCreateWithFlags(&p.ev_copy[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    if (!p.buf[i] && (e = cudaMalloc((void**)&p.buf[i], kHostChunkBytes)) != cudaSuccess) return e;
  }
  p.buf_bytes = kHostChunkBytes;
  return cudaSuccess;
}

static void pipe_release(HostPipe& p) {
  for (int i = 0; i < 2; i++) {
    if (p.buf[i]) cudaFree(p.buf[i]);
    if (p.ev_gen[i]) cudaEventDestroy(p.ev_gen[i]);
    if (p.ev_copy[i]) cudaEventDestroy(p.ev_copy[i]);
  }
  if (p.s_gen) cudaStreamDestroy(p.s_gen);
  if (p.s_copy) cudaStreamDestroy(p.s_copy);
  p = HostPipe();
}

// Generates chunk c on the compute stream while chunk c-1 drains to the host on the copy stream.
// The request is cut at window boundaries into segments (a short first one, then >= 256 MB each);
// every segment gets its own one-shot plan, so the host plans segment s+1 while the chunks of
// segment s are still on their way over the link, and the first byte leaves after planning only
// a handful of windows.  The device staging buffers and streams are kept per device.
static const uint64_t kHostFirstSegmentBytes = 16ull << 20;
static const uint64_t kHostSegmentBytes = 256ull << 20;

static int run_batch_host(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count,
                          void* out_host) {
  if (!descs) return BHW_E_NULL;
  if (nwin <= 0) return BHW_E_ARG;
  if (!out_host && flat_count) return BHW_E_NULL;
  for (int w = 0; w < nwin; w++) {
    int st = validate_desc(&descs[w], true);
    if (st) return st;
  }
  size_t esz = 4;
  int st = batch_elem_bytes(descs, nwin, &esz);
  if (st) return st;
  int first = 0, touched = 0;
  uint64_t local = 0;
  st = shard_windows(descs, nwin, flat_begin, flat_count, &first, &touched, &local);
  if (st) return st;
  if (!touched) return BHW_OK;
  int dev;
  if ((st = current_device(&dev))) return st;
  DeviceState& ds = g_dev[dev];
  std::lock_guard<std::mutex> pipe_lock(ds.pipe_mu);
  HostPipe& p = ds.pipe;
  cudaError_t e = pipe_ensure(p);
  if (e != cudaSuccess) { pipe_release(p); return cuda_fail(e, "host pipeline setup"); }
  const uint64_t chunk = p.buf_bytes / esz;
  uint64_t done = 0;  // samples of the request handed to the pipeline so far
  int c = 0;          // chunks so far (selects the staging buffer)
  int w = first;      // first window of the next segment
  uint64_t seg_local = local;  // where the request enters window w
  while (!st && e == cudaSuccess && done < flat_count) {
    // windows [w, we) of this segment and the samples of the request that fall into them
    const uint64_t target = (done == 0 ? kHostFirstSegmentBytes : kHostSegmentBytes) / esz;
    uint64_t seg_count = 0;
    int we = w;
    while (we < first + touched && seg_count < target) {
      uint64_t n = (1ull << descs[we].phi_width) - (we == w ? seg_local : 0);
      if (n > flat_count - done - seg_count) n = flat_count - done - seg_count;
      seg_count += n;
      we++;
    }
    bhw_plan plan;
    plan.transient = true;
    st = plan_build(plan, descs + w, we - w, seg_local, seg_count, p.s_gen);
    for (uint64_t seg_done = 0; !st && seg_done < seg_count; c++) {
      const int b = c & 1;
      const uint64_t cnt = seg_count - seg_done < chunk ? seg_count - seg_done : chunk;
      if (c >= 2 && (e = cudaStreamWaitEvent(p.s_gen, p.ev_copy[b], 0)) != cudaSuccess) break;  // buffer drained?
      st = plan_execute(plan, seg_local + seg_done, cnt, p.buf[b], p.s_gen);
      if (st) break;
      if ((e = cudaEventRecord(p.ev_gen[b], p.s_gen)) != cudaSuccess) break;
      if ((e = cudaStreamWaitEvent(p.s_copy, p.ev_gen[b], 0)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync((char*)out_host + (done + seg_done) * esz, p.buf[b], cnt * esz,
                               cudaMemcpyDeviceToHost, p.s_copy)) != cudaSuccess) break;
      if ((e = cudaEventRecord(p.ev_copy[b], p.s_copy)) != cudaSuccess) break;
      seg_done += cnt;
    }
    plan_free_device(plan, p.s_gen);  // stream-ordered: after the segment's last kernel
    done += seg_count;
    w = we;
    seg_local = 0;
  }
  cudaError_t e2 = cudaStreamSynchronize(p.s_gen);
  if (e == cudaSuccess) e = e2;
  e2 = cudaStreamSynchronize(p.s_copy);
  if (e == cudaSuccess) e = e2;
  if (st) return st;
  if (e != cudaSuccess) return cuda_fail(e, "host pipeline");
  return BHW_OK;
}

// One window through the direct kernel (BHW_ALGO_DIRECT, and always for DAT_WIDTH > 32).
// `only_fast32`: return 1 without launching when the window is not eligible for the 32-bit kernel.
static int run_direct(const bhw_desc* d, uint64_t n0, uint64_t count, void* out_dev, cudaStream_t stream,
                      bool only_fast32 = false) {
  int dev;
  int st = current_device(&dev);
  if (st) return st;
  DirectArgs a;
  memset(&a, 0, sizeof(a));
  if ((st = resolve_window(d, &a.wp, a.src))) return st;
  for (int u = 0; u < a.wp.nsrc; u++) init_src_core(a.src[u], &a.sc[u]);
  const uint64_t N = 1ull << d->phi_width;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
  if (a.src[0].kind == SRC_TAYLOR) {
    if ((st = get_rom(dev, a.src[0].dw, a.src[0].lut, &a.rom))) return st;
    const uint32_t entries = 1u << a.src[0].lut;
    a.rom_smem_entries = entries * sizeof(I2) <= 32768 ? entries : 0;  // the sine LUT lives in shared memory
  }
  a.n_first = n0 + (uint64_t)d->stream_offset;
  a.count = count;
  if (!count) return BHW_OK;
  cudaError_t e;
  Direct32Args a32;
  DirectTayArgs at;
  const bool fast32 = !a.wp.elem64 && direct32_params(a.wp, a.src, &a32.p);
  const bool tay32 = !fast32 && !a.wp.elem64 && direct_taylor_params(a.wp, a.src, &at.p);
  if (only_fast32 && !fast32 && !tay32) return 1;
  if (tay32) {
    // TAYLOR in 32-bit registers, sine ROM in shared memory, 128-bit stores
    at.rom = a.rom;
    at.n0 = n0;
    at.count = count;
    // whole window: evaluate the units once per sample pair (n, n + N/2)
    at.pair = (n0 == 0 && count == N && N >= 8 && at.p.unit[0].pw == d->phi_width &&
               (at.p.m == 2 || at.p.unit[1].pw == d->phi_width - 1)) ? 1u : 0u;
    if (at.pair && direct_taylor_quad_ok(at.p, d->phi_width)) at.pair = 2u;
    LaunchTimer tm(BHW_KERNEL_DIRECT, stream);
    e = launch_direct_taylor(at, (int32_t*)out_dev, stream);
  } else if (fast32) {
    // register-resident 32-bit stages, 128-bit stores
    a32.n0 = n0;
    a32.count = count;
    // whole window: sample pairs, or all four quarter-window partners, from one set of evaluations
    a32.pair = (n0 == 0 && count == N && N >= 8 && a32.p.pw == d->phi_width) ? (N >= 32 ? 2u : 1u) : 0u;
    LaunchTimer tm(BHW_KERNEL_DIRECT, stream);
    e = launch_direct32(a32, (int32_t*)out_dev, stream);
  } else {
    uint32_t flip = 0, adv = 0;
    a.pair_flip = (n0 == 0 && count == N && N >= 8 && direct_pair_flip(a.wp, a.src, &flip)) ? (0x80000000u | flip) : 0u;
    // from about one CTA of pairs per SM up: all four quarter-window partners from one evaluation
    a.quad_adv = (a.pair_flip && N / 2 > (uint64_t)device_sm_count() * 256u && direct_quad_adv(a.wp, a.src, &adv))
                     ? (0x80000000u | adv) : 0u;
    LaunchTimer tm(BHW_KERNEL_DIRECT, stream);
    e = launch_direct_window(a, out_dev, stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "k_direct_window");
  g_launches++;
  return BHW_OK;
}

// BHW_ALGO_AUTO, one-shot request: does one launch of a register-resident direct kernel beat the
// table path (table build + synthesis, two launches and a plan: ~11 us per call however short)?
// A TAYLOR window of any length: a Taylor evaluation (ROM look-up + two multiplies) per sample is
// cheaper than building, storing and re-reading tables as large as the window.  CORDIC: the direct
// kernel's work is count x (M-1) evaluations of ~DAT_WIDTH stages, quartered for a whole window (four
// samples per evaluation); measured crossover (tools/call_latency --route, us per call direct /
// table): 4-term DW 17 N=2^20 8.2 / 11.0, 2^21 14.4 / 11.3; 2-term DW 16 2^22 10.3 / 12.9; 3-term
// DW 16 2^21 10.3 / 11.3, 2^22 16.4 / 13.4.
static bool auto_prefers_direct(const bhw_desc* d, uint64_t n0, uint64_t count) {
  if (d->model == BHW_MODEL_RTL && d->sin_type == BHW_SIN_TAYLOR) return true;
  const bool whole = n0 == 0 && count == (1ull << d->phi_width);
  const uint64_t work = count * (uint64_t)(d->win_type - 1) * (uint64_t)d->dat_width * (whole ? 1u : 4u);
  return work <= 80000000ull;
}

}  // namespace bhw

using namespace bhw;

extern "C" {

int bhw_generate(const bhw_desc* d, void* out_dev, uint64_t n0, uint64_t count, void* stream) {
  if (!d) return BHW_E_NULL;
  int st = validate_desc(d, true);
  if (st) return st;
  const uint64_t N = 1ull << d->phi_width;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
  if (!out_dev && count) return BHW_E_NULL;
  if (d->out_format == BHW_OUT_INT16) return run_batch(d, 1, n0, count, out_dev, (cudaStream_t)stream);  // plan kernels only
  if (d->dat_width > 32 || d->algo == BHW_ALGO_DIRECT) return run_direct(d, n0, count, out_dev, (cudaStream_t)stream);
  if (d->algo == BHW_ALGO_AUTO && auto_prefers_direct(d, n0, count)) {
    st = run_direct(d, n0, count, out_dev, (cudaStream_t)stream, true);
    if (st != 1) return st;
  }
  return run_batch(d, 1, n0, count, out_dev, (cudaStream_t)stream);
}

int bhw_generate_repeat(const bhw_desc* d, void* out_dev, uint64_t n0, uint64_t count, int reps, uint64_t out_stride,
                        uint64_t out_slots, void* stream) {
  if (reps < 0) return BHW_E_ARG;
  const size_t esz = (size_t)bhw_elem_bytes(d);
  for (int i = 0; i < reps; i++) {
    char* o = (char*)out_dev + (out_slots ? ((uint64_t)i % out_slots) * out_stride * esz : 0);
    int st = bhw_generate(d, o, n0, count, stream);
    if (st) return st;
  }
  return BHW_OK;
}

int bhw_apply(const bhw_desc* d, int mode, const int32_t* x_dev, void* y_dev, uint64_t frames, void* stream_) {
  if (!d) return BHW_E_NULL;
  int st = validate_desc(d, true);
  if (st) return st;
  if (d->dat_width > 32) return BHW_E_DAT_WIDTH;      // the multiplier ports are DAT_WIDTH bits, held in int32 here
  if (d->out_format != BHW_OUT_DEFAULT) return BHW_E_ARG;   // the products have their own containers
  if (mode != BHW_APPLY_EXACT && mode != BHW_APPLY_ROUNDED) return BHW_E_ARG;
  if (!frames) return BHW_OK;
  if (!x_dev || !y_dev) return BHW_E_NULL;
  cudaStream_t stream = (cudaStream_t)stream_;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &cap) == cudaSuccess && cap != cudaStreamCaptureStatusNone) return BHW_E_CAPTURE;
  const uint64_t N = 1ull << d->phi_width;
  bhw_plan plan;
  plan.transient = true;
  st = plan_build(plan, d, 1, 0, N, stream);
  if (!st) {
    if (plan.groups.size() == 1 && plan.win_group.size() == 1 && plan.win_group[0] == 0) {
      // the window never reaches memory: k_synth_group multiplies it into the frames as it is produced
      ApplyInfo ap{x_dev, y_dev, frames, mode, d->dat_width};
      st = plan_execute(plan, 0, N, y_dev, stream, &ap);
    } else {
      // windows the fused kernel does not take: generate into scratch memory, then multiply
      int32_t* w = nullptr;
      cudaError_t e = plan_alloc(plan, (void**)&w, N * sizeof(int32_t), stream);
      if (e != cudaSuccess) st = cuda_fail(e, "alloc(apply scratch)");
      if (!st) st = plan_execute(plan, 0, N, w, stream);
      if (!st) {
        LaunchTimer tm(BHW_KERNEL_APPLY, stream, 0, N * frames * (mode == BHW_APPLY_EXACT ? 8 : 4));
        e = launch_apply_mul(x_dev, w, y_dev, N, frames, mode + 1, d->dat_width, stream);
        if (e != cudaSuccess) st = cuda_fail(e, "k_apply_mul");
        else g_launches++;
      }
      if (w) cudaFreeAsync(w, stream);
    }
  }
  plan_free_device(plan, stream);
  return st;
}

int bhw_generate_host(const bhw_desc* d, void* out_host, uint64_t n0, uint64_t count) {
  if (!d) return BHW_E_NULL;
  int st = validate_desc(d, true);
  if (st) return st;
  const uint64_t N = 1ull << d->phi_width;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
  if (!out_host && count) return BHW_E_NULL;
  // a short request (one staging chunk) that bhw_generate would send to a register-resident
  // direct kernel: one launch + one copy instead of planning, table build and two launches
  if (count && d->dat_width <= 32 && d->out_format == BHW_OUT_DEFAULT && d->algo == BHW_ALGO_AUTO &&
      count * 4 <= kHostChunkBytes && auto_prefers_direct(d, n0, count)) {
    int dev;
    if ((st = current_device(&dev))) return st;
    DeviceState& ds = g_dev[dev];
    std::lock_guard<std::mutex> pipe_lock(ds.pipe_mu);
    HostPipe& p = ds.pipe;
    cudaError_t e = pipe_ensure(p);
    if (e != cudaSuccess) { pipe_release(p); return cuda_fail(e, "host pipeline setup"); }
    st = run_direct(d, n0, count, p.buf[0], p.s_gen, true);
    if (st == BHW_OK) {
      e = cudaMemcpyAsync(out_host, p.buf[0], count * 4, cudaMemcpyDeviceToHost, p.s_gen);
      if (e == cudaSuccess) e = cudaStreamSynchronize(p.s_gen);
      if (e != cudaSuccess) return cuda_fail(e, "host copy");
      return BHW_OK;
    }
    if (st != 1) return st;  // 1: not eligible for the direct kernels, take the pipeline
  }
  return run_batch_host(d, 1, n0, count, out_host);
}

int bhw_generate_batch(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count,
                       void* out_dev, void* stream) {
  return run_batch(descs, nwin, flat_begin, flat_count, out_dev, (cudaStream_t)stream);
}

int bhw_generate_batch_host(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count,
                            void* out_host) {
  return run_batch_host(descs, nwin, flat_begin, flat_count, out_host);
}

int bhw_generate_batch_multi(const bhw_desc* descs, int nwin, int ngpus, void* const* outs_dev) {
  if (!descs || !outs_dev) return BHW_E_NULL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return BHW_E_NO_DEVICE;
  if (ngpus < 1 || ngpus > ndev) return BHW_E_NO_DEVICE;
  uint64_t total = 0;
  int st = bhw_batch_total(descs, nwin, &total);
  if (st) return st;
  int prev = 0;
  cudaGetDevice(&prev);
  // enqueue every shard first, then wait: the devices run concurrently, no collective involved
  for (int g = 0; g < ngpus && !st; g++) {
    uint64_t b, c;
    bhw_shard_range(total, g, ngpus, &b, &c);
    if (cudaSetDevice(g) != cudaSuccess) { st = BHW_E_NO_DEVICE; break; }
    st = run_batch(descs, nwin, b, c, outs_dev[g], nullptr);
  }
  for (int g = 0; g < ngpus; g++) {
    if (cudaSetDevice(g) != cudaSuccess) continue;
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess && !st) st = cuda_fail(e, "cudaDeviceSynchronize");
  }
  cudaSetDevice(prev);
  return st;
}

int bhw_shard_windows(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count,
                      int* first_win, int* nwin_touched, uint64_t* local_begin) {
  if (!descs || !first_win || !nwin_touched || !local_begin) return BHW_E_NULL;
  if (nwin <= 0) return BHW_E_ARG;
  return shard_windows(descs, nwin, flat_begin, flat_count, first_win, nwin_touched, local_begin);
}

int bhw_plan_create(const bhw_desc* descs, int nwin, bhw_plan** plan_out) {
  if (!descs || !plan_out) return BHW_E_NULL;
  *plan_out = nullptr;
  if (nwin <= 0) return BHW_E_ARG;
  for (int w = 0; w < nwin; w++) {
    int st = validate_desc(&descs[w], true);
    if (st) return st;
  }
  bhw_plan* p = new (std::nothrow) bhw_plan();
  if (!p) return BHW_E_ALLOC;
  uint64_t total = 0;
  int st = bhw_batch_total(descs, nwin, &total);
  if (!st) st = plan_build(*p, descs, nwin, 0, total, nullptr);
  if (st) { plan_free_device(*p, nullptr); delete p; return st; }
  *plan_out = p;
  return BHW_OK;
}

int bhw_plan_execute(bhw_plan* plan, uint64_t flat_begin, uint64_t flat_count, void* out_dev, void* stream) {
  if (!plan) return BHW_E_NULL;
  int dev;
  int st = current_device(&dev);
  if (st) return st;
  if (dev != plan->dev) return BHW_E_ARG;
  return plan_execute(*plan, flat_begin, flat_count, out_dev, (cudaStream_t)stream);
}

// undocumented: print how a plan is going to be executed (stderr)
__attribute__((visibility("default"))) int bhw_plan_debug_dump(const bhw_plan* plan) {
  if (!plan) return BHW_E_NULL;
  fprintf(stderr, "plan: nwin=%d total=%llu recs=%zu tables=%zu jobs=%zu uniform_pw=%d all_same=%d elem64=%d\n", plan->nwin,
          (unsigned long long)plan->total, plan->recs.size(), plan->tables.size(), plan->jobs.size(), plan->uniform_pw,
          (int)plan->all_same, (int)plan->elem64);
  for (size_t i = 0; i < plan->tables.size(); i++)
    fprintf(stderr, "  table %zu: kind=%d pw=%d dw=%d entries=%u ptr=%p\n", i, plan->tables[i].canon.kind,
            plan->tables[i].canon.pw, plan->tables[i].canon.dw, plan->tables[i].entries, (void*)plan->tables[i].ptr);
  for (const bhw_plan::BankRun& r : plan->runs)
    fprintf(stderr, "  run: windows [%d,%d) flat_off=%llu m=%u pw=%u mode=%d pair=%d lin=%u ntab=%u smem_words=%u tab0=%p entries0=%u "
            "kstep1=%u idx_rsh1=%u\n", r.w_begin, r.w_end, (unsigned long long)r.flat_off, r.sh.m, r.sh.pw, r.tab_mode,
            (int)r.pair, r.sh.lin, r.sh.ntab, r.sh.smem_words, (const void*)r.sh.tab[0], r.sh.tentries[0], r.sh.kstep[1],
            r.sh.idx_rsh[1]);
  return BHW_OK;
}

int bhw_plan_total(const bhw_plan* plan, uint64_t* total_samples) {
  if (!plan || !total_samples) return BHW_E_NULL;
  *total_samples = plan->total;
  return BHW_OK;
}

int bhw_plan_destroy(bhw_plan* plan) {
  if (!plan) return BHW_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  if (prev != plan->dev) cudaSetDevice(plan->dev);
  plan_free_device(*plan, nullptr);  // cudaFree waits for work that still uses the memory
  if (prev != plan->dev) cudaSetDevice(prev);
  delete plan;
  return BHW_OK;
}

int bhw_sincos(const bhw_desc* d, void* out_sin_dev, void* out_cos_dev, uint64_t n0, uint64_t count,
               void* stream) {
  if (!d) return BHW_E_NULL;
  int st = validate_desc(d, false);
  if (st) return st;
  if (d->out_format != BHW_OUT_DEFAULT) return BHW_E_ARG;   // DT_SIN / DT_COS come in the default container
  const uint64_t N = 1ull << d->phi_width;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
  int dev;
  if ((st = current_device(&dev))) return st;
  SinCosArgs a;
  memset(&a, 0, sizeof(a));
  if ((st = resolve_source(d, 0, &a.src))) return st;
  init_src_core(a.src, &a.sc);
  if (a.src.kind == SRC_TAYLOR && (st = get_rom(dev, a.src.dw, a.src.lut, &a.rom))) return st;
  a.n_first = n0;
  a.count = count;
  // the whole table of a source with an output quadrant mux: four phases per evaluation
  a.quad = (n0 == 0 && a.src.pw >= 3 && count == (1ull << a.src.pw) && a.src.kind != SRC_INQ) ? 1u : 0u;
  if (!count || (!out_sin_dev && !out_cos_dev)) return BHW_OK;
  cudaError_t e;
  {
    LaunchTimer tm(BHW_KERNEL_SINCOS, (cudaStream_t)stream);
    e = launch_sincos(a, out_sin_dev, out_cos_dev, d->dat_width > 32, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "k_sincos");
  g_launches++;
  return BHW_OK;
}

int bhw_atan2_validate(const bhw_atan2_desc* d) { return resolve_atan2(d, nullptr); }

int bhw_atan2(const bhw_atan2_desc* d, const int32_t* x_dev, const int32_t* y_dev, int32_t* phi_dev, uint64_t count,
              void* stream) {
  Atan2Params p;
  int st = resolve_atan2(d, &p);
  if (st) return st;
  if (!count) return BHW_OK;
  if (!x_dev || !y_dev || !phi_dev) return BHW_E_NULL;
  int dev;
  if ((st = current_device(&dev))) return st;
  cudaError_t e;
  {
    LaunchTimer tm(BHW_KERNEL_ATAN2, (cudaStream_t)stream);
    e = launch_atan2(p, x_dev, y_dev, phi_dev, count, (cudaStream_t)stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "k_atan2");
  g_launches++;
  return BHW_OK;
}

int bhw_atan2_host(const bhw_atan2_desc* d, const int32_t* x_host, const int32_t* y_host, int32_t* phi_host,
                   uint64_t count) {
  Atan2Params p;
  int st = resolve_atan2(d, &p);
  if (st) return st;
  if (!count) return BHW_OK;
  if (!x_host || !y_host || !phi_host) return BHW_E_NULL;
  int dev;
  if ((st = current_device(&dev))) return st;
  DeviceState& ds = g_dev[dev];
  std::lock_guard<std::mutex> pipe_lock(ds.pipe_mu);
  HostPipe& hp = ds.pipe;
  cudaError_t e = pipe_ensure(hp);
  if (e != cudaSuccess) { pipe_release(hp); return cuda_fail(e, "host pipeline setup"); }
  // buf[0] = x | y (two halves), buf[1] = phi; one stream, chunk after chunk
  const uint64_t chunk = hp.buf_bytes / 8 - 1;           // one spare pair: stream_quadrant looks one pair ahead
  int32_t* xd = (int32_t*)hp.buf[0];
  int32_t* yd = xd + chunk + 1;
  int32_t* pd = (int32_t*)hp.buf[1];
  for (uint64_t done = 0; done < count && e == cudaSuccess; done += chunk) {
    const uint64_t cnt = count - done < chunk ? count - done : chunk;
    const uint64_t avail = done + cnt < count ? cnt + 1 : cnt;
    if ((e = cudaMemcpyAsync(xd, x_host + done, avail * 4, cudaMemcpyHostToDevice, hp.s_gen)) != cudaSuccess) break;
    if ((e = cudaMemcpyAsync(yd, y_host + done, avail * 4, cudaMemcpyHostToDevice, hp.s_gen)) != cudaSuccess) break;
    if ((e = launch_atan2(p, xd, yd, pd, cnt, hp.s_gen, avail)) != cudaSuccess) break;
    g_launches++;
    e = cudaMemcpyAsync(phi_host + done, pd, cnt * 4, cudaMemcpyDeviceToHost, hp.s_gen);
  }
  const cudaError_t e2 = cudaStreamSynchronize(hp.s_gen);
  if (e == cudaSuccess) e = e2;
  if (e != cudaSuccess) return cuda_fail(e, "bhw_atan2_host");
  return BHW_OK;
}

int bhw_cache_clear(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) return BHW_E_NO_DEVICE;
  int prev = 0;
  cudaGetDevice(&prev);
  for (int g = 0; g < ndev && g < 64; g++) {
    DeviceState& ds = g_dev[g];
    std::lock_guard<std::mutex> pl(ds.pipe_mu);
    std::lock_guard<std::mutex> lk(ds.mu);
    if (ds.roms.empty() && !ds.pipe.s_gen && !ds.pool && !ds.ev_fork) continue;
    cudaSetDevice(g);
    cudaDeviceSynchronize();
    for (auto& kv : ds.roms) cudaFree(kv.second.ptr);
    ds.roms.clear();
    pipe_release(ds.pipe);
    if (ds.pool) { cudaMemPoolDestroy(ds.pool); ds.pool = nullptr; }
    {
      std::lock_guard<std::mutex> fl(ds.fan_mu);
      for (int i = 0; i < kMaxSideStreams; i++) {
        if (ds.side[i]) { cudaStreamDestroy(ds.side[i]); ds.side[i] = nullptr; }
        if (ds.ev_join[i]) { cudaEventDestroy(ds.ev_join[i]); ds.ev_join[i] = nullptr; }
      }
      if (ds.ev_fork) { cudaEventDestroy(ds.ev_fork); ds.ev_fork = nullptr; }
    }
  }
  cudaSetDevice(prev);
  return BHW_OK;
}

int bhw_set_table_cache(int enabled) {
  g_cache_enabled.store(enabled ? 1 : 0);
  return BHW_OK;
}

// undocumented tuning knob: terms from which long pyramid windows get their own spread launch (applies to plans created afterwards)
__attribute__((visibility("default"))) int bhw_debug_set_spread_min_terms(int n) { g_spread_min_terms.store(n); return BHW_OK; }

// undocumented tuning knob: L1 prefetch depth of the spread launches
__attribute__((visibility("default"))) int bhw_debug_set_prefetch_lines(int n) { g_prefetch_lines.store(n); return BHW_OK; }

// undocumented tuning knob: CTAs per SM of a deferred family table build (0 = do not defer)
__attribute__((visibility("default"))) int bhw_debug_set_defer_ctas(int n) {
  if (n < 0 || n > 8) return BHW_E_ARG;
  g_defer_ctas.store(n);
  return BHW_OK;
}

int bhw_set_side_streams(int n) {
  if (n < 0 || n > kMaxSideStreams) return BHW_E_ARG;
  g_side_streams.store(n);
  return BHW_OK;
}

int bhw_timing_enable(int enabled) {
  g_timing.store(enabled ? 1 : 0);
  return BHW_OK;
}

// fold finished spans into the per-class totals (synchronises on each span's end event)
static int timing_collect() {
  std::lock_guard<std::mutex> lk(g_timing_mu);
  int prev = 0;
  cudaGetDevice(&prev);
  int st = BHW_OK;
  for (auto& sp : g_spans) {
    cudaSetDevice(sp.dev);
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(sp.b);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, sp.a, sp.b);
    if (e == cudaSuccess) {
      g_time_ms[sp.cls] += (double)ms; g_time_n[sp.cls]++;
      if (g_launch_log.size() < (1u << 20)) {
        bhw_launch_record rec;
        rec.kernel_class = sp.cls; rec.tag = sp.tag; rec.bytes = sp.bytes; rec.ms = (double)ms;
        g_launch_log.push_back(rec);
      }
    }
    else st = cuda_fail(e, "timing");
    g_event_pool[sp.dev].push_back(sp.a);
    g_event_pool[sp.dev].push_back(sp.b);
  }
  g_spans.clear();
  cudaSetDevice(prev);
  return st;
}

int bhw_timing_reset(void) {
  int st = timing_collect();
  std::lock_guard<std::mutex> lk(g_timing_mu);
  for (int i = 0; i < BHW_KERNEL_CLASSES; i++) { g_time_ms[i] = 0; g_time_n[i] = 0; }
  g_launch_log.clear();
  return st;
}

int bhw_timing_read(int kernel_class, double* total_ms, uint64_t* launches) {
  if (kernel_class < 0 || kernel_class >= BHW_KERNEL_CLASSES) return BHW_E_ARG;
  int st = timing_collect();
  std::lock_guard<std::mutex> lk(g_timing_mu);
  if (total_ms) *total_ms = g_time_ms[kernel_class];
  if (launches) *launches = g_time_n[kernel_class];
  return st;
}

int bhw_timing_launches(bhw_launch_record* out, uint64_t max_records, uint64_t* n_records) {
  if (!n_records) return BHW_E_NULL;
  int st = timing_collect();
  std::lock_guard<std::mutex> lk(g_timing_mu);
  *n_records = g_launch_log.size();
  if (out)
    for (uint64_t i = 0; i < max_records && i < g_launch_log.size(); i++) out[i] = g_launch_log[i];
  return st;
}

uint64_t bhw_launch_count(void) { return g_launches.load(); }

const char* bhw_last_cuda_error(void) { return t_cuda_err.c_str(); }

int bhw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

}  // extern "C"
