// bhw_api.cu - the GPU half of the C ABI (include/bhw.h): planning, trig-table cache,
// descriptor upload, launches, host-buffer pipelines and the single-process multi-GPU form.
//
// There is deliberately no CPU evaluation path in this file: every generate/sincos entry point
// ends in a kernel launch or returns an error.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "bhw_device.cuh"
#include "bhw_launch.h"
#include "bhw_plan.h"

namespace bhw {

static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_cache_enabled{1};
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

// ---- per-device persistent state -------------------------------------------------------------
struct CachedTable {
  int32_t* ptr = nullptr;
  uint32_t entries = 0;
  cudaEvent_t ready = nullptr;  // recorded after the build, waited on by consumers
};
struct CachedRom {
  I2* ptr = nullptr;
  uint32_t entries = 0;
};
struct DeviceState {
  std::mutex call_mu;  // serialises table-path calls on a device while the cache is in use
  std::mutex mu;
  std::map<std::string, CachedTable> tables;  // key: bytes of the canonical SrcParams
  std::map<uint32_t, CachedRom> roms;         // key: dw << 8 | lut
};
static DeviceState g_dev[64];

static std::string key_of(const SrcParams& sp) { return std::string((const char*)&sp, sizeof(sp)); }

static int get_rom(int dev, int dw, int lut, cudaStream_t stream, const I2** out) {
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
  (void)stream;
  *out = it->second.ptr;
  return BHW_OK;
}

// ---- one batch call --------------------------------------------------------------------------
struct PlanTable {
  SrcParams canon;
  uint32_t drop;
  uint32_t entries;
  int32_t* ptr;
  bool build;       // must be built in this call
  bool transient;   // free after the call (cache disabled)
};

struct Plan {
  std::vector<WinRec> recs;
  std::vector<GenRec> gens;
  std::vector<uint32_t> win_rec;
  std::vector<uint64_t> flat_off;
  std::vector<PlanTable> tables;
  std::vector<TabJob> jobs;
  const I2* rom = nullptr;  // at most one Taylor ROM (DW, LUT) per call is supported per window set
  int uniform_pw = -1;
  bool all_same = false;
};

static int find_or_add_table(int dev, Plan& plan, const SrcParams& sp, cudaStream_t stream, int* index) {
  uint32_t drop;
  const SrcParams canon = canonical_source(sp, &drop);
  for (size_t i = 0; i < plan.tables.size(); i++)
    if (!memcmp(&plan.tables[i].canon, &canon, sizeof(canon))) { *index = (int)i; return BHW_OK; }
  PlanTable pt;
  pt.canon = canon;
  pt.drop = drop;
  pt.entries = 1u << canon.pw;
  pt.ptr = nullptr;
  pt.build = true;
  pt.transient = !g_cache_enabled.load();
  const size_t bytes = (size_t)pt.entries * sizeof(int32_t);
  if (pt.transient) {
    BHW_CUDA(cudaMallocAsync((void**)&pt.ptr, bytes, stream));
  } else {
    DeviceState& ds = g_dev[dev];
    std::lock_guard<std::mutex> lk(ds.mu);
    auto it = ds.tables.find(key_of(canon));
    if (it != ds.tables.end()) {
      pt.ptr = it->second.ptr;
      pt.build = false;
      BHW_CUDA(cudaStreamWaitEvent(stream, it->second.ready, 0));
    } else {
      CachedTable ct;
      ct.entries = pt.entries;
      BHW_CUDA(cudaMalloc((void**)&ct.ptr, bytes));
      cudaError_t e = cudaEventCreateWithFlags(&ct.ready, cudaEventDisableTiming);
      if (e != cudaSuccess) { cudaFree(ct.ptr); return cuda_fail(e, "cudaEventCreate"); }
      ds.tables.emplace(key_of(canon), ct);
      pt.ptr = ct.ptr;
    }
  }
  plan.tables.push_back(pt);
  *index = (int)plan.tables.size() - 1;
  return BHW_OK;
}

static int plan_batch(int dev, const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count,
                      cudaStream_t stream, Plan& plan) {
  plan.flat_off.resize((size_t)nwin + 1);
  plan.win_rec.resize((size_t)nwin);
  // pass 1: flat offsets; windows with a byte-identical descriptor share one record
  std::vector<int> rec_desc;       // descriptor index that defines record i
  std::vector<uint64_t> rec_used;  // requested samples that fall into windows of record i
  uint64_t off = 0;
  const int pw0 = descs[0].phi_width;
  bool uniform = true;
  const uint64_t req_end = flat_begin + flat_count;
  for (int w = 0; w < nwin; w++) {
    const bhw_desc& d = descs[w];
    plan.flat_off[w] = off;
    if (d.phi_width < 4 || d.phi_width > 30) return BHW_E_PHI_WIDTH;
    const uint64_t N = 1ull << d.phi_width;
    if (d.phi_width != pw0) uniform = false;
    const uint64_t lo = off > flat_begin ? off : flat_begin;
    const uint64_t hi = off + N < req_end ? off + N : req_end;
    const uint64_t used = lo < hi ? hi - lo : 0;
    off += N;
    int ri = -1;
    if (w > 0 && !memcmp(&d, &descs[w - 1], sizeof(d))) ri = (int)plan.win_rec[w - 1];
    else
      for (size_t i = 0; i < rec_desc.size(); i++)
        if (!memcmp(&d, &descs[rec_desc[i]], sizeof(d))) { ri = (int)i; break; }
    if (ri < 0) { ri = (int)rec_desc.size(); rec_desc.push_back(w); rec_used.push_back(0); }
    plan.win_rec[w] = (uint32_t)ri;
    rec_used[(size_t)ri] += used;
  }
  plan.flat_off[nwin] = off;
  if (flat_begin > off || flat_count > off - flat_begin) return BHW_E_RANGE;
  plan.uniform_pw = uniform ? pw0 : -1;
  plan.all_same = rec_desc.size() == 1;

  // pass 2: validate + resolve every distinct descriptor (touched by the request or not), and
  // give the touched ones their trig tables
  for (size_t i = 0; i < rec_desc.size(); i++) {
    const bhw_desc& d = descs[rec_desc[i]];
    WinParams wp; SrcParams src[2];
    int st = resolve_window(&d, &wp, src);
    if (st) return st;
    if (wp.elem64) return BHW_E_ELEM;  // 64-bit windows go through the per-window direct path
    WinRec r;
    memset(&r, 0, sizeof(r));
    r.n_first = (uint32_t)wp.stream_offset;
    bool generic = d.algo == BHW_ALGO_DIRECT || !fast_tail_exact(wp);
    if (!generic && d.algo == BHW_ALGO_AUTO && rec_used[i]) {
      // a short request into a long window: building the table would cost more than evaluating
      // the requested samples directly
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
      plan.gens.push_back(g);
      if (src[0].kind == SRC_TAYLOR && rec_used[i]) {
        const I2* rom = nullptr;
        if ((st = get_rom(dev, src[0].dw, src[0].lut, stream, &rom))) return st;
        if (plan.rom && plan.rom != rom) return BHW_E_ARG;  // one Taylor (DW, LUT_SIZE) per call
        plan.rom = rom;
      }
    } else {
      fill_fast_rec(wp, r);
      if (rec_used[i]) {
        for (int k = 1; k < wp.m; k++) {
          const TermParams& t = wp.term[k - 1];
          const SrcParams& sp = src[t.src];
          int ti;
          if ((st = find_or_add_table(dev, plan, sp, stream, &ti))) return st;
          const PlanTable& pt = plan.tables[(size_t)ti];
          r.tabp[k] = pt.ptr;
          r.kstep[k] = t.kmul << (32 - sp.pw);
          r.idx_rsh[k] = (uint32_t)(32 - (sp.pw - (int)pt.drop));
        }
      }
    }
    plan.recs.push_back(r);
  }
  uint32_t work = 0;
  for (size_t i = 0; i < plan.tables.size(); i++) {
    PlanTable& pt = plan.tables[i];
    if (!pt.build) continue;
    TabJob j;
    memset(&j, 0, sizeof(j));
    j.sp = pt.canon;
    j.tab = pt.ptr;
    j.entries = pt.entries;
    j.fast = fast32_ok(pt.canon) ? 1u : 0u;
    j.work_begin = work;
    j.work = pt.canon.kind == SRC_INQ ? pt.entries : pt.entries / 4;
    j.rom_off = 0;
    if (pt.canon.kind == SRC_TAYLOR) {
      const I2* rom = nullptr;
      int st = get_rom(dev, pt.canon.dw, pt.canon.lut, stream, &rom);
      if (st) return st;
      if (plan.rom && plan.rom != rom) return BHW_E_ARG;  // one Taylor (DW, LUT_SIZE) per call
      plan.rom = rom;
    }
    work += j.work;
    plan.jobs.push_back(j);
  }
  return BHW_OK;
}

static int current_device(int* dev) {
  cudaError_t e = cudaGetDevice(dev);
  if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return BHW_E_NO_DEVICE; }
  if (*dev < 0 || *dev >= 64) return BHW_E_NO_DEVICE;
  return BHW_OK;
}

static void release_transient(Plan& plan, cudaStream_t stream) {
  for (auto& pt : plan.tables)
    if (pt.transient && pt.ptr) { cudaFreeAsync(pt.ptr, stream); pt.ptr = nullptr; }
}

// Forget cached tables that were planned but whose build launch never happened.
static void drop_unbuilt(int dev, Plan& plan) {
  DeviceState& ds = g_dev[dev];
  std::lock_guard<std::mutex> lk(ds.mu);
  for (auto& pt : plan.tables)
    if (pt.build && !pt.transient) {
      auto it = ds.tables.find(key_of(pt.canon));
      if (it != ds.tables.end()) {
        cudaFree(it->second.ptr);
        cudaEventDestroy(it->second.ready);
        ds.tables.erase(it);
      }
    }
}

// The table-path executor: upload the call's metadata in one copy, build missing tables in one
// launch, synthesise the flat range in one launch.
static int run_batch_i32(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count,
                         void* out_dev, cudaStream_t stream) {
  int dev;
  int st = current_device(&dev);
  if (st) return st;
  // a cached table is visible to other threads from the moment it is planned, so planning and
  // the build launch of one call must not interleave with another call on the same device
  std::unique_lock<std::mutex> call_lock(g_dev[dev].call_mu, std::defer_lock);
  if (g_cache_enabled.load()) call_lock.lock();
  Plan plan;
  st = plan_batch(dev, descs, nwin, flat_begin, flat_count, stream, plan);
  if (st) { release_transient(plan, stream); drop_unbuilt(dev, plan); return st; }
  if (!flat_count) { release_transient(plan, stream); return BHW_OK; }

  // one metadata blob: [recs][gens][jobs][flat_off][win_rec]
  auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
  const size_t o_recs = 0;
  const size_t o_gens = align16(o_recs + plan.recs.size() * sizeof(WinRec));
  const size_t o_jobs = align16(o_gens + plan.gens.size() * sizeof(GenRec));
  const size_t o_off = align16(o_jobs + plan.jobs.size() * sizeof(TabJob));
  const bool need_off = plan.uniform_pw < 0;
  const size_t o_wr = align16(o_off + (need_off ? plan.flat_off.size() * sizeof(uint64_t) : 0));
  const bool need_wr = !plan.all_same;
  const size_t total = align16(o_wr + (need_wr ? plan.win_rec.size() * sizeof(uint32_t) : 0));
  std::vector<char> blob(total);
  memcpy(blob.data() + o_recs, plan.recs.data(), plan.recs.size() * sizeof(WinRec));
  if (!plan.gens.empty()) memcpy(blob.data() + o_gens, plan.gens.data(), plan.gens.size() * sizeof(GenRec));
  if (!plan.jobs.empty()) memcpy(blob.data() + o_jobs, plan.jobs.data(), plan.jobs.size() * sizeof(TabJob));
  if (need_off) memcpy(blob.data() + o_off, plan.flat_off.data(), plan.flat_off.size() * sizeof(uint64_t));
  if (need_wr) memcpy(blob.data() + o_wr, plan.win_rec.data(), plan.win_rec.size() * sizeof(uint32_t));
  char* blob_dev = nullptr;
  cudaError_t e = cudaMallocAsync((void**)&blob_dev, total, stream);
  if (e != cudaSuccess) {
    release_transient(plan, stream);
    drop_unbuilt(dev, plan);
    return cuda_fail(e, "cudaMallocAsync(meta)");
  }
  bool built = plan.jobs.empty();
  // pageable source: the runtime stages it before returning, so `blob` may die with this frame
  e = cudaMemcpyAsync(blob_dev, blob.data(), total, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess && !plan.jobs.empty()) {
    const TabJob& last = plan.jobs.back();
    e = launch_table_build((const TabJob*)(blob_dev + o_jobs), (int)plan.jobs.size(),
                           last.work_begin + last.work, plan.rom, stream);
    if (e == cudaSuccess) { g_launches++; built = true; }
    if (e == cudaSuccess) {
      DeviceState& ds = g_dev[dev];
      std::lock_guard<std::mutex> lk(ds.mu);
      for (auto& pt : plan.tables)
        if (pt.build && !pt.transient) {
          auto it = ds.tables.find(key_of(pt.canon));
          if (it != ds.tables.end()) cudaEventRecord(it->second.ready, stream);
        }
    }
  }
  if (e == cudaSuccess) {
    SynthArgs a;
    a.recs = (const WinRec*)(blob_dev + o_recs);
    a.win_rec = need_wr ? (const uint32_t*)(blob_dev + o_wr) : nullptr;
    a.flat_off = need_off ? (const uint64_t*)(blob_dev + o_off) : nullptr;
    a.gens = (const GenRec*)(blob_dev + o_gens);
    a.rom = plan.rom;
    a.out = out_dev;
    a.flat_begin = flat_begin;
    a.flat_count = flat_count;
    a.nwin = nwin;
    a.uniform_pw = plan.uniform_pw;
    e = launch_synth(a, stream);
    if (e == cudaSuccess) g_launches++;
  }
  cudaFreeAsync(blob_dev, stream);
  release_transient(plan, stream);
  if (!built) drop_unbuilt(dev, plan);
  if (e != cudaSuccess) return cuda_fail(e, "launch");
  return BHW_OK;
}

// One window through the direct kernel (BHW_ALGO_DIRECT, and always for DAT_WIDTH > 32).
static int run_direct(const bhw_desc* d, uint64_t n0, uint64_t count, void* out_dev, cudaStream_t stream) {
  int dev;
  int st = current_device(&dev);
  if (st) return st;
  DirectArgs a;
  memset(&a, 0, sizeof(a));
  if ((st = resolve_window(d, &a.wp, a.src))) return st;
  const uint64_t N = 1ull << d->phi_width;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
  if (a.src[0].kind == SRC_TAYLOR) {
    if ((st = get_rom(dev, a.src[0].dw, a.src[0].lut, stream, &a.rom))) return st;
    const uint32_t entries = 1u << a.src[0].lut;
    a.rom_smem_entries = entries * sizeof(I2) <= 32768 ? entries : 0;  // the sine LUT lives in shared memory
  }
  a.n_first = n0 + (uint64_t)d->stream_offset;
  a.count = count;
  cudaError_t e = launch_direct_window(a, out_dev, stream);
  if (e != cudaSuccess) return cuda_fail(e, "k_direct_window");
  if (count) g_launches++;
  return BHW_OK;
}

static bool batch_is_elem64(const bhw_desc* descs, int nwin, bool* mixed) {
  bool any64 = false, any32 = false;
  for (int i = 0; i < nwin; i++) (descs[i].dat_width > 32 ? any64 : any32) = true;
  *mixed = any64 && any32;
  return any64;
}

static int run_batch(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count, void* out_dev,
                     cudaStream_t stream) {
  if (!descs) return BHW_E_NULL;
  if (nwin <= 0) return BHW_E_ARG;
  if (!out_dev && flat_count) return BHW_E_NULL;
  bool mixed;
  const bool e64 = batch_is_elem64(descs, nwin, &mixed);
  if (mixed) return BHW_E_ELEM;
  if (!e64) return run_batch_i32(descs, nwin, flat_begin, flat_count, out_dev, stream);
  // DAT_WIDTH > 32: window by window through the direct kernel
  uint64_t off = 0;
  for (int w = 0; w < nwin; w++) {
    int st = validate_desc(&descs[w], true);
    if (st) return st;
    off += 1ull << descs[w].phi_width;
  }
  if (flat_begin > off || flat_count > off - flat_begin) return BHW_E_RANGE;
  off = 0;
  for (int w = 0; w < nwin; w++) {
    const uint64_t N = 1ull << descs[w].phi_width;
    const uint64_t lo = off > flat_begin ? off : flat_begin;
    const uint64_t hi = off + N < flat_begin + flat_count ? off + N : flat_begin + flat_count;
    if (lo < hi) {
      int st = run_direct(&descs[w], lo - off, hi - lo, (int64_t*)out_dev + (lo - flat_begin), stream);
      if (st) return st;
    }
    off += N;
  }
  return BHW_OK;
}

// ---- host-buffer pipeline ----------------------------------------------------------------------
// Generates chunk c on the compute stream while chunk c-1 drains to the host on the copy stream.
static int run_batch_host(const bhw_desc* descs, int nwin, uint64_t flat_begin, uint64_t flat_count,
                          void* out_host) {
  if (!descs) return BHW_E_NULL;
  if (nwin <= 0) return BHW_E_ARG;
  if (!out_host && flat_count) return BHW_E_NULL;
  bool mixed;
  const bool e64 = batch_is_elem64(descs, nwin, &mixed);
  if (mixed) return BHW_E_ELEM;
  const size_t esz = e64 ? 8 : 4;
  if (!flat_count) {
    // still validate
    return run_batch(descs, nwin, flat_begin, 0, (void*)descs, nullptr);
  }
  const uint64_t chunk = flat_count < (16ull << 20) ? flat_count : (16ull << 20);  // samples per chunk
  cudaStream_t s_gen = nullptr, s_copy = nullptr;
  cudaEvent_t ev_gen[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
  char* buf[2] = {nullptr, nullptr};
  int st = BHW_OK;
  cudaError_t e = cudaSuccess;
  const int nbuf = flat_count > chunk ? 2 : 1;
  do {
    if ((e = cudaStreamCreateWithFlags(&s_gen, cudaStreamNonBlocking)) != cudaSuccess) break;
    if ((e = cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking)) != cudaSuccess) break;
    for (int i = 0; i < nbuf && e == cudaSuccess; i++) {
      e = cudaMalloc((void**)&buf[i], chunk * esz);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_gen[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) break;
    uint64_t done = 0;
    for (int c = 0; done < flat_count; c++) {
      const int b = c % nbuf;
      const uint64_t cnt = flat_count - done < chunk ? flat_count - done : chunk;
      if (c >= nbuf) {  // buffer b must have drained before it is overwritten
        if ((e = cudaStreamWaitEvent(s_gen, ev_copy[b], 0)) != cudaSuccess) break;
      }
      st = run_batch(descs, nwin, flat_begin + done, cnt, buf[b], s_gen);
      if (st) break;
      if ((e = cudaEventRecord(ev_gen[b], s_gen)) != cudaSuccess) break;
      if ((e = cudaStreamWaitEvent(s_copy, ev_gen[b], 0)) != cudaSuccess) break;
      if ((e = cudaMemcpyAsync((char*)out_host + done * esz, buf[b], cnt * esz, cudaMemcpyDeviceToHost,
                               s_copy)) != cudaSuccess) break;
      if ((e = cudaEventRecord(ev_copy[b], s_copy)) != cudaSuccess) break;
      done += cnt;
    }
  } while (0);
  if (s_gen) { cudaError_t e2 = cudaStreamSynchronize(s_gen); if (e == cudaSuccess) e = e2; }
  if (s_copy) { cudaError_t e2 = cudaStreamSynchronize(s_copy); if (e == cudaSuccess) e = e2; }
  for (int i = 0; i < 2; i++) {
    if (buf[i]) cudaFree(buf[i]);
    if (ev_gen[i]) cudaEventDestroy(ev_gen[i]);
    if (ev_copy[i]) cudaEventDestroy(ev_copy[i]);
  }
  if (s_gen) cudaStreamDestroy(s_gen);
  if (s_copy) cudaStreamDestroy(s_copy);
  if (st) return st;
  if (e != cudaSuccess) return cuda_fail(e, "host pipeline");
  return BHW_OK;
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
  if (d->dat_width > 32 || d->algo == BHW_ALGO_DIRECT) return run_direct(d, n0, count, out_dev, (cudaStream_t)stream);
  return run_batch(d, 1, n0, count, out_dev, (cudaStream_t)stream);
}

int bhw_generate_host(const bhw_desc* d, void* out_host, uint64_t n0, uint64_t count) {
  if (!d) return BHW_E_NULL;
  int st = validate_desc(d, true);
  if (st) return st;
  const uint64_t N = 1ull << d->phi_width;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
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

int bhw_sincos(const bhw_desc* d, void* out_sin_dev, void* out_cos_dev, uint64_t n0, uint64_t count,
               void* stream) {
  if (!d) return BHW_E_NULL;
  int st = validate_desc(d, false);
  if (st) return st;
  const uint64_t N = 1ull << d->phi_width;
  if (n0 > N || count > N - n0) return BHW_E_RANGE;
  int dev;
  if ((st = current_device(&dev))) return st;
  SinCosArgs a;
  memset(&a, 0, sizeof(a));
  if ((st = resolve_source(d, 0, &a.src))) return st;
  if (a.src.kind == SRC_TAYLOR && (st = get_rom(dev, a.src.dw, a.src.lut, (cudaStream_t)stream, &a.rom))) return st;
  a.n_first = n0;
  a.count = count;
  if (!count || (!out_sin_dev && !out_cos_dev)) return BHW_OK;
  cudaError_t e = launch_sincos(a, out_sin_dev, out_cos_dev, d->dat_width > 32, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_fail(e, "k_sincos");
  g_launches++;
  return BHW_OK;
}

int bhw_cache_clear(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) return BHW_E_NO_DEVICE;
  int prev = 0;
  cudaGetDevice(&prev);
  for (int g = 0; g < ndev && g < 64; g++) {
    DeviceState& ds = g_dev[g];
    std::lock_guard<std::mutex> lk(ds.mu);
    if (ds.tables.empty() && ds.roms.empty()) continue;
    cudaSetDevice(g);
    cudaDeviceSynchronize();
    for (auto& kv : ds.tables) { cudaFree(kv.second.ptr); cudaEventDestroy(kv.second.ready); }
    for (auto& kv : ds.roms) cudaFree(kv.second.ptr);
    ds.tables.clear();
    ds.roms.clear();
  }
  cudaSetDevice(prev);
  return BHW_OK;
}

int bhw_set_table_cache(int enabled) {
  g_cache_enabled.store(enabled ? 1 : 0);
  return BHW_OK;
}

uint64_t bhw_launch_count(void) { return g_launches.load(); }

const char* bhw_last_cuda_error(void) { return t_cuda_err.c_str(); }

int bhw_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

}  // extern "C"
