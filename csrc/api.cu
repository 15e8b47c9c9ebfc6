// api.cu — the C ABI of include/bkgpu.h: plan lifecycle (ExecNode::init/open/get_next/close
// inverted into init/open/push/finish/get_next/close), batch staging, result materialisation.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../include/bkgpu.h"
#include "agg.h"
#include "plan.h"
#include "sort.h"
#include "nccl_dl.h"
#include "build_id.h"

using namespace bk;

namespace {

thread_local std::string g_thread_error;

enum State { S_INIT = 0, S_OPEN = 1, S_FINISHED = 2, S_CLOSED = 3 };

struct HostCol {  // one materialised result column (host memory owned by the plan)
    OutCol desc;
    int elem = 0;
    std::vector<uint8_t> values;
    std::vector<uint8_t> validity;  // LSB bitmap; empty = all valid
    void* dev_values = nullptr;     // output_on_device
};

struct EventPair { cudaEvent_t a, b; int64_t bytes; };

// BKGPU_TRACE=1: wall-clock time the HOST spends in each phase of a request (summed per plan, printed by bkgpu_close)
struct HostClock {
    double reset = 0, push = 0, collective = 0, extract_wait = 0, finish = 0; int64_t n = 0;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6; }
};

}  // namespace

struct bkgpu_plan {
    Compiled c;
    int device = 0;
    int sm_count = 148;
    State state = S_INIT;
    std::string last_error;
    std::atomic<int> cancelled{0};
    void* nccl_comm = nullptr;
    int nranks = 1;
    // options
    cudaStream_t stream = nullptr; bool own_stream = false;
    cudaStream_t copy_stream = nullptr;
    int group_cap_log2 = 20;
    int smem_cap_log2 = -1;       // -1 = choose per batch
    int64_t batch_capacity = 1 << 20;
    int64_t chunk_rows = 8 << 20;
    int64_t partial_cap = 1 << 16;
    int force_generic = 0;
    int no_lean = 0;
    int output_on_device = 0;
    int64_t region_base = 0;      // arrival index of this plan's first row (ties across GPUs break by (region, row))
    // aggregate state
    GroupTable gt{};
    uint64_t* d_rows_passed = nullptr;
    uint32_t* d_cursor = nullptr;
    uint64_t* d_partial = nullptr;  // export buffer (this rank)
    uint64_t* d_gather = nullptr;   // nranks export buffers
    size_t d_partial_words = 0, d_gather_words = 0;
    HostClock hclk; bool trace = getenv("BKGPU_TRACE") != nullptr;
    std::vector<uint8_t*> bounce[2]; size_t bounce_rows = 0; cudaEvent_t bounce_done[2] = {nullptr, nullptr}; bool bounce_busy[2] = {false, false};
    int scalar_tma = 1;           // COUNT(*) WHERE int32 <cmp> c runs the TMA-staged kernel (scalar_tma.cu): 0.97 vs 0.78 of HBM (profiles/r02_tma_scalar.md); 0 = the LDG kernel
    int no_bounce = 0;            // 1 = pageable host input goes straight to cudaMemcpyAsync (A/B of the bounce path)
    int no_stream_copy = 0;       // 1 = the bounce copy uses memcpy instead of non-temporal stores (A/B, hostcopy.cpp)
    uint32_t* jr_pairs = nullptr; size_t jr_pairs_cap = 0; uint32_t* jr_cursor = nullptr; size_t jr_cursor_cap = 0;   // PK_JOIN: (probe row, build row) pairs of one batch
    uint8_t* jb_matched = nullptr; size_t jb_matched_cap = 0; bool join_tail_launch = false;   // LEFT / SEMI / ANTI: build rows that found a partner
    SortState* post_sort = nullptr;                         // the post fragment above the aggregate (Compiled::post)
    std::vector<uint8_t*> post_vals, post_nullb, post_bitmap; size_t post_cap = 0;
    uint64_t rows_passed_host = 0;
    int64_t finish_groups = -1;   // groups in the table when the last finish read the result back (-1 = unknown: full re-initialisation)
    uint32_t merge_bound = 0, merge_bound_used = 0;   // groups per rank the all-gather is sized for (learned from earlier runs)
    uint32_t* d_part_cursors = nullptr; int repartition = 0;   // hash repartition of the groups across ranks (option "repartition")
    // merge over NVLink peer memory (option "peer_merge"): this rank's buffer, the peers' mappings of theirs, step counter
    int peer_merge = 0, peer_rank = -1; bool peer_ready = false;
    uint64_t* peer_local = nullptr; std::vector<uint64_t*> peer_ptr; uint64_t** d_peer_ptrs = nullptr; uint64_t peer_seq = 0; size_t peer_seg_words = 0;
    uint32_t* d_peer_timeout = nullptr;
    uint64_t* d_outv = nullptr; uint8_t* d_outn = nullptr; size_t out_cap_alloc = 0;  // extraction buffers (kept across resets)
    std::vector<cudaEvent_t> event_pool;
    uint64_t* h_outv = nullptr; uint8_t* h_outn = nullptr; size_t h_out_cap = 0;   // pinned landing area of the extracted rows
    uint32_t known_groups = 0;
    uint32_t* h_pinned = nullptr;   // [0] groups seen (async copy after every aggregate launch), pinned
    // hash join (K4): retained build side + multimap
    std::vector<uint8_t*> jb_vals, jb_nullbytes, jb_bitmap;   // per plan column (side 1 only)
    std::vector<bool> jb_has_null;
    int64_t jb_rows = 0, jb_cap = 0;
    uint64_t* jt_keys = nullptr; uint32_t* jt_rows = nullptr; uint32_t jt_mask = 0; bool jt_built = false, jt_generic = false;
    JoinFast jf{}; uint32_t* jf_dense = nullptr; uint64_t* jf_packed = nullptr;   // FK -> PK fast path (unique build keys)
    int join_pipeline = 0;        // opt-in: the fused probe issues its lookups one drain ahead (measured equal: the kernel is shared-memory bound, profiles/r02_join_history.md)
    int join_learn_range = 1;     // a re-run plan builds with the key range it saw before (checked by the build kernel): -0.05 ms per C3 request
    int fx_off = 0;               // set when a finished request saw more than 1/64 of its surviving rows take FX's exact path (see agg_finish)
    int lean_bank = 0;            // opt-in: bank-aware dealing of the lean kernel's drain (agg_direct.cuh, BANK)
    int lean_fx = [] { const char* e = getenv("BKGPU_LEAN_FX"); return e ? atoi(e) != 0 : BK_LEAN_FX_DEFAULT; }();   // double sums as fixed-point limbs with native shared atomics (agg_direct.cuh, FX); option lean_fx
    int blocking_sync = -1; bool blocking_wait = false; cudaEvent_t wait_event = nullptr;   // see agg_finish
    bool jf_learned = false; uint64_t jf_learn_min = 0, jf_learn_max = 0;   // key range of the plan's previous build (skips the min/max pass + round trip)
    JoinProbe jp{}; uint32_t* jp_attr = nullptr; uint64_t* jp_packed = nullptr; int jp_key_pos = 0;   // ... fused into the lean aggregate
    size_t jf_dense_cap = 0, jf_packed_cap = 0, jp_attr_cap = 0, jp_packed_cap = 0, j_scratch_cap = 0;
    uint64_t* j_scratch = nullptr;   // [0..1] key min / max, then u32 flags: [4] duplicate build key, [5] fused probe unusable
    int no_fused_probe = 0, no_lean_nulls = 0, no_lean_mm = 0;
    int use_wp = 0;               // 1 = launch the warp-private aggregate kernel where the batch fits it (opt-in: measured at par with k_agg_group_lean, profiles/r02_agg_wp_history.md)
    int wp_kt_log2 = 0;           // log2 words of the warp-private kernel's key table (0 = sized from the cardinality)
    int wp_warps = 0;             // warps per CTA of the warp-private kernel (0 = chosen from the table size; 8, 12 or 16)
    std::vector<uint8_t*> jg_buf; int64_t jg_rows = 0;                            // gathered build columns, one chunk
    std::vector<ColRef> probe_want; std::vector<int> probe_map;   // probe-side columns and their index in c.cols
    // sort / filter state
    SortState* sort = nullptr;
    // host staging for pageable / pinned pushes
    std::vector<void*> stage[2];
    std::vector<void*> stage_valid[2];
    size_t stage_rows = 0;
    cudaEvent_t stage_free[2] = {nullptr, nullptr}, stage_ready[2] = {nullptr, nullptr};
    // results
    std::vector<HostCol> result;
    int64_t result_rows = 0, result_pos = 0;
    // stats
    bkgpu_stats stats{};
    std::vector<EventPair> timed;
    std::vector<EventPair> timed_coll;
    std::vector<std::pair<void*, size_t>> dev_allocs;   // (buffer, size class) — returned to the process-wide cache on free / close

    int fail(int code, const char* fmt, ...) {
        char buf[512]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        last_error = buf; g_thread_error = buf; return code;
    }
    int cuda_fail(cudaError_t e, const char* what) {
        return fail(e == cudaErrorMemoryAllocation ? BKGPU_ENOMEM : BKGPU_ENODEV, "%s: %s", what, cudaGetErrorString(e));
    }
};

#define CK(plan, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return (plan)->cuda_fail(e__, #call); } while (0)

static int thread_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_thread_error = buf;
    return code;
}

// ------------------------------------------------------------------ library
extern "C" const char* bkgpu_version(void) { return "bkgpu 0.2 (sm_100a) src=" BKGPU_SRC_DIGEST; }   // digest of csrc/ + include/ at compile time (csrc/Makefile)

extern "C" int bkgpu_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { g_thread_error = std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e); return BKGPU_ENODEV; }
    return n;
}

extern "C" const char* bkgpu_last_error(bkgpu_plan* p) { return p ? p->last_error.c_str() : g_thread_error.c_str(); }

extern "C" int bkgpu_plan_explain(const uint8_t* desc, size_t len, char* text, size_t text_len) {
    Compiled c; std::string err;
    int rc = compile_plan(desc, len, c, err);
    if (rc != BKGPU_OK) { if (text && text_len) snprintf(text, text_len, "%s", err.c_str()); return thread_fail(rc, err.c_str()); }
    if (text && text_len) snprintf(text, text_len, "%s", c.explain.c_str());
    return BKGPU_OK;
}

// ------------------------------------------------------------------ helpers
// Device buffers come from a process-wide cache of freed buffers (per device, per size class): a store opens and closes one plan per
// request (the reference builds its ExecNode tree per request, src/store/region.cpp:3072), and cudaMalloc / cudaFree of the 40 MB group
// table cost milliseconds and a device-wide synchronisation each.  A buffer enters the cache only after the owning plan's streams have
// drained.  BKGPU_NO_ALLOC_CACHE=1 turns the cache off; bkgpu_release_cache() returns its memory to the driver.
namespace {
struct DevCache {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> free_;
    size_t held = 0;
    const size_t limit = (size_t)8 << 30;
    const bool off = getenv("BKGPU_NO_ALLOC_CACHE") && atoi(getenv("BKGPU_NO_ALLOC_CACHE")) != 0;
    void* take(int device, size_t sc) {
        std::lock_guard<std::mutex> g(mu);
        auto it = free_.find({device, sc});
        if (it == free_.end()) return nullptr;
        void* q = it->second; free_.erase(it); held -= sc;
        return q;
    }
    bool put(int device, size_t sc, void* q) {
        std::lock_guard<std::mutex> g(mu);
        if (off || held + sc > limit) return false;
        free_.insert({{device, sc}, q}); held += sc;
        return true;
    }
    void release_all() {   // (the caller has the right device current for none of them: cudaFree takes any device's pointer)
        std::lock_guard<std::mutex> g(mu);
        for (auto& kv : free_) cudaFree(kv.second);
        free_.clear(); held = 0;
    }
};
DevCache& dev_cache() { static DevCache c; return c; }
size_t size_class(size_t b) {   // eight classes per power of two: at most 12.5 % slack
    if (b < 256) return 256;
    size_t p2 = 1; while (p2 < b) p2 <<= 1;
    const size_t step = p2 >> 3;
    return (b + step - 1) / step * step;
}
}  // namespace
extern "C" void bkgpu_release_cache(void) { dev_cache().release_all(); }

static int dev_alloc(bkgpu_plan* p, void** out, size_t bytes) {
    const size_t sc = size_class(bytes ? bytes : 8);
    void* q = dev_cache().take(p->device, sc);
    if (!q) {
        cudaError_t e = cudaMalloc(&q, sc);
        if (e != cudaSuccess) { cudaGetLastError(); dev_cache().release_all(); e = cudaMalloc(&q, sc); }   // memory held by the cache is memory the device has
        if (e != cudaSuccess) return p->cuda_fail(e, "cudaMalloc");
    }
    *out = q;
    p->dev_allocs.push_back({q, sc});
    return BKGPU_OK;
}
static void dev_free(bkgpu_plan* p, void* ptr) {
    if (!ptr) return;
    size_t sc = 0;
    for (auto it = p->dev_allocs.begin(); it != p->dev_allocs.end(); ++it) if (it->first == ptr) { sc = it->second; p->dev_allocs.erase(it); break; }
    if (sc) {   // work queued on this plan's streams may still use the buffer
        if (p->stream) cudaStreamSynchronize(p->stream);
        if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
        if (dev_cache().put(p->device, sc, ptr)) return;
    }
    cudaFree(ptr);
}
// grow-only device buffer: plans are re-armed with bkgpu_reset and run again; cudaMalloc / cudaFree per run would
// serialise the device every time
static int ensure_buf(bkgpu_plan* p, void** ptr, size_t* cap, size_t bytes) {
    if (*ptr && *cap >= bytes) return BKGPU_OK;
    dev_free(p, *ptr); *ptr = nullptr; *cap = 0;
    const int rc = dev_alloc(p, ptr, bytes);
    if (rc == BKGPU_OK) *cap = bytes;
    return rc;
}
static EventPair* timer_begin(bkgpu_plan* p, std::vector<EventPair>& v, int64_t bytes) {
    EventPair ep{};
    for (cudaEvent_t* e : {&ep.a, &ep.b}) {   // events are recycled: a re-armed plan launches the same kernels again and again
        if (!p->event_pool.empty()) { *e = p->event_pool.back(); p->event_pool.pop_back(); }
        else if (cudaEventCreate(e) != cudaSuccess) return nullptr;
    }
    ep.bytes = bytes;
    cudaEventRecord(ep.a, p->stream);
    v.push_back(ep);
    return &v.back();
}
static void timer_end(bkgpu_plan* p, EventPair* ep) { if (ep) cudaEventRecord(ep->b, p->stream); }

// ------------------------------------------------------------------ lifecycle
extern "C" int bkgpu_init(bkgpu_plan** out, const uint8_t* desc, size_t len, int device, void* nccl_comm) {
    if (!out) return thread_fail(BKGPU_EINVAL, "bkgpu_init: out is NULL");
    *out = nullptr;
    bkgpu_plan* p = new bkgpu_plan();
    std::string err;
    int rc = compile_plan(desc, len, p->c, err);
    if (rc != BKGPU_OK) { g_thread_error = err; delete p; return rc; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        g_thread_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (this library has no CPU fallback)";
        delete p; return BKGPU_ENODEV;
    }
    if (device < 0 || device >= ndev) { g_thread_error = "bkgpu_init: bad device index"; delete p; return BKGPU_EINVAL; }
    p->device = device;
    p->nccl_comm = nccl_comm;
    if (nccl_comm) {
        int n = 0;
        if (nccl_comm_count(nccl_comm, &n) != 0 || n < 1) { g_thread_error = std::string("NCCL: ") + nccl_last_error(); delete p; return BKGPU_ENCCL; }
        p->nranks = n;
    }
    { int n = 0; if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && n > 0) p->sm_count = n; }   // (cudaGetDeviceProperties costs milliseconds per call)
    *out = p;
    return BKGPU_OK;
}

extern "C" int bkgpu_set_option(bkgpu_plan* p, const char* key, int64_t v) {
    if (!p || !key) return thread_fail(BKGPU_EINVAL, "bkgpu_set_option: NULL argument");
    if (p->state != S_INIT) return p->fail(BKGPU_ESTATE, "options must be set before bkgpu_open");
    std::string k = key;
    if (k == "stream") { p->stream = (cudaStream_t)(uintptr_t)v; p->own_stream = false; }
    else if (k == "group_capacity_log2") { if (v < 4 || v > 30) return p->fail(BKGPU_EINVAL, "group_capacity_log2 out of range"); p->group_cap_log2 = (int)v; }
    else if (k == "smem_capacity_log2") { if (v < -1 || v > 13) return p->fail(BKGPU_EINVAL, "smem_capacity_log2 out of range"); p->smem_cap_log2 = (int)v; }
    else if (k == "batch_capacity") { if (v < 1) return p->fail(BKGPU_EINVAL, "batch_capacity must be positive"); p->batch_capacity = v; }
    else if (k == "chunk_rows") { if (v < 1024) return p->fail(BKGPU_EINVAL, "chunk_rows too small"); p->chunk_rows = (v + 7) & ~7ll; }
    else if (k == "partial_capacity") { if (v < 1) return p->fail(BKGPU_EINVAL, "partial_capacity must be positive"); p->partial_cap = v; }
    else if (k == "force_generic") p->force_generic = v != 0;
    else if (k == "no_stream_copy") p->no_stream_copy = v != 0;
    else if (k == "join_pipeline") p->join_pipeline = v != 0;
    else if (k == "lean_bank") p->lean_bank = v != 0;
    else if (k == "lean_fx") p->lean_fx = v != 0;
    else if (k == "blocking_sync") p->blocking_sync = (int)v;
    else if (k == "join_learn_range") p->join_learn_range = v != 0;
    else if (k == "no_lean") p->no_lean = v != 0;
    else if (k == "no_fused_probe") p->no_fused_probe = v != 0;
    else if (k == "repartition") p->repartition = v != 0;
    else if (k == "peer_merge") p->peer_merge = v != 0;
    else if (k == "no_lean_nulls") p->no_lean_nulls = v != 0;
    else if (k == "no_lean_mm") p->no_lean_mm = v != 0;
    else if (k == "use_wp") p->use_wp = v != 0;
    else if (k == "no_bounce") p->no_bounce = v != 0;
    else if (k == "scalar_tma") p->scalar_tma = v != 0;
    else if (k == "wp_warps") { if (v != 0 && v != 8 && v != 12 && v != 16) return p->fail(BKGPU_EINVAL, "wp_warps: 0, 8, 12 or 16"); p->wp_warps = (int)v; }
    else if (k == "wp_kt_log2") { if (v < 0 || v > 14) return p->fail(BKGPU_EINVAL, "wp_kt_log2 out of range"); p->wp_kt_log2 = (int)v; }
    else if (k == "output_on_device") p->output_on_device = v != 0;
    else if (k == "region_base") p->region_base = v;
    else return p->fail(BKGPU_EINVAL, "unknown option '%s'", key);
    return BKGPU_OK;
}

static int alloc_group_table(bkgpu_plan* p) {
    const AggPlan& ap = p->c.ap;
    GroupTable& gt = p->gt;
    int log2 = ap.n_keyw == 0 ? 0 : p->group_cap_log2;
    size_t cap = (size_t)1 << log2;
    gt.cap_mask = (uint32_t)(cap - 1); gt.cap_log2 = (uint32_t)log2;
    int rc;
    if ((rc = dev_alloc(p, (void**)&gt.state, cap * 4))) return rc;
    if ((rc = dev_alloc(p, (void**)&gt.keys, cap * 8 * (size_t)std::max(ap.n_keyw, 1)))) return rc;
    if ((rc = dev_alloc(p, (void**)&gt.lanes, cap * 8 * (size_t)ap.n_lanes))) return rc;
    if ((rc = dev_alloc(p, (void**)&gt.n_groups, 4 * ((size_t)GT_OCC_OFF + cap)))) return rc;   // counter, overflow flag, occupied list
    gt.overflow = gt.n_groups + 1;
    p->d_cursor = gt.n_groups + 2;                       // [2] result cursor [3] merge info
    p->d_rows_passed = (uint64_t*)(gt.n_groups + 4);     // [4..5] rows that passed the filter
    CK(p, cudaMemsetAsync(gt.n_groups, 0, 4 * GT_OCC_OFF, p->stream));
    CK(p, launch_table_init(gt, ap, p->stream));
    p->stats.kernel_launches++;
    return BKGPU_OK;
}

namespace { int cpu_budget(); }   // CPUs this process may use (affinity, cgroup quota): defined with the copy pool

extern "C" int bkgpu_open(bkgpu_plan* p) {
    if (!p) return thread_fail(BKGPU_EINVAL, "bkgpu_open: NULL plan");
    if (p->state != S_INIT) return p->fail(BKGPU_ESTATE, "bkgpu_open called twice");
    CK(p, cudaSetDevice(p->device));
    if (!p->stream) { CK(p, cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking)); p->own_stream = true; }
    CK(p, cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
    if (cudaHostAlloc((void**)&p->h_pinned, 64, cudaHostAllocDefault) == cudaSuccess) memset(p->h_pinned, 0, 64); else p->h_pinned = nullptr;
    for (int i = 0; i < 2; i++) {
        CK(p, cudaEventCreateWithFlags(&p->stage_free[i], cudaEventDisableTiming));
        CK(p, cudaEventCreateWithFlags(&p->stage_ready[i], cudaEventDisableTiming));
    }
    p->blocking_wait = p->blocking_sync >= 0 ? p->blocking_sync != 0 : (p->nranks > 1 && cpu_budget() < 4 * p->nranks);
    int rc = BKGPU_OK;
    if (p->c.kind == PK_AGG || p->c.kind == PK_JOIN_AGG) rc = alloc_group_table(p);
    if (!rc && (p->c.kind == PK_SORT || p->c.kind == PK_FILTER)) rc = sort_open(p->c, p->device, p->stream, p->region_base, &p->sort, p->last_error);
    if (!rc && p->c.post) rc = sort_open(*p->c.post, p->device, p->stream, 0, &p->post_sort, p->last_error);
    if (rc) { g_thread_error = p->last_error; return rc; }
    p->state = S_OPEN;
    return BKGPU_OK;
}

// choose the shared-table capacity for this batch: >= 2x the groups seen so far, 2048 slots when
// nothing is known yet, 0 (straight to the global table) when the groups cannot fit one CTA's 227 KB
static int pick_smem_log2(bkgpu_plan* p, int n_smem_lanes, bool direct, int na) {
    if (p->c.ap.n_keyw == 0) return 0;
    const size_t budget = 220 * 1024;
    auto bytes = [&](int log2) { return direct ? direct_smem_bytes(p->c.ap.n_keyw, n_smem_lanes, log2, na) : agg_smem_bytes(p->c.ap.n_keyw, n_smem_lanes, log2); };
    if (p->smem_cap_log2 >= 0) { int l = p->smem_cap_log2; while (l > 0 && bytes(l) > budget) l--; return l; }
    if (p->h_pinned && p->h_pinned[0] > p->known_groups) p->known_groups = p->h_pinned[0];
    uint32_t g = p->known_groups;
    int log2 = 11;
    while (((uint32_t)1 << log2) < 2 * g && log2 < 14) log2++;
    while (log2 > 0 && bytes(log2) > budget) log2--;
    if (((uint32_t)1 << log2) < g) return 0;  // would mostly miss: skip the shared table
    return log2;
}

// `jp`: fused FK -> PK probe (lean kernel only); returns +1 without launching when this batch does not fit the lean kernel
static int launch_agg_batch(bkgpu_plan* p, const Compiled& c, const DevCol* cols, int64_t nrows, bool vec_ok, const JoinProbe* jp = nullptr) {
    AggArgs a; memset(&a, 0, sizeof a);
    a.plan = c.ap; a.prog = c.prog; a.direct = c.direct; a.gt = p->gt; a.rows_passed = p->d_rows_passed;
    if (c.kind == PK_JOIN_AGG) {
        a.join.enabled = 1; a.join.keys = p->jt_keys; a.join.rows = p->jt_rows; a.join.cap_mask = p->jt_mask;
        a.join.probe_col = c.probe_key_col; a.join.probe_prim = c.cols[(size_t)c.probe_key_col].prim; a.join.cast_prim = c.join_key_prim;
        a.join.join_type = c.join_type; a.join.matched = c.join_type == BK_INNER_JOIN ? nullptr : p->jb_matched; a.join.n_build = p->jb_rows;
        a.join.tail = p->join_tail_launch ? 1 : 0;
    }
    const bool direct = c.has_direct && vec_ok && !p->force_generic && c.kind == PK_AGG;   // (the join's fast-path plan is lowered as PK_AGG)
    const int ncols = (int)c.cols.size();
    if (direct) { for (size_t i = 0; i < c.direct_cols.size(); i++) a.cols[i] = cols[c.direct_cols[i]]; a.n_cols = (int)c.direct_cols.size(); }
    else { for (int i = 0; i < ncols; i++) a.cols[i] = cols[i]; a.n_cols = ncols; }
    // per-batch nullability -> which lanes need per-row work in the shared table
    memset(a.smem_lane, 0xFF, sizeof a.smem_lane);
    a.smem_lane[0] = 0; a.n_smem_lanes = 1; a.alias_mask = 0;
    for (int k = 0; k < a.plan.n_agg; k++) {
        AggSpec& s = a.plan.agg[k];
        if (s.kind == AG_COUNT_STAR) continue;
        bool nullable = c.arg_can_null[(size_t)k] || p->join_tail_launch;   // (the tail's probe-side columns are NULL on every row)
        for (int i = 0; i < ncols; i++) if ((c.arg_cols_mask[(size_t)k] >> i) & 1) nullable = nullable || cols[i].validity != nullptr;
        s.nullable = nullable ? 1 : 0;
        if (s.cnt_lane && s.cnt_owner) {
            if (nullable) { if (a.smem_lane[s.cnt_lane] == 0xFF) a.smem_lane[s.cnt_lane] = (uint8_t)a.n_smem_lanes++; }
            else a.alias_mask |= 1u << s.cnt_lane;
        }
        if (s.kind != AG_COUNT && a.smem_lane[s.acc_lane] == 0xFF) a.smem_lane[s.acc_lane] = (uint8_t)a.n_smem_lanes++;
    }
    for (int l = 1; l < a.plan.n_lanes; l++) if (a.smem_lane[l] != 0xFF) a.alias_mask &= ~(1u << l);
    a.smem_cap_log2 = pick_smem_log2(p, a.n_smem_lanes, direct, c.direct.n_vals);
    a.smem_sentinel = direct && a.plan.n_keyw == 1 ? 1 : 0;
    a.smem_keyw = a.plan.n_keyw;

    if (direct) {  // per value column: the lane operations it feeds
        memset(a.vops, 0, sizeof a.vops);
        for (int v = 0; v < c.direct.n_vals; v++) { a.vops[v].cnt_smem = 0xFF; }
        for (int k = 0; k < a.plan.n_agg; k++) {
            const AggSpec& s = a.plan.agg[k];
            if (s.kind == AG_COUNT_STAR) continue;
            ValOps& vo = a.vops[c.direct.agg_val[k]];
            vo.arg_class = s.arg_vclass;
            if (s.cnt_lane && s.cnt_owner) { vo.cnt_glob = s.cnt_lane; vo.cnt_smem = s.nullable ? a.smem_lane[s.cnt_lane] : 0xFF; }
            if (s.kind != AG_COUNT && s.acc_owner) {
                const int n = vo.n_ops++;
                vo.op[n] = a.plan.lane_op[s.acc_lane]; vo.lane_class[n] = s.vclass; vo.glob_lane[n] = s.acc_lane; vo.smem_lane[n] = a.smem_lane[s.acc_lane];
            }
        }
    }
    a.lean = 0;
    // lean shape: one key word, or a 64-bit key whose second word only carries the (here unused) NULL flag
    const bool lean_keyw = a.plan.n_keyw == 1 || (a.plan.n_keyw == 2 && a.plan.n_group == 1 && a.plan.key_bits[0] == 64 && a.plan.key_null_word[0] == 1);
    if (direct && lean_keyw && a.smem_cap_log2 > 0 && !p->no_lean) {  // does this batch fit the lean kernel?
        bool ok = true;
        const int np = c.direct.n_terms, na = c.direct.n_vals;
        bool any_valid = false;
        for (int i = 0; i < a.n_cols; i++) if (a.cols[i].validity) any_valid = true;
        if (a.cols[np].validity || (any_valid && (jp || p->no_lean_nulls))) ok = false;   // NULL keys (and NULLs under the fused probe) take the general kernel
        for (int t = 0; t < np && ok; t++) {
            const DirectTerm& tm = c.direct.term[t];
            const int64_t cv = (int64_t)tm.cbits;
            ok = a.cols[t].stype == ST_I32 && a.cols[t].prim == BK_INT32 && tm.vclass == VC_I64 && cv >= INT32_MIN && cv <= INT32_MAX;
        }
        const DevCol& kc = a.cols[np];
        ok = ok && ((kc.stype == ST_I32 && kc.prim == BK_INT32) || (kc.stype == ST_U32 && kc.prim == BK_UINT32) || kc.stype == ST_I64 || kc.stype == ST_U64);
        bool mm = false;   // some column feeds MIN / MAX or several lanes: the MM instantiation of the lean kernel
        for (int v = 0; v < na && ok; v++) {
            const DevCol& vc = a.cols[np + 1 + v];
            const ValOps& vo = a.vops[v];
            ok = (vc.stype == ST_F64 || vc.stype == ST_I64 || vc.stype == ST_U64) && vo.n_ops <= 3 && (vo.cnt_smem == 0xFF || vo.cnt_glob != 0);
            const int col_class = vc.stype == ST_F64 ? VC_F64 : (vc.stype == ST_U64 ? VC_U64 : VC_I64);
            for (int k = 0; k < vo.n_ops && ok; k++) {
                const int op = vo.op[k];
                if (op == LN_ADD_F64) ok = vc.stype == ST_F64 && vo.lane_class[k] == VC_F64;
                else if (op == LN_ADD_I64) ok = vc.stype != ST_F64;
                else { ok = vo.lane_class[k] == col_class && vo.arg_class == col_class; mm = true; }   // MIN / MAX in the column's own class
            }
            if (vo.n_ops != 1) mm = true;
        }
        if (mm && (jp || p->no_lean_mm)) ok = false;
        a.lean = ok ? 1 : 0;
        if (a.lean) {
            a.smem_sentinel = 1; a.smem_keyw = 1; a.smem_paired = 1;
            // lane placement for 128-bit shared CAS: ONE double sum shares a 16-byte word with the row count
            // ({rows, sum}: a single atomic per row); otherwise double sums sit in pairs {sumA, sumB}
            memset(a.smem_lane, 0xFF, sizeof a.smem_lane);
            a.smem_lane[0] = 0;
            int nf64 = 0;   // columns that feed exactly one double sum (they sit in {sumA, sumB} pairs)
            for (int v = 0; v < na; v++) if (a.vops[v].n_ops == 1 && a.vops[v].op[0] == LN_ADD_F64) nf64++;
            int next_f = nf64 == 1 ? 1 : 2, next_i = nf64 == 1 ? 2 : 2 + ((nf64 + 1) & ~1);
            for (int v = 0; v < na; v++) {
                ValOps& vo = a.vops[v];
                for (int k = 0; k < vo.n_ops; k++) {
                    const int sl = (vo.n_ops == 1 && vo.op[0] == LN_ADD_F64) ? next_f++ : next_i++;
                    vo.smem_lane[k] = (uint8_t)sl; a.smem_lane[vo.glob_lane[k]] = (uint8_t)sl;
                }
            }
            a.lean_mm = mm ? 1 : 0;
            int next = std::max(next_f, next_i);
            for (int v = 0; v < na; v++) {   // non-NULL counters of the value columns that carry NULLs in this batch
                ValOps& vo = a.vops[v];
                if (vo.cnt_smem == 0xFF) continue;
                vo.cnt_smem = (uint8_t)next; a.smem_lane[vo.cnt_glob] = (uint8_t)next; next++;
            }
            a.lean_nulls = any_valid ? 1 : 0;
            a.n_smem_lanes = (next + 1) & ~1;
            if (a.n_smem_lanes < 2) a.n_smem_lanes = 2;
            a.smem_cap_log2 = pick_smem_log2(p, a.n_smem_lanes, true, na);
            if (a.smem_cap_log2 <= 0) { a.lean = 0; a.smem_paired = 0; return p->fail(BKGPU_ENOMEM, "lean kernel: shared table does not fit"); }
            // FX: plain batches (no NULLs, no MIN / MAX) with at least one double sum, when the extension limbs fit beside table and queues
            if (p->lean_fx && !p->fx_off && !p->lean_bank && !mm && !any_valid && nf64 >= 1 && nrows < ((int64_t)1 << 31)) {
                const size_t base = direct_smem_bytes(a.smem_keyw, a.n_smem_lanes, a.smem_cap_log2, na);
                if (base + fx_ext_bytes(na, a.smem_cap_log2) <= (size_t)224 * 1024) { a.lean_fx = 1; a.fx_ext_off = (uint32_t)base; }   // (227 KB per CTA less the kernel's 1 KB of static shared memory)
            }

        }
    }
    if (jp) { if (!a.lean) return 1; a.jp = *jp; a.jp_pipeline = p->join_pipeline; }
    // warp-private tables (agg_wp.cuh): the plainest lean batches whose groups fit one table per warp.  The capacity follows the
    // cardinality learned from earlier batches / runs of this plan; an unknown cardinality starts with the largest table.
    a.scalar_tma = p->scalar_tma; a.lean_bank = p->lean_bank;
    a.wp = 0;
    if (a.lean && !a.lean_nulls && !a.lean_mm && p->use_wp && c.direct.n_vals <= 2) {
        const int np = c.direct.n_terms, na = c.direct.n_vals;
        const DevCol& kc = a.cols[np];
        bool ok = kc.stype == ST_I32 || kc.stype == ST_U32;
        for (int v = 0; v < na && ok; v++) ok = a.vops[v].n_ops == 1 && (a.vops[v].op[0] == LN_ADD_F64 || a.vops[v].op[0] == LN_ADD_I64);
        if (p->h_pinned && p->h_pinned[0] > p->known_groups) p->known_groups = p->h_pinned[0];
        if (ok) {
            const size_t budget = 226 * 1024;
            const uint32_t g = p->known_groups;
            int warps = 8;
            if (na <= 1 && g) {   // one value column: the tables are small enough for 12 or 16 warps (three / four per scheduler)
                const uint32_t gc = (g + 31) & ~31u;
                int kl4 = 8; while ((1u << kl4) < 4 * gc) kl4++;
                if (wp_smem_bytes(na, gc, kl4, 16) <= budget) warps = 16;
                else if (wp_smem_bytes(na, gc, kl4, 12) <= budget) warps = 12;
            }
            if (p->wp_warps && (na <= 1 || p->wp_warps == 8)) warps = p->wp_warps;
            uint32_t gcap = g ? ((g + 31) & ~31u) : 0;
            auto fits = [&](uint32_t gc, int kl) { return wp_smem_bytes(na, gc, kl, warps) <= budget; };
            auto kt_for = [&](uint32_t gc, int factor) { int kl = 8; while ((1u << kl) < (uint32_t)factor * gc) kl++; return kl; };
            if (gcap == 0) { gcap = 32; while (fits(gcap + 32, kt_for(gcap + 32, 2))) gcap += 32; }   // unknown: the largest table that fits
            int kl = p->wp_kt_log2 > 0 ? p->wp_kt_log2 : kt_for(gcap, 4);          // key table at <= 25 % load when it fits ...
            if (!fits(gcap, kl)) kl = kt_for(gcap, 2);                              // ... else <= 50 %
            if (fits(gcap, kl)) { a.wp = 1; a.wp_gcap = (int)gcap; a.wp_kt_log2 = kl; a.wp_warps = warps; a.wp_dense = 0; a.wp_dense_sub = 0; }
        }
    }
    const int64_t kMax = (int64_t)1 << 30;  // rows per launch (32-bit counters inside a CTA)
    int64_t algo_bytes_per_row = 0;
    for (int i = 0; i < a.n_cols; i++) algo_bytes_per_row += storage_bytes(a.cols[i].stype);
    for (int64_t off = 0; off < nrows; off += kMax) {
        AggArgs b = a;
        b.nrows = std::min(kMax, nrows - off);
        for (int i = 0; i < b.n_cols; i++) {
            b.cols[i].values = (const uint8_t*)a.cols[i].values + (size_t)off * storage_bytes(a.cols[i].stype);
            if (a.cols[i].validity) b.cols[i].validity = a.cols[i].validity + off / 8;
        }
        const char* name = "";
        EventPair* ep = timer_begin(p, p->timed, b.nrows * algo_bytes_per_row);
        cudaError_t e = launch_agg(b, direct, p->sm_count, p->stream, &name);
        timer_end(p, ep);
        if (e != cudaSuccess) return p->cuda_fail(e, "launch_agg");
        snprintf(p->stats.main_kernel_name, sizeof p->stats.main_kernel_name, "%s", b.wp && direct ? "k_agg_group_wp" : (b.lean && direct ? (b.lean_fx ? "k_agg_group_lean_fx" : "k_agg_group_lean") : name));
        p->stats.kernel_launches++;
    }
    return BKGPU_OK;
}

// resolve the plan's columns against one pushed batch
static int bind_columns(bkgpu_plan* p, const std::vector<ColRef>& want, const bkgpu_column* cols, int ncols, int64_t nrows,
                        std::vector<const bkgpu_column*>& bound) {
    bound.assign(want.size(), nullptr);
    for (size_t i = 0; i < want.size(); i++) {
        for (int k = 0; k < ncols; k++) if (cols[k].tuple_id == want[i].tuple_id && cols[k].slot_id == want[i].slot_id) { bound[i] = &cols[k]; break; }
        if (!bound[i]) return p->fail(BKGPU_EINVAL, "batch lacks column %d_%d", want[i].tuple_id, want[i].slot_id);
        if (bound[i]->length != nrows) return p->fail(BKGPU_EINVAL, "column %d_%d has length %lld, batch has %lld rows", want[i].tuple_id,
                                                      want[i].slot_id, (long long)bound[i]->length, (long long)nrows);
        if (nrows > 0 && !bound[i]->values) return p->fail(BKGPU_EINVAL, "column %d_%d has no values buffer", want[i].tuple_id, want[i].slot_id);
        if (prim_storage(bound[i]->prim_type) != prim_storage(want[i].prim))
            return p->fail(BKGPU_EINVAL, "column %d_%d arrives as type %d but the plan declares %d", want[i].tuple_id, want[i].slot_id,
                           bound[i]->prim_type, want[i].prim);
    }
    return BKGPU_OK;
}

static int ensure_stage(bkgpu_plan* p, const std::vector<ColRef>& want, size_t rows) {
    if (p->stage_rows >= rows && p->stage[0].size() == want.size()) return BKGPU_OK;
    for (int b = 0; b < 2; b++) {
        for (void* q : p->stage[b]) dev_free(p, q);
        for (void* q : p->stage_valid[b]) dev_free(p, q);
        p->stage[b].assign(want.size(), nullptr); p->stage_valid[b].assign(want.size(), nullptr);
        for (size_t i = 0; i < want.size(); i++) {
            int rc;
            if ((rc = dev_alloc(p, &p->stage[b][i], rows * (size_t)storage_bytes(prim_storage(want[i].prim))))) return rc;
            if ((rc = dev_alloc(p, &p->stage_valid[b][i], rows / 8 + 8))) return rc;
        }
    }
    p->stage_rows = rows;
    return BKGPU_OK;
}

// pinned bounce buffers for pageable input: one per staging buffer set and column
static int ensure_bounce(bkgpu_plan* p, const std::vector<ColRef>& want, size_t rows) {
    if (p->bounce_rows >= rows && p->bounce[0].size() >= want.size()) return BKGPU_OK;
    for (int b = 0; b < 2; b++) {
        for (uint8_t* q : p->bounce[b]) if (q) cudaFreeHost(q);
        p->bounce[b].assign(want.size(), nullptr);
        for (size_t i = 0; i < want.size(); i++) {
            const size_t bytes = rows * (size_t)storage_bytes(prim_storage(want[i].prim));
            if (cudaHostAlloc((void**)&p->bounce[b][i], bytes ? bytes : 8, cudaHostAllocDefault) != cudaSuccess) return p->fail(BKGPU_ENOMEM, "pinned bounce buffer of %zu bytes", bytes);
        }
        if (!p->bounce_done[b]) CK(p, cudaEventCreateWithFlags(&p->bounce_done[b], cudaEventDisableTiming));
        p->bounce_busy[b] = false;
    }
    p->bounce_rows = rows;
    return BKGPU_OK;
}

typedef int (*BatchFn)(bkgpu_plan*, const DevCol*, int64_t nrows, int64_t row_base, bool vec_ok);

// ---- host-side copy workers: pageable input (what an Arrow RecordBatch of RocksdbVectorizedReader hands over) is copied into
// pinned bounce buffers by several threads — cudaMemcpyAsync from pageable memory goes through the driver's single staging buffer
// at a fraction of the link rate ----
namespace bk { void stream_copy(void* dst, const void* src, size_t bytes); }   // hostcopy.cpp
namespace {
class CopyPool {
  public:
    explicit CopyPool(int n, double poll_ms) : poll_ms_(poll_ms) { for (int i = 0; i < n; i++) th_.emplace_back([this] { run(); }); }
    ~CopyPool() { { std::lock_guard<std::mutex> g(mu_); stop_ = true; stop_pub_.store(true, std::memory_order_release); } cv_.notify_all(); for (auto& t : th_) t.join(); }
    int size() const { return (int)th_.size(); }
    // runs fn(0) .. fn(n - 1) on the workers and the caller; returns when all are done
    void parallel(int n, const std::function<void(int)>& fn) {
        if (n <= 0) return;
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &fn; next_ = 0; total_ = n; done_ = 0; done_pub_.store(0, std::memory_order_release); gen_++;
            gen_pub_.store(gen_, std::memory_order_release);
        }
        cv_.notify_all();
        work();
        if (poll_ms_ > 0)
            for (const double t_end = HostClock::now() + 5.0; done_pub_.load(std::memory_order_acquire) < n && HostClock::now() < t_end;)   // (the last slices finish within microseconds)
                for (int i = 0; i < 32; i++) __builtin_ia32_pause();
        std::unique_lock<std::mutex> g(mu_);
        done_cv_.wait(g, [this] { return done_ == total_; });
        fn_ = nullptr;
    }
  private:
    void work() {
        for (;;) {
            int i;
            const std::function<void(int)>* f;
            { std::lock_guard<std::mutex> g(mu_); if (!fn_ || next_ >= total_) return; i = next_++; f = fn_; }
            (*f)(i);
            { std::lock_guard<std::mutex> g(mu_); ++done_; done_pub_.store(done_, std::memory_order_release); if (done_ == total_) done_cv_.notify_all(); }
        }
    }
    void run() {
        uint64_t seen = 0;
        for (;;) {
            // a push hands over one chunk per millisecond: a worker that went to sleep on the condition variable after each chunk paid a futex
            // wake-up (tens of microseconds, staggered over the threads) on a ~0.3 ms copy — it polls for the next chunk for a short while first
            const double t_end = HostClock::now() + poll_ms_;
            bool got = false;
            while (poll_ms_ > 0 && HostClock::now() < t_end) {
                if (gen_pub_.load(std::memory_order_acquire) != seen || stop_pub_.load(std::memory_order_acquire)) { got = true; break; }
                for (int i = 0; i < 64; i++) __builtin_ia32_pause();
            }
            {
                std::unique_lock<std::mutex> g(mu_);
                if (!got) cv_.wait(g, [&] { return stop_ || gen_ != seen; });
                if (stop_) return;
                seen = gen_;
            }
            work();
        }
    }
    std::vector<std::thread> th_;
    std::mutex mu_; std::condition_variable cv_, done_cv_;
    const std::function<void(int)>* fn_ = nullptr;
    int next_ = 0, total_ = 0, done_ = 0; uint64_t gen_ = 0; bool stop_ = false;
    const double poll_ms_;   // > 0: a worker polls this long for the next chunk before it sleeps on the condition variable (BKGPU_COPY_POLL_US)
    std::atomic<uint64_t> gen_pub_{0}; std::atomic<bool> stop_pub_{false}; std::atomic<int> done_pub_{0};   // what the pollers read without the mutex
};
// CPUs this process may actually use: online CPUs, its affinity mask and its cgroup CPU quota (a container with a 16-CPU quota on a 128-CPU box
// gets throttled for the rest of the scheduling period once its threads have burnt the quota — spinning or oversized pools make the copy SLOWER there)
int cpu_budget() {
    int n = (int)std::thread::hardware_concurrency();
    cpu_set_t set; CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0) { const int a = CPU_COUNT(&set); if (a > 0 && a < n) n = a; }
    if (FILE* f = fopen("/sys/fs/cgroup/cpu.max", "r")) {   // cgroup v2: "<quota|max> <period>"
        char q[32]; long long period = 0;
        if (fscanf(f, "%31s %lld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) { const long long c = atoll(q) / period; if (c > 0 && c < n) n = (int)c; }
        fclose(f);
    } else if (FILE* g = fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) {   // cgroup v1
        long long quota = -1, period = 100000;
        if (fscanf(g, "%lld", &quota) != 1) quota = -1;
        fclose(g);
        if (FILE* h = fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (fscanf(h, "%lld", &period) != 1) period = 100000; fclose(h); }
        if (quota > 0 && period > 0 && quota / period > 0 && quota / period < n) n = (int)(quota / period);
    }
    return n > 0 ? n : 1;
}
CopyPool& copy_pool() {
    static CopyPool pool([] {
        if (const char* e = getenv("BKGPU_COPY_THREADS")) { const int v = atoi(e); if (v > 0) return std::min(v, 64) - 1; }
        return std::max(1, std::min(23, cpu_budget() * 3 / 4 - 1));   // workers + the pushing thread = 3/4 of the budget: measured best under a 16-CPU quota
                                                                       // (12 threads 0.895 of the pinned rate; 8: 0.82, 16: 0.85, 24: 0.81), 24 threads on an unconstrained host
    }(), [] { const char* e = getenv("BKGPU_COPY_POLL_US"); return e ? atof(e) / 1000.0 : 0.0; }());
    return pool;
}
bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
}  // namespace

// feed a batch: device-resident columns go straight to the kernels; host columns stream through
// two sets of device staging buffers (H2D of chunk i+1 overlaps the kernel on chunk i)
static int feed(bkgpu_plan* p, const std::vector<ColRef>& want, const bkgpu_column* cols, int ncols, int64_t nrows, int on_device, BatchFn fn) {
    std::vector<const bkgpu_column*> bound;
    int rc = bind_columns(p, want, cols, ncols, nrows, bound);
    if (rc) return rc;
    std::vector<DevCol> dc(std::max<size_t>(want.size(), 1));
    if (on_device) {
        bool vec_ok = true;
        for (size_t i = 0; i < want.size(); i++) {
            dc[i].values = bound[i]->values; dc[i].validity = bound[i]->validity;
            dc[i].stype = prim_storage(want[i].prim); dc[i].prim = want[i].prim;
            if (((uintptr_t)dc[i].values & 31) != 0) vec_ok = false;
        }
        return fn(p, dc.data(), nrows, 0, vec_ok);
    }
    bool pageable = false;
    for (size_t i = 0; i < want.size() && !p->no_bounce; i++) pageable = pageable || is_pageable(bound[i]->values);
    // (pageable input adds a CPU copy stage of about the link's speed in front of the H2D copy: smaller chunks keep the fill / drain
    //  of that three-stage pipeline at ~2 % of the batch instead of ~8 %)
    const int64_t chunk_cap = pageable ? std::min<int64_t>(p->chunk_rows, (int64_t)2 << 20) : p->chunk_rows;
    const int64_t chunk = std::min<int64_t>(chunk_cap, std::max<int64_t>((nrows + 7) & ~7ll, 8));
    if ((rc = ensure_stage(p, want, (size_t)chunk))) return rc;
    if (pageable && (rc = ensure_bounce(p, want, (size_t)chunk))) return rc;
    int buf = 0;
    cudaEvent_t last_ready = nullptr;
    for (int64_t off = 0; off < nrows; off += chunk, buf ^= 1) {
        if (p->cancelled.load()) return p->fail(BKGPU_ECANCELLED, "cancelled");
        const int64_t n = std::min(chunk, nrows - off);
        CK(p, cudaStreamWaitEvent(p->copy_stream, p->stage_free[buf], 0));
        if (pageable) {   // bounce: worker threads fill this buffer set's pinned copy of the chunk, the link then runs at the pinned rate
            if (p->bounce_busy[buf]) { CK(p, cudaEventSynchronize(p->bounce_done[buf])); p->bounce_busy[buf] = false; }
            struct Piece { uint8_t* dst; const uint8_t* src; size_t bytes; };
            std::vector<Piece> pieces;
            const size_t slice = (size_t)1 << 20;
            for (size_t i = 0; i < want.size(); i++) {
                const size_t eb = (size_t)storage_bytes(prim_storage(want[i].prim)), bytes = (size_t)n * eb;
                const uint8_t* src = (const uint8_t*)bound[i]->values + (size_t)off * eb;
                for (size_t o = 0; o < bytes; o += slice) pieces.push_back({p->bounce[buf][i] + o, src + o, std::min(slice, bytes - o)});
            }
            const bool nt = !p->no_stream_copy;
            copy_pool().parallel((int)pieces.size(), [&](int k) {
                const Piece& pc = pieces[(size_t)k];
                if (nt) stream_copy(pc.dst, pc.src, pc.bytes); else memcpy(pc.dst, pc.src, pc.bytes);
            });
        }
        for (size_t i = 0; i < want.size(); i++) {
            const int st = prim_storage(want[i].prim);
            const size_t eb = (size_t)storage_bytes(st);
            const void* hsrc = pageable ? (const void*)p->bounce[buf][i] : (const void*)((const uint8_t*)bound[i]->values + (size_t)off * eb);
            CK(p, cudaMemcpyAsync(p->stage[buf][i], hsrc, (size_t)n * eb, cudaMemcpyHostToDevice, p->copy_stream));
            p->stats.h2d_bytes += (int64_t)((size_t)n * eb);
            dc[i].values = p->stage[buf][i]; dc[i].stype = st; dc[i].prim = want[i].prim; dc[i].validity = nullptr;
            if (bound[i]->validity) {
                CK(p, cudaMemcpyAsync(p->stage_valid[buf][i], bound[i]->validity + off / 8, (size_t)(n + 7) / 8, cudaMemcpyHostToDevice, p->copy_stream));
                p->stats.h2d_bytes += (n + 7) / 8;
                dc[i].validity = (const uint8_t*)p->stage_valid[buf][i];
            }
        }
        CK(p, cudaEventRecord(p->stage_ready[buf], p->copy_stream));
        if (pageable) { CK(p, cudaEventRecord(p->bounce_done[buf], p->copy_stream)); p->bounce_busy[buf] = true; }
        CK(p, cudaStreamWaitEvent(p->stream, p->stage_ready[buf], 0));
        if ((rc = fn(p, dc.data(), n, off, true))) return rc;
        CK(p, cudaEventRecord(p->stage_free[buf], p->stream));
        last_ready = p->stage_ready[buf];
    }
    // the caller's host buffers are borrowed for the duration of the call only (bkgpu.h): the last host-to-device copy must have
    // read them before bkgpu_push returns (the kernels read the staging buffers, they may still be running)
    if (last_ready) CK(p, cudaEventSynchronize(last_ready));
    return BKGPU_OK;
}

static int agg_batch(bkgpu_plan* p, const DevCol* cols, int64_t nrows, int64_t, bool vec_ok) { return launch_agg_batch(p, p->c, cols, nrows, vec_ok); }
static int sort_batch(bkgpu_plan* p, const DevCol* cols, int64_t nrows, int64_t, bool) {
    int rc = sort_push(p->sort, cols, nrows, p->stream, &p->stats, p->last_error);
    if (rc) g_thread_error = p->last_error;
    return rc;
}

// ---- K4: retain the build (outer / driver) table, build the multimap at the first probe batch ----
static int join_retain_build(bkgpu_plan* p, const bkgpu_column* cols, int ncols, int64_t nrows, int on_device) {
    const Compiled& c = p->c;
    if (p->jt_built) return p->fail(BKGPU_ESTATE, "build-side rows arrived after probing started (the reference fetches the whole driver table first, join_node.cpp:920-1022)");
    const size_t nc = c.cols.size();
    if (p->jb_vals.empty()) { p->jb_vals.assign(nc, nullptr); p->jb_nullbytes.assign(nc, nullptr); p->jb_bitmap.assign(nc, nullptr); p->jb_has_null.assign(nc, false); }
    std::vector<ColRef> want; std::vector<int> idx;
    for (size_t i = 0; i < nc; i++) if (c.col_side[i] == 1) { want.push_back(c.cols[i]); idx.push_back((int)i); }
    std::vector<const bkgpu_column*> bound;
    int rc = bind_columns(p, want, cols, ncols, nrows, bound);
    if (rc) return rc;
    if (p->jb_rows + nrows > 0xFFFFFFF0ll) return p->fail(BKGPU_EUNSUPPORTED, "build side larger than 2^32 rows");
    if (p->jb_rows + nrows > p->jb_cap) {
        const int64_t cap = std::max<int64_t>(p->jb_rows + nrows, std::max<int64_t>(p->jb_cap * 2, 1 << 16));
        for (int i : idx) {
            const size_t eb = (size_t)storage_bytes(prim_storage(c.cols[(size_t)i].prim));
            uint8_t *nv = nullptr, *nn = nullptr;
            if ((rc = dev_alloc(p, (void**)&nv, (size_t)cap * eb))) return rc;
            if ((rc = dev_alloc(p, (void**)&nn, (size_t)cap))) return rc;
            if (p->jb_rows) {
                CK(p, cudaMemcpyAsync(nv, p->jb_vals[(size_t)i], (size_t)p->jb_rows * eb, cudaMemcpyDeviceToDevice, p->stream));
                CK(p, cudaMemcpyAsync(nn, p->jb_nullbytes[(size_t)i], (size_t)p->jb_rows, cudaMemcpyDeviceToDevice, p->stream));
                CK(p, cudaStreamSynchronize(p->stream));
            }
            dev_free(p, p->jb_vals[(size_t)i]); dev_free(p, p->jb_nullbytes[(size_t)i]);
            p->jb_vals[(size_t)i] = nv; p->jb_nullbytes[(size_t)i] = nn;
        }
        p->jb_cap = cap;
    }
    for (size_t w = 0; w < want.size(); w++) {
        const int i = idx[w];
        const size_t eb = (size_t)storage_bytes(prim_storage(c.cols[(size_t)i].prim));
        uint8_t* dv = p->jb_vals[(size_t)i] + (size_t)p->jb_rows * eb;
        uint8_t* dn = p->jb_nullbytes[(size_t)i] + p->jb_rows;
        if (on_device) {
            CK(p, cudaMemcpyAsync(dv, bound[w]->values, (size_t)nrows * eb, cudaMemcpyDeviceToDevice, p->stream));
            CK(p, launch_unpack_validity(bound[w]->validity, nrows, dn, p->stream));
            p->stats.kernel_launches++;
        } else {
            CK(p, cudaMemcpyAsync(dv, bound[w]->values, (size_t)nrows * eb, cudaMemcpyHostToDevice, p->stream));
            std::vector<uint8_t> nb((size_t)nrows, 0);
            if (bound[w]->validity) for (int64_t r = 0; r < nrows; r++) nb[(size_t)r] = ((bound[w]->validity[r >> 3] >> (r & 7)) & 1) ? 0 : 1;
            CK(p, cudaMemcpyAsync(dn, nb.data(), (size_t)nrows, cudaMemcpyHostToDevice, p->stream));
            CK(p, cudaStreamSynchronize(p->stream));
            p->stats.h2d_bytes += (int64_t)((size_t)nrows * (eb + 1));
        }
        if (bound[w]->validity) p->jb_has_null[(size_t)i] = true;
    }
    p->jb_rows += nrows;
    return BKGPU_OK;
}

// Fused probe (JoinProbe in agg.h): usable when the only build-side column the aggregate reads is its 4-byte GROUP BY
// key and the probe key is a 32-bit integer column whose cast to the comparison type keeps the canonical image.
static int join_compose_probe(bkgpu_plan* p, uint32_t* flag) {
    const Compiled& c = p->c;
    p->jp = JoinProbe{};
    if (!c.jfast || !c.jfast->has_direct || c.jfast->direct.n_keys != 1) return BKGPU_OK;
    const Compiled& f = *c.jfast;
    const size_t kpos = (size_t)f.direct.n_terms;
    int attr_main = -1;
    for (size_t i = 0; i < f.direct_cols.size(); i++) {
        const int m = c.jfast_of_main[(size_t)f.direct_cols[i]];
        if (c.col_side[(size_t)m] != 1) continue;
        if (i != kpos) return BKGPU_OK;   // a predicate or an aggregate argument comes from the build side: gather path
        attr_main = m;
    }
    if (attr_main < 0) return BKGPU_OK;
    const int ast = prim_storage(c.cols[(size_t)attr_main].prim);
    if ((ast != ST_I32 && ast != ST_U32) || (c.cols[(size_t)attr_main].prim != BK_INT32 && c.cols[(size_t)attr_main].prim != BK_UINT32)) return BKGPU_OK;
    const int pp = c.cols[(size_t)c.probe_key_col].prim, ct = c.join_key_prim;
    const bool ident = (pp == BK_INT32 && (ct == BK_INT32 || ct == BK_INT64)) || (pp == BK_UINT32 && (ct == BK_UINT32 || ct == BK_UINT64 || ct == BK_INT64));
    if (!ident) return BKGPU_OK;
    int rc;
    JoinProbe jp{};
    jp.mode = p->jf.mode; jp.key_signed = pp == BK_INT32 ? 1 : 0;
    jp.dense_min = p->jf.dense_min; jp.dense_size = p->jf.dense_size; jp.bias = p->jf.bias; jp.packed_mask = p->jf.packed_mask;
    if (jp.mode == 1) {
        if ((rc = ensure_buf(p, (void**)&p->jp_attr, &p->jp_attr_cap, (size_t)jp.dense_size * 4))) return rc;
        jp.attr = p->jp_attr;
        jp.present = (uint64_t)p->jb_rows == jp.dense_size ? nullptr : p->jf.dense;   // unique, NULL-free keys filling the range: every key exists
    } else {
        if ((rc = ensure_buf(p, (void**)&p->jp_packed, &p->jp_packed_cap, ((size_t)jp.packed_mask + 1) * 8))) return rc;
        jp.packed = p->jp_packed;
    }
    CK(p, launch_join_compose(p->jf, (const uint32_t*)p->jb_vals[(size_t)attr_main], p->jp_attr, p->jp_packed, flag, p->stream));
    p->stats.kernel_launches++;
    p->jp = jp; p->jp_key_pos = (int)kpos;   // (the caller drops it again when `flag` comes back set)
    return BKGPU_OK;
}

static int join_build_table(bkgpu_plan* p) {
    const Compiled& c = p->c;
    const size_t nc = c.cols.size();
    if (p->jb_vals.empty()) { p->jb_vals.assign(nc, nullptr); p->jb_nullbytes.assign(nc, nullptr); p->jb_bitmap.assign(nc, nullptr); p->jb_has_null.assign(nc, false); }
    int rc;
    for (size_t i = 0; i < nc; i++) {
        if (c.col_side[i] != 1) continue;
        // (a bitmap kept from an earlier run of the plan must not mask this run's rows: a column without NULLs has none)
        dev_free(p, p->jb_bitmap[i]); p->jb_bitmap[i] = nullptr;
        if (!p->jb_has_null[i]) continue;
        if ((rc = dev_alloc(p, (void**)&p->jb_bitmap[i], (size_t)(p->jb_rows + 7) / 8 + 8))) return rc;
        CK(p, launch_pack_validity(p->jb_nullbytes[i], p->jb_rows, p->jb_bitmap[i], p->stream));
        p->stats.kernel_launches++;
    }
    DevCol key{};
    const size_t ki = (size_t)c.build_key_col;
    key.values = p->jb_vals[ki]; key.validity = p->jb_bitmap[ki]; key.stype = prim_storage(c.cols[ki].prim); key.prim = c.cols[ki].prim;
    p->jt_built = true; p->jt_generic = false;   // the multimap of the general probe is built on first use (join_generic_table)
    // ---- FK -> PK fast path: are the build keys unique, and is their range dense or their type <= 32 bits? ----
    p->jf = JoinFast{}; p->jp = JoinProbe{};
    if (p->jg_buf.empty()) p->jg_buf.assign(nc, nullptr);
    bool build_nulls = false;
    for (size_t i = 0; i < nc; i++) if (c.col_side[i] == 1 && p->jb_has_null[i]) build_nulls = true;   // gathered columns must be NULL-free
    bool retry_measured = false;
again:
    if (c.jfast && !build_nulls && p->jb_rows > 0) {
        const int cls = host_prim_class(c.join_key_prim);
        const uint64_t bias = cls == VC_I64 ? 0x8000000000000000ull : 0ull;
        if ((rc = ensure_buf(p, (void**)&p->j_scratch, &p->j_scratch_cap, 64))) return rc;
        uint64_t* mm = p->j_scratch; uint32_t* dup = (uint32_t*)(p->j_scratch + 2);
        // the key range: measured (one pass over the keys + a round trip), or — when this plan has built before — the range it saw then,
        // checked by the build kernel itself (a key outside it raises a flag and the build is redone with a measured range)
        uint64_t h_mm[2];
        bool guessed = p->jf_learned && !retry_measured && p->join_learn_range;
        if (guessed) { h_mm[0] = p->jf_learn_min; h_mm[1] = p->jf_learn_max; }
        else {
            CK(p, launch_join_minmax(key, c.cols[ki].prim, c.join_key_prim, p->jb_rows, bias, mm, p->stream));
            CK(p, cudaMemcpyAsync(h_mm, mm, 16, cudaMemcpyDeviceToHost, p->stream));
            CK(p, cudaStreamSynchronize(p->stream));
            p->stats.kernel_launches++;
        }
        JoinFast jf{}; jf.bias = bias;
        const uint64_t range = h_mm[1] >= h_mm[0] ? h_mm[1] - h_mm[0] + 1 : 0;
        const int kb = storage_bytes(prim_storage(c.join_key_prim));
        if (range > 0 && range <= (uint64_t)4 * (uint64_t)p->jb_rows + 1024 && range <= (1ull << 30)) {
            jf.mode = 1; jf.dense_min = h_mm[0]; jf.dense_size = range;
            if ((rc = ensure_buf(p, (void**)&p->jf_dense, &p->jf_dense_cap, (size_t)range * 4))) return rc;
            jf.dense = p->jf_dense;
        } else if (kb <= 4) {
            uint32_t pc = 1024; while ((int64_t)pc < 2 * p->jb_rows) pc <<= 1;
            jf.mode = 2; jf.packed_mask = pc - 1;
            if ((rc = ensure_buf(p, (void**)&p->jf_packed, &p->jf_packed_cap, (size_t)pc * 8))) return rc;
            jf.packed = p->jf_packed;
        }
        if (jf.mode) {
            CK(p, launch_join_build_fast(key, c.cols[ki].prim, c.join_key_prim, p->jb_rows, jf, p->jf_dense, p->jf_packed, dup, p->stream));
            p->stats.kernel_launches += 1;
            p->jf = jf;
            if ((rc = join_compose_probe(p, dup + 1))) return rc;   // enqueued behind the build: all flags come back in one round trip
            uint32_t h_flags[3] = {0, 0, 0};
            CK(p, cudaMemcpyAsync(h_flags, dup, 12, cudaMemcpyDeviceToHost, p->stream));
            CK(p, cudaStreamSynchronize(p->stream));
            if (h_flags[2] && guessed) { retry_measured = true; goto again; }   // this run's keys left the learned range
            if (h_flags[0]) { p->jf = JoinFast{}; p->jp = JoinProbe{}; }   // duplicate build keys: not a PK, the general probe runs
            else if (h_flags[1]) p->jp = JoinProbe{};
            p->jf_learned = true; p->jf_learn_min = h_mm[0]; p->jf_learn_max = h_mm[1];
        } else if (guessed) { retry_measured = true; goto again; }   // (the learned range no longer fits this row count)
    }
    return BKGPU_OK;
}

// the general probe's multimap (duplicate build keys, unmatched probe rows): built only when a batch needs it
static int join_generic_table(bkgpu_plan* p) {
    if (p->jt_generic) return BKGPU_OK;
    const Compiled& c = p->c;
    int rc;
    uint32_t cap = 1024;
    while ((int64_t)cap < 2 * p->jb_rows) cap <<= 1;
    dev_free(p, p->jt_keys); dev_free(p, p->jt_rows);
    if ((rc = dev_alloc(p, (void**)&p->jt_keys, (size_t)cap * 8))) return rc;
    if ((rc = dev_alloc(p, (void**)&p->jt_rows, (size_t)cap * 4))) return rc;
    p->jt_mask = cap - 1;
    DevCol key{};
    const size_t ki = (size_t)c.build_key_col;
    key.values = p->jb_vals[ki]; key.validity = p->jb_bitmap[ki]; key.stype = prim_storage(c.cols[ki].prim); key.prim = c.cols[ki].prim;
    CK(p, launch_join_build(key, c.cols[ki].prim, c.join_key_prim, p->jb_rows, p->jt_keys, p->jt_rows, p->jt_mask, p->stream));
    p->stats.kernel_launches++;
    if (c.join_type != BK_INNER_JOIN) {   // one "found a partner" byte per build row
        if ((rc = ensure_buf(p, (void**)&p->jb_matched, &p->jb_matched_cap, (size_t)std::max<int64_t>(p->jb_rows, 1)))) return rc;
        CK(p, cudaMemsetAsync(p->jb_matched, 0, (size_t)std::max<int64_t>(p->jb_rows, 1), p->stream));
    }
    p->jt_generic = true;
    return BKGPU_OK;
}

static int join_probe_batch(bkgpu_plan* p, const DevCol* probe_cols, int64_t nrows, int64_t, bool) {
    const Compiled& c = p->c;
    std::vector<DevCol> all(c.cols.size());
    for (size_t i = 0; i < c.cols.size(); i++) {
        if (c.col_side[i] == 1) { all[i].values = p->jb_vals[i]; all[i].validity = p->jb_bitmap[i]; all[i].stype = prim_storage(c.cols[i].prim); all[i].prim = c.cols[i].prim; }
    }
    for (size_t w = 0; w < p->probe_map.size(); w++) all[(size_t)p->probe_map[w]] = probe_cols[w];
    // ---- FK -> PK fast path: gather the build columns next to the probe rows, then the ordinary (lean) aggregate ----
    if (p->jf.mode != 0 && p->jp.mode != 0 && c.jfast && !p->force_generic && !p->no_fused_probe) {
        // ---- fused probe: the lean aggregate reads the foreign key and looks the group key up itself ----
        const Compiled& f = *c.jfast;
        std::vector<DevCol> fc(f.cols.size());
        const int key_fcol = f.direct_cols[(size_t)p->jp_key_pos];
        bool vec_ok = true;
        for (size_t i = 0; i < f.cols.size(); i++) {
            fc[i] = (int)i == key_fcol ? all[(size_t)c.probe_key_col] : all[(size_t)c.jfast_of_main[i]];
            if (((uintptr_t)fc[i].values & 31) != 0) vec_ok = false;
        }
        const int rc = launch_agg_batch(p, f, fc.data(), nrows, vec_ok, &p->jp);
        if (rc <= 0) return rc;   // launched (0) or failed (< 0); +1: this batch does not fit the lean kernel
    }
    if (p->jf.mode != 0 && c.jfast && !p->force_generic) {
        const Compiled& f = *c.jfast;
        const int64_t chunk = 16 << 20;   // the gathered chunk is re-read immediately: keep it L2-sized
        for (int64_t off = 0; off < nrows; off += chunk) {
            const int64_t n = std::min(chunk, nrows - off);
            GatherCols gc{}; std::vector<DevCol> fc(f.cols.size());
            int rc;
            for (size_t i = 0; i < f.cols.size(); i++) {
                const size_t m = (size_t)c.jfast_of_main[i];
                const size_t eb = (size_t)storage_bytes(prim_storage(c.cols[m].prim));
                if (c.col_side[m] == 1) {
                    if (!p->jg_buf[m] || p->jg_rows < n) {
                        dev_free(p, p->jg_buf[m]); p->jg_buf[m] = nullptr;
                        if ((rc = dev_alloc(p, (void**)&p->jg_buf[m], (size_t)chunk * eb))) return rc;
                    }
                    gc.src[gc.n] = p->jb_vals[m]; gc.dst[gc.n] = p->jg_buf[m]; gc.elem[gc.n] = (int32_t)eb; gc.n++;
                    fc[i].values = p->jg_buf[m]; fc[i].validity = nullptr; fc[i].stype = prim_storage(c.cols[m].prim); fc[i].prim = c.cols[m].prim;
                } else {
                    fc[i] = all[m];
                    fc[i].values = (const uint8_t*)all[m].values + (size_t)off * eb;
                    if (all[m].validity) fc[i].validity = all[m].validity + off / 8;
                }
            }
            p->jg_rows = chunk;
            DevCol pk = all[(size_t)c.probe_key_col];
            pk.values = (const uint8_t*)pk.values + (size_t)off * (size_t)storage_bytes(pk.stype);
            if (pk.validity) pk.validity += off / 8;
            CK(p, cudaMemsetAsync(p->d_cursor + 1, 0, 4, p->stream));
            CK(p, launch_join_gather(pk, c.cols[(size_t)c.probe_key_col].prim, c.join_key_prim, n, p->jf, gc, p->d_cursor + 1, p->stream));
            p->stats.kernel_launches++;
            uint32_t miss = 0;
            CK(p, cudaMemcpyAsync(&miss, p->d_cursor + 1, 4, cudaMemcpyDeviceToHost, p->stream));
            CK(p, cudaStreamSynchronize(p->stream));
            if (miss) {   // some probe row has no partner (or a NULL key): this chunk takes the general probe
                std::vector<DevCol> sl(all);
                for (size_t i = 0; i < sl.size(); i++) if (c.col_side[i] == 0) {
                    sl[i].values = (const uint8_t*)all[i].values + (size_t)off * (size_t)storage_bytes(all[i].stype);
                    if (all[i].validity) sl[i].validity = all[i].validity + off / 8;
                }
                if ((rc = join_generic_table(p))) return rc;
                if ((rc = launch_agg_batch(p, c, sl.data(), n, false))) return rc;
            } else {
                bool vec_ok = true;
                for (auto& d : fc) if (((uintptr_t)d.values & 31) != 0) vec_ok = false;
                if ((rc = launch_agg_batch(p, f, fc.data(), n, vec_ok))) return rc;
            }
        }
        return BKGPU_OK;
    }
    { int rc = join_generic_table(p); if (rc) return rc; }
    return launch_agg_batch(p, c, all.data(), nrows, false);
}

// LEFT / SEMI / ANTI_SEMI: once every probe batch has run, the preserved (build) side's rows the join type asks for go through the
// aggregate — unmatched rows NULL-extended (LEFT), matched rows (SEMI), unmatched rows (ANTI_SEMI); join_node.cpp:1200-1276
static int join_tail(bkgpu_plan* p) {
    const Compiled& c = p->c;
    if (p->jb_rows == 0) return BKGPU_OK;
    int rc;
    if (!p->jt_built && (rc = join_build_table(p))) return rc;
    if ((rc = join_generic_table(p))) return rc;
    std::vector<DevCol> all(c.cols.size());
    for (size_t i = 0; i < c.cols.size(); i++) {
        all[i].stype = prim_storage(c.cols[i].prim); all[i].prim = c.cols[i].prim;
        if (c.col_side[i] == 1) { all[i].values = p->jb_vals[i]; all[i].validity = p->jb_bitmap[i]; }   // (probe-side columns are never read: row = -1 is NULL)
    }
    p->join_tail_launch = true;
    rc = launch_agg_batch(p, c, all.data(), p->jb_rows, false);
    p->join_tail_launch = false;
    return rc;
}

// ---- a JOIN that returns rows (PK_JOIN): pairs -> gather -> the sink fragment (p->post_sort) ----
static int join_rows_emit(bkgpu_plan* p, const std::vector<DevCol>& all, int64_t n_probe, bool tail) {
    const Compiled& c = p->c;
    int rc;
    AggArgs a; memset(&a, 0, sizeof a);
    a.plan = c.ap; a.prog = c.prog; a.nrows = n_probe;
    const int ncols = (int)c.cols.size();
    for (int i = 0; i < ncols; i++) a.cols[i] = all[(size_t)i];
    a.n_cols = ncols;
    a.join.enabled = 1; a.join.keys = p->jt_keys; a.join.rows = p->jt_rows; a.join.cap_mask = p->jt_mask;
    a.join.probe_col = c.probe_key_col; a.join.probe_prim = c.cols[(size_t)c.probe_key_col].prim; a.join.cast_prim = c.join_key_prim;
    a.join.join_type = c.join_type; a.join.matched = c.join_type == BK_INNER_JOIN ? nullptr : p->jb_matched; a.join.n_build = p->jb_rows;
    a.join.tail = tail ? 1 : 0;
    if ((rc = ensure_buf(p, (void**)&p->jr_cursor, &p->jr_cursor_cap, 64))) return rc;
    size_t want = (size_t)std::max<int64_t>(tail ? p->jb_rows : n_probe, 1 << 12);
    uint32_t found = 0;
    for (int attempt = 0; attempt < 2; attempt++) {   // a batch with more pairs than the buffer holds is counted, the buffer grown, the batch rerun
        if (want > 0x7FFFFFF0ull) return p->fail(BKGPU_ETOOBIG, "a probe batch joins to more than 2^31 rows: push smaller batches");
        if ((rc = ensure_buf(p, (void**)&p->jr_pairs, &p->jr_pairs_cap, want * 8))) return rc;
        CK(p, cudaMemsetAsync(p->jr_cursor, 0, 4, p->stream));
        // (a rerun of a LEFT batch sets the same matched[] flags again: they only ever go from 0 to 1)
        CK(p, launch_join_pairs(a, p->jr_pairs, (uint32_t)want, p->jr_cursor, p->sm_count, p->stream));
        CK(p, cudaMemcpyAsync(&found, p->jr_cursor, 4, cudaMemcpyDeviceToHost, p->stream));
        CK(p, cudaStreamSynchronize(p->stream));
        p->stats.kernel_launches++;
        if ((size_t)found <= want) break;
        want = found;
    }
    if (!found) return BKGPU_OK;
    const Compiled& pc = *c.post;
    const size_t nc = pc.cols.size();
    if (p->post_cap < found || p->post_vals.size() != nc) {
        for (auto& v : {&p->post_vals, &p->post_nullb, &p->post_bitmap}) { for (uint8_t* q : *v) dev_free(p, q); v->assign(nc, nullptr); }
        const size_t cap = std::max<size_t>((size_t)found + found / 4, 1 << 16);
        for (size_t k = 0; k < nc; k++) {
            if ((rc = dev_alloc(p, (void**)&p->post_vals[k], cap * (size_t)storage_bytes(prim_storage(pc.cols[k].prim)) + 64))) return rc;
            if ((rc = dev_alloc(p, (void**)&p->post_nullb[k], cap + 64))) return rc;
            if ((rc = dev_alloc(p, (void**)&p->post_bitmap[k], cap / 8 + 64))) return rc;
        }
        p->post_cap = cap;
    }
    std::vector<DevCol> dc(std::max<size_t>(nc, 1));
    for (size_t k = 0; k < nc; k++) {   // sink column k = output column k = device column k of the join (plan.cpp::lower_join_rows)
        const int st = prim_storage(pc.cols[k].prim);
        CK(p, launch_join_rows_gather(all[k], p->jr_pairs, c.col_side[k] == 1 ? 1 : 0, found, storage_bytes(st), p->post_vals[k], p->post_nullb[k], p->stream));
        CK(p, launch_pack_validity(p->post_nullb[k], found, p->post_bitmap[k], p->stream));
        dc[k].values = p->post_vals[k]; dc[k].validity = p->post_bitmap[k]; dc[k].stype = st; dc[k].prim = pc.cols[k].prim;
    }
    p->stats.kernel_launches += 2 * (int64_t)nc;
    p->rows_passed_host += found;
    if ((rc = sort_push(p->post_sort, dc.data(), (int64_t)found, p->stream, &p->stats, p->last_error))) { g_thread_error = p->last_error; return rc; }
    CK(p, cudaStreamSynchronize(p->stream));   // the gather buffers are reused by the next batch
    return BKGPU_OK;
}
static void join_all_cols(bkgpu_plan* p, const DevCol* probe_cols, std::vector<DevCol>& all) {
    const Compiled& c = p->c;
    all.assign(c.cols.size(), DevCol{});
    for (size_t i = 0; i < c.cols.size(); i++) {
        all[i].stype = prim_storage(c.cols[i].prim); all[i].prim = c.cols[i].prim;
        if (c.col_side[i] == 1) { all[i].values = p->jb_vals[i]; all[i].validity = p->jb_bitmap[i]; }
    }
    if (probe_cols) for (size_t w = 0; w < p->probe_map.size(); w++) all[(size_t)p->probe_map[w]] = probe_cols[w];
}
static int join_rows_batch(bkgpu_plan* p, const DevCol* probe_cols, int64_t nrows, int64_t, bool) {
    int rc;
    if ((rc = join_generic_table(p))) return rc;
    std::vector<DevCol> all;
    join_all_cols(p, probe_cols, all);
    return join_rows_emit(p, all, nrows, false);
}
static int join_rows_finish(bkgpu_plan* p) {
    const Compiled& c = p->c;
    int rc;
    if (c.join_type != BK_INNER_JOIN && p->jb_rows > 0) {   // LEFT: the preserved rows that found no partner, NULL-extended
        if (!p->jt_built && (rc = join_build_table(p))) return rc;
        if ((rc = join_generic_table(p))) return rc;
        std::vector<DevCol> all;
        join_all_cols(p, nullptr, all);
        if ((rc = join_rows_emit(p, all, 0, true))) return rc;
    }
    std::vector<SortOutCol> cols; int64_t rows = 0;
    if ((rc = sort_finish(p->post_sort, nullptr, 1, p->stream, &p->stats, cols, &rows, p->last_error))) { g_thread_error = p->last_error; return rc; }
    p->result.clear();
    for (size_t k = 0; k < cols.size(); k++) {
        SortOutCol& sc = cols[k];
        HostCol hc; hc.desc = c.out_cols[k]; hc.elem = sc.elem;
        hc.values = std::move(sc.values); hc.validity = std::move(sc.validity);
        p->result.push_back(std::move(hc));
    }
    p->result_rows = rows; p->result_pos = 0;
    p->stats.rows_filtered = 0;
    return BKGPU_OK;
}

static int join_push(bkgpu_plan* p, const bkgpu_column* cols, int ncols, int64_t nrows, int on_device) {
    const Compiled& c = p->c;
    bool is_build = false;
    for (int k = 0; k < ncols; k++) if (cols[k].tuple_id == c.build_tuple) is_build = true;
    if (is_build) return join_retain_build(p, cols, ncols, nrows, on_device);
    if (!p->jt_built) { int rc = join_build_table(p); if (rc) return rc; }
    if (p->probe_want.empty()) for (size_t i = 0; i < c.cols.size(); i++) if (c.col_side[i] == 0) { p->probe_want.push_back(c.cols[i]); p->probe_map.push_back((int)i); }
    return feed(p, p->probe_want, cols, ncols, nrows, on_device, c.kind == PK_JOIN ? join_rows_batch : join_probe_batch);
}

static int bkgpu_push_impl(bkgpu_plan* p, const bkgpu_column* cols, int ncols, int64_t nrows, int on_device);
extern "C" int bkgpu_push(bkgpu_plan* p, const bkgpu_column* cols, int ncols, int64_t nrows, int on_device) {
    const double t0 = HostClock::now();
    const int rc = bkgpu_push_impl(p, cols, ncols, nrows, on_device);
    if (p) p->hclk.push += HostClock::now() - t0;
    return rc;
}
static int bkgpu_push_impl(bkgpu_plan* p, const bkgpu_column* cols, int ncols, int64_t nrows, int on_device) {
    if (!p) return thread_fail(BKGPU_EINVAL, "bkgpu_push: NULL plan");
    if (p->state != S_OPEN) return p->fail(BKGPU_ESTATE, "bkgpu_push needs an open, unfinished plan");
    if (p->cancelled.load()) return p->fail(BKGPU_ECANCELLED, "cancelled");
    if (nrows < 0 || ncols < 0 || (ncols > 0 && !cols)) return p->fail(BKGPU_EINVAL, "bkgpu_push: bad arguments");
    CK(p, cudaSetDevice(p->device));
    p->stats.rows_scanned += nrows;
    if (nrows == 0) return BKGPU_OK;
    switch (p->c.kind) {
        case PK_AGG: {
            int rc = feed(p, p->c.cols, cols, ncols, nrows, on_device, agg_batch);
            if (rc) return rc;
            if (p->c.ap.n_keyw > 0 && p->smem_cap_log2 < 0 && p->h_pinned)   // learn the cardinality for later batches' table size;
                CK(p, cudaMemcpyAsync(p->h_pinned, p->gt.n_groups, 4, cudaMemcpyDeviceToHost, p->stream));  // no sync: a stale value only costs speed
            return BKGPU_OK;
        }
        case PK_JOIN_AGG: case PK_JOIN: return join_push(p, cols, ncols, nrows, on_device);
        case PK_SORT: case PK_FILTER: return feed(p, p->c.cols, cols, ncols, nrows, on_device, sort_batch);
        default: return p->fail(BKGPU_EUNSUPPORTED, "plan kind %d has no push path yet", p->c.kind);
    }
}

// ---- result materialisation: device images -> typed host columns ----
static void put_value(HostCol& hc, int64_t row, uint64_t bits, bool isnull, int prim) {
    uint8_t* dst = hc.values.data() + (size_t)row * (size_t)hc.elem;
    if (isnull) { hc.validity[(size_t)row >> 3] &= (uint8_t)~(1u << (row & 7)); return; }
    switch (prim_storage(prim)) {
        case ST_I32: { int32_t v = (int32_t)(int64_t)bits; memcpy(dst, &v, 4); } break;
        case ST_U32: { uint32_t v = (uint32_t)bits; memcpy(dst, &v, 4); } break;
        case ST_F32: { double d; memcpy(&d, &bits, 8); float f = (float)d; memcpy(dst, &f, 4); } break;
        case ST_U8: dst[0] = bits ? 1 : 0; break;
        default: memcpy(dst, &bits, 8); break;
    }
}

// One-time setup of the peer-memory merge: allocate this rank's buffer, trade CUDA IPC handles through the communicator, map the peers'.
static int peer_setup(bkgpu_plan* p, size_t seg_words) {
    if (p->peer_ready) return p->peer_seg_words == seg_words ? BKGPU_OK : p->fail(BKGPU_ESTATE, "peer merge: the partial-state size changed after setup");
    const int n = p->nranks;
    if (nccl_comm_rank(p->nccl_comm, &p->peer_rank) != 0) return p->fail(BKGPU_ENCCL, "ncclCommUserRank: %s", nccl_last_error());
    const size_t words = 2 * (size_t)n * seg_words + 2 * (size_t)n;
    int rc;
    if ((rc = dev_alloc(p, (void**)&p->peer_local, words * 8))) return rc;
    CK(p, cudaMemsetAsync(p->peer_local, 0, words * 8, p->stream));
    cudaIpcMemHandle_t mine;
    CK(p, cudaIpcGetMemHandle(&mine, p->peer_local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    uint64_t *d_mine = nullptr, *d_all = nullptr;
    if ((rc = dev_alloc(p, (void**)&d_mine, 64))) return rc;
    if ((rc = dev_alloc(p, (void**)&d_all, 64 * (size_t)n))) return rc;
    CK(p, cudaMemcpyAsync(d_mine, &mine, 64, cudaMemcpyHostToDevice, p->stream));
    if (nccl_all_gather(p->nccl_comm, d_mine, d_all, 8, p->stream) != 0) return p->fail(BKGPU_ENCCL, "ncclAllGather (IPC handles): %s", nccl_last_error());
    std::vector<cudaIpcMemHandle_t> all((size_t)n);
    CK(p, cudaMemcpyAsync(all.data(), d_all, 64 * (size_t)n, cudaMemcpyDeviceToHost, p->stream));
    CK(p, cudaStreamSynchronize(p->stream));
    p->peer_ptr.assign((size_t)n, nullptr);
    for (int r = 0; r < n; r++) {
        if (r == p->peer_rank) { p->peer_ptr[(size_t)r] = p->peer_local; continue; }
        void* q = nullptr;
        CK(p, cudaIpcOpenMemHandle(&q, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess));
        p->peer_ptr[(size_t)r] = (uint64_t*)q;
    }
    if ((rc = dev_alloc(p, (void**)&p->d_peer_ptrs, 8 * (size_t)n))) return rc;
    if ((rc = dev_alloc(p, (void**)&p->d_peer_timeout, 8))) return rc;
    CK(p, cudaMemcpyAsync(p->d_peer_ptrs, p->peer_ptr.data(), 8 * (size_t)n, cudaMemcpyHostToDevice, p->stream));
    CK(p, cudaMemsetAsync(p->d_peer_timeout, 0, 8, p->stream));
    // nobody may write into a buffer before its owner has zeroed it: one more trip through the communicator is the barrier
    if (nccl_all_gather(p->nccl_comm, d_mine, d_all, 8, p->stream) != 0) return p->fail(BKGPU_ENCCL, "ncclAllGather (barrier): %s", nccl_last_error());
    CK(p, cudaStreamSynchronize(p->stream));
    dev_free(p, d_mine); dev_free(p, d_all);
    p->peer_seg_words = seg_words; p->peer_ready = true;
    return BKGPU_OK;
}

// groups per rank in the exchanged partial state: never more than the group table can hold (every rank runs the same plan with
// the same options, so all ranks agree on it) — a 2^14-slot table ships 0.8 MB per rank instead of the 3 MB of the default cap
static uint32_t eff_pcap(const bkgpu_plan* p) {
    const int64_t table = p->c.ap.n_keyw == 0 ? 1 : (int64_t)1 << p->group_cap_log2;
    return (uint32_t)std::min<int64_t>(p->partial_cap, table);
}

// grow-only device buffers of the multi-GPU exchange
static int ensure_words(bkgpu_plan* p, uint64_t** buf, size_t* have, size_t want) {
    if (*have >= want) return BKGPU_OK;
    dev_free(p, *buf); *buf = nullptr; *have = 0;
    int rc = dev_alloc(p, (void**)buf, want * 8);
    if (rc) return rc;
    *have = want;
    return BKGPU_OK;
}

// the collective step of bkgpu_finish: regions -> one set per GPU, the per-GPU partial tables meet once and are folded by K3.
// Default: compact rows, one ncclAllGather of 1 + bound x (key + lane words) per rank — `bound` = the largest group count any rank
// held in earlier runs of this plan (2048 on the first run); the merge reports the true maximum and the exchange is repeated once
// with a larger bound when it did not fit.  Options: peer_merge (NVLink peer memory), repartition (all-to-all by key owner).
static int agg_collective(bkgpu_plan* p, bool* rows_mode) {
    const AggPlan& ap = p->c.ap;
    GroupTable& gt = p->gt;
    *rows_mode = false;
    const uint32_t pcap = eff_pcap(p);
    int rc;
    const bool repart = p->repartition && ap.n_keyw > 0;   // (a scalar aggregate has one group: nothing to partition)
    EventPair* ep = timer_begin(p, p->timed_coll, 0);
    if (!(p->peer_merge && !repart) && !repart) {
        *rows_mode = true;
        if (p->peer_rank < 0 && nccl_comm_rank(p->nccl_comm, &p->peer_rank) != 0) return p->fail(BKGPU_ENCCL, "ncclCommUserRank: %s", nccl_last_error());
        uint32_t bound = p->merge_bound ? p->merge_bound : 2048u;
        if (ap.n_keyw == 0) bound = 1;
        bound = std::min<uint32_t>(bound, pcap);
        const size_t rw = (size_t)(ap.n_keyw + ap.n_lanes);
        const size_t words = 1 + rw * bound;
        if ((rc = ensure_words(p, &p->d_partial, &p->d_partial_words, words))) return rc;
        if ((rc = ensure_words(p, &p->d_gather, &p->d_gather_words, words * (size_t)p->nranks))) return rc;
        CK(p, launch_partial_export_rows(gt, ap, p->d_partial, bound, p->stream));
        if (nccl_all_gather(p->nccl_comm, p->d_partial, p->d_gather, words, p->stream) != 0) return p->fail(BKGPU_ENCCL, "ncclAllGather: %s", nccl_last_error());
        CK(p, launch_partial_merge_rows(gt, ap, p->d_gather, words, bound, p->nranks, p->peer_rank, p->d_cursor + 1, p->stream));
        p->merge_bound_used = bound;
        p->stats.kernel_launches += 2;
        timer_end(p, ep);
        return BKGPU_OK;
    }
    const size_t words = 1 + (size_t)(ap.n_keyw + ap.n_lanes) * pcap;
    if ((rc = ensure_words(p, &p->d_partial, &p->d_partial_words, words * (repart ? (size_t)p->nranks : 1)))) return rc;
    if ((rc = ensure_words(p, &p->d_gather, &p->d_gather_words, words * (size_t)p->nranks))) return rc;
    if (p->peer_merge && !repart && (rc = peer_setup(p, words))) return rc;
    if (p->peer_merge && !repart) {
        const uint64_t seq = ++p->peer_seq;
        CK(p, launch_peer_exchange(gt, ap, p->d_peer_ptrs, p->peer_local, p->nranks, p->peer_rank, words, pcap, seq, p->d_cursor, p->d_peer_timeout, p->stream));
        CK(p, launch_table_init(gt, ap, p->stream, 1));
        CK(p, launch_partial_merge(gt, ap, p->peer_local + (size_t)(seq & 1) * (size_t)p->nranks * words, words, pcap, p->nranks, p->stream));
        timer_end(p, ep);
        p->stats.kernel_launches += 5;
        if (p->h_pinned) CK(p, cudaMemcpyAsync(p->h_pinned + 6, p->d_peer_timeout, 4, cudaMemcpyDeviceToHost, p->stream));   // checked after the result's synchronisation
        return BKGPU_OK;
    }
    // hash repartition: every rank keeps only the groups it owns — an all-to-all of per-owner segments instead of the
    // all-gather; the merged table (and the result) of a rank is its partition, the union over ranks is the answer
    if (!p->d_part_cursors && (rc = dev_alloc(p, (void**)&p->d_part_cursors, 4 * (size_t)p->nranks))) return rc;
    CK(p, launch_partial_export_parts(gt, ap, p->d_partial, words, pcap, p->d_part_cursors, p->nranks, p->stream));
    if (nccl_all_to_all(p->nccl_comm, p->d_partial, p->d_gather, words, p->nranks, p->stream) != 0) return p->fail(BKGPU_ENCCL, "all-to-all: %s", nccl_last_error());
    CK(p, launch_table_init(gt, ap, p->stream, 1));   // (keeps the overflow flag an export beyond partial_capacity raised)
    CK(p, launch_partial_merge(gt, ap, p->d_gather, words, pcap, p->nranks, p->stream));
    timer_end(p, ep);
    p->stats.kernel_launches += 4;
    return BKGPU_OK;
}

// [LIMIT ->] [SORT ->] [HAVING ->] above the aggregate: the extracted rows (still on the device) become typed columns and run through
// the post fragment's sort / filter kernels; only the final rows travel to the host
static int agg_post(bkgpu_plan* p, uint32_t n_out, uint32_t out_cap) {
    const Compiled& pc = *p->c.post;
    const size_t nc = pc.cols.size();
    int rc;
    if (p->post_cap < std::max<size_t>(n_out, 1) || p->post_vals.size() != nc) {
        for (auto& v : {&p->post_vals, &p->post_nullb, &p->post_bitmap}) { for (uint8_t* q : *v) dev_free(p, q); v->assign(nc, nullptr); }
        const size_t cap = std::max<size_t>(n_out, 2048);
        for (size_t c = 0; c < nc; c++) {
            if ((rc = dev_alloc(p, (void**)&p->post_vals[c], cap * (size_t)storage_bytes(prim_storage(pc.cols[c].prim)) + 64))) return rc;
            if ((rc = dev_alloc(p, (void**)&p->post_nullb[c], cap + 64))) return rc;
            if ((rc = dev_alloc(p, (void**)&p->post_bitmap[c], cap / 8 + 64))) return rc;
        }
        p->post_cap = cap;
    }
    PostCols pcs; memset(&pcs, 0, sizeof pcs);
    pcs.n = (int)nc;
    int img = 0;
    for (size_t c = 0; c < nc; c++) {   // (post columns = the aggregate's output columns, in order: plan.cpp::lower_post)
        pcs.img[c] = img; pcs.stype[c] = prim_storage(pc.cols[c].prim); pcs.values[c] = p->post_vals[c]; pcs.null_bytes[c] = p->post_nullb[c];
        img += p->c.out_cols[c].kind == 1 ? 2 : 1;
    }
    CK(p, launch_images_to_columns(p->d_outv, p->d_outn, out_cap, n_out, pcs, p->stream));
    std::vector<DevCol> dc(std::max<size_t>(nc, 1));
    for (size_t c = 0; c < nc; c++) {
        if (n_out) CK(p, launch_pack_validity(p->post_nullb[c], n_out, p->post_bitmap[c], p->stream));
        dc[c].values = p->post_vals[c]; dc[c].validity = p->post_bitmap[c]; dc[c].stype = prim_storage(pc.cols[c].prim); dc[c].prim = pc.cols[c].prim;
    }
    p->stats.kernel_launches += 1 + (int64_t)nc;
    if ((rc = sort_reset(p->post_sort, p->stream, p->last_error))) { g_thread_error = p->last_error; return rc; }
    bkgpu_stats scratch{};   // (the post fragment's kernels are not the request's "main kernel")
    if (n_out && (rc = sort_push(p->post_sort, dc.data(), n_out, p->stream, &scratch, p->last_error))) { g_thread_error = p->last_error; return rc; }
    std::vector<SortOutCol> cols; int64_t rows = 0;
    if ((rc = sort_finish(p->post_sort, nullptr, 1, p->stream, &scratch, cols, &rows, p->last_error))) { g_thread_error = p->last_error; return rc; }
    p->stats.kernel_launches += scratch.kernel_launches;
    p->result.clear();
    for (size_t c = 0; c < cols.size(); c++) {
        SortOutCol& sc = cols[c];
        HostCol hc; hc.desc = p->c.out_cols[c]; hc.elem = sc.elem;
        hc.values = std::move(sc.values); hc.validity = std::move(sc.validity);
        p->result.push_back(std::move(hc));
    }
    p->result_rows = rows; p->result_pos = 0;
    return BKGPU_OK;
}

static int agg_finish(bkgpu_plan* p) {
    const AggPlan& ap = p->c.ap;
    GroupTable& gt = p->gt;
    uint32_t host_counts[2] = {0, 0};
    uint64_t* finish_hv = nullptr; uint8_t* finish_hn = nullptr; uint32_t finish_n_out = 0, finish_out_cap = 0;
    const bool multi = p->nccl_comm && p->nranks > 1;
    bool rows_mode = false;
    // device images: one per group expr, per aggregate its final (+2 for an AVG blob)
    int n_img = ap.n_group;
    for (int k = 0; k < ap.n_agg; k++) if (!ap.agg[k].hidden) n_img += ap.agg[k].kind == AG_AVG ? 3 : 1;
    int rc;
  for (int mtry = 0; mtry < 3; mtry++) {   // (a second trip only when some rank held more groups than the exchange was sized for)
    { const double t0 = HostClock::now(); if (multi && (rc = agg_collective(p, &rows_mode))) return rc; p->hclk.collective += HostClock::now() - t0; }
    uint32_t n_out = 0, out_cap = 0, merge_max = 0;
    uint64_t* hv = nullptr; uint8_t* hn = nullptr;
    bool redo_merge = false;
    for (int attempt = 0; attempt < 2; attempt++) {
        // speculative extraction into the buffers kept from earlier runs: counts, cursor and rows come back
        // in ONE synchronisation; only a result larger than the buffers costs a second round
        uint32_t want = std::max<uint32_t>(std::max<uint32_t>(p->known_groups, 1), attempt ? std::max<uint32_t>(host_counts[0], 1) : 1);
        if (ap.n_keyw == 0) want = 1;
        if (p->out_cap_alloc < want) {
            dev_free(p, p->d_outv); p->d_outv = nullptr; p->d_outn = nullptr;
            size_t cap = std::max<size_t>(want, 2048);
            if ((rc = dev_alloc(p, (void**)&p->d_outv, cap * 9 * (size_t)n_img + 64))) return rc;   // values, then the null bytes: ONE copy brings both back
            p->d_outn = (uint8_t*)(p->d_outv + cap * (size_t)n_img);
            p->out_cap_alloc = cap;
        }
        out_cap = (uint32_t)p->out_cap_alloc;
        CK(p, launch_extract(gt, ap, p->d_outv, p->d_outn, out_cap, p->d_cursor, p->c.emit_default ? 1 : 0, p->stream));
        p->stats.kernel_launches++;
        const size_t n_words = (size_t)out_cap * (size_t)n_img;
        if (p->h_out_cap < n_words) {   // pinned: the four copies below are truly asynchronous and land in one synchronisation
            if (p->h_outv) cudaFreeHost(p->h_outv);
            p->h_outv = nullptr; p->h_outn = nullptr; p->h_out_cap = 0;
            CK(p, cudaHostAlloc((void**)&p->h_outv, n_words * 9 + 64, cudaHostAllocDefault));
            p->h_out_cap = n_words;
        }
        hv = p->h_outv; hn = p->h_outn = (uint8_t*)(p->h_outv + n_words);
        // two copies, one synchronisation: the 32-byte counter block (groups, overflow, rows extracted, merge info, rows passed) and
        // the extracted rows (values + null bytes, contiguous on both sides)
        uint32_t local_ctr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint32_t* hc3 = p->h_pinned ? p->h_pinned + 8 : local_ctr;   // [8] groups [9] overflow [10] rows extracted [11] merge info [12..13] rows passed
        CK(p, cudaMemcpyAsync(hc3, gt.n_groups, 32, cudaMemcpyDeviceToHost, p->stream));
        if ((const uint8_t*)p->d_outn == (const uint8_t*)(p->d_outv + n_words)) CK(p, cudaMemcpyAsync(hv, p->d_outv, n_words * 9, cudaMemcpyDeviceToHost, p->stream));
        else {
            CK(p, cudaMemcpyAsync(hv, p->d_outv, n_words * 8, cudaMemcpyDeviceToHost, p->stream));
            CK(p, cudaMemcpyAsync(hn, p->d_outn, n_words, cudaMemcpyDeviceToHost, p->stream));
        }
        {   // the one wait of the request.  With several ranks on a host whose CPU budget is small (8 ranks under a 16-CPU cgroup quota: two CPUs per
            // rank for the python thread, NCCL's proxy thread and this wait) a spinning wait burns the quota the other ranks' launches need:
            // there the thread sleeps on a blocking-sync event instead (option blocking_sync: -1 auto, 0 spin, 1 sleep)
            const double t0 = HostClock::now();
            if (p->blocking_wait) {
                if (!p->wait_event) CK(p, cudaEventCreateWithFlags(&p->wait_event, cudaEventBlockingSync | cudaEventDisableTiming));
                CK(p, cudaEventRecord(p->wait_event, p->stream));
                CK(p, cudaEventSynchronize(p->wait_event));
            } else CK(p, cudaStreamSynchronize(p->stream));
            p->hclk.extract_wait += HostClock::now() - t0;
        }
        host_counts[0] = hc3[0]; host_counts[1] = hc3[1]; n_out = hc3[2]; merge_max = hc3[3];
        memcpy(&p->rows_passed_host, hc3 + 4, 8);
        // FX: values beyond the sampled scale are added exactly, one global atomic each; a plan whose double columns keep producing many of them
        // (a column spanning dozens of binades, outliers dominating the sample) goes back to the CAS kernel for its later launches
        if (hc3[GT_FX_EXACT] > 4096 && (uint64_t)hc3[GT_FX_EXACT] * 64 > (uint64_t)std::max<int64_t>(p->rows_passed_host, 1)) p->fx_off = 1;
        if (p->peer_ready && p->h_pinned && p->h_pinned[6]) return p->fail(BKGPU_ENCCL, "peer merge: a rank did not publish its partial state within the time limit");
        if (rows_mode) {
            const uint32_t mx = merge_max;
            if (mx > p->merge_bound_used) {   // nothing was merged (k_partial_merge_rows): exchange again, sized for what the ranks really hold
                if (mx > eff_pcap(p)) return p->fail(BKGPU_ETOOBIG, "a rank holds %u groups, more than partial_capacity %lld: raise partial_capacity", mx, (long long)p->partial_cap);
                p->merge_bound = (mx + 63u) & ~63u;
                redo_merge = true;
                break;
            }
            const uint32_t learned = (std::max<uint32_t>(mx, 1u) + 63u) & ~63u;
            if (learned > p->merge_bound) p->merge_bound = learned;
        }
        p->stats.d2h_bytes += (int64_t)(n_words * 9 + 12);
        if (host_counts[1]) return p->fail(BKGPU_ETOOBIG, "group table overflow (capacity 2^%d slots / partial_capacity %lld): raise group_capacity_log2",
                                           (int)gt.cap_log2, (long long)p->partial_cap);
        if (host_counts[0] > p->known_groups) p->known_groups = host_counts[0];
        p->finish_groups = host_counts[0];
        if (n_out <= out_cap) break;   // everything fitted
    }
    if (!redo_merge) { finish_hv = hv; finish_hn = hn; finish_n_out = n_out; finish_out_cap = out_cap; break; }
    if (mtry == 2) return p->fail(BKGPU_ENCCL, "partial-state exchange did not converge");
  }
    uint64_t* hv = finish_hv; uint8_t* hn = finish_hn; uint32_t n_out = finish_n_out, out_cap = finish_out_cap;
    if (n_out > out_cap) n_out = out_cap;
    if (p->c.post) return agg_post(p, n_out, out_cap);
    int64_t rows = n_out;
    int64_t skip = p->c.offset > 0 ? std::min<int64_t>(p->c.offset, rows) : 0;
    int64_t lim = p->c.agg_limit;   // AggNode::get_next stops at its limit (agg_node.cpp:555)
    if (p->c.limit >= 0 && (lim < 0 || p->c.limit < lim)) lim = p->c.limit;
    int64_t keep = rows - skip;
    if (lim >= 0 && keep > lim) keep = lim;
    p->result.clear();
    int img = 0;
    for (const OutCol& oc : p->c.out_cols) {
        HostCol hc; hc.desc = oc;
        hc.elem = oc.kind == 1 ? 16 : storage_bytes(prim_storage(oc.prim));
        hc.values.assign((size_t)std::max<int64_t>(keep, 1) * (size_t)hc.elem, 0);
        hc.validity.assign((size_t)(keep + 7) / 8 + 1, 0xFF);
        bool any_null = false;
        for (int64_t r = 0; r < keep; r++) {
            const size_t src = (size_t)(r + skip);
            if (oc.kind == 1) {
                if (hn[(size_t)img * out_cap + src]) { any_null = true; hc.validity[(size_t)r >> 3] &= (uint8_t)~(1u << (r & 7)); }
                memcpy(hc.values.data() + (size_t)r * 16, &hv[(size_t)img * out_cap + src], 8);
                memcpy(hc.values.data() + (size_t)r * 16 + 8, &hv[(size_t)(img + 1) * out_cap + src], 8);
            } else {
                bool isnull = hn[(size_t)img * out_cap + src] != 0;
                any_null |= isnull;
                put_value(hc, r, hv[(size_t)img * out_cap + src], isnull, oc.prim);
            }
        }
        if (!any_null) hc.validity.clear();
        img += oc.kind == 1 ? 2 : 1;
        p->result.push_back(std::move(hc));
    }
    p->result_rows = keep; p->result_pos = 0;
    return BKGPU_OK;
}

static void resolve_timers(bkgpu_plan* p) {
    for (auto& ep : p->timed) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) { p->stats.main_kernel_ms += ms; p->stats.main_kernel_launches++; p->stats.main_kernel_bytes += ep.bytes; }
        p->event_pool.push_back(ep.a); p->event_pool.push_back(ep.b);
    }
    p->timed.clear();
    for (auto& ep : p->timed_coll) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) p->stats.collective_ms += ms;
        p->event_pool.push_back(ep.a); p->event_pool.push_back(ep.b);
    }
    p->timed_coll.clear();
}

static int bkgpu_finish_impl(bkgpu_plan* p);
extern "C" int bkgpu_finish(bkgpu_plan* p) {
    const double t0 = HostClock::now();
    const int rc = bkgpu_finish_impl(p);
    if (p) { p->hclk.finish += HostClock::now() - t0; p->hclk.n++; }
    return rc;
}
static int bkgpu_finish_impl(bkgpu_plan* p) {
    if (!p) return thread_fail(BKGPU_EINVAL, "bkgpu_finish: NULL plan");
    if (p->state != S_OPEN) return p->fail(BKGPU_ESTATE, "bkgpu_finish needs an open plan");
    if (p->cancelled.load()) return p->fail(BKGPU_ECANCELLED, "cancelled");
    CK(p, cudaSetDevice(p->device));
    int rc = BKGPU_OK;
    switch (p->c.kind) {
        case PK_AGG: case PK_JOIN_AGG: {
            if (p->c.kind == PK_JOIN_AGG && p->c.join_type != BK_INNER_JOIN && (rc = join_tail(p))) break;
            rc = agg_finish(p);
            if (!rc) {
                p->stats.rows_filtered = p->stats.rows_scanned - (int64_t)p->rows_passed_host;   // (came back with the counter block)
            }
        } break;
        case PK_JOIN: rc = join_rows_finish(p); break;
        case PK_SORT: case PK_FILTER: {
            std::vector<SortOutCol> cols; int64_t rows = 0;
            rc = sort_finish(p->sort, p->nccl_comm, p->nranks, p->stream, &p->stats, cols, &rows, p->last_error);
            if (rc) { g_thread_error = p->last_error; break; }
            p->result.clear();
            for (auto& sc : cols) {
                HostCol hc; hc.desc = {sc.tuple_id, sc.slot_id, sc.prim, 0}; hc.elem = sc.elem;
                hc.values = std::move(sc.values); hc.validity = std::move(sc.validity);
                p->result.push_back(std::move(hc));
            }
            p->result_rows = rows; p->result_pos = 0;
        } break;
        default: rc = p->fail(BKGPU_EUNSUPPORTED, "plan kind %d cannot finish", p->c.kind);
    }
    cudaStreamSynchronize(p->stream);
    resolve_timers(p);
    if (rc) return rc;
    p->stats.rows_returned = p->result_rows;
    p->state = S_FINISHED;
    return BKGPU_OK;
}

extern "C" int bkgpu_get_next(bkgpu_plan* p, bkgpu_column* out_cols, int* ncols, int64_t* nrows, int* eos) {
    if (!p || !ncols || !nrows || !eos) return thread_fail(BKGPU_EINVAL, "bkgpu_get_next: NULL argument");
    if (p->state != S_FINISHED) return p->fail(BKGPU_ESTATE, "bkgpu_get_next before bkgpu_finish");
    if (p->cancelled.load()) { *eos = 1; *nrows = 0; *ncols = 0; return BKGPU_OK; }  // cancelled: eos, like the reference's operators
    const int have = (int)p->result.size();
    if (*ncols < have || (have > 0 && !out_cols)) return p->fail(BKGPU_EINVAL, "bkgpu_get_next: need room for %d columns", have);
    const int64_t n = std::min(p->batch_capacity, p->result_rows - p->result_pos);
    for (int i = 0; i < have; i++) {
        HostCol& hc = p->result[(size_t)i];
        bkgpu_column& o = out_cols[i];
        o.tuple_id = hc.desc.tuple_id; o.slot_id = hc.desc.slot_id; o.prim_type = hc.desc.prim; o.elem_size = hc.elem;
        o.values = hc.values.data() + (size_t)p->result_pos * (size_t)hc.elem;
        o.length = n;
        o.validity = nullptr;
        if (!hc.validity.empty()) {
            if (p->result_pos == 0) o.validity = hc.validity.data();
            else {  // re-base the bitmap for a non-first batch (batch_capacity is normally a multiple of 8)
                if (p->result_pos % 8 == 0) o.validity = hc.validity.data() + p->result_pos / 8;
                else return p->fail(BKGPU_EINVAL, "batch_capacity must be a multiple of 8 when results hold NULLs");
            }
        }
    }
    *ncols = have; *nrows = n;
    p->result_pos += n;
    *eos = p->result_pos >= p->result_rows ? 1 : 0;
    return BKGPU_OK;
}

// Re-arm an executed plan for the next request of the same fragment (prepared-statement reuse): tables
// are cleared, allocations, streams and staging buffers are kept.
static int bkgpu_reset_impl(bkgpu_plan* p);
extern "C" int bkgpu_reset(bkgpu_plan* p) {
    const double t0 = HostClock::now();
    const int rc = bkgpu_reset_impl(p);
    if (p) p->hclk.reset += HostClock::now() - t0;
    return rc;
}
static int bkgpu_reset_impl(bkgpu_plan* p) {
    if (!p) return thread_fail(BKGPU_EINVAL, "bkgpu_reset: NULL plan");
    if (p->state != S_OPEN && p->state != S_FINISHED) return p->fail(BKGPU_ESTATE, "bkgpu_reset needs an opened plan");
    CK(p, cudaSetDevice(p->device));
    p->cancelled.store(0);
    if (p->c.kind == PK_AGG || p->c.kind == PK_JOIN_AGG) {
        // a finished run knows how many groups its table holds (the count came back with the result): clear those slots through the
        // occupied list instead of re-initialising the whole capacity (2^20 slots = 44 MB by default)
        if (p->state == S_FINISHED && p->finish_groups >= 0 && p->c.ap.n_keyw > 0) CK(p, launch_table_clear(p->gt, p->c.ap, (uint32_t)p->finish_groups, p->stream));
        else CK(p, launch_table_init(p->gt, p->c.ap, p->stream));
        p->finish_groups = -1;
        p->stats.kernel_launches++;
    }
    p->jb_rows = 0; p->jt_built = false; p->jt_generic = false; std::fill(p->jb_has_null.begin(), p->jb_has_null.end(), false);
    for (size_t i = 0; i < p->jb_bitmap.size(); i++) { dev_free(p, p->jb_bitmap[i]); p->jb_bitmap[i] = nullptr; }   // build-side validity of the previous run
    if (p->sort) { int rc = sort_reset(p->sort, p->stream, p->last_error); if (rc) { g_thread_error = p->last_error; return rc; } }
    if (p->c.kind == PK_JOIN && p->post_sort) { int rc = sort_reset(p->post_sort, p->stream, p->last_error); if (rc) { g_thread_error = p->last_error; return rc; } }
    p->result.clear(); p->result_rows = 0; p->result_pos = 0;
    bkgpu_stats z{}; z.kernel_launches = p->stats.kernel_launches; p->stats = z;
    p->state = S_OPEN;
    return BKGPU_OK;
}

extern "C" void bkgpu_cancel(bkgpu_plan* p) { if (p) p->cancelled.store(1); }

extern "C" void bkgpu_close(bkgpu_plan* p) {
    if (!p) return;
    if (p->trace && p->hclk.n) {
        const double n = (double)p->hclk.n;
        fprintf(stderr, "[bkgpu trace dev %d] %lld requests: host ms per request  reset %.4f  push %.4f  finish %.4f (collective enqueue %.4f, wait for the result %.4f)\n",
                p->device, (long long)p->hclk.n, p->hclk.reset / n, p->hclk.push / n, p->hclk.finish / n, p->hclk.collective / n, p->hclk.extract_wait / n);
    }
    cudaSetDevice(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
    resolve_timers(p);
    if (p->sort) sort_close(p->sort);
    if (p->post_sort) sort_close(p->post_sort);
    for (size_t r = 0; r < p->peer_ptr.size(); r++) if ((int)r != p->peer_rank && p->peer_ptr[r]) cudaIpcCloseMemHandle(p->peer_ptr[r]);
    for (cudaEvent_t e : p->event_pool) cudaEventDestroy(e);
    for (auto& q : p->dev_allocs) if (!dev_cache().put(p->device, q.second, q.first)) cudaFree(q.first);   // (both streams were drained above)
    for (int i = 0; i < 2; i++) { if (p->stage_free[i]) cudaEventDestroy(p->stage_free[i]); if (p->stage_ready[i]) cudaEventDestroy(p->stage_ready[i]); }
    for (int i = 0; i < 2; i++) { for (uint8_t* q : p->bounce[i]) if (q) cudaFreeHost(q); if (p->bounce_done[i]) cudaEventDestroy(p->bounce_done[i]); }
    if (p->wait_event) cudaEventDestroy(p->wait_event);
    if (p->h_pinned) cudaFreeHost(p->h_pinned);
    if (p->h_outv) cudaFreeHost(p->h_outv);
    if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
    if (p->own_stream && p->stream) cudaStreamDestroy(p->stream);
    delete p;
}

extern "C" int bkgpu_get_stats(bkgpu_plan* p, bkgpu_stats* out) {
    if (!p || !out) return thread_fail(BKGPU_EINVAL, "bkgpu_get_stats: NULL argument");
    *out = p->stats;
    return BKGPU_OK;
}

// ------------------------------------------------------------------ resident regions
// A region's columns can be registered once and stay in HBM across queries (the reference's analogue: its column store and parquet
// cache keep hot regions decoded in memory, include/column/file_manager.h:252-272): 180 GB hold ~7e9 rows of the C2 shape, and a query
// over a resident region moves no input over PCIe at all.  Keyed by (device, region id); re-registering an id replaces it.
namespace {
struct ResidentRegion { int64_t nrows = 0; size_t bytes = 0; std::vector<bkgpu_column> cols; std::vector<void*> allocs; };
std::mutex g_region_mu;
std::map<std::pair<int, int64_t>, ResidentRegion> g_regions;
void free_region(ResidentRegion& r) { for (void* q : r.allocs) cudaFree(q); r.allocs.clear(); r.cols.clear(); r.bytes = 0; }
}  // namespace

extern "C" int bkgpu_region_register(int device, int64_t region_id, const bkgpu_column* cols, int ncols, int64_t nrows, int on_device) {
    if (!cols || ncols <= 0 || nrows < 0) return thread_fail(BKGPU_EINVAL, "bkgpu_region_register: bad arguments");
    if (cudaSetDevice(device) != cudaSuccess) return thread_fail(BKGPU_ENODEV, "cudaSetDevice failed");
    ResidentRegion r; r.nrows = nrows;
    cudaStream_t st = nullptr;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return thread_fail(BKGPU_ENODEV, "cudaStreamCreate failed");
    int rc = BKGPU_OK;
    for (int i = 0; i < ncols && rc == BKGPU_OK; i++) {
        const bkgpu_column& c = cols[i];
        const int stype = prim_storage(c.prim_type);
        if (stype < 0 || c.length != nrows || (nrows > 0 && !c.values)) { rc = thread_fail(BKGPU_EINVAL, "bkgpu_region_register: column %d_%d is malformed", c.tuple_id, c.slot_id); break; }
        const size_t vb = (size_t)nrows * (size_t)storage_bytes(stype), nb = c.validity ? (size_t)(nrows + 7) / 8 : 0;
        void *dv = nullptr, *dn = nullptr;
        if (cudaMalloc(&dv, vb + 64) != cudaSuccess) { rc = thread_fail(BKGPU_ENOMEM, "cudaMalloc of %zu bytes failed", vb); break; }
        r.allocs.push_back(dv);
        if (nb) { if (cudaMalloc(&dn, nb + 64) != cudaSuccess) { rc = thread_fail(BKGPU_ENOMEM, "cudaMalloc of %zu bytes failed", nb); break; } r.allocs.push_back(dn); }
        const cudaMemcpyKind k = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
        if (vb && cudaMemcpyAsync(dv, c.values, vb, k, st) != cudaSuccess) { rc = thread_fail(BKGPU_ENODEV, "copy of column %d_%d failed", c.tuple_id, c.slot_id); break; }
        if (nb && cudaMemcpyAsync(dn, c.validity, nb, k, st) != cudaSuccess) { rc = thread_fail(BKGPU_ENODEV, "copy of column %d_%d failed", c.tuple_id, c.slot_id); break; }
        bkgpu_column d = c; d.values = dv; d.validity = (const uint8_t*)dn;
        r.cols.push_back(d); r.bytes += vb + nb;
    }
    if (rc == BKGPU_OK && cudaStreamSynchronize(st) != cudaSuccess) rc = thread_fail(BKGPU_ENODEV, "bkgpu_region_register: copy failed");
    cudaStreamDestroy(st);
    if (rc != BKGPU_OK) { free_region(r); return rc; }
    std::lock_guard<std::mutex> g(g_region_mu);
    auto key = std::make_pair(device, region_id);
    auto it = g_regions.find(key);
    if (it != g_regions.end()) { free_region(it->second); g_regions.erase(it); }
    g_regions.emplace(key, std::move(r));
    return BKGPU_OK;
}
extern "C" int bkgpu_region_evict(int device, int64_t region_id) {
    std::lock_guard<std::mutex> g(g_region_mu);
    auto it = g_regions.find(std::make_pair(device, region_id));
    if (it == g_regions.end()) return thread_fail(BKGPU_EINVAL, "region %lld is not resident on device %d", (long long)region_id, device);
    cudaSetDevice(device);
    free_region(it->second);
    g_regions.erase(it);
    return BKGPU_OK;
}
extern "C" int bkgpu_region_info(int device, int64_t region_id, int64_t* nrows, size_t* bytes) {
    std::lock_guard<std::mutex> g(g_region_mu);
    auto it = g_regions.find(std::make_pair(device, region_id));
    if (it == g_regions.end()) return thread_fail(BKGPU_EINVAL, "region %lld is not resident on device %d", (long long)region_id, device);
    if (nrows) *nrows = it->second.nrows;
    if (bytes) *bytes = it->second.bytes;
    return BKGPU_OK;
}
// child->get_next() over a resident region: the same as bkgpu_push of its columns with on_device = 1
extern "C" int bkgpu_push_region(bkgpu_plan* p, int64_t region_id) {
    if (!p) return thread_fail(BKGPU_EINVAL, "bkgpu_push_region: NULL plan");
    std::vector<bkgpu_column> cols; int64_t nrows = 0;
    {
        std::lock_guard<std::mutex> g(g_region_mu);
        auto it = g_regions.find(std::make_pair(p->device, region_id));
        if (it == g_regions.end()) return p->fail(BKGPU_EINVAL, "region %lld is not resident on device %d", (long long)region_id, p->device);
        cols = it->second.cols; nrows = it->second.nrows;
    }
    return bkgpu_push(p, cols.data(), (int)cols.size(), nrows, 1);
}

// ------------------------------------------------------------------ partial state
extern "C" int bkgpu_partial_capacity(bkgpu_plan* p, size_t* bytes) {
    if (!p || !bytes) return thread_fail(BKGPU_EINVAL, "bkgpu_partial_capacity: NULL argument");
    if (p->c.kind == PK_AGG || p->c.kind == PK_JOIN_AGG) { *bytes = 8 * (1 + (size_t)(p->c.ap.n_keyw + p->c.ap.n_lanes) * (size_t)eff_pcap(p)); return BKGPU_OK; }
    if (p->c.kind == PK_SORT && p->sort) { *bytes = sort_partial_bytes(p->sort); return BKGPU_OK; }
    return p->fail(BKGPU_EUNSUPPORTED, "plan kind %d has no partial state", p->c.kind);
}
extern "C" int bkgpu_partial_export(bkgpu_plan* p, void* dev_dst, size_t bytes) {
    if (!p || !dev_dst) return thread_fail(BKGPU_EINVAL, "bkgpu_partial_export: NULL argument");
    if (p->state != S_OPEN && p->state != S_FINISHED) return p->fail(BKGPU_ESTATE, "bkgpu_partial_export needs an open plan");
    size_t need = 0; int rc = bkgpu_partial_capacity(p, &need);
    if (rc) return rc;
    if (bytes < need) return p->fail(BKGPU_EINVAL, "partial buffer too small: %zu < %zu", bytes, need);
    CK(p, cudaSetDevice(p->device));
    if (p->c.kind == PK_SORT) { rc = sort_partial_export(p->sort, dev_dst, p->stream, p->last_error); if (rc) g_thread_error = p->last_error; return rc; }
    // compact rows (the layout the in-library all-gather ships): [u64 groups][groups x (key words + lane words)]
    CK(p, launch_partial_export_rows(p->gt, p->c.ap, (uint64_t*)dev_dst, eff_pcap(p), p->stream));
    p->stats.kernel_launches += 1;
    uint64_t held = 0; uint32_t ov = 0;
    CK(p, cudaMemcpyAsync(&held, dev_dst, 8, cudaMemcpyDeviceToHost, p->stream));
    CK(p, cudaMemcpyAsync(&ov, p->gt.overflow, 4, cudaMemcpyDeviceToHost, p->stream));
    CK(p, cudaStreamSynchronize(p->stream));
    if (ov) return p->fail(BKGPU_ETOOBIG, "group table overflow (capacity 2^%d slots): raise group_capacity_log2", (int)p->gt.cap_log2);
    if (held > eff_pcap(p)) return p->fail(BKGPU_ETOOBIG, "this rank holds %llu groups, more than partial_capacity=%lld", (unsigned long long)held, (long long)p->partial_cap);
    return BKGPU_OK;
}
extern "C" int bkgpu_partial_merge(bkgpu_plan* p, const void* dev_src, size_t bytes_per_rank, int nranks) {
    if (!p || !dev_src || nranks < 1) return thread_fail(BKGPU_EINVAL, "bkgpu_partial_merge: bad argument");
    if (p->state != S_OPEN && p->state != S_FINISHED) return p->fail(BKGPU_ESTATE, "bkgpu_partial_merge needs an open plan");
    size_t need = 0; int rc = bkgpu_partial_capacity(p, &need);
    if (rc) return rc;
    if (bytes_per_rank != need) return p->fail(BKGPU_EINVAL, "bytes_per_rank %zu != partial capacity %zu", bytes_per_rank, need);
    CK(p, cudaSetDevice(p->device));
    if (p->c.kind == PK_SORT) {
        std::vector<SortOutCol> cols; int64_t rows = 0;
        rc = sort_partial_merge(p->sort, dev_src, nranks, p->stream, cols, &rows, p->last_error);
        if (rc) { g_thread_error = p->last_error; return rc; }
        p->result.clear();
        for (auto& sc : cols) { HostCol hc; hc.desc = {sc.tuple_id, sc.slot_id, sc.prim, 0}; hc.elem = sc.elem; hc.values = std::move(sc.values); hc.validity = std::move(sc.validity); p->result.push_back(std::move(hc)); }
        p->result_rows = rows; p->result_pos = 0; p->stats.rows_returned = rows; p->state = S_FINISHED;
        return BKGPU_OK;
    }
    CK(p, launch_table_init(p->gt, p->c.ap, p->stream, 1));
    CK(p, launch_partial_merge_rows(p->gt, p->c.ap, (const uint64_t*)dev_src, need / 8, eff_pcap(p), nranks, -1 /* fold every segment */, p->d_cursor + 1, p->stream));
    p->stats.kernel_launches += 2;
    void* comm = p->nccl_comm; int nr = p->nranks; p->nccl_comm = nullptr; p->nranks = 1;  // already merged: finalize locally
    rc = agg_finish(p);
    p->nccl_comm = comm; p->nranks = nr;
    if (rc) return rc;
    p->stats.rows_returned = p->result_rows; p->state = S_FINISHED;
    return BKGPU_OK;
}

// ------------------------------------------------------------------ NCCL plumbing
extern "C" int bkgpu_nccl_unique_id(uint8_t id_out[128]) {
    if (!id_out) return thread_fail(BKGPU_EINVAL, "bkgpu_nccl_unique_id: NULL");
    if (nccl_unique_id(id_out) != 0) return thread_fail(BKGPU_ENCCL, "%s", nccl_last_error());
    return BKGPU_OK;
}
extern "C" int bkgpu_nccl_comm_create(void** comm_out, const uint8_t id[128], int nranks, int rank, int device) {
    if (!comm_out || !id) return thread_fail(BKGPU_EINVAL, "bkgpu_nccl_comm_create: NULL");
    if (cudaSetDevice(device) != cudaSuccess) return thread_fail(BKGPU_ENODEV, "cudaSetDevice failed");
    if (nccl_comm_create(comm_out, id, nranks, rank) != 0) return thread_fail(BKGPU_ENCCL, "%s", nccl_last_error());
    return BKGPU_OK;
}
extern "C" void bkgpu_nccl_comm_destroy(void* comm) { if (comm) nccl_comm_destroy(comm); }

// ------------------------------------------------------------------ memory helpers
extern "C" void* bkgpu_host_alloc(size_t bytes) { void* p = nullptr; return cudaHostAlloc(&p, bytes ? bytes : 8, cudaHostAllocDefault) == cudaSuccess ? p : nullptr; }
extern "C" void bkgpu_host_free(void* p) { if (p) cudaFreeHost(p); }
extern "C" void* bkgpu_device_alloc(int device, size_t bytes) {
    void* p = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    return cudaMalloc(&p, bytes ? bytes : 8) == cudaSuccess ? p : nullptr;
}
extern "C" void bkgpu_device_free(int device, void* p) { if (p && cudaSetDevice(device) == cudaSuccess) cudaFree(p); }
extern "C" int bkgpu_memcpy_h2d(int device, void* dst, const void* src, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess || cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return thread_fail(BKGPU_ENODEV, "cudaMemcpy H2D failed");
    return BKGPU_OK;
}
extern "C" int bkgpu_memcpy_d2h(int device, void* dst, const void* src, size_t bytes) {
    if (cudaSetDevice(device) != cudaSuccess || cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return thread_fail(BKGPU_ENODEV, "cudaMemcpy D2H failed");
    return BKGPU_OK;
}
