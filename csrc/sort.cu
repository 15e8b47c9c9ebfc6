// sort.cu — K5 (ORDER BY / top-k) and K1 (filter-only stream compaction) for sm_100a.
//
// Replaces SortNode::open + Sorter / TopNSorter (src/exec/sort_node.cpp:278-346, src/runtime/sorter.cpp:54-114,
// src/runtime/topn_sorter.cpp:25-103) and the row-copy half of FilterNode::get_next (src/exec/filter_node.cpp:736-795).
//
// Ordering is the reference's MemRowCompare (src/mem_row/mem_row_compare.cpp:18-38) made total: every row
// carries the composite (class, key image, arrival index) where class orders NULL keys first / last
// (is_null_first), the key image is an order-preserving 64-bit transform (descending = complemented)
// and the arrival index breaks ties the way TopNSorter does (include/runtime/topn_sorter.h:96-106).
//
// ORDER BY ... LIMIT k (one key): selection, not sorting.  Per batch:
//   sample 4096 random positions -> threshold at a rank that leaves >= k rows below it with overwhelming
//   probability -> ONE pass over the key column compacts the candidates (warp-ballot compaction, 8 B/row) ->
//   the same step recurses on the shrinking candidate arrays until <= 8192 remain -> one CTA sorts them in
//   shared memory together with the k rows kept from earlier batches -> payload columns are gathered for
//   the k survivors only.
// ORDER BY without LIMIT / several keys: LSD radix sort (8 bits per pass, stable) of (key image, row id),
//   least significant ORDER BY key first, then a gather of every column.
// Filter-only fragments: predicate -> ordered compaction (block scan) -> gather, honouring LIMIT in input order.
#include "sort.h"
#include <algorithm>
#include <string.h>
#include "dev_common.cuh"
#include "interp.cuh"
#include "nccl_dl.h"

namespace bk {

namespace {

constexpr int SMALL_N = 8192;        // one CTA sorts this many composite keys in shared memory
constexpr int SAMPLE_N = 4096;

struct KeySpec {       // how a row's sort key is obtained on the device
    int direct;        // 1: plain column `col` of class `vclass`; 0: program output register `out_reg`
    int col, vclass, out_reg;
    int desc;          // ORDER BY ... DESC
    int pred_out;      // predicate register (-1 = none); generic path only
};

struct RowArgs {
    DevCol cols[MAX_COLS];
    int32_t n_cols;
    int64_t nrows;
    uint64_t row_base;   // arrival index of row 0 of this batch
    Program prog;
    KeySpec key;
};

// order-preserving 64-bit image (ascending); DESC complements it
__device__ __forceinline__ uint64_t key_image(uint64_t v, int vclass, int desc) {
    uint64_t t;
    if (vclass == VC_F64) t = (v >> 63) ? ~v : (v | 0x8000000000000000ull);
    else if (vclass == VC_I64) t = v ^ 0x8000000000000000ull;
    else t = v;
    return desc ? ~t : t;
}

// cls: 0 = row filtered out, 1 = key present, 2 = key is NULL
__device__ __forceinline__ int row_key(const RowArgs& a, int64_t row, uint64_t& image) {
    if (a.key.direct) {
        const DevCol& c = a.cols[a.key.col];
        if (elem_is_null(c, row)) { image = 0; return 2; }
        image = key_image(load_elem(c, row), a.key.vclass, a.key.desc);
        return 1;
    }
    uint64_t out[8]; uint32_t out_null;
    run_program(a.prog, a.cols, row, out, out_null);
    if (a.key.pred_out >= 0 && (((out_null >> a.key.pred_out) & 1u) || out[a.key.pred_out] == 0)) return 0;
    if ((out_null >> a.key.out_reg) & 1u) { image = 0; return 2; }
    image = key_image(out[a.key.out_reg], a.key.vclass, a.key.desc);
    return 1;
}

__device__ __forceinline__ bool comp_le(uint64_t k, uint64_t i, uint64_t tk, uint64_t ti) { return k < tk || (k == tk && i <= ti); }
__device__ __forceinline__ bool comp_lt(uint64_t k, uint64_t i, uint64_t tk, uint64_t ti) { return k < tk || (k == tk && i < ti); }

__device__ __forceinline__ uint64_t mix64d(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}

// ---- level-1 kernels: straight from the batch columns ----
// sample: position i of SAMPLE_N is drawn uniformly inside its stratum [i*n/S, (i+1)*n/S)
__global__ void k_sample_rows(RowArgs a, int want_cls, uint64_t* skey, uint64_t* sidx, uint32_t* scount, uint64_t salt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= SAMPLE_N) return;
    const int64_t lo = (int64_t)(((__int128)i * a.nrows) / SAMPLE_N), hi = (int64_t)(((__int128)(i + 1) * a.nrows) / SAMPLE_N);
    if (hi <= lo) return;
    const int64_t row = lo + (int64_t)(mix64d(salt + i) % (uint64_t)(hi - lo));
    uint64_t img;
    if (row_key(a, row, img) != want_cls) return;
    const uint32_t pos = atomicAdd(scount, 1u);
    skey[pos] = img; sidx[pos] = a.row_base + (uint64_t)row;
}

// compaction of the rows whose composite key is <= (tk, ti): warp-ballot positions, one atomic per warp
__global__ void __launch_bounds__(256) k_collect_rows(RowArgs a, int want_cls, const uint64_t* thr, uint64_t* okey, uint64_t* oidx,
                                                      uint32_t* ocount, uint32_t cap, uint32_t* class_count) {
    const int lane = threadIdx.x & 31;
    const uint64_t tk = thr[0], ti = thr[1];   // threshold produced on the device by k_sort_small: no host round trip
    uint32_t seen = 0;
    for (int64_t base = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x - lane); base < a.nrows; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = base + lane;
        uint64_t img = 0; bool take = false;
        if (row < a.nrows) {
            const int cls = row_key(a, row, img);
            if (cls == want_cls) { seen++; take = comp_le(img, a.row_base + (uint64_t)row, tk, ti); }
        }
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, take);
        if (b) {
            uint32_t wbase = 0;
            if (lane == 0) wbase = atomicAdd(ocount, (uint32_t)__popc(b));
            wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
            const uint32_t pos = wbase + __popc(b & ((1u << lane) - 1u));
            if (take && pos < cap) { okey[pos] = img; oidx[pos] = a.row_base + (uint64_t)row; }
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) seen += __shfl_xor_sync(0xFFFFFFFFu, seen, d);
    if (lane == 0 && seen && class_count) atomicAdd(class_count, seen);
}

// level 1 for the direct shape (plain NULL-free key column, config C5): four keys per lane and load (LDG.256 / LDG.128),
// two loads in flight; candidates are rare (~1 % of the rows), so the compaction slow path is off the hot loop
template <int KEY_BYTES>
__global__ void __launch_bounds__(256) k_collect_rows_vec(RowArgs a, const uint64_t* thr, uint64_t* okey, uint64_t* oidx, uint32_t* ocount,
                                                          uint32_t cap, uint32_t* class_count) {
    // candidates are staged per CTA in shared memory and leave with ONE global atomic per CTA: ~1M candidates
    // hammering a single global counter cost 0.8 ms on their own (profiles/r01_topk_history.md)
    constexpr uint32_t STAGE = 2048;
    __shared__ uint64_t s_key[STAGE], s_idx[STAGE];
    __shared__ uint32_t s_n, s_base;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const uint64_t tk = thr[0], ti = thr[1];
    const DevCol& c = a.cols[a.key.col];
    const int vclass = a.key.vclass, desc = a.key.desc;
    const bool sext = c.stype == ST_I32;
    const int64_t nquads = a.nrows >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    auto emit = [&](uint64_t img, uint64_t idx) {
        const uint32_t pos = atomicAdd(&s_n, 1u);
        if (pos < STAGE) { s_key[pos] = img; s_idx[pos] = idx; }
        else { const uint32_t g = atomicAdd(ocount, 1u); if (g < cap) { okey[g] = img; oidx[g] = idx; } }   // stage full: rare
    };
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += 2 * stride) {
        uint64_t v[8];
        const bool second = q + stride < nquads;
        if (KEY_BYTES == 8) {
            const U64x4 r0 = ldg256_u64((const uint8_t*)c.values + q * 32);
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = r0.v[j];
            if (second) { const U64x4 r1 = ldg256_u64((const uint8_t*)c.values + (q + stride) * 32);
#pragma unroll
                for (int j = 0; j < 4; j++) v[4 + j] = r1.v[j]; }
        } else {
            const U32x4 r0 = ldg128_u32((const uint8_t*)c.values + q * 16);
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = sext ? (uint64_t)(int64_t)(int32_t)r0.v[j] : (uint64_t)r0.v[j];
            if (second) { const U32x4 r1 = ldg128_u32((const uint8_t*)c.values + (q + stride) * 16);
#pragma unroll
                for (int j = 0; j < 4; j++) v[4 + j] = sext ? (uint64_t)(int64_t)(int32_t)r1.v[j] : (uint64_t)r1.v[j]; }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j >= 4 && !second) break;
            const uint64_t img = key_image(v[j], vclass, desc);
            if (img > tk) continue;                                    // the common case: not a candidate
            const uint64_t idx = a.row_base + (uint64_t)((j < 4 ? q : q + stride) * 4 + (j & 3));
            if (comp_le(img, idx, tk, ti)) emit(img, idx);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (a.nrows & 3)) {   // ragged tail
        const int64_t row = (a.nrows & ~(int64_t)3) + threadIdx.x;
        const uint64_t img = key_image(load_elem(c, row), vclass, desc);
        if (comp_le(img, a.row_base + (uint64_t)row, tk, ti)) emit(img, a.row_base + (uint64_t)row);
    }
    __syncthreads();
    const uint32_t n = min(s_n, STAGE);
    if (threadIdx.x == 0) s_base = n ? atomicAdd(ocount, n) : 0;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) { const uint32_t g = s_base + i; if (g < cap) { okey[g] = s_key[i]; oidx[g] = s_idx[i]; } }
    if (blockIdx.x == 0 && threadIdx.x == 0 && class_count) atomicAdd(class_count, (uint32_t)min(a.nrows, (int64_t)0xFFFFFFFFll));  // no NULLs, no filter: every row is of this class
}

// ---- level >= 2 kernels: on (key, idx) candidate arrays ----
__global__ void k_sample_pairs(const uint64_t* key, const uint64_t* idx, const uint32_t* n_ptr, uint32_t cap, uint64_t* skey, uint64_t* sidx, uint32_t* scount, uint64_t salt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= SAMPLE_N) return;
    const uint32_t n = min(*n_ptr, cap);
    const uint64_t lo = (uint64_t)i * n / SAMPLE_N, hi = (uint64_t)(i + 1) * n / SAMPLE_N;
    if (hi <= lo) return;
    const uint64_t p = lo + mix64d(salt + i) % (hi - lo);
    const uint32_t pos = atomicAdd(scount, 1u);
    skey[pos] = key[p]; sidx[pos] = idx[p];
}
__global__ void __launch_bounds__(256) k_collect_pairs(const uint64_t* key, const uint64_t* idx, const uint32_t* n_ptr, const uint64_t* thr,
                                                       uint64_t* okey, uint64_t* oidx, uint32_t* ocount, uint32_t cap) {
    const int lane = threadIdx.x & 31;
    const uint32_t n = min(*n_ptr, cap);
    const uint64_t tk = thr[0], ti = thr[1];
    for (uint32_t base = blockIdx.x * blockDim.x + threadIdx.x - lane; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t i = base + lane;
        const bool take = i < n && comp_le(key[i], idx[i], tk, ti);
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, take);
        if (b) {
            uint32_t wbase = 0;
            if (lane == 0) wbase = atomicAdd(ocount, (uint32_t)__popc(b));
            wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
            const uint32_t pos = wbase + __popc(b & ((1u << lane) - 1u));
            if (take && pos < cap) { okey[pos] = key[i]; oidx[pos] = idx[i]; }
        }
    }
}

// one CTA: bitonic sort of n <= SMALL_N composite keys in shared memory (ascending); optionally reads a
// second segment (the rows kept from earlier batches).  Writes the first `keep` entries and the r-th (0-based) one.
__global__ void __launch_bounds__(1024) k_sort_small(const uint64_t* k1, const uint64_t* i1, const uint32_t* n1p, uint32_t n1max,
                                                     const uint64_t* k2, const uint64_t* i2, uint32_t n2,
                                                     uint64_t* okey, uint64_t* oidx, uint32_t keep, uint32_t* out_n,
                                                     uint32_t rank, uint64_t* rank_out,
                                                     const uint32_t* pop_ptr, uint32_t k_want, uint32_t widen, uint32_t room, const uint64_t* cap_thr) {
    extern __shared__ __align__(16) unsigned char sm[];
    uint64_t* sk = (uint64_t*)sm; uint64_t* si = sk + SMALL_N;
    uint32_t n1 = n1p ? *n1p : n1max; if (n1 > n1max) n1 = n1max;
    const uint32_t n = n1 + n2;
    uint32_t m = 1; while (m < n) m <<= 1; if (m < 2) m = 2;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        if (i < n1) { sk[i] = k1[i]; si[i] = i1[i]; }
        else if (i < n) { sk[i] = k2[i - n1]; si[i] = i2[i - n1]; }
        else { sk[i] = ~0ull; si[i] = ~0ull; }
    }
    __syncthreads();
    for (uint32_t size = 2; size <= m; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                const uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                const bool up = (lo & size) == 0;
                const uint64_t ka = sk[lo], ia = si[lo], kb = sk[hi], ib = si[hi];
                const bool a_gt_b = comp_lt(kb, ib, ka, ia);
                if (a_gt_b == up) { sk[lo] = kb; si[lo] = ib; sk[hi] = ka; si[hi] = ia; }
            }
            __syncthreads();
        }
    }
    const uint32_t nk = n < keep ? n : keep;
    if (okey) for (uint32_t i = threadIdx.x; i < nk; i += blockDim.x) { okey[i] = sk[i]; oidx[i] = si[i]; }
    if (threadIdx.x == 0) {
        if (out_n) *out_n = nk;
        if (rank_out) {
            // threshold for the next compaction: the sample (n entries of a population of *pop_ptr) at a rank that
            // leaves >= k_want population members below it with overwhelming probability; +inf when the population
            // already fits `room` or the sample is too small to say anything
            uint64_t tk = ~0ull, ti = ~0ull;
            uint32_t r = rank;
            if (pop_ptr) {
                const double pop = (double)*pop_ptr;
                if (pop <= (double)room) r = 0xFFFFFFFFu;
                else { double w = ((double)k_want * (double)SAMPLE_N / pop * 1.5 + 32.0) * (double)widen; r = w >= (double)(SAMPLE_N - 1) ? 0xFFFFFFFFu : (uint32_t)w; }
            }
            if (n > 0 && r < n) { tk = sk[r]; ti = si[r]; }
            if (cap_thr && comp_lt(cap_thr[0], cap_thr[1], tk, ti)) { tk = cap_thr[0]; ti = cap_thr[1]; }   // never looser than the k-th row kept so far
            rank_out[0] = tk; rank_out[1] = ti; rank_out[2] = n;
        }
    }
}

// one CTA: the r-th smallest (0-based) of n <= SAMPLE_N composite keys by radix selection — most significant byte
// first over the 16 bytes of (key, idx), a 256-bin shared histogram per byte, stopping as soon as the selected bin
// holds one element (random 64-bit keys: 2 rounds).  Same threshold rules as k_sort_small's rank output; replaces
// a full bitonic sort of the sample (~70 us) on the top-k chain (profiles/r01_topk_history.md).
__global__ void __launch_bounds__(1024) k_select_rank(const uint64_t* key, const uint64_t* idx, const uint32_t* np, uint32_t nmax,
                                                      uint32_t rank, uint64_t* rank_out,
                                                      const uint32_t* pop_ptr, uint32_t k_want, uint32_t widen, uint32_t room, const uint64_t* cap_thr) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_bin, s_rank, s_cnt;
    __shared__ uint64_t s_res[2];
    constexpr int PER = SAMPLE_N / 1024;
    uint32_t n = *np; if (n > nmax) n = nmax; if (n > (uint32_t)SAMPLE_N) n = SAMPLE_N;
    uint32_t r = rank;
    if (pop_ptr) {
        const double pop = (double)*pop_ptr;
        if (pop <= (double)room) r = 0xFFFFFFFFu;
        else { double w = ((double)k_want * (double)SAMPLE_N / pop * 1.5 + 32.0) * (double)widen; r = w >= (double)(SAMPLE_N - 1) ? 0xFFFFFFFFu : (uint32_t)w; }
    }
    if (threadIdx.x == 0) { s_res[0] = ~0ull; s_res[1] = ~0ull; }
    if (n > 0 && r < n) {   // (uniform across the CTA)
        uint64_t k[PER], ix[PER]; uint32_t alive = 0;
#pragma unroll
        for (int e = 0; e < PER; e++) {
            const uint32_t i = threadIdx.x + e * 1024;
            k[e] = 0; ix[e] = 0;
            if (i < n) { k[e] = key[i]; ix[e] = idx[i]; alive |= 1u << e; }
        }
        uint32_t rr = r;
        for (int byte = 15; byte >= 0; byte--) {
            if (threadIdx.x < 256) hist[threadIdx.x] = 0;
            __syncthreads();
            uint32_t dig[PER];
#pragma unroll
            for (int e = 0; e < PER; e++) {
                dig[e] = (uint32_t)((byte >= 8 ? k[e] >> ((byte - 8) * 8) : ix[e] >> (byte * 8)) & 255ull);
                if ((alive >> e) & 1u) atomicAdd(&hist[dig[e]], 1u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {   // warp 0: the bin that holds rank rr
                uint32_t h[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; j++) { h[j] = hist[threadIdx.x * 8 + j]; sum += h[j]; }
                uint32_t incl = sum;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d); if ((int)threadIdx.x >= d) incl += o; }
                uint32_t cum = incl - sum;
                if (rr >= cum && rr < incl) {
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (rr < cum + h[j]) { s_bin = threadIdx.x * 8 + j; s_rank = rr - cum; s_cnt = h[j]; break; }
                        cum += h[j];
                    }
                }
            }
            __syncthreads();
            const uint32_t bin = s_bin;
#pragma unroll
            for (int e = 0; e < PER; e++) if (dig[e] != bin) alive &= ~(1u << e);
            rr = s_rank;
            if (s_cnt == 1 || byte == 0) break;
        }
#pragma unroll
        for (int e = 0; e < PER; e++) if ((alive >> e) & 1u) { s_res[0] = k[e]; s_res[1] = ix[e]; }   // one element (or identical ones)
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t tk = s_res[0], ti = s_res[1];
        uint64_t capped = 0;
        if (cap_thr && comp_lt(cap_thr[0], cap_thr[1], tk, ti)) { tk = cap_thr[0]; ti = cap_thr[1]; capped = 1ull << 62; }   // never looser than the k-th row kept so far
        rank_out[0] = tk; rank_out[1] = ti; rank_out[2] = (uint64_t)n | capped;   // bit 62: the threshold in force is the pool's k-th key
    }
}

// ---- payload: gather rows of the current batch / move rows kept from earlier batches ----
struct GatherArgs {
    DevCol cols[MAX_COLS]; int32_t n_cols; uint64_t row_base; int64_t nrows;
    uint8_t* dst_vals[MAX_COLS]; uint8_t* dst_null[MAX_COLS];
    const uint8_t* old_vals[MAX_COLS]; const uint8_t* old_null[MAX_COLS];
    const uint64_t* old_idx; uint32_t old_n;
};
__global__ void k_gather_topk(GatherArgs g, const uint64_t* idx, const uint32_t* np) {
    const uint32_t n = *np;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t id = idx[i];
        const bool from_batch = g.nrows > 0 && id >= g.row_base && id < g.row_base + (uint64_t)g.nrows;
        int64_t src = -1;
        if (from_batch) src = (int64_t)(id - g.row_base);
        else {  // binary search in the previous survivors (sorted by composite key, not by idx): linear scan is fine for k rows
            for (uint32_t j = 0; j < g.old_n; j++) if (g.old_idx[j] == id) { src = j; break; }
        }
        for (int c = 0; c < g.n_cols; c++) {
            const int eb = g.cols[c].stype == ST_U8 ? 1 : (g.cols[c].stype == ST_BLOB16 ? 16 : ((g.cols[c].stype == ST_I32 || g.cols[c].stype == ST_U32 || g.cols[c].stype == ST_F32) ? 4 : 8));
            const uint8_t* sv; uint8_t isnull;
            if (from_batch) { sv = (const uint8_t*)g.cols[c].values + (size_t)src * eb; isnull = elem_is_null(g.cols[c], src) ? 1 : 0; }
            else { sv = g.old_vals[c] + (size_t)src * eb; isnull = g.old_null[c][src]; }
            uint8_t* dv = g.dst_vals[c] + (size_t)i * eb;
            for (int b = 0; b < eb; b++) dv[b] = sv[b];
            g.dst_null[c][i] = isnull;
        }
    }
}

// ---- filter-only: ordered compaction.  Pass 1 counts per block, pass 2 (after a scan of the counts) writes. ----
__global__ void __launch_bounds__(256) k_filter_count(RowArgs a, uint32_t* block_counts, int64_t rows_per_block) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, a.nrows);
    uint32_t cnt = 0;
    for (int64_t row = r0 + threadIdx.x; row < r1; row += blockDim.x) {
        uint64_t img; const int cls = a.key.pred_out < 0 ? 1 : row_key(a, row, img);
        cnt += cls != 0;
    }
    __shared__ uint32_t s[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int i = 0; i < 8; i++) t += s[i]; block_counts[blockIdx.x] = t; }
}
__global__ void k_scan_counts(uint32_t* counts, uint32_t n, uint64_t* total) {  // single CTA exclusive scan (n <= a few thousand blocks)
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? counts[i] : 0;
        // inclusive warp scan
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, d); if ((threadIdx.x & 31) >= d) x += y; }
        __shared__ uint32_t ws[32];
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0, xw = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xFFFFFFFFu, xw, d); if (threadIdx.x >= d) xw += y; }
            ws[threadIdx.x] = xw - w;
        }
        __syncthreads();
        const uint64_t excl = carry + ws[threadIdx.x >> 5] + (x - v);
        if (i < n) counts[i] = (uint32_t)excl;  // offsets fit 32 bits per launch (<= 2^31 rows)
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) k_filter_write(RowArgs a, const uint32_t* block_offsets, int64_t rows_per_block, uint32_t* sel, uint64_t out_base, uint64_t cap) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, a.nrows);
    __shared__ uint32_t warp_cnt[8];
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = block_offsets[blockIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = r0; base < r1; base += blockDim.x) {
        const int64_t row = base + threadIdx.x;
        bool take = false;
        if (row < r1) { uint64_t img; take = a.key.pred_out < 0 ? true : row_key(a, row, img) != 0; }
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, take);
        if (lane == 0) warp_cnt[warp] = __popc(b);
        __syncthreads();
        uint32_t before = running;
        for (int w = 0; w < warp; w++) before += warp_cnt[w];
        const uint64_t pos = out_base + before + __popc(b & ((1u << lane) - 1u));
        if (take && pos < cap) sel[pos] = (uint32_t)row;
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < 8; w++) t += warp_cnt[w]; running += t; }
        __syncthreads();
    }
}
__global__ void k_gather_sel(GatherArgs g, const uint32_t* sel, uint64_t n, uint64_t dst_off) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int64_t src = sel[i];
        for (int c = 0; c < g.n_cols; c++) {
            const int eb = g.cols[c].stype == ST_U8 ? 1 : (g.cols[c].stype == ST_BLOB16 ? 16 : ((g.cols[c].stype == ST_I32 || g.cols[c].stype == ST_U32 || g.cols[c].stype == ST_F32) ? 4 : 8));
            const uint8_t* sv = (const uint8_t*)g.cols[c].values + (size_t)src * eb;
            uint8_t* dv = g.dst_vals[c] + (size_t)(dst_off + i) * eb;
            if (eb == 8) *(uint64_t*)dv = *(const uint64_t*)sv; else if (eb == 4) *(uint32_t*)dv = *(const uint32_t*)sv;
            else if (eb == 16) { ((uint64_t*)dv)[0] = ((const uint64_t*)sv)[0]; ((uint64_t*)dv)[1] = ((const uint64_t*)sv)[1]; }   // AVG intermediate blob (payload only)
            else *dv = *sv;
            g.dst_null[c][dst_off + i] = elem_is_null(g.cols[c], src) ? 1 : 0;
        }
    }
}

// ---- full sort: materialise (image, row id), LSD radix sort, gather ----
// mode 0: order-preserving image of the key (NULL -> 0); mode 1: the NULL rank (0 sorts first) — sorted by one
// extra radix pass AFTER the eight value passes so NULL keys come strictly first / last (is_null_first)
__global__ void k_sort_images(RowArgs a, int key_reg, int vclass, int desc, int null_first, int mode, uint64_t* img, uint32_t* rid) {
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < a.nrows; row += (int64_t)gridDim.x * blockDim.x) {
        uint64_t out[8]; uint32_t out_null;
        run_program(a.prog, a.cols, row, out, out_null);
        const bool isnull = (out_null >> key_reg) & 1u;
        if (mode == 0) img[row] = isnull ? 0ull : key_image(out[key_reg], vclass, desc);
        else img[row] = isnull ? (null_first ? 0ull : 1ull) : (null_first ? 1ull : 0ull);
        if (rid) rid[row] = (uint32_t)row;
    }
}
// ------------------------------------------------------------------------------------------------------------
// Full ORDER BY (Sorter::sort, src/runtime/sorter.cpp:54-114): stable LSD radix sort of (key image, row id) PAIRS, 8 bits per
// pass, one read and one write of the pairs per pass ("Onesweep": chained scan with decoupled look-back):
//   k_rs_hist : ONE pass over the images gives the histograms of all eight digits; a digit whose histogram has a single
//               non-empty bin is skipped (small-range keys, the NULL-rank image, descending flags ...)
//   k_rs_pass : a CTA takes the next 4096-pair tile (atomic ticket: tiles start in order, so the look-back never waits for a
//               tile that has not started), ranks its pairs per digit (warp match + per-warp counters: stable), publishes the
//               tile's digit counts, looks back over the preceding tiles' published counts for its scatter bases, reorders the
//               tile through shared memory and writes digit-contiguous (coalesced) runs
// The previous version gathered key[perm[i]] (a random 8-byte read per element and pass) and scattered 32 elements at a time.
// ------------------------------------------------------------------------------------------------------------
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 12;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 3072 pairs per CTA tile (36 KB of shared memory for the reorder): four CTAs per SM
constexpr uint32_t RS_FLAG_AGG = 1u << 30, RS_FLAG_PREFIX = 2u << 30, RS_VAL_MASK = (1u << 30) - 1u;

__global__ void __launch_bounds__(256) k_rs_hist(const uint64_t* key, uint32_t n, uint32_t* hist /* [8][256] */) {
    __shared__ uint32_t h[8][256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t k = key[i];
#pragma unroll
        for (int d = 0; d < 8; d++) atomicAdd(&h[d][(k >> (8 * d)) & 0xFF], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) { const uint32_t v = (&h[0][0])[i]; if (v) atomicAdd(hist + i, v); }
}
// per digit: exclusive scan of its histogram -> global scatter bases; active[d] = more than one non-empty bin
__global__ void k_rs_bases(const uint32_t* hist, uint32_t* bases, uint32_t* active) {
    const int d = blockIdx.x;
    __shared__ uint32_t s[256];
    const uint32_t v = hist[d * 256 + threadIdx.x];
    s[threadIdx.x] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0, nz = 0;
        for (int i = 0; i < 256; i++) { const uint32_t c = s[i]; s[i] = run; run += c; nz += c != 0; }
        active[d] = nz > 1;
    }
    __syncthreads();
    bases[d * 256 + threadIdx.x] = s[threadIdx.x];
}
// null BYTES (1 = NULL) of a retained column -> Arrow validity bitmap (bit set = valid)
__global__ void k_pack_null_bytes(const uint8_t* nb, int64_t n, uint8_t* bitmap, uint32_t* any_null) {
    const int64_t nbytes = (n + 7) / 8;
    bool seen = false;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbytes; b += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        for (int j = 0; j < 8; j++) { const int64_t r = b * 8 + j; if (r < n) { if (!nb[r]) v |= 1u << j; else seen = true; } }
        bitmap[b] = (uint8_t)v;
    }
    if (seen) *any_null = 1u;   // (lets the sort skip the NULL-rank pass of a key column without NULLs)
}
// key image of a plain column key, without the expression interpreter
__global__ void k_sort_images_direct(DevCol c, int64_t n, int vclass, int desc, uint64_t* img) {
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n; row += (int64_t)gridDim.x * blockDim.x)
        img[row] = elem_is_null(c, row) ? 0ull : key_image(load_elem(c, row), vclass, desc);
}
__global__ void k_rs_iota(uint32_t* ids, uint32_t n) { for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) ids[i] = i; }
__global__ void k_rs_gather_u64(const uint64_t* src, const uint32_t* ids, uint64_t* dst, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[ids[i]];
}
__global__ void k_rs_gather_bytes(const uint8_t* src, const uint32_t* ids, uint8_t* dst, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[ids[i]];
}

__global__ void __launch_bounds__(RS_THREADS, 4) k_rs_pass(const uint64_t* key_in, const uint32_t* id_in, uint64_t* key_out, uint32_t* id_out, uint32_t n, int shift,
                                                        const uint32_t* global_base /* [256] of this digit */, uint32_t* status /* [ntiles][256], zeroed */,
                                                        uint32_t* ticket) {
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t* s_key = (uint64_t*)rs_smem;                        // [RS_TILE]
    uint32_t* s_id = (uint32_t*)(s_key + RS_TILE);               // [RS_TILE]
    uint32_t* wcnt = s_id + RS_TILE;                             // [8][256] per-warp digit counts -> running offsets inside the digit's run
    uint32_t* dstart = wcnt + 8 * 256;                           // [256] start of the digit inside the reordered tile
    uint32_t* tbase = dstart + 256;                              // [256] global address of the tile's first pair of the digit
    __shared__ uint32_t s_tile;
    __shared__ uint32_t wsum[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    for (int i = threadIdx.x; i < 8 * 256; i += RS_THREADS) wcnt[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t t0 = tile * RS_TILE;
    const uint32_t tile_n = min((uint32_t)RS_TILE, n - t0);
    // ---- load: warp w owns the contiguous chunk [t0 + w * 32 * RS_ITEMS, + 32 * RS_ITEMS), RS_ITEMS rounds of 32 consecutive pairs ----
    uint64_t k[RS_ITEMS]; uint32_t id[RS_ITEMS];
    const uint32_t c0 = w * (RS_ITEMS * 32);
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const uint32_t li = c0 + r * 32 + lane;
        if (li < tile_n) { k[r] = key_in[t0 + li]; id[r] = id_in[t0 + li]; } else { k[r] = ~0ull; id[r] = 0; }
    }
    // ---- early counts: the warp's digit histogram (order-free atomics), so the tile's counts are published BEFORE the ranking and the
    //      following tiles' look-back finds them sooner ----
    uint32_t* wc = wcnt + w * 256;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++)
        if (c0 + r * 32 + lane < tile_n) atomicAdd(wc + (uint32_t)((k[r] >> shift) & 0xFF), 1u);
    __syncthreads();
    // ---- thread d: the digit's count per warp -> per-warp offsets, tile count; publish; where the digit starts in the reordered tile ----
    const int d_own = threadIdx.x;
    uint32_t cnt;
    volatile uint32_t* st = status + (size_t)tile * 256 + d_own;
    {
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < 8; ww++) { const uint32_t c = wcnt[ww * 256 + d_own]; wcnt[ww * 256 + d_own] = run; run += c; }
        cnt = run;
        *st = (tile == 0 ? RS_FLAG_PREFIX : RS_FLAG_AGG) | cnt;   // (flag and count travel in ONE word: no fence around the publication)
        uint32_t x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[w] = x;
        __syncthreads();
        uint32_t wpre = 0;
        for (int ww = 0; ww < w; ww++) wpre += wsum[ww];
        dstart[d_own] = wpre + x - cnt;
    }
    __syncthreads();
    // ---- rank in element order (stable) and place: peers of the digit in this round + the digit's running offset.  The peer mask
    //      comes from eight ballots, one per digit bit: MATCH.ANY kept the XU pipe 72 % busy (profiles/r02_radix_history.md) ----
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const bool live = c0 + r * 32 + lane < tile_n;
        const uint32_t d = (uint32_t)((k[r] >> shift) & 0xFF);
        uint32_t peers = __ballot_sync(0xFFFFFFFFu, live);
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const uint32_t bit = (d >> b) & 1u;
            peers &= ~(__ballot_sync(0xFFFFFFFFu, bit != 0) ^ (0u - bit));
        }
        uint32_t before = 0;
        if (live) before = wc[d];
        __syncwarp();
        const uint32_t mine = __popc(peers & lt_mask);
        if (live) {
            if ((peers >> lane) == 1u) wc[d] = before + mine + 1u;   // the highest peer lane advances the offset by the group's size
            const uint32_t pos = dstart[d] + before + mine;
            s_key[pos] = k[r]; s_id[pos] = id[r];
        }
        __syncwarp();
    }
    // ---- chained scan with decoupled look-back over the preceding tiles (they all started before this one: ticket order) ----
    {
        uint32_t excl = 0;
        if (tile != 0) {
            // four predecessors per round trip: the loads are independent, the walk consumes them nearest first
            for (int64_t p = (int64_t)tile - 1;;) {
                uint32_t v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = p - j >= 0 ? *(volatile const uint32_t*)(status + (size_t)(p - j) * 256 + d_own) : RS_FLAG_PREFIX;
                int j = 0; bool done = false;
#pragma unroll
                for (; j < 4; j++) {
                    if ((v[j] & ~RS_VAL_MASK) == 0) break;            // not published yet: poll again from this tile
                    excl += v[j] & RS_VAL_MASK;
                    if (v[j] & RS_FLAG_PREFIX) { done = true; break; }
                }
                if (done) break;
                p -= j;
            }
            *st = RS_FLAG_PREFIX | (excl + cnt);
        }
        tbase[d_own] = global_base[d_own] + excl;
    }
    __syncthreads();
    // ---- write the digit-contiguous runs ----
    for (uint32_t i = threadIdx.x; i < tile_n; i += RS_THREADS) {
        const uint64_t kk = s_key[i];
        const uint32_t d = (uint32_t)((kk >> shift) & 0xFF);
        const uint32_t dst = tbase[d] + (i - dstart[d]);
        key_out[dst] = kk; id_out[dst] = s_id[i];
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
struct Pool {               // the k best rows seen so far of one class (keys present / NULL keys)
    uint64_t* key[2] = {nullptr, nullptr}; uint64_t* idx[2] = {nullptr, nullptr};
    uint8_t* vals[2][MAX_COLS] = {}; uint8_t* nulls[2][MAX_COLS] = {};
    int cur = 0; uint32_t n = 0;
};

struct SortState {
    Compiled c;
    int device = 0, sm_count = 148;
    int ncols = 0;
    int64_t k = -1;                 // rows to keep; -1 = all
    uint64_t row_base = 0, region_base = 0;
    bool topk = false;              // single key + limit: selection path
    bool used_vec = false;
    Pool pool[2];
    // scratch for selection
    uint64_t *cand_key[2] = {nullptr, nullptr}, *cand_idx[2] = {nullptr, nullptr}; uint32_t cand_cap = 0;
    uint64_t *samp_key = nullptr, *samp_idx = nullptr;
    uint32_t* d_counts = nullptr;   // [0] sample count, [1] candidate count A, [2] candidate count B, [3] class count, [4] pool out n
    uint64_t* d_rank = nullptr;     // {key, idx, n}
    uint8_t* h_stage = nullptr; size_t h_stage_cap = 0;   // pinned landing area of the top-k rows
    uint32_t* h_counts = nullptr;                         // pinned copy of d_counts
    // full sort / filter-only: retained rows
    std::vector<uint8_t*> ret_vals, ret_null; int64_t ret_rows = 0, ret_cap = 0;
    std::vector<uint8_t*> key_img;  // per ORDER BY key: retained images (full sort)
    uint8_t* ret_keep = nullptr;
    std::vector<void*> allocs;
    int64_t total_seen = 0;
    std::vector<SortOutCol> pending_out; int64_t pending_rows = 0;
};

namespace {

int fail(std::string& err, int code, const char* msg) { err = msg; return code; }
int cuda_fail(std::string& err, cudaError_t e, const char* what) { err = std::string(what) + ": " + cudaGetErrorString(e); return e == cudaErrorMemoryAllocation ? BKGPU_ENOMEM : BKGPU_ENODEV; }
#define SCK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(err, e__, #call); } while (0)

template <class T> int dalloc(SortState* s, T** p, size_t bytes, std::string& err) {
    cudaError_t e = cudaMalloc((void**)p, bytes ? bytes : 8);
    if (e != cudaSuccess) return cuda_fail(err, e, "cudaMalloc");
    s->allocs.push_back(*p);
    return 0;
}
void dfree(SortState* s, void* p) {
    if (!p) return;
    auto it = std::find(s->allocs.begin(), s->allocs.end(), p);
    if (it != s->allocs.end()) s->allocs.erase(it);
    cudaFree(p);
}
int elem_bytes_of(int prim) { return storage_bytes(prim_storage(prim)); }

void fill_row_args(const SortState* s, const DevCol* cols, int64_t nrows, RowArgs& a) {
    memset(&a, 0, sizeof a);
    for (int i = 0; i < s->ncols; i++) a.cols[i] = cols[i];
    a.n_cols = s->ncols; a.nrows = nrows; a.row_base = s->row_base; a.prog = s->c.prog;
    a.key.pred_out = s->c.ap.pred_out;
    if (!s->c.sort_keys.empty()) {
        const SortKey& k = s->c.sort_keys[0];
        a.key.out_reg = k.out_reg; a.key.vclass = host_prim_class(k.prim); a.key.desc = k.asc ? 0 : 1;
        a.key.direct = s->c.has_direct ? 1 : 0; a.key.col = s->c.has_direct ? s->c.direct_cols[0] : 0;
    }
}

int grid_for(int64_t n, int per_block, int sm) { int64_t g = (n + per_block - 1) / per_block; int64_t cap = (int64_t)sm * 8; return (int)std::max<int64_t>(1, std::min(g, cap)); }

// ---- selection of the k smallest composite keys of one class out of the batch; result merged into the pool ----
int select_topk_class(SortState* s, const RowArgs& ra, int cls, cudaStream_t st, bkgpu_stats* stats, std::string& err) {
    Pool& P = s->pool[cls - 1];
    const uint32_t k = (uint32_t)s->k;
    uint32_t* cnt = s->d_counts;     // [0] sample count  [1] level-1 candidates  [2] level-2  [3] rows of this class  [4] kept rows  [5] level-3
    uint64_t* thr = s->d_rank;       // three {key, idx, n} triples: thresholds of the three levels
    const uint32_t old_n = std::min<uint32_t>(P.n, k);
    const uint32_t room = SMALL_N - old_n;
    const uint32_t trim = std::min<uint32_t>(room, std::max<uint32_t>(2 * k, 2048) - std::min<uint32_t>(old_n, 1024));
    const int nxt = P.cur ^ 1;
    const uint64_t* pool_thr = nullptr;
    if (P.n >= k) {   // later rows must beat the k-th composite key kept so far
        SCK(cudaMemcpyAsync(thr + 12, P.key[P.cur] + (k - 1), 8, cudaMemcpyDeviceToDevice, st));
        SCK(cudaMemcpyAsync(thr + 13, P.idx[P.cur] + (k - 1), 8, cudaMemcpyDeviceToDevice, st));
        pool_thr = thr + 12;
    }
    uint32_t h[8];
    // The whole chain (sample -> threshold -> compact, three times, then the one-CTA sort and the payload gather)
    // is enqueued without a host round trip; counts are checked once at the end.  A statistically unlucky
    // threshold (too tight / too loose) repeats the chain with wider ranks.
    for (int attempt = 0; attempt < 6; attempt++) {
        const uint32_t widen = 1u << attempt;
        SCK(cudaMemsetAsync(cnt, 0, 32, st));
        const bool big = ra.nrows > (int64_t)s->cand_cap / 2;
        if (big) {
            k_sample_rows<<<(SAMPLE_N + 255) / 256, 256, 0, st>>>(ra, cls, s->samp_key, s->samp_idx, cnt, 0x5bd1e995ull + attempt * 7919);
            double w = ((double)k * (double)SAMPLE_N / (double)ra.nrows * 1.5 + 32.0) * (double)widen;
            uint32_t rank = w >= (double)(SAMPLE_N - 1) ? 0xFFFFFFFFu : (uint32_t)w;
            k_select_rank<<<1, 1024, 0, st>>>(s->samp_key, s->samp_idx, cnt, SAMPLE_N, rank, thr, nullptr, 0, 0, 0, pool_thr);
            stats->kernel_launches += 2;
        } else if (pool_thr) {
            SCK(cudaMemcpyAsync(thr, pool_thr, 16, cudaMemcpyDeviceToDevice, st));
        } else SCK(cudaMemsetAsync(thr, 0xFF, 16, st));
        const DevCol& kc = ra.cols[ra.key.col];
        const bool vec = cls == 1 && ra.key.direct && !kc.validity && (((uintptr_t)kc.values & 31) == 0) && kc.prim != BK_INT8 && kc.prim != BK_INT16 &&
                         kc.prim != BK_UINT8 && kc.prim != BK_UINT16 && (kc.stype == ST_I64 || kc.stype == ST_U64 || kc.stype == ST_F64 || kc.stype == ST_I32 || kc.stype == ST_U32);
        s->used_vec = vec;
        if (vec) {
            const int g = std::max(1, std::min(s->sm_count * 8, (int)((ra.nrows / 4 + 511) / 512)));
            if (kc.stype == ST_I32 || kc.stype == ST_U32) k_collect_rows_vec<4><<<g, 256, 0, st>>>(ra, thr, s->cand_key[0], s->cand_idx[0], cnt + 1, s->cand_cap, cnt + 3);
            else k_collect_rows_vec<8><<<g, 256, 0, st>>>(ra, thr, s->cand_key[0], s->cand_idx[0], cnt + 1, s->cand_cap, cnt + 3);
        } else
        k_collect_rows<<<grid_for(ra.nrows, 256, s->sm_count), 256, 0, st>>>(ra, cls, thr, s->cand_key[0], s->cand_idx[0], cnt + 1, s->cand_cap, cnt + 3);
        // level 2: cand[0] (cnt[1]) -> cand[1] (cnt[2]);  level 3: cand[1] (cnt[2]) -> cand[0] (cnt[5])
        const int src_cnt[2] = {1, 2}, dst_cnt[2] = {2, 5};
        for (int lvl = 0; lvl < 2; lvl++) {
            const int from = lvl, to = lvl ^ 1;
            SCK(cudaMemsetAsync(cnt, 0, 4, st));
            k_sample_pairs<<<(SAMPLE_N + 255) / 256, 256, 0, st>>>(s->cand_key[from], s->cand_idx[from], cnt + src_cnt[lvl], s->cand_cap, s->samp_key, s->samp_idx, cnt,
                                                                   0x9e3779b9ull + attempt * 104729 + lvl);
            // (levels keep trimming until about two k-fulls are left: the one-CTA sort below costs 35 us for <= 2048 keys, 150 us for 8192)
            k_select_rank<<<1, 1024, 0, st>>>(s->samp_key, s->samp_idx, cnt, SAMPLE_N, 0, thr + 3 * (lvl + 1), cnt + src_cnt[lvl], k, widen, trim, nullptr);
            k_collect_pairs<<<grid_for(lvl == 0 ? (int64_t)s->cand_cap / 4 : 65536, 256, s->sm_count), 256, 0, st>>>(
                s->cand_key[from], s->cand_idx[from], cnt + src_cnt[lvl], thr + 3 * (lvl + 1), s->cand_key[to], s->cand_idx[to], cnt + dst_cnt[lvl], s->cand_cap);
            stats->kernel_launches += 3;
        }
        // final: candidates + kept rows sorted by one CTA, keep k, fetch the payload of the survivors
        k_sort_small<<<1, 1024, SMALL_N * 16, st>>>(s->cand_key[0], s->cand_idx[0], cnt + 5, room, P.key[P.cur], P.idx[P.cur], old_n,
                                                    P.key[nxt], P.idx[nxt], k, cnt + 4, 0, nullptr, nullptr, 0, 0, 0, nullptr);
        GatherArgs g; memset(&g, 0, sizeof g);
        for (int i = 0; i < s->ncols; i++) { g.cols[i] = ra.cols[i]; g.dst_vals[i] = P.vals[nxt][i]; g.dst_null[i] = P.nulls[nxt][i]; g.old_vals[i] = P.vals[P.cur][i]; g.old_null[i] = P.nulls[P.cur][i]; }
        g.n_cols = s->ncols; g.row_base = ra.row_base; g.nrows = ra.nrows; g.old_idx = P.idx[P.cur]; g.old_n = old_n;
        k_gather_topk<<<std::max(1, (int)((k + 127) / 128)), 128, 0, st>>>(g, P.idx[nxt], cnt + 4);
        stats->kernel_launches += 3;
        if (!s->h_counts) SCK(cudaHostAlloc((void**)&s->h_counts, 64, cudaHostAllocDefault));
        SCK(cudaMemcpyAsync(s->h_counts, cnt, 32, cudaMemcpyDeviceToHost, st));
        if (big && pool_thr) SCK(cudaMemcpyAsync(s->h_counts + 8, thr + 2, 8, cudaMemcpyDeviceToHost, st));
        SCK(cudaStreamSynchronize(st));
        memcpy(h, s->h_counts, 32);
        const uint32_t c1 = h[1], c2 = h[2], cls_rows = h[3], c3 = h[5];
        // fewer than k candidates are complete only when the threshold in force was the pool's k-th key (every row at or below it
        // was collected); a tighter SAMPLE threshold may have cut rows that belong between it and the pool's key: widen and retry
        uint64_t thr_info = 0;
        if (big && pool_thr) memcpy(&thr_info, s->h_counts + 8, 8);
        const bool pool_in_force = pool_thr != nullptr && (!big || ((thr_info >> 62) & 1ull));
        const bool ok1 = c1 <= s->cand_cap && (c1 >= k || c1 >= cls_rows || pool_in_force);
        const bool ok2 = c2 <= s->cand_cap && (c2 >= std::min(k, c1));
        const bool ok3 = c3 <= room && (c3 >= std::min(k, c2));
        if (ok1 && ok2 && ok3) { P.cur = nxt; P.n = h[4]; return 0; }
    }
    return fail(err, BKGPU_ETOOBIG, "top-k selection did not converge (candidate buffers too small for this key distribution)");
}

int ensure_retained(SortState* s, int64_t need, cudaStream_t st, std::string& err) {
    if (need <= s->ret_cap) return 0;
    int64_t cap = std::max<int64_t>(need, std::max<int64_t>(s->ret_cap * 2, 1 << 16));
    for (int c = 0; c < s->ncols; c++) {
        const int eb = elem_bytes_of(s->c.cols[(size_t)c].prim);
        uint8_t *nv = nullptr, *nn = nullptr; int rc;
        if ((rc = dalloc(s, &nv, (size_t)cap * eb, err))) return rc;
        if ((rc = dalloc(s, &nn, (size_t)cap, err))) return rc;
        if (s->ret_rows) {
            SCK(cudaMemcpyAsync(nv, s->ret_vals[(size_t)c], (size_t)s->ret_rows * eb, cudaMemcpyDeviceToDevice, st));
            SCK(cudaMemcpyAsync(nn, s->ret_null[(size_t)c], (size_t)s->ret_rows, cudaMemcpyDeviceToDevice, st));
        }
        SCK(cudaStreamSynchronize(st));
        dfree(s, s->ret_vals[(size_t)c]); dfree(s, s->ret_null[(size_t)c]);
        s->ret_vals[(size_t)c] = nv; s->ret_null[(size_t)c] = nn;
    }
    s->ret_cap = cap;
    return 0;
}

// rows that pass the predicate are appended, in input order, to the retained columns (filter-only and full sort)
int retain_batch(SortState* s, const RowArgs& ra, cudaStream_t st, bkgpu_stats* stats, std::string& err) {
    const int64_t limit_left = (s->c.kind == PK_FILTER && s->c.limit >= 0) ? std::max<int64_t>(s->c.limit - s->ret_rows, 0) : INT64_MAX;
    if (limit_left == 0) return 0;   // FilterNode stops after `limit` passing rows (filter_node.cpp:786-791)
    const int64_t rows_per_block = 4096;
    const uint32_t nblocks = (uint32_t)((ra.nrows + rows_per_block - 1) / rows_per_block);
    uint32_t* counts = nullptr; uint64_t* d_total = nullptr; uint32_t* sel = nullptr; int rc;
    if ((rc = dalloc(s, &counts, (size_t)nblocks * 4, err))) return rc;
    if ((rc = dalloc(s, &d_total, 8, err))) return rc;
    k_filter_count<<<nblocks, 256, 0, st>>>(ra, counts, rows_per_block);
    k_scan_counts<<<1, 1024, 0, st>>>(counts, nblocks, d_total);
    uint64_t total = 0;
    SCK(cudaMemcpyAsync(&total, d_total, 8, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    const uint64_t take = std::min<uint64_t>(total, (uint64_t)limit_left);
    if (take > 0) {
        if ((rc = dalloc(s, &sel, (size_t)take * 4, err))) return rc;
        k_filter_write<<<nblocks, 256, 0, st>>>(ra, counts, rows_per_block, sel, 0, take);
        if ((rc = ensure_retained(s, s->ret_rows + (int64_t)take, st, err))) return rc;
        GatherArgs g; memset(&g, 0, sizeof g);
        for (int i = 0; i < s->ncols; i++) { g.cols[i] = ra.cols[i]; g.dst_vals[i] = s->ret_vals[(size_t)i]; g.dst_null[i] = s->ret_null[(size_t)i]; }
        g.n_cols = s->ncols;
        k_gather_sel<<<grid_for((int64_t)take, 256, s->sm_count), 256, 0, st>>>(g, sel, take, (uint64_t)s->ret_rows);
        SCK(cudaStreamSynchronize(st));
        s->ret_rows += (int64_t)take;
        stats->kernel_launches += 2;
    }
    stats->kernel_launches += 2;
    stats->rows_filtered += ra.nrows - (int64_t)total;
    dfree(s, counts); dfree(s, d_total); dfree(s, sel);
    return 0;
}

int copy_out(SortState* s, uint8_t* const* vals, uint8_t* const* nulls, const uint32_t* perm_dev, int64_t n, int64_t skip, cudaStream_t st,
             std::vector<SortOutCol>& out, std::string& err) {
    // perm_dev == nullptr: rows [skip, skip+n) in place; else gather through a device permutation first
    out.clear();
    for (int c = 0; c < s->ncols; c++) {
        const ColRef& cr = s->c.cols[(size_t)c];
        SortOutCol oc; oc.tuple_id = cr.tuple_id; oc.slot_id = cr.slot_id; oc.prim = cr.prim; oc.elem = elem_bytes_of(cr.prim);
        oc.values.assign((size_t)std::max<int64_t>(n, 1) * oc.elem, 0);
        std::vector<uint8_t> nb((size_t)std::max<int64_t>(n, 1));
        const uint8_t* sv = vals[c]; const uint8_t* sn = nulls[c];
        uint8_t *tv = nullptr, *tn = nullptr;
        if (perm_dev && n > 0) {
            int rc;
            if ((rc = dalloc(s, &tv, (size_t)n * oc.elem, err))) return rc;
            if ((rc = dalloc(s, &tn, (size_t)n, err))) return rc;
            GatherArgs g; memset(&g, 0, sizeof g);
            g.cols[0].values = sv; g.cols[0].validity = nullptr; g.cols[0].stype = prim_storage(cr.prim); g.cols[0].prim = cr.prim;
            g.n_cols = 1; g.dst_vals[0] = tv; g.dst_null[0] = tn;
            k_gather_sel<<<grid_for(n, 256, s->sm_count), 256, 0, st>>>(g, perm_dev + skip, (uint64_t)n, 0);
            // null bytes travel separately (k_gather_sel reads validity bitmaps, retained columns hold null BYTES)
            k_rs_gather_bytes<<<grid_for(n, 256, s->sm_count), 256, 0, st>>>(sn, perm_dev + skip, tn, (uint64_t)n);
            SCK(cudaMemcpyAsync(oc.values.data(), tv, (size_t)n * oc.elem, cudaMemcpyDeviceToHost, st));
            SCK(cudaMemcpyAsync(nb.data(), tn, (size_t)n, cudaMemcpyDeviceToHost, st));
            SCK(cudaStreamSynchronize(st));
            dfree(s, tv); dfree(s, tn);
        } else if (n > 0) {
            SCK(cudaMemcpyAsync(oc.values.data(), sv + (size_t)skip * oc.elem, (size_t)n * oc.elem, cudaMemcpyDeviceToHost, st));
            SCK(cudaMemcpyAsync(nb.data(), sn + skip, (size_t)n, cudaMemcpyDeviceToHost, st));
            SCK(cudaStreamSynchronize(st));
        }
        bool any = false;
        for (int64_t i = 0; i < n; i++) any |= nb[(size_t)i] != 0;
        if (any) {
            oc.validity.assign((size_t)(n + 7) / 8 + 1, 0xFF);
            for (int64_t i = 0; i < n; i++) if (nb[(size_t)i]) oc.validity[(size_t)i >> 3] &= (uint8_t)~(1u << (i & 7));
        }
        out.push_back(std::move(oc));
    }
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------
int sort_open(const Compiled& c, int device, cudaStream_t stream, int64_t region_base, SortState** out, std::string& err) {
    (void)stream;
    SortState* s = new SortState();
    s->c = c; s->device = device; s->ncols = (int)c.cols.size();
    s->row_base = s->region_base = (uint64_t)region_base;
    { int n = 0; if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && n > 0) s->sm_count = n; }
    s->k = c.kind == PK_SORT ? c.limit : -1;
    s->topk = c.kind == PK_SORT && c.sort_keys.size() == 1 && s->k >= 0 && s->k <= 4096;
    s->ret_vals.assign((size_t)s->ncols, nullptr); s->ret_null.assign((size_t)s->ncols, nullptr);
    int rc = 0;
    if (s->topk) {
        s->cand_cap = 1u << 22;   // 4M candidate pairs (64 MB per array pair)
        const size_t kk = (size_t)std::max<int64_t>(s->k, 1);
        for (int b = 0; b < 2 && !rc; b++) {
            if ((rc = dalloc(s, &s->cand_key[b], (size_t)s->cand_cap * 8, err))) break;
            if ((rc = dalloc(s, &s->cand_idx[b], (size_t)s->cand_cap * 8, err))) break;
        }
        if (!rc) rc = dalloc(s, &s->samp_key, SAMPLE_N * 8, err);
        if (!rc) rc = dalloc(s, &s->samp_idx, SAMPLE_N * 8, err);
        if (!rc) rc = dalloc(s, &s->d_counts, 64, err);
        if (!rc) rc = dalloc(s, &s->d_rank, 128, err);
        for (int p = 0; p < 2 && !rc; p++) for (int b = 0; b < 2 && !rc; b++) {
            if ((rc = dalloc(s, &s->pool[p].key[b], kk * 8, err))) break;
            if ((rc = dalloc(s, &s->pool[p].idx[b], kk * 8, err))) break;
            for (int ci = 0; ci < s->ncols && !rc; ci++) {
                if ((rc = dalloc(s, &s->pool[p].vals[b][ci], kk * (size_t)elem_bytes_of(c.cols[(size_t)ci].prim), err))) break;
                rc = dalloc(s, &s->pool[p].nulls[b][ci], kk, err);
            }
        }
        if (!rc) { cudaError_t e = cudaFuncSetAttribute(k_sort_small, cudaFuncAttributeMaxDynamicSharedMemorySize, SMALL_N * 16); if (e != cudaSuccess) rc = cuda_fail(err, e, "cudaFuncSetAttribute"); }
    }
    if (rc) { sort_close(s); return rc; }
    *out = s;
    return 0;
}

int sort_reset(SortState* s, cudaStream_t, std::string&) {
    for (auto& p : s->pool) { p.n = 0; p.cur = 0; }
    s->ret_rows = 0; s->total_seen = 0;
    s->row_base = s->region_base;
    return 0;
}

int sort_push(SortState* s, const DevCol* cols, int64_t nrows, cudaStream_t stream, bkgpu_stats* stats, std::string& err) {
    if (nrows > 0x7FFFFFFFll) return fail(err, BKGPU_EUNSUPPORTED, "a single pushed batch of a SORT/FILTER plan is limited to 2^31-1 rows");
    RowArgs ra; fill_row_args(s, cols, nrows, ra);
    int rc = 0;
    if (s->topk) {
        if (s->k == 0) { s->row_base += (uint64_t)nrows; return 0; }
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, stream);
        rc = select_topk_class(s, ra, 1, stream, stats, err);
        const bool need_nulls = ra.key.direct ? cols[ra.key.col].validity != nullptr : true;
        if (!rc && need_nulls) rc = select_topk_class(s, ra, 2, stream, stats, err);
        cudaEventRecord(e1, stream); cudaEventSynchronize(e1);
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1);
        stats->main_kernel_ms += ms; stats->main_kernel_launches += 1;
        stats->main_kernel_bytes += nrows * (ra.key.direct ? storage_bytes(cols[ra.key.col].stype) : 8);
        snprintf(stats->main_kernel_name, sizeof stats->main_kernel_name, "%s", s->used_vec ? "topk_select(k_collect_rows_vec)" : "topk_select(k_collect_rows)");
    } else {
        rc = retain_batch(s, ra, stream, stats, err);
        snprintf(stats->main_kernel_name, sizeof stats->main_kernel_name, "%s", s->c.kind == PK_FILTER ? "k_filter_write" : "radix_sort(k_rs_pass)");
    }
    s->row_base += (uint64_t)nrows; s->total_seen += nrows;
    return rc;
}

// layout of the exported partial (top-k): [u64 n][k keys][k idx][per column: k values (8-byte slots)][per column: k null bytes padded to 8]
size_t sort_partial_bytes(SortState* s) {
    if (!s->topk) return 0;
    const size_t k = (size_t)std::max<int64_t>(s->k, 1);
    return 8 * (1 + 2 * 2 * k) + (size_t)s->ncols * 2 * (k * 8 + ((k + 7) & ~(size_t)7));
}

// host image of the k best rows of this rank: class-tagged, ready to be merged with other ranks'
struct HostRows {
    std::vector<uint8_t> cls; std::vector<uint64_t> key, idx;
    std::vector<std::vector<uint8_t>> vals, nulls;   // per column, row-major fixed width
};

// the k surviving rows of both pools (keys present / NULL keys) land in ONE pinned staging block: every copy is truly
// asynchronous and the whole result costs a single synchronisation
static int topk_host_rows(SortState* s, HostRows& h, cudaStream_t st, std::string& err) {
    h.vals.assign((size_t)s->ncols, {}); h.nulls.assign((size_t)s->ncols, {});
    size_t row_bytes = 16;
    for (int c = 0; c < s->ncols; c++) row_bytes += (size_t)elem_bytes_of(s->c.cols[(size_t)c].prim) + 1;
    const size_t n_tot = (size_t)s->pool[0].n + (size_t)s->pool[1].n;
    const size_t need = n_tot * row_bytes + 64 * (size_t)(s->ncols + 2) * 2;
    if (s->h_stage_cap < need) {
        if (s->h_stage) cudaFreeHost(s->h_stage);
        s->h_stage = nullptr; s->h_stage_cap = 0;
        SCK(cudaHostAlloc((void**)&s->h_stage, need * 2, cudaHostAllocDefault));
        s->h_stage_cap = need * 2;
    }
    struct Seg { size_t key, idx; std::vector<size_t> val, nul; } seg[2];
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 63) & ~(size_t)63; return o; };
    for (int p = 0; p < 2; p++) {
        Pool& P = s->pool[p];
        const size_t n = P.n;
        if (!n) continue;
        seg[p].key = take(n * 8); seg[p].idx = take(n * 8);
        SCK(cudaMemcpyAsync(s->h_stage + seg[p].key, P.key[P.cur], n * 8, cudaMemcpyDeviceToHost, st));
        SCK(cudaMemcpyAsync(s->h_stage + seg[p].idx, P.idx[P.cur], n * 8, cudaMemcpyDeviceToHost, st));
        for (int c = 0; c < s->ncols; c++) {
            const size_t eb = (size_t)elem_bytes_of(s->c.cols[(size_t)c].prim);
            seg[p].val.push_back(take(n * eb)); seg[p].nul.push_back(take(n));
            SCK(cudaMemcpyAsync(s->h_stage + seg[p].val.back(), P.vals[P.cur][c], n * eb, cudaMemcpyDeviceToHost, st));
            SCK(cudaMemcpyAsync(s->h_stage + seg[p].nul.back(), P.nulls[P.cur][c], n, cudaMemcpyDeviceToHost, st));
        }
    }
    SCK(cudaStreamSynchronize(st));
    for (int p = 0; p < 2; p++) {
        const size_t n = s->pool[p].n;
        if (!n) continue;
        const uint64_t* key = (const uint64_t*)(s->h_stage + seg[p].key); const uint64_t* idx = (const uint64_t*)(s->h_stage + seg[p].idx);
        for (size_t i = 0; i < n; i++) { h.cls.push_back((uint8_t)(p + 1)); h.key.push_back(key[i]); h.idx.push_back(idx[i]); }
        for (int c = 0; c < s->ncols; c++) {
            const size_t eb = (size_t)elem_bytes_of(s->c.cols[(size_t)c].prim);
            const uint8_t* v = s->h_stage + seg[p].val[(size_t)c]; const uint8_t* nb = s->h_stage + seg[p].nul[(size_t)c];
            h.vals[(size_t)c].insert(h.vals[(size_t)c].end(), v, v + n * eb);
            h.nulls[(size_t)c].insert(h.nulls[(size_t)c].end(), nb, nb + n);
        }
    }
    return 0;
}
static void finish_host_rows(SortState* s, const HostRows& h, std::vector<SortOutCol>& out, int64_t* nrows) {
    const bool null_first = s->c.sort_keys[0].null_first;
    std::vector<uint32_t> order(h.key.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = (uint32_t)i;
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const int ra = h.cls[a] == 2 ? (null_first ? 0 : 2) : 1, rb = h.cls[b] == 2 ? (null_first ? 0 : 2) : 1;
        if (ra != rb) return ra < rb;
        if (h.key[a] != h.key[b]) return h.key[a] < h.key[b];
        return h.idx[a] < h.idx[b];
    });
    int64_t n = (int64_t)order.size();
    if (s->k >= 0 && n > s->k) n = s->k;
    int64_t skip = std::min<int64_t>(s->c.offset, n);
    n -= skip;
    out.clear();
    for (int c = 0; c < s->ncols; c++) {
        const ColRef& cr = s->c.cols[(size_t)c];
        SortOutCol oc; oc.tuple_id = cr.tuple_id; oc.slot_id = cr.slot_id; oc.prim = cr.prim; oc.elem = elem_bytes_of(cr.prim);
        oc.values.assign((size_t)std::max<int64_t>(n, 1) * oc.elem, 0);
        bool any = false;
        std::vector<uint8_t> valid((size_t)(n + 7) / 8 + 1, 0xFF);
        for (int64_t i = 0; i < n; i++) {
            const uint32_t src = order[(size_t)(i + skip)];
            memcpy(oc.values.data() + (size_t)i * oc.elem, h.vals[(size_t)c].data() + (size_t)src * oc.elem, (size_t)oc.elem);
            if (h.nulls[(size_t)c][src]) { any = true; valid[(size_t)i >> 3] &= (uint8_t)~(1u << (i & 7)); }
        }
        if (any) oc.validity = std::move(valid);
        out.push_back(std::move(oc));
    }
    *nrows = n;
}

static void pack_rows(SortState* s, const HostRows& h, std::vector<uint64_t>& buf) {
    const size_t k = (size_t)std::max<int64_t>(s->k, 1), kk = 2 * k;   // both classes
    const size_t nullw = (k + 7) / 8;
    buf.assign(sort_partial_bytes(s) / 8, 0);
    const size_t n = std::min(h.key.size(), kk);
    buf[0] = n;
    uint64_t* keys = buf.data() + 1; uint64_t* idx = keys + kk;
    uint64_t* col0 = idx + kk;
    for (size_t i = 0; i < n; i++) { keys[i] = h.key[i]; idx[i] = h.idx[i] | ((uint64_t)(h.cls[i] == 2) << 63); }
    for (int c = 0; c < s->ncols; c++) {
        const int eb = elem_bytes_of(s->c.cols[(size_t)c].prim);
        uint64_t* v = col0 + (size_t)c * 2 * (k + nullw);
        uint8_t* nb = (uint8_t*)(v + kk);
        for (size_t i = 0; i < n; i++) { memcpy(&v[i], h.vals[(size_t)c].data() + i * eb, (size_t)eb); nb[i] = h.nulls[(size_t)c][i]; }
    }
}
static void unpack_rows(SortState* s, const uint64_t* buf, HostRows& h) {
    const size_t k = (size_t)std::max<int64_t>(s->k, 1), kk = 2 * k, nullw = (k + 7) / 8;
    const size_t n = (size_t)buf[0];
    const uint64_t* keys = buf + 1; const uint64_t* idx = keys + kk; const uint64_t* col0 = idx + kk;
    if (h.vals.empty()) { h.vals.assign((size_t)s->ncols, {}); h.nulls.assign((size_t)s->ncols, {}); }
    for (size_t i = 0; i < n; i++) { h.key.push_back(keys[i]); h.idx.push_back(idx[i] & ~(1ull << 63)); h.cls.push_back((idx[i] >> 63) ? 2 : 1); }
    for (int c = 0; c < s->ncols; c++) {
        const int eb = elem_bytes_of(s->c.cols[(size_t)c].prim);
        const uint64_t* v = col0 + (size_t)c * 2 * (k + nullw);
        const uint8_t* nb = (const uint8_t*)(v + kk);
        for (size_t i = 0; i < n; i++) {
            const uint8_t* p = (const uint8_t*)&v[i];
            h.vals[(size_t)c].insert(h.vals[(size_t)c].end(), p, p + eb);
            h.nulls[(size_t)c].push_back(nb[i]);
        }
    }
}

int sort_partial_export(SortState* s, void* dev_dst, cudaStream_t st, std::string& err) {
    if (!s->topk) return fail(err, BKGPU_EUNSUPPORTED, "only ORDER BY ... LIMIT plans have a partial state");
    HostRows h; int rc = topk_host_rows(s, h, st, err);
    if (rc) return rc;
    std::vector<uint64_t> buf; pack_rows(s, h, buf);
    SCK(cudaMemcpyAsync(dev_dst, buf.data(), buf.size() * 8, cudaMemcpyHostToDevice, st));
    SCK(cudaStreamSynchronize(st));
    return 0;
}

int sort_partial_merge(SortState* s, const void* dev_src, int nranks, cudaStream_t st, std::vector<SortOutCol>& out, int64_t* nrows, std::string& err) {
    if (!s->topk) return fail(err, BKGPU_EUNSUPPORTED, "only ORDER BY ... LIMIT plans have a partial state");
    const size_t words = sort_partial_bytes(s) / 8;
    std::vector<uint64_t> all(words * (size_t)nranks);
    SCK(cudaMemcpyAsync(all.data(), dev_src, all.size() * 8, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    HostRows h;
    for (int r = 0; r < nranks; r++) unpack_rows(s, all.data() + (size_t)r * words, h);
    finish_host_rows(s, h, out, nrows);   // SelectManagerNode's merge of per-region sorted runs (select_manager_node.cpp:50-51)
    return 0;
}

int sort_finish(SortState* s, void* nccl_comm, int nranks, cudaStream_t st, bkgpu_stats* stats, std::vector<SortOutCol>& out, int64_t* nrows, std::string& err) {
    if (s->topk) {
        HostRows h; int rc = topk_host_rows(s, h, st, err);
        if (rc) return rc;
        if (nccl_comm && nranks > 1) {   // per-GPU top-k rows meet in one all-gather; every rank finishes the merge
            std::vector<uint64_t> buf; pack_rows(s, h, buf);
            uint64_t *d_send = nullptr, *d_recv = nullptr;
            if ((rc = dalloc(s, &d_send, buf.size() * 8, err))) return rc;
            if ((rc = dalloc(s, &d_recv, buf.size() * 8 * (size_t)nranks, err))) return rc;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            SCK(cudaMemcpyAsync(d_send, buf.data(), buf.size() * 8, cudaMemcpyHostToDevice, st));
            cudaEventRecord(e0, st);
            if (nccl_all_gather(nccl_comm, d_send, d_recv, buf.size(), st) != 0) { err = std::string("ncclAllGather: ") + nccl_last_error(); return BKGPU_ENCCL; }
            cudaEventRecord(e1, st);
            rc = sort_partial_merge(s, d_recv, nranks, st, out, nrows, err);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1); stats->collective_ms += ms; cudaEventDestroy(e0); cudaEventDestroy(e1);
            dfree(s, d_send); dfree(s, d_recv);
            return rc;
        }
        finish_host_rows(s, h, out, nrows);
        return 0;
    }
    if (nccl_comm && nranks > 1) return fail(err, BKGPU_EUNSUPPORTED, "multi-GPU merge is implemented for ORDER BY ... LIMIT (top-k) and aggregates");
    if (s->c.kind == PK_FILTER) {
        int64_t n = s->ret_rows;
        int64_t skip = std::min<int64_t>(s->c.offset, n);
        n -= skip;
        *nrows = n;
        return copy_out(s, s->ret_vals.data(), s->ret_null.data(), nullptr, n, skip, st, out, err);
    }
    // ---- full ORDER BY: stable LSD radix sort of (key image, row id) pairs, least significant ORDER BY key first ----
    const int64_t n = s->ret_rows;
    if (n >= (int64_t)RS_VAL_MASK) return fail(err, BKGPU_EUNSUPPORTED, "full sort is limited to 2^30 rows per GPU");
    if (n == 0) { *nrows = 0; return copy_out(s, s->ret_vals.data(), s->ret_null.data(), nullptr, 0, 0, st, out, err); }
    const uint32_t un = (uint32_t)n;
    const uint32_t ntiles = (un + RS_TILE - 1) / RS_TILE;
    uint64_t *img_row = nullptr, *img[2] = {nullptr, nullptr}; uint32_t *ids[2] = {nullptr, nullptr}, *hist = nullptr, *status = nullptr; int rc;
    if ((rc = dalloc(s, &img_row, (size_t)n * 8, err))) return rc;
    if ((rc = dalloc(s, &img[0], (size_t)n * 8, err))) return rc;
    if ((rc = dalloc(s, &img[1], (size_t)n * 8, err))) return rc;
    if ((rc = dalloc(s, &ids[0], (size_t)n * 4, err))) return rc;
    if ((rc = dalloc(s, &ids[1], (size_t)n * 4, err))) return rc;
    if ((rc = dalloc(s, &hist, (size_t)(8 * 256 * 2 + 8 + 8) * 4, err))) return rc;       // histograms, bases, active flags, tickets
    if ((rc = dalloc(s, &status, (size_t)ntiles * 256 * 8 * 4, err))) return rc;           // one look-back array per pass
    uint32_t* bases = hist + 8 * 256; uint32_t* active = bases + 8 * 256; uint32_t* tickets = active + 8;
    const size_t rs_smem = (size_t)RS_TILE * 12 + (8 * 256 + 256 + 256) * 4;
    SCK(cudaFuncSetAttribute(k_rs_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem));
    RowArgs ra; memset(&ra, 0, sizeof ra);
    for (int c = 0; c < s->ncols; c++) { ra.cols[c].values = s->ret_vals[(size_t)c]; ra.cols[c].validity = nullptr; ra.cols[c].stype = prim_storage(s->c.cols[(size_t)c].prim); ra.cols[c].prim = s->c.cols[(size_t)c].prim; }
    // retained columns carry null BYTES: pack them into bitmaps on the device so the key program sees NULLs
    std::vector<uint8_t*> bitmaps((size_t)s->ncols, nullptr);
    uint32_t* null_flags = nullptr;
    if ((rc = dalloc(s, &null_flags, sizeof(uint32_t) * MAX_COLS, err))) return rc;
    SCK(cudaMemsetAsync(null_flags, 0, sizeof(uint32_t) * MAX_COLS, st));
    for (int c = 0; c < s->ncols; c++) {
        if ((rc = dalloc(s, &bitmaps[(size_t)c], (size_t)(n + 7) / 8 + 64, err))) return rc;
        k_pack_null_bytes<<<grid_for((n + 7) / 8, 256, s->sm_count), 256, 0, st>>>(s->ret_null[(size_t)c], n, bitmaps[(size_t)c], null_flags + c);
        ra.cols[c].validity = bitmaps[(size_t)c];
    }
    uint32_t h_null_flags[MAX_COLS] = {0};
    SCK(cudaMemcpyAsync(h_null_flags, null_flags, sizeof(uint32_t) * (size_t)s->ncols, cudaMemcpyDeviceToHost, st));
    SCK(cudaStreamSynchronize(st));
    ra.n_cols = s->ncols; ra.nrows = n; ra.prog = s->c.prog; ra.key.pred_out = -1;
    k_rs_iota<<<grid_for(n, 256, s->sm_count), 256, 0, st>>>(ids[0], un);
    int cur = 0; bool first = true;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    int64_t pass_bytes = 0;
    // sorts the pairs (img[cur], ids[cur]) by the 64-bit image: one histogram pass, then one Onesweep pass per digit that varies
    auto sort_pairs = [&]() -> int {
        SCK(cudaMemsetAsync(hist, 0, (size_t)(8 * 256 * 2 + 8 + 8) * 4, st));
        k_rs_hist<<<std::min(s->sm_count * 4, (int)((un + 1023) / 1024)), 256, 0, st>>>(img[cur], un, hist);
        k_rs_bases<<<8, 256, 0, st>>>(hist, bases, active);
        uint32_t h_active[8];
        SCK(cudaMemcpyAsync(h_active, active, 32, cudaMemcpyDeviceToHost, st));
        SCK(cudaStreamSynchronize(st));
        int n_active = 0;
        for (int d = 0; d < 8; d++) n_active += h_active[d] != 0;
        if (n_active) SCK(cudaMemsetAsync(status, 0, (size_t)ntiles * 256 * 4 * (size_t)n_active, st));
        stats->kernel_launches += 2; pass_bytes += (int64_t)n * 8;
        int slot = 0;
        for (int d = 0; d < 8; d++) {
            if (!h_active[d]) continue;
            k_rs_pass<<<ntiles, RS_THREADS, rs_smem, st>>>(img[cur], ids[cur], img[cur ^ 1], ids[cur ^ 1], un, 8 * d, bases + 256 * d,
                                                        status + (size_t)slot * ntiles * 256, tickets + d);
            cur ^= 1; slot++;
            stats->kernel_launches++; pass_bytes += (int64_t)n * 24;
        }
        return 0;
    };
    for (int ki = (int)s->c.sort_keys.size() - 1; ki >= 0; ki--) {
        const SortKey& sk = s->c.sort_keys[(size_t)ki];
        // a plain column key (config C5's shape): its image comes straight from the column, and without NULLs in it the NULL-rank pass is void
        const bool direct_key = s->c.has_direct && s->c.sort_keys.size() == 1 && !s->c.direct_cols.empty();
        const int dcol = direct_key ? s->c.direct_cols[0] : -1;
        for (int mode = 0; mode < 2; mode++) {   // the key's value image first, then its NULL rank (stable: NULLs end up strictly first / last)
            if (mode == 1 && direct_key && !h_null_flags[dcol]) continue;
            uint64_t* dst = first ? img[cur] : img_row;
            if (mode == 0 && direct_key) k_sort_images_direct<<<grid_for(n, 256, s->sm_count), 256, 0, st>>>(ra.cols[dcol], n, host_prim_class(sk.prim), sk.asc ? 0 : 1, dst);
            else
            k_sort_images<<<grid_for(n, 256, s->sm_count), 256, 0, st>>>(ra, sk.out_reg, host_prim_class(sk.prim), sk.asc ? 0 : 1, sk.null_first ? 1 : 0, mode, dst, nullptr);
            if (!first) k_rs_gather_u64<<<grid_for(n, 256, s->sm_count), 256, 0, st>>>(img_row, ids[cur], img[cur], un);   // images in the current order
            first = false;
            stats->kernel_launches += 2;
            if ((rc = sort_pairs())) return rc;
        }
    }
    cudaEventRecord(e1, st);
    SCK(cudaStreamSynchronize(st));
    { float ms = 0; cudaEventElapsedTime(&ms, e0, e1); stats->main_kernel_ms += ms; stats->main_kernel_launches += 1; stats->main_kernel_bytes += pass_bytes; }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    snprintf(stats->main_kernel_name, sizeof stats->main_kernel_name, "%s", "radix_sort(k_rs_pass)");
    int64_t keep = n; if (s->k >= 0 && keep > s->k) keep = s->k;
    int64_t skip = std::min<int64_t>(s->c.offset, keep); keep -= skip;
    *nrows = keep;
    rc = copy_out(s, s->ret_vals.data(), s->ret_null.data(), ids[cur], keep, skip, st, out, err);
    dfree(s, img_row); dfree(s, img[0]); dfree(s, img[1]); dfree(s, ids[0]); dfree(s, ids[1]); dfree(s, hist); dfree(s, status);
    for (auto b : bitmaps) dfree(s, b);
    dfree(s, null_flags);
    return rc;
}

void sort_close(SortState* s) {
    if (!s) return;
    for (void* p : s->allocs) cudaFree(p);
    if (s->h_stage) cudaFreeHost(s->h_stage);
    if (s->h_counts) cudaFreeHost(s->h_counts);
    delete s;
}

}  // namespace bk
