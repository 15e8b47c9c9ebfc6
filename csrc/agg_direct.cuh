// agg_direct.cuh — the specialised filter+aggregate kernels (K1+K2) for the canonical shape:
//   predicate = conjunction of NP terms `column <cmp> constant`, GROUP BY one column (or none),
//   aggregate arguments = NA plain columns.  args.cols[] is ordered [NP predicate][key][NA value]
//   columns by the host, so every slot index below is a compile-time constant.
//
// Structure of one warp iteration (128 rows = 4 consecutive rows per lane):
//   1. LOAD     one 128-bit (4-byte types) or 256-bit (8-byte types) non-allocating load per column
//   2. FILTER   predicate in registers -> 4-bit pass mask per lane          (FilterNode::need_copy)
//   3. COMPACT  warp prefix sum of popc(pass); surviving rows are written to a per-warp queue in
//               shared memory (key + NA values + null bits)                 [warp-ballot stream compaction]
//   4. CONSUME  all 32 lanes drain the queue: hash-probe the per-CTA table, then shared-memory
//               atomics on the group's lanes                                (AggFnCall::update)
// Step 3 removes the selectivity-dependent lane divergence from step 4 and keeps the kernel small
// enough for the instruction cache (the first version unrolled the aggregate step 8x per thread,
// grew to 15k SASS instructions and stalled 97% of its issue slots on instruction fetch: see
// profiles/r01_agg_v1_summary.md).
#pragma once
#include "agg_kernels.cuh"

namespace bk {

constexpr int DIRECT_THREADS = 512;
constexpr int ROWS_PER_LANE = 4;
constexpr int QCAP = 32 * ROWS_PER_LANE;  // queue entries per warp

struct alignas(16) U32x4 { uint32_t v[4]; };
__device__ __forceinline__ U32x4 ldg128_u32(const void* p) {
    U32x4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]) : "l"(p));
    return r;
}

// ragged tail of a batch (fewer than four rows left): element loads, out of line
static __device__ __noinline__ void load_quad_tail(const DevCol& c, int64_t row0, int64_t nrows, uint64_t* v, uint32_t* nm_out) {
    uint32_t nm = 0;
    for (int j = 0; j < 4; j++) {
        v[j] = 0;
        if (row0 + j < nrows) { v[j] = load_elem(c, row0 + j); if (elem_is_null(c, row0 + j)) nm |= 1u << j; }
    }
    *nm_out = nm;
}

// four consecutive rows [4q, 4q+4) of one column -> canonical images + 4-bit null mask
__device__ __forceinline__ void load_quad(const DevCol& c, int64_t q, int64_t nrows, uint64_t (&v)[4], uint32_t& nm) {
    const int64_t row0 = q * 4;
    if (row0 + 4 <= nrows) {
        switch (c.stype) {
            case ST_I32: { U32x4 r = ldg128_u32((const uint8_t*)c.values + q * 16);
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = (uint64_t)(int64_t)(int32_t)r.v[j]; } break;
            case ST_U32: { U32x4 r = ldg128_u32((const uint8_t*)c.values + q * 16);
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = (uint64_t)r.v[j]; } break;
            case ST_F32: { U32x4 r = ldg128_u32((const uint8_t*)c.values + q * 16);
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = f64_bits((double)__uint_as_float(r.v[j])); } break;
            case ST_U8: { uint32_t r = __ldg((const uint32_t*)c.values + q);
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = (r >> (8 * j)) & 0xFFu; } break;
            default: { U64x4 r = ldg256_u64((const uint8_t*)c.values + q * 32);
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = r.v[j]; } break;
        }
        nm = 0;
        if (c.validity) nm = (~((uint32_t)__ldg(c.validity + (row0 >> 3)) >> (row0 & 4))) & 0xFu;
    } else {
        load_quad_tail(c, row0, nrows, v, &nm);
        return;
    }
    if (c.prim == BK_INT8 || c.prim == BK_INT16 || c.prim == BK_UINT8 || c.prim == BK_UINT16 || c.prim == BK_BOOL) {
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = narrow_prim(v[j], c.prim);
    }
}

// `column <cmp> constant` over four rows -> 4-bit mask.  The class switch is hoisted out of the row
// loop; the operator is applied branch-free from (lt, eq, gt).  IEEE semantics for DOUBLE.
__device__ __forceinline__ uint32_t term_mask(const DirectTerm& t, const uint64_t (&v)[4]) {
    uint32_t lt = 0, eq = 0, gt = 0;
    if (t.vclass == VC_F64) {
        const double y = bits_f64(t.cbits);
#pragma unroll
        for (int j = 0; j < 4; j++) { const double x = bits_f64(v[j]); lt |= (x < y ? 1u : 0u) << j; eq |= (x == y ? 1u : 0u) << j; gt |= (x > y ? 1u : 0u) << j; }
    } else if (t.vclass == VC_U64) {
        const uint64_t y = t.cbits;
#pragma unroll
        for (int j = 0; j < 4; j++) { lt |= (v[j] < y ? 1u : 0u) << j; eq |= (v[j] == y ? 1u : 0u) << j; gt |= (v[j] > y ? 1u : 0u) << j; }
    } else {
        const int64_t y = (int64_t)t.cbits;
#pragma unroll
        for (int j = 0; j < 4; j++) { const int64_t x = (int64_t)v[j]; lt |= (x < y ? 1u : 0u) << j; eq |= (x == y ? 1u : 0u) << j; gt |= (x > y ? 1u : 0u) << j; }
    }
    switch (t.cmp) {
        case BK_FT_EQ: return eq;
        case BK_FT_NE: return ~eq & 0xFu;
        case BK_FT_LT: return lt;
        case BK_FT_LE: return lt | eq;
        case BK_FT_GT: return gt;
        default: return gt | eq;
    }
}

// read-modify-write of one 8-byte lane in shared memory with a CAS loop (min / max of any class)
static __device__ __noinline__ void smem_rmw(int op, uint64_t* p, uint64_t v) {
    unsigned long long* q = (unsigned long long*)p;
    unsigned long long cur = *(volatile unsigned long long*)q;
    for (;;) {
        const unsigned long long nw = lane_combine(op, cur, v);
        if (nw == cur) break;
        const unsigned long long prev = atomicCAS(q, cur, nw);
        if (prev == cur) break;
        cur = prev;
    }
}
// double add: LDS + DADD + ATOMS.CAS loop (sm_100a has no native 64-bit floating add in shared memory)
__device__ __forceinline__ void smem_add_f64(uint64_t* p, double v) {
    unsigned long long* q = (unsigned long long*)p;
    unsigned long long cur = *(volatile unsigned long long*)q;
    for (;;) {
        const unsigned long long nw = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)cur) + v);
        const unsigned long long prev = atomicCAS(q, cur, nw);
        if (prev == cur) break;
        cur = prev;
    }
}
__device__ __forceinline__ void smem_lane_update(int op, uint64_t* p, uint64_t v) {
    if (op == LN_ADD_F64) smem_add_f64(p, bits_f64(v));
    else if (op == LN_ADD_I64) smem_add_u64(p, v);
    else smem_rmw(op, p, v);
}
// register accumulation of four rows for the less common lane operations (min / max), out of line
static __device__ __noinline__ uint64_t combine4_generic(int op, uint64_t acc, const uint64_t* v, uint32_t ok, int arg_class, int lane_class) {
    for (int j = 0; j < 4; j++)
        if ((ok >> j) & 1u) acc = lane_combine(op, acc, to_lane_class(v[j], arg_class, lane_class));
    return acc;
}

// a row whose group does not fit the shared table (or when no shared table is in use): update the
// global table directly.  Kept out of line: it is the rare path and must not bloat the hot loop.
template <int NA>
__device__ __noinline__ void global_update_row(const AggArgs& a, const uint64_t* key, const uint64_t* vals, uint32_t nullbits) {
    const AggPlan& ap = a.plan;
    const GroupTable& gt = a.gt;
    const uint32_t gcap = gt.cap_mask + 1;
    const int slot = table_upsert<false, 0>(gt.state, gt.keys, gt.cap_mask, key, ap.n_keyw,
                                            ap.n_keyw == 1 ? hash_key1(key[0]) : hash_key(key, ap.n_keyw), (int)gcap, gt.n_groups);
    if (slot < 0) { atomicExch(gt.overflow, 1u); return; }
    atomicAdd((unsigned long long*)(gt.lanes + slot), 1ull);
#pragma unroll
    for (int s = 0; s < NA; s++) {
        if ((nullbits >> s) & 1u) continue;
        const ValOps vo = a.vops[s];
        if (vo.cnt_glob) atomicAdd((unsigned long long*)(gt.lanes + (size_t)vo.cnt_glob * gcap + slot), 1ull);
        for (int k = 0; k < vo.n_ops; k++)
            lane_atomic<false>(vo.op[k], gt.lanes + (size_t)vo.glob_lane[k] * gcap + slot, to_lane_class(vals[s], vo.arg_class, vo.lane_class[k]));
    }
}

// ------------------------------------------------------------------------------------------
// GROUP BY one column
// ------------------------------------------------------------------------------------------
template <int NP, int NA>
__global__ void __launch_bounds__(DIRECT_THREADS, 1) k_agg_group_direct(const __grid_constant__ AggArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const AggPlan& ap = a.plan;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool use_smem = a.smem_cap_log2 > 0;
    SmemTable st{};
    size_t table_bytes = 0;
    if (use_smem) {
        st = smem_table_init(smem_raw, a);
        table_bytes = (((size_t)(ap.n_keyw + a.n_smem_lanes) * 8 + 4) << a.smem_cap_log2);
    }
    // per-warp queue: [key words kw x QCAP][values NA x QCAP][null bits QCAP]
    const int kw = ap.n_keyw;  // 1 or 2
    const size_t qwords = (size_t)(kw + NA) * QCAP;
    uint64_t* qbase = (uint64_t*)(smem_raw + ((table_bytes + 15) & ~(size_t)15)) + (size_t)warp * (qwords + QCAP / 8);
    uint64_t* qkey = qbase;
    uint64_t* qval = qbase + (size_t)kw * QCAP;
    uint8_t* qnull = (uint8_t*)(qbase + qwords);
    const uint64_t kmask = ap.key_bits[0] >= 64 ? ~0ull : ((1ull << ap.key_bits[0]) - 1ull);
    const int null_word = ap.key_null_word[0] == 0xFF ? 0 : ap.key_null_word[0];
    const uint64_t null_bit = 1ull << ap.key_null_shift[0];
    uint32_t passed = 0;
    const int64_t nquads = (a.nrows + 3) >> 2;
    const int64_t stride = (int64_t)gridDim.x * DIRECT_THREADS;
    // all lanes of a warp run the same number of iterations (the queue is warp-collective)
    for (int64_t q0 = (int64_t)blockIdx.x * DIRECT_THREADS + warp * 32; q0 < nquads; q0 += stride) {
        const int64_t q = q0 + lane;
        uint64_t kv[4]; uint32_t knm = 0;
        uint64_t vv[NA > 0 ? NA : 1][4]; uint32_t vnm[NA > 0 ? NA : 1];
        uint32_t pass = 0;
        if (q < nquads) {
            const int64_t left = a.nrows - q * 4;
            pass = left >= 4 ? 0xFu : ((1u << left) - 1u);
#pragma unroll
            for (int t = 0; t < NP; t++) {
                uint64_t pv[4]; uint32_t pnm;
                load_quad(a.cols[t], q, a.nrows, pv, pnm);
                pass &= term_mask(a.direct.term[t], pv) & ~pnm;  // NULL or false drops the row
            }
            load_quad(a.cols[NP], q, a.nrows, kv, knm);
#pragma unroll
            for (int s = 0; s < NA; s++) load_quad(a.cols[NP + 1 + s], q, a.nrows, vv[s], vnm[s]);
        }
        // ---- compact the surviving rows of this warp into its queue ----
        const int cnt = __popc(pass);
        int base = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int n = __shfl_up_sync(0xFFFFFFFFu, base, d); if (lane >= d) base += n; }
        const int total = __shfl_sync(0xFFFFFFFFu, base, 31);
        base -= cnt;
        passed += cnt;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (!((pass >> j) & 1u)) continue;
            const bool knull = (knm >> j) & 1u;
            qkey[base] = knull ? (null_word == 0 ? null_bit : 0ull) : (kv[j] & kmask);
            if (kw == 2) qkey[QCAP + base] = knull && null_word == 1 ? null_bit : 0ull;
            uint32_t nb = 0;
#pragma unroll
            for (int s = 0; s < NA; s++) { qval[s * QCAP + base] = vv[s][j]; nb |= ((vnm[s] >> j) & 1u) << s; }
            qnull[base] = (uint8_t)nb;
            base++;
        }
        __syncwarp();
        // ---- drain: every lane takes queue entries, full warps regardless of selectivity ----
        for (int e = lane; e < total; e += 32) {
            uint64_t key[2];
            key[0] = qkey[e]; key[1] = kw == 2 ? qkey[QCAP + e] : 0ull;
            const uint32_t nb = qnull[e];
            int slot = -1;
            if (use_smem) {
                const uint32_t h = kw == 1 ? hash_key1(key[0]) : hash_key(key, 2);
                slot = kw == 1 ? table_upsert<true, 1>(st.state, st.keys, st.cap_mask, key, 1, h >> 7, 16, nullptr)
                               : table_upsert<true, 2>(st.state, st.keys, st.cap_mask, key, 2, h >> 7, 16, nullptr);
            }
            if (slot >= 0) {
                const uint32_t cap = st.cap_mask + 1;
                atomicAdd((uint32_t*)(st.lanes + slot), 1u);  // lane 0 = row count (< 2^32 rows per CTA and launch)
#pragma unroll
                for (int s = 0; s < NA; s++) {
                    if ((nb >> s) & 1u) continue;
                    const ValOps vo = a.vops[s];
                    const uint64_t v = qval[s * QCAP + e];
                    if (vo.cnt_smem != 0xFF) atomicAdd((uint32_t*)(st.lanes + (size_t)vo.cnt_smem * cap + slot), 1u);
                    for (int k = 0; k < vo.n_ops; k++)
                        smem_lane_update(vo.op[k], st.lanes + (size_t)vo.smem_lane[k] * cap + slot, to_lane_class(v, vo.arg_class, vo.lane_class[k]));
                }
            } else {
                uint64_t vals[NA > 0 ? NA : 1];
#pragma unroll
                for (int s = 0; s < NA; s++) vals[s] = qval[s * QCAP + e];
                global_update_row<NA>(a, key, vals, nb);
            }
        }
        __syncwarp();
    }
    if (use_smem) smem_table_flush(st, a);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) passed += __shfl_xor_sync(0xFFFFFFFFu, passed, d);
    if (lane == 0 && passed) atomicAdd((unsigned long long*)a.rows_passed, (unsigned long long)passed);
}

// ------------------------------------------------------------------------------------------
// no GROUP BY: accumulate in registers (per value column up to 3 lane operations), warp shuffle
// reduction, one set of global atomics per warp
// ------------------------------------------------------------------------------------------
template <int NP, int NA>
__global__ void __launch_bounds__(DIRECT_THREADS, 1) k_agg_scalar_direct(const __grid_constant__ AggArgs a) {
    const int lane = threadIdx.x & 31;
    uint64_t rows = 0;
    uint64_t acc[NA > 0 ? NA : 1][3], cnt[NA > 0 ? NA : 1];
#pragma unroll
    for (int s = 0; s < NA; s++) {
        cnt[s] = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) acc[s][k] = k < a.vops[s].n_ops ? lane_identity(a.vops[s].op[k]) : 0;
    }
    const int64_t nquads = (a.nrows + 3) >> 2;
    for (int64_t q = (int64_t)blockIdx.x * DIRECT_THREADS + threadIdx.x; q < nquads; q += (int64_t)gridDim.x * DIRECT_THREADS) {
        const int64_t left = a.nrows - q * 4;
        uint32_t pass = left >= 4 ? 0xFu : ((1u << left) - 1u);
#pragma unroll
        for (int t = 0; t < NP; t++) {
            uint64_t pv[4]; uint32_t pnm;
            load_quad(a.cols[t], q, a.nrows, pv, pnm);
            pass &= term_mask(a.direct.term[t], pv) & ~pnm;
        }
        rows += __popc(pass);
#pragma unroll
        for (int s = 0; s < NA; s++) {
            uint64_t v[4]; uint32_t nm;
            load_quad(a.cols[NP + s], q, a.nrows, v, nm);
            const uint32_t ok = pass & ~nm;
            cnt[s] += __popc(ok);
            const ValOps vo = a.vops[s];
#pragma unroll
            for (int k = 0; k < 3; k++) {
                if (k >= vo.n_ops) break;
                const int op = vo.op[k];
                if (op == LN_ADD_F64 && vo.arg_class == VC_F64) {
                    double d = bits_f64(acc[s][k]);
#pragma unroll
                    for (int j = 0; j < 4; j++) if ((ok >> j) & 1u) d += bits_f64(v[j]);
                    acc[s][k] = f64_bits(d);
                } else if (op == LN_ADD_I64) {
#pragma unroll
                    for (int j = 0; j < 4; j++) if ((ok >> j) & 1u) acc[s][k] += v[j];
                } else acc[s][k] = combine4_generic(op, acc[s][k], v, ok, vo.arg_class, vo.lane_class[k]);
            }
        }
    }
    const GroupTable& gt = a.gt;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xFFFFFFFFu, rows, d);
#pragma unroll
    for (int s = 0; s < NA; s++) {
        const ValOps vo = a.vops[s];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt[s] += __shfl_xor_sync(0xFFFFFFFFu, cnt[s], d);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (k >= vo.n_ops) break;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) acc[s][k] = lane_combine(vo.op[k], acc[s][k], __shfl_xor_sync(0xFFFFFFFFu, acc[s][k], d));
        }
    }
    if (lane == 0) {  // the single group lives in slot 0 (capacity 1): lane l is gt.lanes[l]
        if (rows) { atomicAdd((unsigned long long*)gt.lanes, (unsigned long long)rows); atomicAdd((unsigned long long*)a.rows_passed, (unsigned long long)rows); }
#pragma unroll
        for (int s = 0; s < NA; s++) {
            const ValOps vo = a.vops[s];
            if (cnt[s] == 0) continue;
            if (vo.cnt_glob) atomicAdd((unsigned long long*)(gt.lanes + vo.cnt_glob), (unsigned long long)cnt[s]);
            for (int k = 0; k < vo.n_ops; k++) lane_atomic<false>(vo.op[k], gt.lanes + vo.glob_lane[k], acc[s][k]);
        }
    }
}

// ------------------------------------------------------------------------------------------
template <class K>
static inline int direct_grid(K kernel, size_t smem, int sm_count, int64_t nrows) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, DIRECT_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int64_t want = ((nrows + 3) / 4 + DIRECT_THREADS - 1) / DIRECT_THREADS;
    const int64_t full = (int64_t)per_sm * sm_count;  // persistent grid: whole CTAs per SM x 148 SMs
    return (int)(want < full ? want : full);
}

template <int NP, int NA>
static inline cudaError_t launch_direct(const AggArgs& a, int sm_count, size_t smem, cudaStream_t s, bool grouped) {
    if (grouped) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_agg_group_direct<NP, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k_agg_group_direct<NP, NA><<<direct_grid(k_agg_group_direct<NP, NA>, smem, sm_count, a.nrows), DIRECT_THREADS, smem, s>>>(a);
    } else {
        k_agg_scalar_direct<NP, NA><<<direct_grid(k_agg_scalar_direct<NP, NA>, 0, sm_count, a.nrows), DIRECT_THREADS, 0, s>>>(a);
    }
    return cudaGetLastError();
}
template <int NP>
static inline cudaError_t launch_direct_np(const AggArgs& a, int na, int sm_count, size_t smem, cudaStream_t s, bool grouped) {
    switch (na) {
        case 0: return launch_direct<NP, 0>(a, sm_count, smem, s, grouped);
        case 1: return launch_direct<NP, 1>(a, sm_count, smem, s, grouped);
        case 2: return launch_direct<NP, 2>(a, sm_count, smem, s, grouped);
        case 3: return launch_direct<NP, 3>(a, sm_count, smem, s, grouped);
        default: return launch_direct<NP, 4>(a, sm_count, smem, s, grouped);
    }
}

}  // namespace bk
