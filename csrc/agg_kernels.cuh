// agg_kernels.cuh — device templates of the filter+aggregate kernels (see agg.cu for the overview).
#pragma once
#include "agg.h"
#include "dev_common.cuh"

namespace bk {

// ------------------------------------------------------------------------------------------
// accumulate one row into a table slot (shared or global)
//
// Shared tables hold only the lanes that need per-row work in THIS batch (args.smem_lane maps a
// global lane to its shared lane or 0xFF): the non-NULL counter of an aggregate whose argument
// cannot be NULL in this batch always equals the group's row count, so it is not touched per row;
// the flush adds the shared row count to it (args.alias_mask).
// ------------------------------------------------------------------------------------------
template <bool SHARED, class ArgFn>
__device__ __forceinline__ void accumulate_row(const AggArgs& args, uint64_t* lanes, uint32_t cap, int slot, ArgFn arg) {
    const AggPlan& ap = args.plan;
    lane_atomic<SHARED>(LN_ADD_I64, lanes + slot, 1ull);  // lane 0: group row count
#pragma unroll 1
    for (int i = 0; i < ap.n_agg; i++) {
        const AggSpec a = ap.agg[i];
        if (a.kind == AG_COUNT_STAR) continue;
        uint64_t v; bool isnull;
        arg(i, v, isnull);
        if (isnull) continue;
        if (a.cnt_lane && (!SHARED || a.nullable)) {
            const int l = SHARED ? args.smem_lane[a.cnt_lane] : a.cnt_lane;
            lane_atomic<SHARED>(LN_ADD_I64, lanes + (size_t)l * cap + slot, 1ull);
        }
        if (a.kind == AG_COUNT) continue;
        v = to_lane_class(v, a.arg_vclass, a.vclass);
        const int l = SHARED ? args.smem_lane[a.acc_lane] : a.acc_lane;
        lane_atomic<SHARED>(ap.lane_op[a.acc_lane], lanes + (size_t)l * cap + slot, v);
    }
}

// merge the lane values of one group (a shared-table slot or an imported partial) into the global
// table: AggFnCall::merge (src/expr/agg_fn_call.cpp:779-822) lane by lane.  lane_val(l, v) yields
// the value for global lane l or returns false when the source has nothing for it.
template <class LaneVal>
__device__ __forceinline__ void merge_group(const AggPlan& ap, const GroupTable& gt, const uint64_t* key, LaneVal lane_val) {
    int slot = 0;
    if (ap.n_keyw > 0) {
        slot = table_upsert<false, 0>(gt.state, gt.keys, gt.cap_mask, key, ap.n_keyw,
                                      ap.n_keyw == 1 ? hash_key1(key[0]) : hash_key(key, ap.n_keyw),
                                      (int)(gt.cap_mask + 1), gt.n_groups);
        if (slot < 0) { atomicExch(gt.overflow, 1u); return; }
    }
    const uint32_t cap = gt.cap_mask + 1;
#pragma unroll 1
    for (int l = 0; l < ap.n_lanes; l++) {
        uint64_t v;
        if (!lane_val(l, v)) continue;
        const int op = ap.lane_op[l];
        if (v == lane_identity(op)) continue;
        lane_atomic<false>(op, gt.lanes + (size_t)l * cap + slot, v);
    }
}

// ------------------------------------------------------------------------------------------
// shared-memory table carve-up: [keys n_keyw x cap][lanes n_smem_lanes x cap][state cap]
// ------------------------------------------------------------------------------------------
struct SmemTable {
    uint32_t* state; uint64_t* keys; uint64_t* lanes; uint32_t cap_mask;
};
__device__ __forceinline__ SmemTable smem_table_init(unsigned char* raw, const AggArgs& args) {
    const AggPlan& ap = args.plan;
    SmemTable t;
    const uint32_t cap = 1u << args.smem_cap_log2;
    t.cap_mask = cap - 1;
    t.keys = (uint64_t*)raw;
    t.lanes = t.keys + (size_t)ap.n_keyw * cap;
    t.state = (uint32_t*)(t.lanes + (size_t)args.n_smem_lanes * cap);
    for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) t.state[i] = 0u;
    for (int l = 0; l < ap.n_lanes; l++) {
        const int sl = args.smem_lane[l];
        if (sl == 0xFF) continue;
        const uint64_t id = lane_identity(ap.lane_op[l]);
        for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) t.lanes[(size_t)sl * cap + i] = id;
    }
    __syncthreads();
    return t;
}
__device__ __forceinline__ void smem_table_flush(const SmemTable& t, const AggArgs& args) {
    __syncthreads();
    const AggPlan& ap = args.plan;
    const uint32_t cap = t.cap_mask + 1;
    for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) {
        if (t.state[i] != 2u) continue;
        uint64_t key[MAX_KEYW];
        for (int w = 0; w < ap.n_keyw; w++) key[w] = t.keys[(size_t)w * cap + i];
        merge_group(ap, args.gt, key, [&](int l, uint64_t& v) {
            const int sl = args.smem_lane[l];
            if (sl != 0xFF) { v = t.lanes[(size_t)sl * cap + i]; return true; }
            if ((args.alias_mask >> l) & 1u) { v = t.lanes[i]; return true; }  // counter that follows the row count
            return false;
        });
    }
}

// ------------------------------------------------------------------------------------------
// "direct" kernels: predicate = conjunction of `column <cmp> constant`; key and aggregate
// arguments are plain columns.  args.cols[] is ordered [NP predicate columns][NK key columns]
// [NA value columns] by the host, so every slot index below is a compile-time constant.
// ------------------------------------------------------------------------------------------
template <int NS>
struct Oct {
    uint64_t v[NS > 0 ? NS : 1][8];
    uint32_t nm[NS > 0 ? NS : 1];
};

template <int NS>
__device__ __forceinline__ uint32_t load_slots(const AggArgs& a, int64_t q, Oct<NS>& o) {
    const int64_t row0 = q * 8;
    if (row0 + 8 <= a.nrows) {
#pragma unroll
        for (int s = 0; s < NS; s++) load_oct(a.cols[s], q, o.v[s], o.nm[s]);
        return 0xFFu;
    }
    const int rem = (int)(a.nrows - row0);  // ragged tail: element loads
#pragma unroll
    for (int s = 0; s < NS; s++) {
        o.nm[s] = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j < rem) {
                o.v[s][j] = load_elem(a.cols[s], row0 + j);
                if (elem_is_null(a.cols[s], row0 + j)) o.nm[s] |= 1u << j;
            } else o.v[s][j] = 0;
        }
    }
    return (1u << rem) - 1u;
}

template <int NP, int NS>
__device__ __forceinline__ uint32_t direct_pred(const AggArgs& a, const Oct<NS>& o, uint32_t live) {
    uint32_t pass = live;
#pragma unroll
    for (int t = 0; t < NP; t++) {
        const DirectTerm term = a.direct.term[t];
        uint32_t m = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) m |= (cmp_vals(term.cmp, term.vclass, o.v[t][j], term.cbits) ? 1u : 0u) << j;
        pass &= m & ~o.nm[t];  // NULL or false drops the row (filter_node.cpp:726-734)
    }
    return pass;
}

// ---- no GROUP BY: accumulate in registers, one global merge per CTA ----
template <int NP, int NA>
__global__ void __launch_bounds__(256) k_agg_scalar_direct(const AggArgs a) {
    constexpr int NS = NP + NA;
    const AggPlan& ap = a.plan;
    uint64_t rows = 0;
    constexpr int DA = DIRECT_MAX_AGG;  // the host lowers plans with more aggregates to the generic kernel
    uint64_t acc[DA], cnt[DA];
#pragma unroll
    for (int i = 0; i < DA; i++) { acc[i] = i < ap.n_agg ? lane_identity(ap.lane_op[ap.agg[i].acc_lane]) : 0; cnt[i] = 0; }
    uint64_t passed = 0;
    const int64_t nq = (a.nrows + 7) >> 3;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        Oct<NS> o;
        const uint32_t live = load_slots<NS>(a, q, o);
        const uint32_t pass = direct_pred<NP, NS>(a, o, live);
        const int np = __popc(pass);
        rows += np; passed += np;
        if (NA > 0 && pass) {
#pragma unroll
            for (int i = 0; i < DA; i++) {
                if (i >= ap.n_agg) break;
                const AggSpec g = ap.agg[i];
                if (g.kind == AG_COUNT_STAR) continue;
                const int vs = a.direct.agg_val[i];
#pragma unroll
                for (int s = 0; s < NA; s++) {
                    if (s != vs) continue;
                    const uint32_t ok = pass & ~o.nm[NP + s];
                    cnt[i] += __popc(ok);
                    if (g.kind == AG_COUNT) continue;
                    const int op = ap.lane_op[g.acc_lane];
#pragma unroll
                    for (int j = 0; j < 8; j++)
                        if (ok & (1u << j)) acc[i] = lane_combine(op, acc[i], to_lane_class(o.v[NP + s][j], g.arg_vclass, g.vclass));
                }
            }
        }
    }
    // warp reduce, then one set of global atomics per warp
    const GroupTable& gt = a.gt;
    const uint32_t cap = gt.cap_mask + 1;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { rows += __shfl_xor_sync(0xFFFFFFFFu, rows, d); passed += __shfl_xor_sync(0xFFFFFFFFu, passed, d); }
#pragma unroll
    for (int i = 0; i < DA; i++) {
        if (i >= ap.n_agg) break;
        const AggSpec g = ap.agg[i];
        const int op = ap.lane_op[g.acc_lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            cnt[i] += __shfl_xor_sync(0xFFFFFFFFu, cnt[i], d);
            acc[i] = lane_combine(op, acc[i], __shfl_xor_sync(0xFFFFFFFFu, acc[i], d));
        }
    }
    if ((threadIdx.x & 31) == 0) {
        if (rows) atomicAdd((unsigned long long*)gt.lanes, (unsigned long long)rows);
#pragma unroll
        for (int i = 0; i < DA; i++) {
            if (i >= ap.n_agg) break;
            const AggSpec g = ap.agg[i];
            if (g.kind == AG_COUNT_STAR || cnt[i] == 0) continue;
            if (g.cnt_lane) atomicAdd((unsigned long long*)(gt.lanes + (size_t)g.cnt_lane * cap), (unsigned long long)cnt[i]);
            if (g.kind != AG_COUNT) lane_atomic<false>(ap.lane_op[g.acc_lane], gt.lanes + (size_t)g.acc_lane * cap, acc[i]);
        }
        if (passed) atomicAdd((unsigned long long*)a.rows_passed, (unsigned long long)passed);
    }
}

// ---- one GROUP BY column: per-CTA shared table + global merge ----
template <int NP, int NA>
__global__ void __launch_bounds__(256) k_agg_group_direct(const AggArgs a) {
    constexpr int NS = NP + 1 + NA;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const AggPlan& ap = a.plan;
    const GroupTable& gt = a.gt;
    SmemTable st;
    const bool use_smem = a.smem_cap_log2 > 0;
    if (use_smem) st = smem_table_init(smem_raw, a);
    const uint32_t gcap = gt.cap_mask + 1;
    const uint64_t kmask = ap.key_bits[0] >= 64 ? ~0ull : ((1ull << ap.key_bits[0]) - 1ull);
    uint32_t passed = 0;
    const int64_t nq = (a.nrows + 7) >> 3;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
        Oct<NS> o;
        const uint32_t live = load_slots<NS>(a, q, o);
        const uint32_t pass = direct_pred<NP, NS>(a, o, live);
        passed += __popc(pass);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (!(pass & (1u << j))) continue;
            uint64_t key[2] = {0, 0};
            const bool knull = (o.nm[NP] >> j) & 1u;
            if (!knull) key[0] = o.v[NP][j] & kmask;
            else key[ap.key_null_word[0] ? 1 : 0] |= 1ull << ap.key_null_shift[0];
            auto arg = [&](int i, uint64_t& v, bool& isnull) {
                const int vs = a.direct.agg_val[i];
                v = 0; isnull = true;
#pragma unroll
                for (int s = 0; s < NA; s++)
                    if (s == vs) { v = o.v[NP + 1 + s][j]; isnull = (o.nm[NP + 1 + s] >> j) & 1u; }
            };
            const uint32_t h = ap.n_keyw == 1 ? hash_key1(key[0]) : hash_key(key, 2);
            int slot = -1;
            if (use_smem) {
                slot = ap.n_keyw == 1
                    ? table_upsert<true, 1>(st.state, st.keys, st.cap_mask, key, 1, h >> 7, 16, nullptr)
                    : table_upsert<true, 2>(st.state, st.keys, st.cap_mask, key, 2, h >> 7, 16, nullptr);
            }
            if (slot >= 0) accumulate_row<true>(a, st.lanes, st.cap_mask + 1, slot, arg);
            else {  // group does not fit the shared table: update the global table directly
                slot = table_upsert<false, 0>(gt.state, gt.keys, gt.cap_mask, key, ap.n_keyw, h, (int)gcap, gt.n_groups);
                if (slot < 0) atomicExch(gt.overflow, 1u);
                else accumulate_row<false>(a, gt.lanes, gcap, slot, arg);
            }
        }
    }
    if (use_smem) smem_table_flush(st, a);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) passed += __shfl_xor_sync(0xFFFFFFFFu, passed, d);
    if ((threadIdx.x & 31) == 0 && passed) atomicAdd((unsigned long long*)a.rows_passed, (unsigned long long)passed);
}

template <int NP, int NA>
static inline cudaError_t launch_direct(const AggArgs& a, int grid, size_t smem, cudaStream_t s, bool grouped) {
    if (grouped) {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(k_agg_group_direct<NP, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        k_agg_group_direct<NP, NA><<<grid, 256, smem, s>>>(a);
    } else {
        k_agg_scalar_direct<NP, NA><<<grid, 256, 0, s>>>(a);
    }
    return cudaGetLastError();
}
template <int NP>
static inline cudaError_t launch_direct_np(const AggArgs& a, int na, int grid, size_t smem, cudaStream_t s, bool grouped) {
    switch (na) {
        case 0: return launch_direct<NP, 0>(a, grid, smem, s, grouped);
        case 1: return launch_direct<NP, 1>(a, grid, smem, s, grouped);
        case 2: return launch_direct<NP, 2>(a, grid, smem, s, grouped);
        case 3: return launch_direct<NP, 3>(a, grid, smem, s, grouped);
        default: return launch_direct<NP, 4>(a, grid, smem, s, grouped);
    }
}


}  // namespace bk
