// agg_kernels.cuh — device pieces shared by the aggregate kernels: row accumulation into a table slot,
// group merge (K3), shared-table carve-up and flush (see agg.cu for the overview).
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
        if (a.cnt_lane && a.cnt_owner && (!SHARED || a.nullable)) {
            const int l = SHARED ? args.smem_lane[a.cnt_lane] : a.cnt_lane;
            lane_atomic<SHARED>(LN_ADD_I64, lanes + (size_t)l * cap + slot, 1ull);
        }
        if (a.kind == AG_COUNT || !a.acc_owner) continue;
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
constexpr uint64_t EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;  // "free slot" marker of sentinel-mode shared tables (one-word keys)

// word index of shared lane `sl` of `slot`: SoA, or 16-byte pairs so that two lanes of a group sit in one
// 128-bit word (updated together by one ATOMS.CAS.128 in the lean kernel)
__device__ __forceinline__ size_t smem_lane_word(const AggArgs& a, int sl, uint32_t slot, uint32_t cap) {
    return a.smem_paired ? (((size_t)(sl >> 1) * cap + slot) * 2 + (size_t)(sl & 1)) : ((size_t)sl * cap + slot);
}

struct SmemTable {
    uint32_t* state; uint64_t* keys; uint64_t* lanes; uint32_t cap_mask;
};
__device__ __forceinline__ SmemTable smem_table_init(unsigned char* raw, const AggArgs& args) {
    const AggPlan& ap = args.plan;
    SmemTable t;
    const uint32_t cap = 1u << args.smem_cap_log2;
    t.cap_mask = cap - 1;
    t.keys = (uint64_t*)raw;
    t.lanes = t.keys + (size_t)args.smem_keyw * cap;
    t.state = (uint32_t*)(t.lanes + (size_t)args.n_smem_lanes * cap);
    if (args.smem_sentinel) { for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) t.keys[i] = EMPTY_KEY; }
    else { for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) t.state[i] = 0u; }
    for (int l = 0; l < ap.n_lanes; l++) {
        const int sl = args.smem_lane[l];
        if (sl == 0xFF) continue;
        const uint64_t id = lane_identity(ap.lane_op[l]);
        for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) t.lanes[smem_lane_word(args, sl, i, cap)] = id;
    }
    __syncthreads();
    return t;
}
__device__ __forceinline__ void smem_table_flush(const SmemTable& t, const AggArgs& args) {
    __syncthreads();
    const AggPlan& ap = args.plan;
    const uint32_t cap = t.cap_mask + 1;
    for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) {
        if (args.smem_sentinel ? (t.keys[i] == EMPTY_KEY) : (t.state[i] != 2u)) continue;
        uint64_t key[MAX_KEYW];
        for (int w = 0; w < ap.n_keyw; w++) key[w] = w < args.smem_keyw ? t.keys[(size_t)w * cap + i] : 0ull;  // sentinel mode: further words are 0
        merge_group(ap, args.gt, key, [&](int l, uint64_t& v) {
            const int sl = args.smem_lane[l];
            if (sl != 0xFF) { v = t.lanes[smem_lane_word(args, sl, i, cap)]; return true; }
            if ((args.alias_mask >> l) & 1u) { v = t.lanes[smem_lane_word(args, 0, i, cap)]; return true; }  // counter that follows the row count
            return false;
        });
    }
}


}  // namespace bk
