// agg.cu — K1+K2: fused scan -> filter -> (GROUP BY) aggregate kernels for sm_100a.
//
// Replaces, for one column batch, the reference's hot loop
//     AggNode::open:  do { child->get_next(batch); process_row_batch(batch); } while (!eos)
//         /root/reference/src/exec/agg_node.cpp:447-485,507-545
// with its per-row FilterNode::need_copy (src/exec/filter_node.cpp:726-734), ExecNode::encode_exprs_key
// (src/exec/exec_node.cpp:555-571) and AggFnCall::update (src/expr/agg_fn_call.cpp:496-555).
//
// Data flow: every referenced column is read exactly once with 256-bit non-allocating loads
// (8 rows per thread and load); the predicate is evaluated in registers; passing rows update a
// per-CTA open-addressed hash table in shared memory (native 32-bit ATOMS for counts, CAS loop for
// double sums); each CTA merges its table once into the global group table (RED.ADD.F64 / RED.ADD.64).
// Rows whose group does not fit the shared table go straight to the global table, so any
// cardinality is handled; the global table is also the per-GPU partial state that NCCL ships.
#include "agg_kernels.cuh"
#include "interp.cuh"

namespace bk {

// ------------------------------------------------------------------------------------------
// generic path: the lowered expression program (interp.cuh) is interpreted per row (any predicate,
// computed keys and arguments, multi-column keys).
// ------------------------------------------------------------------------------------------
// one (possibly joined) row through the program into the tables
struct InterpCtx { SmemTable st; bool grouped, use_smem; uint32_t gcap; };
// how: 0 = filter, then aggregate the row; 1 = filter only (SEMI / ANTI probe: does the joined row satisfy the conditions?);
//      2 = aggregate without the filter (the NULL-extended / preserved rows of the join's tail)
__device__ __forceinline__ uint32_t interp_row(const AggArgs& a, const InterpCtx& cx, int64_t row, int64_t brow, int how = 0) {
    const AggPlan& ap = a.plan;
    const GroupTable& gt = a.gt;
    uint64_t out[MAX_GROUP + MAX_AGG + 1];
    uint32_t out_null;
    run_program(a.prog, a.cols, row, out, out_null, brow);
    if (how != 2 && ap.pred_out >= 0 && (((out_null >> ap.pred_out) & 1u) || out[ap.pred_out] == 0)) return 0;
    if (how == 1) return 1;
    auto arg = [&](int i, uint64_t& v, bool& isnull) {
        const int r = ap.agg[i].arg_out;
        if (r == 0xFF) { v = 0; isnull = true; return; }
        v = out[r]; isnull = (out_null >> r) & 1u;
    };
    if (!cx.grouped) {  // single group: slot 0 of the global table
        accumulate_row<false>(a, gt.lanes, cx.gcap, 0, arg);
        return 1;
    }
    uint64_t key[MAX_KEYW];
    for (int w = 0; w < ap.n_keyw; w++) key[w] = 0;
    for (int g = 0; g < ap.n_group; g++) {
        const int r = ap.key_out[g];
        if ((out_null >> r) & 1u) key[ap.key_null_word[g]] |= 1ull << ap.key_null_shift[g];
        else {
            const uint64_t m = ap.key_bits[g] >= 64 ? ~0ull : ((1ull << ap.key_bits[g]) - 1ull);
            key[ap.key_word[g]] |= (out[r] & m) << ap.key_shift[g];
        }
    }
    const uint32_t h = ap.n_keyw == 1 ? hash_key1(key[0]) : hash_key(key, ap.n_keyw);
    int slot = -1;
    if (cx.use_smem) slot = table_upsert<true, 0>(cx.st.state, cx.st.keys, cx.st.cap_mask, key, ap.n_keyw, h >> 7, 16, nullptr);
    if (slot >= 0) accumulate_row<true>(a, cx.st.lanes, cx.st.cap_mask + 1, slot, arg);
    else {
        slot = table_upsert<false, 0>(gt.state, gt.keys, gt.cap_mask, key, ap.n_keyw, h, (int)cx.gcap, gt.n_groups);
        if (slot < 0) atomicExch(gt.overflow, 1u);
        else accumulate_row<false>(a, gt.lanes, cx.gcap, slot, arg);
    }
    return 1;
}

__global__ void __launch_bounds__(256) k_agg_interp(const __grid_constant__ AggArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const AggPlan& ap = a.plan;
    InterpCtx cx;
    cx.grouped = ap.n_keyw > 0;
    cx.use_smem = cx.grouped && a.smem_cap_log2 > 0;
    cx.gcap = a.gt.cap_mask + 1;
    cx.st = SmemTable{};
    if (cx.use_smem) cx.st = smem_table_init(smem_raw, a);
    uint32_t passed = 0;
    if (a.join.enabled && a.join.tail) {   // after the last probe batch: the preserved (build) side's rows the join type asks for
        const int jt = a.join.join_type;
        for (int64_t br = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; br < a.join.n_build; br += (int64_t)gridDim.x * blockDim.x) {
            const bool m = a.join.matched[br] != 0;
            if (jt == 4 /*SEMI*/ ? m : !m) passed += interp_row(a, cx, -1, br, 2);
        }
    } else
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < a.nrows; row += (int64_t)gridDim.x * blockDim.x) {
        if (!a.join.enabled) { passed += interp_row(a, cx, row, -1); continue; }
        // K4 probe: every build row whose cast key equals this row's (Joiner::encode_hash_key + FlatMap seek,
        // src/exec/joiner.cpp:608-622, join_node.cpp:1290-1292); a NULL key never matches (SQL / Acero hashjoin)
        const DevCol& pc = a.cols[a.join.probe_col];
        if (elem_is_null(pc, row)) continue;
        const uint64_t img = cast_prim(load_elem(pc, row), a.join.probe_prim, a.join.cast_prim);
        uint32_t slot = hash_key1(img) & a.join.cap_mask;
        for (;;) {
            const uint32_t br = a.join.rows[slot];
            if (br == 0xFFFFFFFFu) break;
            if (a.join.keys[slot] == img) {
                if (a.join.join_type == 3 /*INNER*/ || !a.join.matched) passed += interp_row(a, cx, row, (int64_t)br);
                else if (a.join.join_type == 1 /*LEFT*/) { if (interp_row(a, cx, row, (int64_t)br)) { passed++; a.join.matched[br] = 1; } }
                else if (interp_row(a, cx, row, (int64_t)br, 1)) a.join.matched[br] = 1;   // SEMI / ANTI_SEMI: only whether a partner exists
            }
            slot = (slot + 1) & a.join.cap_mask;
        }
    }
    if (cx.use_smem) smem_table_flush(cx.st, a);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) passed += __shfl_xor_sync(0xFFFFFFFFu, passed, d);
    if ((threadIdx.x & 31) == 0 && passed) atomicAdd((unsigned long long*)a.rows_passed, (unsigned long long)passed);
}

// ---- a JOIN that returns rows (JoinNode::get_next_for_hash_inner_join / _other_join, join_node.cpp:1200-1326): the probe emits (probe row,
// build row) pairs, a gather per output column follows.  LEFT: matched[] is set here, the tail emits (NONE, build row) for the rest.
constexpr uint32_t JOIN_NO_ROW = 0xFFFFFFFFu;
__global__ void __launch_bounds__(256) k_join_pairs(const __grid_constant__ AggArgs a, uint32_t* pairs, uint32_t cap, uint32_t* cursor) {
    InterpCtx cx; cx.grouped = false; cx.use_smem = false; cx.gcap = 0; cx.st = SmemTable{};
    auto emit = [&](uint32_t prow, uint32_t brow) {
        const uint32_t pos = atomicAdd(cursor, 1u);
        if (pos < cap) { pairs[2 * (size_t)pos] = prow; pairs[2 * (size_t)pos + 1] = brow; }   // (past the capacity only the count matters: the host grows the buffer and reruns)
    };
    if (a.join.tail) {
        for (int64_t br = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; br < a.join.n_build; br += (int64_t)gridDim.x * blockDim.x)
            if (!a.join.matched[br]) emit(JOIN_NO_ROW, (uint32_t)br);
        return;
    }
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < a.nrows; row += (int64_t)gridDim.x * blockDim.x) {
        const DevCol& pc = a.cols[a.join.probe_col];
        if (elem_is_null(pc, row)) continue;
        const uint64_t img = cast_prim(load_elem(pc, row), a.join.probe_prim, a.join.cast_prim);
        uint32_t slot = hash_key1(img) & a.join.cap_mask;
        for (;;) {
            const uint32_t br = a.join.rows[slot];
            if (br == 0xFFFFFFFFu) break;
            if (a.join.keys[slot] == img && interp_row(a, cx, row, (int64_t)br, 1)) {
                if (a.join.matched) a.join.matched[br] = 1;
                emit((uint32_t)row, br);
            }
            slot = (slot + 1) & a.join.cap_mask;
        }
    }
}
// one output column of the joined rows: element `which` of each pair addresses `src`; JOIN_NO_ROW (the NULL-extended side) gives NULL
__global__ void k_join_rows_gather(DevCol src, const uint32_t* pairs, int which, uint32_t n, int eb, uint8_t* dst, uint8_t* dst_null) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t r = pairs[2 * (size_t)i + which];
        const bool isnull = r == JOIN_NO_ROW || elem_is_null(src, r);
        dst_null[i] = isnull ? 1 : 0;
        const uint8_t* s = (const uint8_t*)src.values + (size_t)r * eb;
        uint8_t* d = dst + (size_t)i * eb;
        if (isnull) { for (int b = 0; b < eb; b++) d[b] = 0; continue; }
        switch (eb) {
            case 1: *d = *s; break;
            case 4: *(uint32_t*)d = *(const uint32_t*)s; break;
            case 8: *(uint64_t*)d = *(const uint64_t*)s; break;
            default: for (int b = 0; b < eb; b++) d[b] = s[b]; break;
        }
    }
}

// K4 build: insert (cast key image, row) of every non-NULL build row (Joiner::construct_hash_map, joiner.cpp:624-631)
__global__ void k_join_build(DevCol key, int from_prim, int cast_to, int64_t nrows, uint64_t* keys, uint32_t* rows, uint32_t cap_mask) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
        if (elem_is_null(key, r)) continue;
        const uint64_t img = cast_prim(load_elem(key, r), from_prim, cast_to);
        uint32_t slot = hash_key1(img) & cap_mask;
        for (;;) {
            if (atomicCAS(rows + slot, 0xFFFFFFFFu, (uint32_t)r) == 0xFFFFFFFFu) { keys[slot] = img; break; }
            slot = (slot + 1) & cap_mask;
        }
    }
}

// ------------------------------------------------------------------------------------------
// table maintenance: init, partial export / import (K3), result extraction + finalize
// ------------------------------------------------------------------------------------------
__global__ void k_table_init(GroupTable gt, AggPlan ap, int keep_overflow) {
    const uint32_t cap = gt.cap_mask + 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += gridDim.x * blockDim.x) {
        gt.state[i] = ap.n_keyw == 0 ? 2u : 0u;
        for (int l = 0; l < ap.n_lanes; l++) gt.lanes[(size_t)l * cap + i] = lane_identity(ap.lane_op[l]);
    }
    // (re-initialisation between the export of a partial state and the merge keeps the overflow flag: a table that overflowed
    //  while rows were pushed, or an export that met more groups than the exchange format holds, must still fail the plan)
    if (blockIdx.x == 0 && threadIdx.x < GT_OCC_OFF && !keep_overflow) gt.n_groups[threadIdx.x] = 0;   // fresh request: every counter of the block
    if (blockIdx.x == 0 && threadIdx.x == 0 && keep_overflow) *gt.n_groups = 0;                          // mid-request re-initialisation (multi-GPU merge): the group count only
}

// walk the groups a table holds: the occupied list of a GROUP BY table (slots in insertion order), slot 0 of a scalar aggregate
#define BK_FOR_EACH_GROUP(gt, ap, i)                                                                                   \
    const uint32_t bk_n_ = (ap).n_keyw == 0 ? 1u : *(gt).n_groups;                                                    \
    for (uint32_t bk_j_ = blockIdx.x * blockDim.x + threadIdx.x, i = 0; bk_j_ < bk_n_ && ((i = (ap).n_keyw == 0 ? 0u : (gt).n_groups[GT_OCC_OFF + bk_j_]), true); bk_j_ += gridDim.x * blockDim.x)

// re-initialisation of a table whose `n` occupied slots are known (bkgpu_reset after a finished run): touches n slots, not the capacity
__global__ void k_table_clear(GroupTable gt, AggPlan ap, uint32_t n) {
    const uint32_t cap = gt.cap_mask + 1;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t i = gt.n_groups[GT_OCC_OFF + j];
        gt.state[i] = 0u;
        for (int l = 0; l < ap.n_lanes; l++) gt.lanes[(size_t)l * cap + i] = lane_identity(ap.lane_op[l]);
    }
    // (the counters can go in the same launch: nobody here reads them — `n` came by value, the occupied list sits behind them)
    if (blockIdx.x == 0 && threadIdx.x < GT_OCC_OFF) gt.n_groups[threadIdx.x] = 0;
}

// Partial state layout (fixed capacity `pcap` groups): [u64 n_groups][keys n_keyw x pcap][lanes n_lanes x pcap]
__global__ void k_partial_export(GroupTable gt, AggPlan ap, uint64_t* dst, uint32_t pcap, uint32_t* cursor) {
    const uint32_t cap = gt.cap_mask + 1;
    uint64_t* dkeys = dst + 1;
    uint64_t* dlanes = dkeys + (size_t)ap.n_keyw * pcap;
    BK_FOR_EACH_GROUP(gt, ap, i) {
        if (ap.n_keyw == 0 && gt.lanes[i] == 0) continue;  // no row reached the single group
        const uint32_t pos = atomicAdd(cursor, 1u);
        if (pos >= pcap) { atomicExch(gt.overflow, 1u); continue; }
        for (int w = 0; w < ap.n_keyw; w++) dkeys[(size_t)w * pcap + pos] = gt.keys[(size_t)w * cap + i];
        for (int l = 0; l < ap.n_lanes; l++) dlanes[(size_t)l * pcap + pos] = gt.lanes[(size_t)l * cap + i];
    }
}
__global__ void k_partial_count(uint64_t* dst, const uint32_t* cursor, uint32_t pcap) { dst[0] = *cursor < pcap ? *cursor : pcap; }

// Compact partial state for the all-gather (the default multi-GPU merge): dst[0] = the number of groups this rank holds (its low
// half doubles as the export cursor, so it may exceed `bound`), followed by min(groups, bound) rows of n_keyw + n_lanes words.
// Only `1 + bound * row words` cross NVLink; `bound` follows the group counts seen in earlier runs (api.cu).
__global__ void k_partial_export_rows(GroupTable gt, AggPlan ap, uint64_t* dst, uint32_t bound) {
    const uint32_t cap = gt.cap_mask + 1;
    const int rw = ap.n_keyw + ap.n_lanes;
    BK_FOR_EACH_GROUP(gt, ap, i) {
        if (ap.n_keyw == 0 && gt.lanes[i] == 0) continue;  // no row reached the single group
        const uint32_t pos = atomicAdd((uint32_t*)dst, 1u);
        if (pos >= bound) continue;                         // (the merge sees count > bound and asks for a second, larger exchange)
        uint64_t* row = dst + 1 + (size_t)pos * rw;
        for (int w = 0; w < ap.n_keyw; w++) row[w] = gt.keys[(size_t)w * cap + i];
        for (int l = 0; l < ap.n_lanes; l++) row[ap.n_keyw + l] = gt.lanes[(size_t)l * cap + i];
    }
}
// K3 on the gathered rows: every rank folds the OTHER ranks' groups into its own table (its own groups are already there), so
// nothing is re-initialised.  When some rank held more groups than `bound` nothing is merged anywhere (every rank reads the same
// headers) and *max_count tells the host to repeat the exchange with a larger bound.
__global__ void k_partial_merge_rows(GroupTable gt, AggPlan ap, const uint64_t* src, size_t words_per_rank, uint32_t bound, int nranks, int self, uint32_t* max_count) {
    uint32_t mx = 0;
    for (int r = 0; r < nranks; r++) { const uint32_t c = (uint32_t)src[(size_t)r * words_per_rank]; mx = c > mx ? c : mx; }
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *max_count = mx;
    if (mx > bound) return;
    const int r = blockIdx.y;
    if (r == self) return;
    const uint64_t* base = src + (size_t)r * words_per_rank;
    const uint32_t n = (uint32_t)base[0];
    const int rw = ap.n_keyw + ap.n_lanes;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint64_t* row = base + 1 + (size_t)i * rw;
        uint64_t key[MAX_KEYW];
        for (int w = 0; w < ap.n_keyw; w++) key[w] = row[w];
        merge_group(ap, gt, key, [&](int l, uint64_t& v) { v = row[ap.n_keyw + l]; return true; });
    }
}

// ---- merge over NVLink peer memory (option "peer_merge"): export, exchange and merge without a collective library call ----
// Every rank owns one buffer [2 parities][nranks segments of the partial-state layout] + [2][nranks] sequence flags, mapped into
// all peers through CUDA IPC.  k_peer_export writes this rank's groups straight into segment `rank` of EVERY peer's buffer,
// k_peer_publish releases them (count, system-scope fence, sequence flag), k_peer_wait acquires the peers' flags, then the
// ordinary K3 merge runs on the local buffer.  Parity = step & 1: a peer one step ahead writes the other half.
__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) { uint64_t v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__global__ void k_peer_export(GroupTable gt, AggPlan ap, uint64_t* const* peers, int nranks, int rank, size_t seg_words, size_t parity_off, uint32_t pcap, uint32_t* cursor) {
    const uint32_t cap = gt.cap_mask + 1;
    BK_FOR_EACH_GROUP(gt, ap, i) {
        if (ap.n_keyw == 0 && gt.lanes[i] == 0) continue;
        const uint32_t pos = atomicAdd(cursor, 1u);
        if (pos >= pcap) { atomicExch(gt.overflow, 1u); continue; }
        for (int r = 0; r < nranks; r++) {
            uint64_t* dkeys = peers[r] + parity_off + (size_t)rank * seg_words + 1;
            uint64_t* dlanes = dkeys + (size_t)ap.n_keyw * pcap;
            for (int w = 0; w < ap.n_keyw; w++) dkeys[(size_t)w * pcap + pos] = gt.keys[(size_t)w * cap + i];
            for (int l = 0; l < ap.n_lanes; l++) dlanes[(size_t)l * pcap + pos] = gt.lanes[(size_t)l * cap + i];
        }
    }
}
__global__ void k_peer_publish(uint64_t* const* peers, int nranks, int rank, size_t seg_words, size_t parity_off, size_t flag_off, const uint32_t* cursor, uint32_t pcap, uint64_t seq) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    const uint32_t n = *cursor < pcap ? *cursor : pcap;
    peers[r][parity_off + (size_t)rank * seg_words] = n;     // (the entries themselves were written by the kernel before this one)
    __threadfence_system();
    st_release_sys(peers[r] + flag_off + rank, seq);
}
__global__ void k_peer_wait(const uint64_t* flags, int nranks, uint64_t seq, long long limit_cycles, uint32_t* timed_out) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + r) < seq) {
        if (clock64() - t0 > limit_cycles) { atomicExch(timed_out, 1u); break; }   // a peer died: report instead of spinning forever
        __nanosleep(200);
    }
    __threadfence_system();
}
cudaError_t launch_peer_exchange(const GroupTable& gt, const AggPlan& ap, uint64_t* const* d_peers, uint64_t* local, int nranks, int rank, size_t seg_words,
                                 uint32_t pcap, uint64_t seq, uint32_t* cursor, uint32_t* timed_out, cudaStream_t s) {
    if (nranks > 1024) return cudaErrorInvalidValue;
    const size_t parity_off = (size_t)(seq & 1) * (size_t)nranks * seg_words;
    const size_t flag_off = 2 * (size_t)nranks * seg_words + (size_t)(seq & 1) * (size_t)nranks;
    const uint32_t cap = gt.cap_mask + 1;
    int grid = (int)((cap + 255) / 256); if (grid > 1184) grid = 1184;
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    k_peer_export<<<grid, 256, 0, s>>>(gt, ap, d_peers, nranks, rank, seg_words, parity_off, pcap, cursor);
    k_peer_publish<<<1, 1024, 0, s>>>(d_peers, nranks, rank, seg_words, parity_off, flag_off, cursor, pcap, seq);
    k_peer_wait<<<1, 1024, 0, s>>>(local + flag_off, nranks, seq, 4000000000ll, timed_out);   // ~2 s at 1.9 GHz
    return cudaGetLastError();
}

// hash repartition (the reference's exchange between fragments, src/exec/exchange_sender_node.cpp:867-957): every group of this
// rank's table goes to the segment of the rank that OWNS its key; segment layout = the partial state layout above
__device__ __forceinline__ uint32_t owner_of(const uint64_t* key, int kw, int nranks) {
    uint64_t h = 0xC2B2AE3D27D4EB4Full;
    for (int i = 0; i < kw; i++) { h ^= key[i]; h *= 0x9FB21C651E98DF25ull; h ^= h >> 29; }
    return (uint32_t)((h >> 17) % (uint64_t)nranks);
}
__global__ void k_partial_export_parts(GroupTable gt, AggPlan ap, uint64_t* dst, size_t words_per_seg, uint32_t pcap, uint32_t* cursors, int nranks) {
    const uint32_t cap = gt.cap_mask + 1;
    BK_FOR_EACH_GROUP(gt, ap, i) {
        uint64_t key[MAX_KEYW];
        for (int w = 0; w < ap.n_keyw; w++) key[w] = gt.keys[(size_t)w * cap + i];
        const uint32_t o = owner_of(key, ap.n_keyw, nranks);
        const uint32_t pos = atomicAdd(cursors + o, 1u);
        if (pos >= pcap) { atomicExch(gt.overflow, 1u); continue; }
        uint64_t* dkeys = dst + (size_t)o * words_per_seg + 1;
        uint64_t* dlanes = dkeys + (size_t)ap.n_keyw * pcap;
        for (int w = 0; w < ap.n_keyw; w++) dkeys[(size_t)w * pcap + pos] = key[w];
        for (int l = 0; l < ap.n_lanes; l++) dlanes[(size_t)l * pcap + pos] = gt.lanes[(size_t)l * cap + i];
    }
}
__global__ void k_partial_counts(uint64_t* dst, size_t words_per_seg, const uint32_t* cursors, uint32_t pcap, int nranks) {
    const int r = threadIdx.x;
    if (r < nranks) dst[(size_t)r * words_per_seg] = cursors[r] < pcap ? cursors[r] : pcap;
}

// K3: fold `nranks` exported partials into the (re-initialised) global table
__global__ void k_partial_merge(GroupTable gt, AggPlan ap, const uint64_t* src, size_t words_per_rank, uint32_t pcap, int nranks) {
    for (int r = 0; r < nranks; r++) {
        const uint64_t* base = src + (size_t)r * words_per_rank;
        const uint32_t n = (uint32_t)base[0] < pcap ? (uint32_t)base[0] : pcap;   // (a count beyond the segment's capacity never indexes past it)
        const uint64_t* skeys = base + 1;
        const uint64_t* slanes = skeys + (size_t)ap.n_keyw * pcap;
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            uint64_t key[MAX_KEYW];
            for (int w = 0; w < ap.n_keyw; w++) key[w] = skeys[(size_t)w * pcap + i];
            merge_group(ap, gt, key, [&](int l, uint64_t& v) { v = slanes[(size_t)l * pcap + i]; return true; });
        }
    }
}

// Result extraction: one output row per occupied slot.  Output columns (canonical 64-bit images +
// one null byte each), in this order: group expressions, then per aggregate its final value and,
// for AVG, the intermediate {sum, count} pair (AggFnCall::finalize, agg_fn_call.cpp:927-990).
__global__ void k_extract(GroupTable gt, AggPlan ap, uint64_t* outv, uint8_t* outn, uint32_t out_cap, uint32_t* cursor, int emit_default) {
    const uint32_t cap = gt.cap_mask + 1;
    BK_FOR_EACH_GROUP(gt, ap, i) {
        const uint64_t nrows = gt.lanes[i];
        if (ap.n_keyw == 0 && nrows == 0 && !emit_default) continue;
        const uint32_t pos = atomicAdd(cursor, 1u);
        if (pos >= out_cap) continue;
        int c = 0;
        for (int g = 0; g < ap.n_group; g++, c++) {
            const bool isnull = ap.key_null_word[g] != 0xFF &&
                                ((gt.keys[(size_t)ap.key_null_word[g] * cap + i] >> ap.key_null_shift[g]) & 1ull);
            uint64_t v = gt.keys[(size_t)ap.key_word[g] * cap + i] >> ap.key_shift[g];
            if (ap.key_bits[g] < 64) v &= (1ull << ap.key_bits[g]) - 1ull;
            // sign-extend narrow signed keys back to their canonical image
            if (ap.key_bits[g] == 32 && prim_class(ap.key_prim[g]) == VC_I64) v = (uint64_t)(int64_t)(int32_t)v;
            outv[(size_t)c * out_cap + pos] = isnull ? 0 : v;
            outn[(size_t)c * out_cap + pos] = isnull ? 1 : 0;
        }
        for (int k = 0; k < ap.n_agg; k++) {
            const AggSpec a = ap.agg[k];
            if (a.hidden) continue;
            const uint64_t cnt = a.cnt_lane ? gt.lanes[(size_t)a.cnt_lane * cap + i] : nrows;
            const uint64_t acc = gt.lanes[(size_t)a.acc_lane * cap + i];
            uint64_t v = 0; uint8_t isnull = 0;
            switch (a.kind) {
                case AG_COUNT_STAR: v = nrows; break;
                case AG_COUNT: v = cnt; break;
                case AG_COUNT_MERGE: v = acc; break;
                case AG_AVG:
                    if (cnt == 0) isnull = 1; else v = f64_bits(__ddiv_rn(bits_f64(acc), (double)(int64_t)cnt));
                    break;
                default: if (cnt == 0) isnull = 1; else v = acc; break;  // SUM / MIN / MAX: NULL without input
            }
            outv[(size_t)c * out_cap + pos] = v; outn[(size_t)c * out_cap + pos] = isnull; c++;
            if (a.kind == AG_AVG) {  // intermediate blob {double sum; int64 count}
                // the blank row of an empty scalar aggregate initialises only its counts (agg_node.cpp:489-503): blob stays NULL
                outv[(size_t)c * out_cap + pos] = acc; outn[(size_t)c * out_cap + pos] = (ap.n_keyw == 0 && nrows == 0) ? 1 : 0; c++;
                outv[(size_t)c * out_cap + pos] = cnt; outn[(size_t)c * out_cap + pos] = 0; c++;
            }
        }
    }
}

// The aggregate's extracted rows (canonical 64-bit images + null bytes, column-major with stride out_cap) as typed Arrow-layout
// columns on the device: the input of the post fragment ([LIMIT ->] [SORT ->] [HAVING ->] above the aggregate).
__global__ void k_images_to_columns(const uint64_t* outv, const uint8_t* outn, uint32_t out_cap, uint32_t n, PostCols pc) {
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        for (int c = 0; c < pc.n; c++) {
            const int img = pc.img[c];
            const uint64_t v = outv[(size_t)img * out_cap + r];
            pc.null_bytes[c][r] = outn[(size_t)img * out_cap + r];
            switch (pc.stype[c]) {
                case ST_I32: ((int32_t*)pc.values[c])[r] = (int32_t)(int64_t)v; break;
                case ST_U32: ((uint32_t*)pc.values[c])[r] = (uint32_t)v; break;
                case ST_F32: ((float*)pc.values[c])[r] = (float)bits_f64(v); break;
                case ST_U8: ((uint8_t*)pc.values[c])[r] = v ? 1 : 0; break;
                case ST_BLOB16: ((uint64_t*)pc.values[c])[2 * (size_t)r] = v; ((uint64_t*)pc.values[c])[2 * (size_t)r + 1] = outv[(size_t)(img + 1) * out_cap + r]; break;
                default: ((uint64_t*)pc.values[c])[r] = v; break;
            }
        }
    }
}
cudaError_t launch_images_to_columns(const uint64_t* outv, const uint8_t* outn, uint32_t out_cap, uint32_t n, const PostCols& pc, cudaStream_t s) {
    if (!n) return cudaSuccess;
    int grid = (int)((n + 255) / 256); if (grid > 1184) grid = 1184;
    k_images_to_columns<<<grid, 256, 0, s>>>(outv, outn, out_cap, n, pc);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------
size_t agg_smem_bytes(int smem_keyw, int n_smem_lanes, int cap_log2) {
    if (cap_log2 <= 0 || smem_keyw == 0) return 0;
    return ((size_t)(smem_keyw + n_smem_lanes) * 8 + 4) << cap_log2;
}

// shared memory of the direct GROUP BY kernel: the table plus one compaction queue per warp
size_t direct_smem_bytes(int smem_keyw, int n_smem_lanes, int cap_log2, int na) {
    const size_t table = (agg_smem_bytes(smem_keyw, n_smem_lanes, cap_log2) + 127) & ~(size_t)127;
    const size_t q_direct = ((size_t)(2 + na) * 128 + 16) * 8;  // QCAP = 128 entries per warp, up to 2 key words
    const size_t q_lean = ((((size_t)(1 + na) * 160 + 20) + 15) & ~(size_t)15) * 8 + 8;   // LEAN_QCAP = 160-entry ring, one key word, 128-byte aligned (+ slack for the base)
    const size_t queue = q_direct > q_lean ? q_direct : q_lean;
    const int warps = (DIRECT_THREADS > LEAN_THREADS ? DIRECT_THREADS : LEAN_THREADS) / 32;   // sized for the larger of the two CTA shapes
    return table + queue * warps;
}

// FX (agg_direct.cuh): one low-extension limb (4 bytes) per value column and slot, behind the queues
size_t fx_ext_bytes(int na, int cap_log2) { return ((size_t)4 * (size_t)na) << cap_log2; }

template <class K>
static int occupancy_grid(K kernel, size_t smem, int sm_count) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * sm_count;  // persistent grid: a whole number of CTAs per SM (148 SMs on B200)
}

cudaError_t launch_agg(const AggArgs& a, bool direct, int sm_count, cudaStream_t s, const char** kernel_name) {
    const bool grouped = a.plan.n_keyw > 0;
    const size_t smem = grouped ? agg_smem_bytes(a.smem_keyw, a.n_smem_lanes, a.smem_cap_log2) : 0;
    if (a.nrows <= 0) return cudaSuccess;
    if (direct && !grouped && a.scalar_tma) {
        cudaError_t e = cudaSuccess;
        if (launch_count_where_tma(a, sm_count, s, &e)) { *kernel_name = "k_count_where_tma"; return e; }
    }
    if (direct) {
        const size_t dsmem = grouped ? direct_smem_bytes(a.smem_keyw, a.n_smem_lanes, a.smem_cap_log2, a.direct.n_vals) + (a.lean_fx ? fx_ext_bytes(a.direct.n_vals, a.smem_cap_log2) : 0) : 0;
        *kernel_name = grouped ? "k_agg_group_direct" : "k_agg_scalar_direct";
        switch (a.direct.n_terms) {
            case 0: return launch_direct_np0(a, a.direct.n_vals, sm_count, dsmem, s, grouped);
            case 1: return launch_direct_np1(a, a.direct.n_vals, sm_count, dsmem, s, grouped);
            default: return launch_direct_np2(a, a.direct.n_vals, sm_count, dsmem, s, grouped);
        }
    }
    *kernel_name = "k_agg_interp";
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_agg_interp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int grid = occupancy_grid(k_agg_interp, smem, sm_count);
    int64_t want = (a.nrows + 255) / 256;
    if (want < grid) grid = (int)want;
    k_agg_interp<<<grid, 256, smem, s>>>(a);
    return cudaGetLastError();
}

// ---- FK -> PK fast path (see JoinFast in agg.h) ----
__global__ void k_join_minmax(DevCol key, int from_prim, int cast_to, int64_t nrows, uint64_t bias, uint64_t* mm) {
    uint64_t lo = ~0ull, hi = 0;
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r0 < nrows; r0 += 4 * T) {   // four independent loads in flight per thread
        uint64_t v[4]; bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t r = r0 + u * T;
            ok[u] = r < nrows && !elem_is_null(key, r);
            v[u] = ok[u] ? load_elem(key, r) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (!ok[u]) continue;
            const uint64_t x = cast_prim(v[u], from_prim, cast_to) ^ bias;
            lo = x < lo ? x : lo; hi = x > hi ? x : hi;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const uint64_t l2 = __shfl_xor_sync(0xFFFFFFFFu, lo, d), h2 = __shfl_xor_sync(0xFFFFFFFFu, hi, d);
        lo = l2 < lo ? l2 : lo; hi = h2 > hi ? h2 : hi;
    }
    if ((threadIdx.x & 31) == 0) { atomicMin((unsigned long long*)mm, (unsigned long long)lo); atomicMax((unsigned long long*)(mm + 1), (unsigned long long)hi); }
}
__global__ void k_join_build_fast(DevCol key, int from_prim, int cast_to, int64_t nrows, JoinFast jf, uint32_t* dense_w, uint64_t* packed_w, uint32_t* dup_flag) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t)gridDim.x * blockDim.x) {
        if (elem_is_null(key, r)) continue;   // a NULL key never matches
        const uint64_t img = cast_prim(load_elem(key, r), from_prim, cast_to);
        if (jf.mode == 1) {
            const uint64_t off = (img ^ jf.bias) - jf.dense_min;
            if (off >= jf.dense_size) { atomicExch(dup_flag + 2, 1u); continue; }   // outside the range learned from an earlier run of the plan: the host rebuilds
            if (atomicExch(dense_w + off, (uint32_t)r) != 0xFFFFFFFFu) atomicExch(dup_flag, 1u);
        } else {
            const uint32_t k32 = (uint32_t)img;
            const uint64_t v = ((uint64_t)k32 << 32) | (uint32_t)r;
            uint32_t slot = (k32 * 0x9E3779B1u) & jf.packed_mask;
            for (;;) {
                const uint64_t old = atomicCAS((unsigned long long*)(packed_w + slot), ~0ull, (unsigned long long)v);
                if (old == ~0ull) break;
                if ((uint32_t)(old >> 32) == k32) { atomicExch(dup_flag, 1u); break; }   // duplicate key: not a PK
                slot = (slot + 1) & jf.packed_mask;
            }
        }
    }
}
__device__ __forceinline__ uint32_t join_lookup(const JoinFast& jf, uint64_t img) {
    if (jf.mode == 1) {
        const uint64_t off = (img ^ jf.bias) - jf.dense_min;
        return off < jf.dense_size ? __ldg(jf.dense + off) : 0xFFFFFFFFu;
    }
    const uint32_t k32 = (uint32_t)img;
    uint32_t slot = (k32 * 0x9E3779B1u) & jf.packed_mask;
    for (;;) {
        const uint64_t v = __ldg((const unsigned long long*)(jf.packed + slot));
        if (v == ~0ull) return 0xFFFFFFFFu;
        if ((uint32_t)(v >> 32) == k32) return (uint32_t)v;
        slot = (slot + 1) & jf.packed_mask;
    }
}
// gather the build-side columns to probe-row alignment; four independent rows per thread keep the random
// lookups and gathers in flight together
__global__ void __launch_bounds__(256) k_join_gather(DevCol pk, int from_prim, int cast_to, int64_t nrows, JoinFast jf, GatherCols gc, uint32_t* miss_flag) {
    const int64_t T = (int64_t)gridDim.x * blockDim.x;
    bool miss = false;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < nrows; base += 4 * T) {
        uint32_t br[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t r = base + u * T;
            br[u] = 0xFFFFFFFEu;   // beyond the batch
            if (r < nrows) br[u] = elem_is_null(pk, r) ? 0xFFFFFFFFu : join_lookup(jf, cast_prim(load_elem(pk, r), from_prim, cast_to));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int64_t r = base + u * T;
            if (br[u] == 0xFFFFFFFEu) continue;
            if (br[u] == 0xFFFFFFFFu) { miss = true; continue; }
            for (int c = 0; c < gc.n; c++) {
                if (gc.elem[c] == 8) ((uint64_t*)gc.dst[c])[r] = __ldg((const unsigned long long*)gc.src[c] + br[u]);
                else if (gc.elem[c] == 4) ((uint32_t*)gc.dst[c])[r] = __ldg((const uint32_t*)gc.src[c] + br[u]);
                else gc.dst[c][r] = __ldg(gc.src[c] + br[u]);
            }
        }
    }
    if (__any_sync(0xFFFFFFFFu, miss) && (threadIdx.x & 31) == 0) atomicExch(miss_flag, 1u);
}
// K4 fused probe (JoinProbe): compose the grouped-by build attribute onto the key index built by k_join_build_fast
__global__ void k_join_compose(JoinFast jf, const uint32_t* attr_by_row, uint32_t* attr_of_key, uint64_t* packed_attr, uint32_t* bad_flag) {
    const uint64_t n = jf.mode == 1 ? jf.dense_size : (uint64_t)jf.packed_mask + 1;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        if (jf.mode == 1) {
            const uint32_t r = jf.dense[i];
            attr_of_key[i] = r == 0xFFFFFFFFu ? 0u : __ldg(attr_by_row + r);
        } else {
            const uint64_t e = jf.packed[i];
            uint64_t o = ~0ull;
            if (e != ~0ull) {
                o = (e & 0xFFFFFFFF00000000ull) | __ldg(attr_by_row + (uint32_t)e);
                if (o == ~0ull) atomicExch(bad_flag, 1u);   // (key, attribute) = (0xFFFFFFFF, 0xFFFFFFFF) would read as a free slot
            }
            packed_attr[i] = o;
        }
    }
}
cudaError_t launch_join_compose(const JoinFast& jf, const uint32_t* attr_by_row, uint32_t* attr_of_key, uint64_t* packed_attr, uint32_t* bad_flag, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(bad_flag, 0, 4, s);
    if (e != cudaSuccess) return e;
    const uint64_t n = jf.mode == 1 ? jf.dense_size : (uint64_t)jf.packed_mask + 1;
    int grid = (int)((n + 255) / 256); if (grid > 148 * 16) grid = 148 * 16; if (grid < 1) grid = 1;
    k_join_compose<<<grid, 256, 0, s>>>(jf, attr_by_row, attr_of_key, packed_attr, bad_flag);
    return cudaGetLastError();
}
cudaError_t launch_join_minmax(const DevCol& key, int from_prim, int cast_prim_, int64_t nrows, uint64_t bias, uint64_t* mm, cudaStream_t s) {
    const uint64_t init[2] = {~0ull, 0ull};
    cudaError_t e = cudaMemcpyAsync(mm, init, 16, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess || nrows == 0) return e;
    int grid = (int)((nrows + 255) / 256); if (grid > 148 * 8) grid = 148 * 8;
    k_join_minmax<<<grid, 256, 0, s>>>(key, from_prim, cast_prim_, nrows, bias, mm);
    return cudaGetLastError();
}
cudaError_t launch_join_build_fast(const DevCol& key, int from_prim, int cast_prim_, int64_t nrows, const JoinFast& jf, uint32_t* dense_w, uint64_t* packed_w, uint32_t* dup_flag, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(dup_flag, 0, 12, s);   // [0] duplicate key, [1] fused probe unusable (set by the compose step), [2] key outside the guessed range
    if (e != cudaSuccess) return e;
    if (jf.mode == 1) e = cudaMemsetAsync(dense_w, 0xFF, jf.dense_size * 4, s); else e = cudaMemsetAsync(packed_w, 0xFF, ((size_t)jf.packed_mask + 1) * 8, s);
    if (e != cudaSuccess || nrows == 0) return e;
    int grid = (int)((nrows + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    k_join_build_fast<<<grid, 256, 0, s>>>(key, from_prim, cast_prim_, nrows, jf, dense_w, packed_w, dup_flag);
    return cudaGetLastError();
}
cudaError_t launch_join_gather(const DevCol& probe_key, int from_prim, int cast_prim_, int64_t nrows, const JoinFast& jf, const GatherCols& gc, uint32_t* miss_flag, cudaStream_t s) {
    if (nrows == 0) return cudaSuccess;
    int grid = (int)((nrows + 1023) / 1024); if (grid > 148 * 8) grid = 148 * 8; if (grid < 1) grid = 1;
    k_join_gather<<<grid, 256, 0, s>>>(probe_key, from_prim, cast_prim_, nrows, jf, gc, miss_flag);
    return cudaGetLastError();
}

__global__ void k_unpack_validity(const uint8_t* bitmap, int64_t n, uint8_t* null_bytes) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        null_bytes[i] = bitmap ? (((bitmap[i >> 3] >> (i & 7)) & 1) ? 0 : 1) : 0;
}
__global__ void k_pack_validity(const uint8_t* null_bytes, int64_t n, uint8_t* bitmap) {
    const int64_t nbytes = (n + 7) >> 3;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nbytes; b += (int64_t)gridDim.x * blockDim.x) {
        uint8_t v = 0;
        for (int j = 0; j < 8; j++) { const int64_t i = b * 8 + j; if (i >= n || !null_bytes[i]) v |= (uint8_t)(1u << j); }
        bitmap[b] = v;
    }
}
cudaError_t launch_unpack_validity(const uint8_t* bitmap, int64_t n, uint8_t* null_bytes, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    int grid = (int)((n + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    k_unpack_validity<<<grid, 256, 0, s>>>(bitmap, n, null_bytes);
    return cudaGetLastError();
}
cudaError_t launch_pack_validity(const uint8_t* null_bytes, int64_t n, uint8_t* bitmap, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    int grid = (int)(((n + 7) / 8 + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    k_pack_validity<<<grid, 256, 0, s>>>(null_bytes, n, bitmap);
    return cudaGetLastError();
}
cudaError_t launch_join_pairs(const AggArgs& a, uint32_t* pairs, uint32_t cap, uint32_t* cursor, int sm_count, cudaStream_t s) {
    const int64_t n = a.join.tail ? a.join.n_build : a.nrows;
    if (n == 0) return cudaSuccess;
    int64_t g64 = (n + 255) / 256; if (g64 > (int64_t)sm_count * 8) g64 = (int64_t)sm_count * 8;
    const int grid = g64 < 1 ? 1 : (int)g64;
    k_join_pairs<<<grid, 256, 0, s>>>(a, pairs, cap, cursor);
    return cudaGetLastError();
}
cudaError_t launch_join_rows_gather(const DevCol& src, const uint32_t* pairs, int which, uint32_t n, int eb, uint8_t* dst, uint8_t* dst_null, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    int grid = (int)((n + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    k_join_rows_gather<<<grid, 256, 0, s>>>(src, pairs, which, n, eb, dst, dst_null);
    return cudaGetLastError();
}
cudaError_t launch_join_build(const DevCol& key, int from_prim, int cast_prim_, int64_t nrows, uint64_t* keys, uint32_t* rows, uint32_t cap_mask, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(rows, 0xFF, (size_t)(cap_mask + 1) * 4, s);
    if (e != cudaSuccess || nrows == 0) return e;
    int grid = (int)((nrows + 255) / 256); if (grid > 148 * 16) grid = 148 * 16;
    k_join_build<<<grid, 256, 0, s>>>(key, from_prim, cast_prim_, nrows, keys, rows, cap_mask);
    return cudaGetLastError();
}
cudaError_t launch_table_init(const GroupTable& gt, const AggPlan& ap, cudaStream_t s, int keep_overflow) {
    const uint32_t cap = gt.cap_mask + 1;
    int grid = (int)((cap + 255) / 256); if (grid > 1184) grid = 1184;
    k_table_init<<<grid, 256, 0, s>>>(gt, ap, keep_overflow);
    return cudaGetLastError();
}
cudaError_t launch_table_clear(const GroupTable& gt, const AggPlan& ap, uint32_t n_occupied, cudaStream_t s) {
    int grid = (int)((n_occupied + 255) / 256); if (grid > 1184) grid = 1184; if (grid < 1) grid = 1;
    k_table_clear<<<grid, 256, 0, s>>>(gt, ap, n_occupied);
    return cudaGetLastError();
}
cudaError_t launch_partial_export_rows(const GroupTable& gt, const AggPlan& ap, uint64_t* dst, uint32_t bound, cudaStream_t s) {
    const uint32_t cap = gt.cap_mask + 1;
    int grid = (int)((cap + 255) / 256); if (grid > 148) grid = 148;   // (the kernel walks the occupied list, grid-stride)
    cudaError_t e = cudaMemsetAsync(dst, 0, 8, s);
    if (e != cudaSuccess) return e;
    k_partial_export_rows<<<grid, 256, 0, s>>>(gt, ap, dst, bound);
    return cudaGetLastError();
}
cudaError_t launch_partial_merge_rows(const GroupTable& gt, const AggPlan& ap, const uint64_t* src, size_t words_per_rank, uint32_t bound, int nranks, int self,
                                      uint32_t* max_count, cudaStream_t s) {
    int gx = (int)((bound + 255) / 256); if (gx > 64) gx = 64; if (gx < 1) gx = 1;
    k_partial_merge_rows<<<dim3((unsigned)gx, (unsigned)nranks), 256, 0, s>>>(gt, ap, src, words_per_rank, bound, nranks, self, max_count);
    return cudaGetLastError();
}
cudaError_t launch_partial_export(const GroupTable& gt, const AggPlan& ap, uint64_t* dst, uint32_t pcap, uint32_t* cursor, cudaStream_t s) {
    const uint32_t cap = gt.cap_mask + 1;
    int grid = (int)((cap + 255) / 256); if (grid > 1184) grid = 1184;
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    k_partial_export<<<grid, 256, 0, s>>>(gt, ap, dst, pcap, cursor);
    k_partial_count<<<1, 1, 0, s>>>(dst, cursor, pcap);
    return cudaGetLastError();
}
cudaError_t launch_partial_export_parts(const GroupTable& gt, const AggPlan& ap, uint64_t* dst, size_t words_per_seg, uint32_t pcap, uint32_t* cursors, int nranks, cudaStream_t s) {
    if (nranks > 1024) return cudaErrorInvalidValue;
    const uint32_t cap = gt.cap_mask + 1;
    int grid = (int)((cap + 255) / 256); if (grid > 1184) grid = 1184;
    cudaError_t e = cudaMemsetAsync(cursors, 0, sizeof(uint32_t) * (size_t)nranks, s);
    if (e != cudaSuccess) return e;
    k_partial_export_parts<<<grid, 256, 0, s>>>(gt, ap, dst, words_per_seg, pcap, cursors, nranks);
    k_partial_counts<<<1, 1024, 0, s>>>(dst, words_per_seg, cursors, pcap, nranks);
    return cudaGetLastError();
}
cudaError_t launch_partial_merge(const GroupTable& gt, const AggPlan& ap, const uint64_t* src, size_t words_per_rank, uint32_t pcap, int nranks, cudaStream_t s) {
    int grid = (int)((pcap + 255) / 256); if (grid > 1184) grid = 1184; if (grid < 1) grid = 1;
    k_partial_merge<<<grid, 256, 0, s>>>(gt, ap, src, words_per_rank, pcap, nranks);
    return cudaGetLastError();
}
cudaError_t launch_extract(const GroupTable& gt, const AggPlan& ap, uint64_t* outv, uint8_t* outn, uint32_t out_cap, uint32_t* cursor, int emit_default, cudaStream_t s) {
    const uint32_t cap = gt.cap_mask + 1;
    int grid = (int)((cap + 255) / 256); if (grid > 148) grid = 148;   // (the kernel walks the occupied list, grid-stride)
    cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    k_extract<<<grid, 256, 0, s>>>(gt, ap, outv, outn, out_cap, cursor, emit_default);
    return cudaGetLastError();
}

}  // namespace bk
