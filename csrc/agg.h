// agg.h — kernel argument block and host launchers of the filter+aggregate kernels (agg.cu).
#pragma once
#include <cuda_runtime.h>
#include "dev_types.h"

namespace bk {

// hash join (K4): open-addressed multimap from the cast build key image to the build row
struct JoinArgs {
    const uint64_t* keys;    // [cap] key images
    const uint32_t* rows;    // [cap] build row, 0xFFFFFFFF = free slot
    uint32_t cap_mask;
    int32_t enabled;
    int32_t probe_col;       // index into cols of the probe key
    int32_t probe_prim, cast_prim;
    // LEFT / SEMI / ANTI_SEMI (the build side = the reference's outer / driver table is the preserved side, join_node.cpp:1200-1276,
    // joiner.cpp:633-685): `matched[build row]` is set by the probe; the TAIL launch (tail = 1) walks the n_build build rows and emits the
    // unmatched ones NULL-extended (LEFT), the matched ones (SEMI) or the unmatched ones (ANTI_SEMI)
    int32_t join_type;       // pb::JoinType: 1 LEFT, 3 INNER, 4 SEMI, 5 ANTI_SEMI (RIGHT arrives as LEFT with the children swapped)
    int32_t tail;
    uint8_t* matched;
    int64_t n_build;
};

// FK -> PK join fused into the lean aggregate: the GROUP BY key is a build-side (dimension) attribute reached through
// the probe row's foreign key.  At build time the attribute is composed onto the key index (attr_of_key), so the probe
// is ONE random 4-byte read per fact row (served from L2: a 10M-key dimension is 40 MB); rows without a partner drop
// out (inner join), which the filter mask already expresses.
struct JoinProbe {
    int32_t mode;                // 0 = off, 1 = dense index, 2 = packed (key32 << 32 | attr32) open-addressed table
    int32_t key_signed;          // probe key column is a signed 32-bit type (canonical image sign-extends)
    const uint32_t* attr;        // mode 1: attribute by (key image ^ bias) - dense_min
    const uint32_t* present;     // mode 1: build row per key, 0xFFFFFFFF = no such key; nullptr = every key of the range exists
    uint64_t dense_min, dense_size, bias;
    const uint64_t* packed;      // mode 2
    uint32_t packed_mask;
};

struct AggArgs {
    DevCol cols[MAX_COLS];   // direct kernels: ordered [predicate][key][value] columns
    int32_t n_cols;
    int32_t smem_cap_log2;   // log2 slots of the per-CTA shared table (0 = no shared table)
    int64_t nrows;
    AggPlan plan;
    DirectPlan direct;
    Program prog;            // generic path only
    GroupTable gt;
    uint64_t* rows_passed;   // device counter: rows that survived the filter
    JoinArgs join;
    uint8_t smem_lane[MAX_LANES];  // global lane -> shared lane of this batch, 0xFF = not held in shared memory
    int32_t n_smem_lanes;
    uint32_t alias_mask;     // global lanes that receive the shared row count at flush time
    ValOps vops[4];          // direct kernels: per value column
    int32_t smem_keyw;       // key words per slot of the shared table (1 in sentinel mode, else plan.n_keyw)
    int32_t smem_paired;     // shared lanes are interleaved in 16-byte pairs {lane 2p, lane 2p+1} per slot (lean kernel: ATOMS.CAS.128)
    int32_t lean;            // batch qualifies for k_agg_group_lean (see agg_direct.cuh)
    int32_t smem_sentinel;   // shared table of one-word keys: the key word doubles as slot state (EMPTY_KEY = free)
    JoinProbe jp;            // lean kernel only: the key column holds the probe-side foreign key (see JoinProbe)
    int32_t lean_nulls;      // lean kernel: some predicate / value column of this batch carries a validity bitmap
    int32_t lean_mm;         // lean kernel: some value column feeds MIN / MAX lanes or more than one lane
    // warp-private kernel (agg_wp.cuh): chosen by the host for the plainest lean batches
    int32_t scalar_tma;      // 1 = try the TMA-staged scalar kernel (scalar_tma.cu)
    int32_t jp_pipeline;     // 1 = the fused probe issues its dimension lookups one drain ahead (agg_direct.cuh)
    int32_t lean_bank;       // 1 = the lean kernel's drain deals entries to lanes by the bank group of their home slot (option lean_bank)
    int32_t lean_fx;         // 1 = the lean kernel accumulates its double sums as fixed-point limbs with native 32-bit shared atomics (FX, agg_direct.cuh)
    uint32_t fx_ext_off;     // FX: byte offset (in the CTA's dynamic shared memory) of the low-extension limb arrays [value column][slot]
    int32_t wp;              // 1 = launch k_agg_group_wp
    int32_t wp_gcap;         // dense group ids per warp table (multiple of 32)
    int32_t wp_kt_log2;      // log2 words of the CTA's key -> id table
    int32_t wp_warps;        // warps per CTA (one accumulator table each): 8, 12 or 16
    int32_t wp_dense;        // 1 = no key table: id = key - wp_dense_sub (keys known to lie in a range of wp_gcap values)
    uint32_t wp_dense_sub;
};

// shared memory of k_agg_group_wp: key table + id counter + one accumulator set {sums 8 B x na, cnt 4 B} x gcap per warp
inline size_t wp_warp_bytes(int na, uint32_t gcap) { return ((size_t)gcap * (8u * (uint32_t)na + 4u) + 15) & ~(size_t)15; }
inline size_t wp_smem_bytes(int na, uint32_t gcap, int kt_log2, int warps) { return (kt_log2 < 0 ? 0 : ((size_t)8 << kt_log2)) + 16 + wp_warp_bytes(na, gcap) * (size_t)warps; }

// typed columns built from the aggregate's extracted images (input of the post fragment)
struct PostCols { int32_t n; int32_t img[MAX_COLS]; int32_t stype[MAX_COLS]; uint8_t* values[MAX_COLS]; uint8_t* null_bytes[MAX_COLS]; };
cudaError_t launch_images_to_columns(const uint64_t* outv, const uint8_t* outn, uint32_t out_cap, uint32_t n, const PostCols& pc, cudaStream_t s);
size_t agg_smem_bytes(int smem_keyw, int n_smem_lanes, int cap_log2);
// scalar_tma.cu: TMA-staged COUNT(*) WHERE int32 <cmp> c (experiment, option "scalar_tma"); false = not this kernel's shape
bool launch_count_where_tma(const AggArgs& a, int sm_count, cudaStream_t s, cudaError_t* err);
cudaError_t launch_agg(const AggArgs& a, bool direct, int sm_count, cudaStream_t s, const char** kernel_name);
cudaError_t launch_direct_np0(const AggArgs& a, int na, int sm_count, size_t smem, cudaStream_t s, bool grouped);
cudaError_t launch_direct_np1(const AggArgs& a, int na, int sm_count, size_t smem, cudaStream_t s, bool grouped);
cudaError_t launch_direct_np2(const AggArgs& a, int na, int sm_count, size_t smem, cudaStream_t s, bool grouped);
size_t direct_smem_bytes(int smem_keyw, int n_smem_lanes, int cap_log2, int na);
size_t fx_ext_bytes(int na, int cap_log2);
cudaError_t launch_join_build(const DevCol& key, int from_prim, int cast_prim, int64_t nrows, uint64_t* keys, uint32_t* rows, uint32_t cap_mask, cudaStream_t s);
// FK -> PK join fast path: unique build keys, looked up through a dense array (small key range) or a packed
// (key32 << 32 | row) table; build-side columns are gathered to probe-row alignment
struct JoinFast {
    int32_t mode;               // 0 = none, 1 = dense array, 2 = packed 32-bit-key table
    const uint32_t* dense; uint64_t dense_min; uint64_t dense_size; uint64_t bias;
    const uint64_t* packed; uint32_t packed_mask;
};
struct GatherCols { int32_t n; const uint8_t* src[MAX_COLS]; uint8_t* dst[MAX_COLS]; int32_t elem[MAX_COLS]; };
cudaError_t launch_join_minmax(const DevCol& key, int from_prim, int cast_prim, int64_t nrows, uint64_t bias, uint64_t* mm, cudaStream_t s);
cudaError_t launch_join_build_fast(const DevCol& key, int from_prim, int cast_prim, int64_t nrows, const JoinFast& jf, uint32_t* dense_w, uint64_t* packed_w, uint32_t* dup_flag, cudaStream_t s);
cudaError_t launch_join_gather(const DevCol& probe_key, int from_prim, int cast_prim, int64_t nrows, const JoinFast& jf, const GatherCols& gc, uint32_t* miss_flag, cudaStream_t s);
cudaError_t launch_join_compose(const JoinFast& jf, const uint32_t* attr_by_row, uint32_t* attr_of_key, uint64_t* packed_attr, uint32_t* bad_flag, cudaStream_t s);
cudaError_t launch_unpack_validity(const uint8_t* bitmap, int64_t n, uint8_t* null_bytes, cudaStream_t s);
cudaError_t launch_pack_validity(const uint8_t* null_bytes, int64_t n, uint8_t* bitmap, cudaStream_t s);
// a JOIN that returns rows: (probe row, build row) pairs of the joined rows that pass the conditions, then one gather per output column
cudaError_t launch_join_pairs(const AggArgs& a, uint32_t* pairs, uint32_t cap, uint32_t* cursor, int sm_count, cudaStream_t s);
cudaError_t launch_join_rows_gather(const DevCol& src, const uint32_t* pairs, int which, uint32_t n, int elem_bytes, uint8_t* dst, uint8_t* dst_null, cudaStream_t s);
cudaError_t launch_table_init(const GroupTable& gt, const AggPlan& ap, cudaStream_t s, int keep_overflow = 0);
cudaError_t launch_table_clear(const GroupTable& gt, const AggPlan& ap, uint32_t n_occupied, cudaStream_t s);
cudaError_t launch_partial_export_rows(const GroupTable& gt, const AggPlan& ap, uint64_t* dst, uint32_t bound, cudaStream_t s);
cudaError_t launch_partial_merge_rows(const GroupTable& gt, const AggPlan& ap, const uint64_t* src, size_t words_per_rank, uint32_t bound, int nranks, int self,
                                      uint32_t* max_count, cudaStream_t s);
cudaError_t launch_partial_export(const GroupTable& gt, const AggPlan& ap, uint64_t* dst, uint32_t pcap, uint32_t* cursor, cudaStream_t s);
cudaError_t launch_peer_exchange(const GroupTable& gt, const AggPlan& ap, uint64_t* const* d_peers, uint64_t* local, int nranks, int rank, size_t seg_words,
                                 uint32_t pcap, uint64_t seq, uint32_t* cursor, uint32_t* timed_out, cudaStream_t s);
cudaError_t launch_partial_export_parts(const GroupTable& gt, const AggPlan& ap, uint64_t* dst, size_t words_per_seg, uint32_t pcap, uint32_t* cursors, int nranks, cudaStream_t s);
cudaError_t launch_partial_merge(const GroupTable& gt, const AggPlan& ap, const uint64_t* src, size_t words_per_rank, uint32_t pcap, int nranks, cudaStream_t s);
cudaError_t launch_extract(const GroupTable& gt, const AggPlan& ap, uint64_t* outv, uint8_t* outn, uint32_t out_cap, uint32_t* cursor, int emit_default, cudaStream_t s);

}  // namespace bk
