/*
 * bk_oracle.h — CPU restatement of BaikalDB's ROW ENGINE for the analytical hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (baikaldb_b200/, include/bkgpu.h) never links, imports or executes it.
 *
 * PARITY PIN STATUS: the ExprValue arithmetic below is pinned against the reference's
 * known-answer tests (test/test_expr_value.cpp:456-653) and the Arrow fixture of
 * test/test_arrow_compute.cpp:51-277 (tests/test_oracle_golden.py).  For the operator
 * results themselves (FilterNode/AggNode/SortNode/JoinNode) the reference holds no
 * unit-level golden vectors and its row engine cannot be built here (brpc, braft,
 * protobuf+protoc, boost, rocksdb ... absent): "parity unpinned" at operator level;
 * it is defined as agreement of this restatement, the Acero plan the reference's
 * vectorized engine would build (oracle/acero_oracle.py) and the GPU path.
 */
#ifndef BK_ORACLE_H_
#define BK_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bko_column {
    int32_t tuple_id, slot_id, prim_type, elem_size;
    const void* values;
    const uint8_t* validity; /* Arrow LSB bitmap, 1 = valid; NULL = no nulls */
    int64_t length;
} bko_column;

typedef struct bko_result {
    int32_t ncols;
    int64_t nrows;
    bko_column* cols; /* buffers owned by the result */
    int64_t rows_scanned, rows_filtered;
} bko_result;

/* Run the plan (same word stream as include/bkgpu_plan.h) over the given scan columns
 * with the row engine's semantics, single-threaded like one bthread per fragment
 * (src/exec/agg_node.cpp:447-485).  Returns 0 or a negative code with `err` filled. */
int  bko_execute(const uint8_t* plan, size_t len, const bko_column* in_cols, int n_in,
                 bko_result** out, char* err, size_t errlen);
void bko_free_result(bko_result* r);

/* ExprValue known-answer hooks (include/common/expr_value.h). `bits` is the raw 8-byte
 * union image. */
int64_t  bko_ev_compare(int type_a, uint64_t bits_a, int type_b, uint64_t bits_b, int diff_type);
uint64_t bko_ev_cast(int from_type, uint64_t bits, int to_type);
/* MutTableKey memcomparable encoding of one value (include/common/mut_table_key.h:60-160);
 * returns the number of bytes written to out[8]. */
int      bko_key_encode(int type, uint64_t bits, uint8_t out[8]);

#ifdef __cplusplus
}
#endif
#endif
