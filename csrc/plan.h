// plan.h — host side: parse the plan word stream (include/bkgpu_plan.h), run the reference's type
// inference, and lower the operator subtree to device programs.  Pure C++, no CUDA calls, so
// bkgpu_plan_explain() works on a machine without a GPU.
#pragma once
#include <stdint.h>
#include <memory>
#include <string>
#include <vector>
#include "dev_types.h"

namespace bk {

struct HExpr {
    int node_type = 0, col_type = 0;
    std::vector<HExpr> ch;
    int tuple_id = 0, slot_id = 0;
    int fn_op = 0;
    std::string name;
    std::vector<int> arg_types;
    int return_type = 0;
    // literal: value image in `lit_prim`
    bool lit_null = false;
    uint64_t lit_bits = 0;
    int lit_prim = 0;
    std::string lit_str;              // STRING literal text (lit_prim == BK_STRING) until type inference folds it into an image
    int final_slot = 0, inter_slot = 0;
    bool distinct = false;            // AGG_EXPR count_distinct / sum_distinct / avg_distinct: name holds the base function
    bool is_constant = false;
};

struct HNode {
    int node_type = 0;
    int64_t limit = -1;
    std::vector<HNode> ch;
    int tuple_id = -1;
    std::vector<HExpr> conjuncts;            // filter / join conditions
    int agg_tuple_id = -1;
    std::vector<HExpr> group_exprs, agg_fns; // agg
    std::vector<HExpr> order_exprs;          // sort
    std::vector<int> is_asc, is_null_first;
    int join_type = 0;
    int64_t offset = 0;
};

struct HTuple { int tuple_id; std::vector<std::pair<int, int>> slots; /* (slot_id, prim) */ };

struct ColRef { int tuple_id, slot_id, prim; };

struct OutCol {
    int tuple_id, slot_id, prim;
    int kind;       // 0 = plain value column, 1 = AVG intermediate blob (two device images: sum, count)
};

enum PlanKind { PK_AGG = 1, PK_FILTER = 2, PK_SORT = 3, PK_JOIN_AGG = 4, PK_JOIN = 5 };

struct SortKey { int out_reg; int prim; bool asc, null_first; };

struct Compiled {
    int kind = 0;
    std::vector<HTuple> tuples;
    // main (probe-side / only) scan
    int scan_tuple = -1;
    std::vector<ColRef> cols;          // columns the device program references (index = DevCol index)
    std::vector<int> col_side;         // per column: 0 = probe / only scan tuple, 1 = build side of a join
    Program prog;                      // outputs: [predicate][keys...][agg args...] / sort keys
    int n_const = 0;
    // aggregate
    AggPlan ap;
    bool is_merge = false, emit_default = false;
    bool has_direct = false;
    DirectPlan direct;
    std::vector<int> direct_cols;      // indices into `cols`, in [terms][key][values] order
    std::vector<int> arg_cols_mask;    // per aggregate: bitmask of cols its argument reads (nullability per batch)
    std::vector<bool> arg_can_null;    // per aggregate: NULL possible even when all its columns are valid
    std::vector<OutCol> out_cols;
    int64_t agg_limit = -1;
    // filter-only / sort: rows of the scan tuple are returned
    int64_t limit = -1, offset = 0;
    std::vector<SortKey> sort_keys;
    // hash join (build side = outer child, probe side = inner child; join_node.cpp:920-1022)
    int build_tuple = -1;
    int build_key_col = -1;            // index into cols (side 1)
    int probe_key_col = -1;            // index into cols
    int join_type = 0;
    int join_key_prim = 0;
    // FK -> PK fast path: the same aggregate lowered over the virtual joined tuple (build-side columns are gathered
    // to probe-row alignment first); usable while every probe row has exactly one match
    std::shared_ptr<Compiled> jfast;
    std::vector<int> jfast_of_main;    // for every column of jfast->cols: index of the same column in this->cols
    // operators ABOVE the aggregate ([LIMIT ->] [SORT ->] [HAVING ->] AGG, exec_node.cpp:347-394): a second, small fragment of kind
    // PK_SORT / PK_FILTER whose "scan tuple" is the aggregate's output row (its columns = this->out_cols, in order)
    std::shared_ptr<Compiled> post;
    std::string explain;
};

// returns 0 or a negative BKGPU_E* code; `err` receives the message
int compile_plan(const uint8_t* desc, size_t len, Compiled& out, std::string& err);

// host-side ExprValue::cast_to on canonical images (x86 semantics == the reference's build)
uint64_t host_cast_prim(uint64_t v, int from, int to);
// literal.cpp: the text of a literal as a DATETIME / TIMESTAMP / DATE / TIME image (ExprValue::cast_to from STRING)
uint64_t parse_literal(const char* text, size_t length, int to_prim);
int host_prim_class(int prim);
int prim_storage(int prim);       // SType of the column buffer carrying `prim`
int storage_bytes(int stype);

}  // namespace bk
