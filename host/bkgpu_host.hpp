// bkgpu_host.hpp — C++ host side above the C ABI (include/bkgpu.h): the operator interface of the
// reference, mirrored so that a BaikalDB maintainer (or a test) drives the GPU path exactly the way
// Region::select drives any ExecNode (src/store/region.cpp:3069-3216):
//
//      ExecNode::create_tree -> root->open(&state) -> while (!eos) root->get_next(&state, &batch, &eos) -> root->close(&state)
//
// Mirrors (names, argument meaning, error convention):
//   ExecNode::init/open/get_next/close, add_child/replace_child     include/exec/exec_node.h:79-153
//   RuntimeState (error_code/error_msg, is_cancelled, counters)      include/runtime/runtime_state.h:312-324
//   pb::Plan / pb::PlanNode / pb::Expr / pb::ExprNode (pre-order)    proto/plan.proto:495-511, proto/expr.proto:67-84
//   MockScanNode feeding synthetic rows into a tree                  test/test_window.cpp:117-125,257-262
// Batches are COLUMN batches (what Chunk / RowBatch::transfer_rowbatch_to_arrow produce in the vectorized engine,
// include/runtime/row_batch.h:182-201), not MemRows.  Header-only, C++17, links against libbkgpu.so.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "../include/bkgpu.h"

namespace bkgpu {

// ---------------------------------------------------------------- plan description (pb::Plan mirror)
struct Expr {
    int node_type = 0, col_type = 0;
    std::vector<Expr> children;
    int tuple_id = 0, slot_id = 0;
    int64_t int_val = 0; double double_val = 0; bool bool_val = false;
    int fn_op = 0; std::string name; std::vector<int> arg_types; int return_type = 0;
    int final_slot_id = 0, intermediate_slot_id = 0;

    static Expr slot_ref(int tuple, int slot, int type) { Expr e; e.node_type = BK_SLOT_REF; e.col_type = type; e.tuple_id = tuple; e.slot_id = slot; return e; }
    static Expr int_literal(int64_t v) { Expr e; e.node_type = BK_INT_LITERAL; e.col_type = BK_INT64; e.int_val = v; return e; }
    static Expr double_literal(double v) { Expr e; e.node_type = BK_DOUBLE_LITERAL; e.col_type = BK_DOUBLE; e.double_val = v; return e; }
    static Expr null_literal() { Expr e; e.node_type = BK_NULL_LITERAL; e.col_type = BK_NULL_TYPE; return e; }
    static Expr fn(int fn_op, const char* name, std::vector<Expr> args) { Expr e; e.node_type = BK_FUNCTION_CALL; e.fn_op = fn_op; e.name = name; e.children = std::move(args); return e; }
    static Expr predicate(int node_type, int fn_op, const char* name, std::vector<Expr> args) {
        Expr e; e.node_type = node_type; e.col_type = BK_BOOL; e.fn_op = fn_op; e.name = name; e.children = std::move(args); return e;
    }
    // AGG_EXPR: name in {count_star,count,sum,avg,min,max}; intermediate != final only for AVG (expr.proto:59-60)
    static Expr agg(const char* name, int agg_tuple, int final_slot, int inter_slot, std::vector<Expr> args) {
        Expr e; e.node_type = BK_AGG_EXPR; e.name = name; e.tuple_id = agg_tuple; e.final_slot_id = final_slot; e.intermediate_slot_id = inter_slot;
        e.children = std::move(args); return e;
    }
    int count() const { int n = 1; for (auto& c : children) n += c.count(); return n; }
};

struct PlanNode {
    int node_type = 0;
    int64_t limit = -1;
    std::vector<PlanNode> children;
    int tuple_id = 0; int64_t table_id = 0;              // SCAN / SORT
    std::vector<Expr> conjuncts;                         // FILTER / JOIN conditions
    int agg_tuple_id = -1; std::vector<Expr> group_exprs, agg_fn_calls;
    std::vector<Expr> order_exprs; std::vector<bool> is_asc, is_null_first;
    int join_type = BK_INNER_JOIN; int64_t offset = 0;
    int count() const { int n = 1; for (auto& c : children) n += c.count(); return n; }
};

struct TupleDescriptor { int tuple_id; std::vector<std::pair<int, int>> slots; /* (slot_id, pb::PrimitiveType) */ };

class PlanWriter {   // the word stream of include/bkgpu_plan.h
public:
    std::vector<uint8_t> bytes;
    void w(int32_t v) { put(&v, 4); }
    void w64(int64_t v) { put(&v, 8); }
    void f64(double v) { put(&v, 8); }
    void str(const std::string& s) { w((int32_t)s.size()); put(s.data(), s.size()); static const char z[4] = {0, 0, 0, 0}; put(z, (4 - s.size() % 4) % 4); }
    void expr(const Expr& e) { w(e.count()); enode(e); }
    void node(const PlanNode& n) {
        w(n.node_type); w((int32_t)n.children.size()); w64(n.limit);
        switch (n.node_type) {
            case BK_SCAN_NODE: w(n.tuple_id); w64(n.table_id); break;
            case BK_WHERE_FILTER_NODE: case BK_TABLE_FILTER_NODE: case BK_HAVING_FILTER_NODE:
                w((int32_t)n.conjuncts.size()); for (auto& e : n.conjuncts) expr(e); break;
            case BK_AGG_NODE: case BK_MERGE_AGG_NODE:
                w(n.agg_tuple_id); w((int32_t)n.group_exprs.size()); for (auto& e : n.group_exprs) expr(e);
                w((int32_t)n.agg_fn_calls.size()); for (auto& e : n.agg_fn_calls) expr(e); break;
            case BK_SORT_NODE:
                w(n.tuple_id); w((int32_t)n.order_exprs.size());
                for (size_t i = 0; i < n.order_exprs.size(); i++) { expr(n.order_exprs[i]); w(n.is_asc[i] ? 1 : 0); w(n.is_null_first[i] ? 1 : 0); }
                break;
            case BK_JOIN_NODE: w(n.join_type); w((int32_t)n.conjuncts.size()); for (auto& e : n.conjuncts) expr(e); break;
            case BK_LIMIT_NODE: w64(n.offset); break;
            default: break;   // PACKET / SELECT_MANAGER: no payload
        }
        for (auto& c : n.children) node(c);
    }
private:
    void put(const void* p, size_t n) { const uint8_t* b = (const uint8_t*)p; bytes.insert(bytes.end(), b, b + n); }
    void enode(const Expr& e) {
        w(e.node_type); w(e.col_type); w((int32_t)e.children.size());
        switch (e.node_type) {
            case BK_SLOT_REF: w(e.tuple_id); w(e.slot_id); break;
            case BK_NULL_LITERAL: break;
            case BK_BOOL_LITERAL: w(e.bool_val ? 1 : 0); break;
            case BK_INT_LITERAL: w64(e.int_val); break;
            case BK_DOUBLE_LITERAL: f64(e.double_val); break;
            case BK_AGG_EXPR: str(e.name); w(e.tuple_id); w(e.final_slot_id); w(e.intermediate_slot_id); break;
            default: w(e.fn_op); str(e.name); w((int32_t)e.arg_types.size()); for (int a : e.arg_types) w(a); w(e.return_type); break;
        }
        for (auto& c : e.children) enode(c);
    }
};

struct Plan {
    std::vector<TupleDescriptor> tuples;   // sorted by tuple_id
    PlanNode root;
    std::vector<uint8_t> serialize() const {
        PlanWriter out;
        out.w((int32_t)BKGPU_PLAN_MAGIC); out.w(BKGPU_PLAN_VERSION); out.w((int32_t)tuples.size()); out.w(root.count());
        for (auto& t : tuples) { out.w(t.tuple_id); out.w((int32_t)t.slots.size()); for (auto& s : t.slots) { out.w(s.first); out.w(s.second); } }
        out.node(root);
        return out.bytes;
    }
};

// ---------------------------------------------------------------- batches and state
struct Column {
    int tuple_id = 0, slot_id = 0, prim_type = 0, elem_size = 0;
    std::vector<uint8_t> values;
    std::vector<uint8_t> validity;   // Arrow LSB bitmap; empty = no NULLs
    int64_t length = 0;
    template <class T> static Column from(int tuple, int slot, int prim, const std::vector<T>& v) {
        Column c; c.tuple_id = tuple; c.slot_id = slot; c.prim_type = prim; c.elem_size = (int)sizeof(T); c.length = (int64_t)v.size();
        c.values.resize(v.size() * sizeof(T)); if (!v.empty()) memcpy(c.values.data(), v.data(), c.values.size()); return c;
    }
    template <class T> T at(int64_t i) const { T v; memcpy(&v, values.data() + (size_t)i * sizeof(T), sizeof(T)); return v; }
    bool is_null(int64_t i) const { return !validity.empty() && !((validity[(size_t)i >> 3] >> (i & 7)) & 1); }
};

struct RowBatch {   // include/runtime/row_batch.h: here a batch of columns
    std::vector<Column> columns;
    int64_t size() const { return columns.empty() ? 0 : columns[0].length; }
    void clear() { columns.clear(); }
};

struct RuntimeState {
    int device = 0;
    void* nccl_comm = nullptr;
    int64_t row_batch_capacity = 1 << 20;
    int error_code = 0; std::string error_msg;
    int64_t num_scan_rows = 0, num_filter_rows = 0;
    bool cancelled = false;
    void cancel() { cancelled = true; }
    bool is_cancelled() const { return cancelled; }
};

// ---------------------------------------------------------------- operators
class ExecNode {
public:
    virtual ~ExecNode() = default;
    virtual int init(const Plan&) { return 0; }
    virtual int open(RuntimeState* state) { for (auto& c : _children) { int rc = c->open(state); if (rc < 0) return rc; } return 0; }
    virtual int get_next(RuntimeState* state, RowBatch* batch, bool* eos) = 0;
    virtual void close(RuntimeState* state) { for (auto& c : _children) c->close(state); }
    void add_child(std::unique_ptr<ExecNode> c) { _children.push_back(std::move(c)); }
    void replace_child(size_t i, std::unique_ptr<ExecNode> c) { _children[i] = std::move(c); }   // ExecNode::replace_child
protected:
    std::vector<std::unique_ptr<ExecNode>> _children;
};

// leaf: a scan that hands out prepared column batches (the MockScanNode of test/test_window.cpp:117-125)
class ColumnScanNode : public ExecNode {
public:
    explicit ColumnScanNode(std::vector<RowBatch> batches) : _batches(std::move(batches)) {}
    int get_next(RuntimeState*, RowBatch* batch, bool* eos) override {
        batch->clear();
        if (_pos < _batches.size()) *batch = std::move(_batches[_pos++]);
        *eos = _pos >= _batches.size();
        return 0;
    }
private:
    std::vector<RowBatch> _batches; size_t _pos = 0;
};

// the fused AGG -> [FILTER] -> SCAN / SORT / AGG -> JOIN / FILTER subtree on the GPU
class GpuExecNode : public ExecNode {
public:
    ~GpuExecNode() override { if (_h) bkgpu_close(_h); }
    int init(const Plan& plan) override { _plan = plan.serialize(); return 0; }
    int open(RuntimeState* state) override {
        int rc = bkgpu_init(&_h, _plan.data(), _plan.size(), state->device, state->nccl_comm);
        if (rc != 0) return fail(state, rc, nullptr);
        if ((rc = bkgpu_set_option(_h, "batch_capacity", state->row_batch_capacity)) != 0 || (rc = bkgpu_open(_h)) != 0) return fail(state, rc, _h);
        for (auto& child : _children) {   // children in driver-table-first order (the join's outer child is child 0)
            if (child->open(state) < 0) return -1;
            bool eos = false;
            while (!eos) {                 // the child-pull loop of AggNode::open (agg_node.cpp:447-485)
                if (state->is_cancelled()) { bkgpu_cancel(_h); return 0; }
                RowBatch b;
                if (child->get_next(state, &b, &eos) < 0) return -1;
                if (b.columns.empty()) continue;
                std::vector<bkgpu_column> cols(b.columns.size());
                for (size_t i = 0; i < cols.size(); i++) {
                    const Column& c = b.columns[i];
                    cols[i] = bkgpu_column{c.tuple_id, c.slot_id, c.prim_type, c.elem_size, c.values.data(), c.validity.empty() ? nullptr : c.validity.data(), c.length};
                }
                if ((rc = bkgpu_push(_h, cols.data(), (int)cols.size(), b.size(), 0)) != 0) return fail(state, rc, _h);
            }
        }
        if ((rc = bkgpu_finish(_h)) != 0) return fail(state, rc, _h);
        bkgpu_stats st; bkgpu_get_stats(_h, &st);
        state->num_scan_rows += st.rows_scanned; state->num_filter_rows += st.rows_filtered;   // region.cpp:3140-3143
        return 0;
    }
    int get_next(RuntimeState* state, RowBatch* batch, bool* eos) override {
        batch->clear();
        if (state->is_cancelled()) { *eos = true; return 0; }
        bkgpu_column cols[64]; int ncols = 64; int64_t nrows = 0; int e = 0;
        int rc = bkgpu_get_next(_h, cols, &ncols, &nrows, &e);
        if (rc != 0) return fail(state, rc, _h);
        for (int i = 0; i < ncols; i++) {
            Column c; c.tuple_id = cols[i].tuple_id; c.slot_id = cols[i].slot_id; c.prim_type = cols[i].prim_type; c.elem_size = cols[i].elem_size; c.length = nrows;
            c.values.assign((const uint8_t*)cols[i].values, (const uint8_t*)cols[i].values + (size_t)nrows * (size_t)cols[i].elem_size);
            if (cols[i].validity) c.validity.assign(cols[i].validity, cols[i].validity + (nrows + 7) / 8);
            batch->columns.push_back(std::move(c));
        }
        *eos = e != 0;
        return 0;
    }
    void close(RuntimeState* state) override { if (_h) { bkgpu_close(_h); _h = nullptr; } ExecNode::close(state); }
private:
    int fail(RuntimeState* state, int rc, bkgpu_plan* h) { state->error_code = rc; state->error_msg = bkgpu_last_error(h); return -1; }   // negative return + message
    std::vector<uint8_t> _plan;
    bkgpu_plan* _h = nullptr;
};

}  // namespace bkgpu
