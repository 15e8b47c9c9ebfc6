// plan.cpp — plan word stream -> typed trees -> device programs.
//
// Mirrors what the reference does at plan time, in this order:
//   ExecNode::create_tree / ExprNode::create_tree      src/exec/exec_node.cpp:347-394, src/expr/expr_node.cpp:401-445
//   ScalarFnCall::type_inferer                         src/expr/scalar_fn_call.cpp:40-120
//   FunctionManager::complete_fn                       src/expr/fn_manager.cpp:316-409
//   AggFnCall::type_inferer                            src/expr/agg_fn_call.cpp:87-122
//   Literal::get_value / cast_to_col_type              include/expr/literal.h:196-206
// and then lowers instead of interpreting: expressions become postfix bytecode whose operand
// classes are fixed here (no run-time type tags on the device).
#include "plan.h"
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <map>
#include "../include/bkgpu.h"
#include "datetime.h"

namespace bk {

// ---------------------------------------------------------------- type helpers
static bool is_int_t(int t) { return t >= BK_INT8 && t <= BK_UINT64; }
static bool is_uint_t(int t) { return t >= BK_UINT8 && t <= BK_UINT64; }
static bool is_double_t(int t) { return t == BK_FLOAT || t == BK_DOUBLE; }
static bool is_datetime_family(int t) { return t == BK_DATETIME || t == BK_TIMESTAMP || t == BK_DATE || t == BK_TIME; }
static bool is_numeric_path_type(int t) { return t == BK_BOOL || is_int_t(t) || is_double_t(t) || is_datetime_family(t); }

int host_prim_class(int prim) {
    switch (prim) {
        case BK_FLOAT: case BK_DOUBLE: return VC_F64;
        case BK_BOOL: case BK_UINT8: case BK_UINT16: case BK_UINT32: case BK_UINT64:
        case BK_TIMESTAMP: case BK_DATE: case BK_DATETIME: return VC_U64;
        default: return VC_I64;
    }
}
int prim_storage(int prim) {  // src/common/common.cpp:514-544
    switch (prim) {
        case BK_INT8: case BK_INT16: case BK_INT32: case BK_TIME: return ST_I32;
        case BK_INT64: return ST_I64;
        case BK_UINT8: case BK_UINT16: case BK_UINT32: case BK_TIMESTAMP: case BK_DATE: return ST_U32;
        case BK_UINT64: case BK_DATETIME: return ST_U64;
        case BK_FLOAT: return ST_F32;
        case BK_DOUBLE: return ST_F64;
        case BK_BOOL: return ST_U8;
        case BK_STRING: return ST_BLOB16;
        default: return -1;
    }
}
int storage_bytes(int st) {
    switch (st) { case ST_I32: case ST_U32: case ST_F32: return 4; case ST_U8: return 1; case ST_BLOB16: return 16; default: return 8; }
}
static uint64_t dbits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static double bitsd(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }

// ExprValue::cast_to on canonical images, executed with the host's (x86-64) conversion semantics —
// the same ones the reference binary gets from its static_casts (expr_value.h:340-410,502-611).
uint64_t host_cast_prim(uint64_t v, int from, int to) {
    if (from != to && dt_is_family(from) && dt_is_family(to)) return dt_family_cast(v, from, to);   // (callers reject TIME sources first)
    int fc = host_prim_class(from);
    if (to == BK_DOUBLE || to == BK_FLOAT) {
        double d = fc == VC_F64 ? bitsd(v) : (fc == VC_U64 ? (double)v : (double)(int64_t)v);
        if (to == BK_FLOAT) d = (double)(float)d;
        return dbits(d);
    }
    if (to == BK_BOOL) return fc == VC_F64 ? (bitsd(v) != 0.0) : (v != 0);
    if (fc == VC_F64) {
        double d = bitsd(v);
        switch (to) {
            case BK_INT8: return (uint64_t)(int64_t)(int8_t)d;
            case BK_INT16: return (uint64_t)(int64_t)(int16_t)d;
            case BK_INT32: case BK_TIME: return (uint64_t)(int64_t)(int32_t)d;
            case BK_INT64: return (uint64_t)(int64_t)d;
            case BK_UINT8: return (uint64_t)(uint8_t)d;
            case BK_UINT16: return (uint64_t)(uint16_t)d;
            case BK_UINT32: case BK_TIMESTAMP: case BK_DATE: return (uint64_t)(uint32_t)d;
            default: return (uint64_t)d;
        }
    }
    switch (to) {
        case BK_INT8: return (uint64_t)(int64_t)(int8_t)v;
        case BK_INT16: return (uint64_t)(int64_t)(int16_t)v;
        case BK_INT32: case BK_TIME: return (uint64_t)(int64_t)(int32_t)v;
        case BK_UINT8: return (uint64_t)(uint8_t)v;
        case BK_UINT16: return (uint64_t)(uint16_t)v;
        case BK_UINT32: case BK_TIMESTAMP: case BK_DATE: return (uint64_t)(uint32_t)v;
        default: return v;
    }
}

// ---------------------------------------------------------------- reader
struct Reader {
    const int32_t* w; size_t n, pos = 0; bool fail = false; std::string err;
    int32_t rd() { if (pos >= n) { bad("truncated plan"); return 0; } return w[pos++]; }
    int64_t rd64() { uint32_t lo = (uint32_t)rd(), hi = (uint32_t)rd(); return (int64_t)(((uint64_t)hi << 32) | lo); }
    std::string rdstr() {
        int32_t len = rd();
        size_t words = ((size_t)(len < 0 ? 0 : len) + 3) / 4;
        if (len < 0 || len > 4096 || pos + words > n) { bad("bad string in plan"); return ""; }
        std::string s((const char*)(w + pos), (size_t)len); pos += words; return s;
    }
    void bad(const char* m) { if (!fail) { fail = true; err = m; } }
};

static bool is_literal_node(int nt) {
    return (nt >= BK_NULL_LITERAL && nt <= BK_STRING_LITERAL) || nt == BK_TIMESTAMP_LITERAL || nt == BK_DATETIME_LITERAL || nt == BK_DATE_LITERAL || nt == BK_TIME_LITERAL;
}

static void parse_enode(Reader& r, HExpr& e, int& remaining, int depth) {
    if (remaining <= 0 || depth > 64) { r.bad("expr node list does not match its node count"); return; }
    remaining--;
    e.node_type = r.rd(); e.col_type = r.rd();
    int nch = r.rd();
    switch (e.node_type) {
        case BK_SLOT_REF: e.tuple_id = r.rd(); e.slot_id = r.rd(); break;
        case BK_NULL_LITERAL: e.lit_null = true; e.lit_prim = BK_NULL_TYPE; break;
        case BK_BOOL_LITERAL: e.lit_bits = r.rd() ? 1 : 0; e.lit_prim = BK_BOOL; break;
        case BK_INT_LITERAL: e.lit_bits = (uint64_t)r.rd64(); e.lit_prim = BK_INT64; break;
        case BK_DOUBLE_LITERAL: e.lit_bits = (uint64_t)r.rd64(); e.lit_prim = BK_DOUBLE; break;
        case BK_STRING_LITERAL: e.lit_str = r.rdstr(); e.lit_prim = BK_STRING; break;   // (only where inference folds it into a date/time image)
        // DeriveExprNode.int_val carries the image (Literal::init, include/expr/literal.h:95-114)
        case BK_DATETIME_LITERAL: e.lit_bits = (uint64_t)r.rd64(); e.lit_prim = BK_DATETIME; break;
        case BK_TIMESTAMP_LITERAL: e.lit_bits = (uint64_t)(uint32_t)r.rd64(); e.lit_prim = BK_TIMESTAMP; break;
        case BK_DATE_LITERAL: e.lit_bits = (uint64_t)(uint32_t)r.rd64(); e.lit_prim = BK_DATE; break;
        case BK_TIME_LITERAL: e.lit_bits = (uint64_t)(int64_t)(int32_t)r.rd64(); e.lit_prim = BK_TIME; break;
        case BK_AGG_EXPR:
            e.name = r.rdstr(); e.tuple_id = r.rd(); e.final_slot = r.rd(); e.inter_slot = r.rd();
            // count_distinct / sum_distinct / avg_distinct are COUNT / SUM / AVG (name_type_map, agg_fn_call.cpp:32-40) that a MERGE_AGG node UPDATES
            // from its input rows instead of merging intermediates (AggFnCall::merge, agg_fn_call.cpp:719-727): the planner put the argument
            // into the GROUP BY of the aggregate below (select_planner.cpp:640-667), so each distinct value arrives once
            if (e.name == "count_distinct" || e.name == "sum_distinct" || e.name == "avg_distinct") { e.distinct = true; e.name.resize(e.name.size() - 9); }
            break;
        case BK_FUNCTION_CALL: case BK_IS_NULL_PREDICATE: case BK_IN_PREDICATE: case BK_NOT_PREDICATE:
        case BK_AND_PREDICATE: case BK_OR_PREDICATE: case BK_XOR_PREDICATE: case BK_IS_TRUE_PREDICATE: {
            e.fn_op = r.rd(); e.name = r.rdstr();
            int na = r.rd();
            if (na < 0 || na > 8) { r.bad("bad n_arg_types"); return; }
            for (int i = 0; i < na; i++) e.arg_types.push_back(r.rd());
            e.return_type = r.rd();
        } break;
        case BK_LIKE_PREDICATE: case BK_ROW_EXPR:
            r.bad("LIKE / ROW expressions are outside the GPU path"); return;
        default: r.bad("unknown expr node type"); return;
    }
    if (nch < 0 || nch > 1024) { r.bad("bad expr num_children"); return; }
    e.ch.resize((size_t)nch);
    for (int i = 0; i < nch && !r.fail; i++) parse_enode(r, e.ch[(size_t)i], remaining, depth + 1);
}
static void parse_expr(Reader& r, HExpr& e) {
    int n = r.rd();
    parse_enode(r, e, n, 0);
    if (!r.fail && n != 0) r.bad("expr node count mismatch");
}
static void parse_node(Reader& r, HNode& nd, int& remaining, int depth) {
    if (remaining <= 0 || depth > 32) { r.bad("plan node list does not match its node count"); return; }
    remaining--;
    nd.node_type = r.rd();
    int nch = r.rd();
    nd.limit = r.rd64();
    auto exprs = [&](std::vector<HExpr>& v) {
        int n = r.rd();
        if (n < 0 || n > 256) { r.bad("bad expression count"); return; }
        v.resize((size_t)n);
        for (int i = 0; i < n && !r.fail; i++) parse_expr(r, v[(size_t)i]);
    };
    switch (nd.node_type) {
        case BK_SCAN_NODE: nd.tuple_id = r.rd(); (void)r.rd64(); break;
        case BK_WHERE_FILTER_NODE: case BK_TABLE_FILTER_NODE: case BK_HAVING_FILTER_NODE: exprs(nd.conjuncts); break;
        case BK_AGG_NODE: case BK_MERGE_AGG_NODE: nd.agg_tuple_id = r.rd(); exprs(nd.group_exprs); exprs(nd.agg_fns); break;
        case BK_SORT_NODE: {
            nd.tuple_id = r.rd();
            int n = r.rd();
            if (n < 0 || n > 64) { r.bad("bad order expr count"); return; }
            nd.order_exprs.resize((size_t)n);
            for (int i = 0; i < n && !r.fail; i++) {
                parse_expr(r, nd.order_exprs[(size_t)i]);
                nd.is_asc.push_back(r.rd()); nd.is_null_first.push_back(r.rd());
            }
        } break;
        case BK_JOIN_NODE: nd.join_type = r.rd(); exprs(nd.conjuncts); break;
        case BK_LIMIT_NODE: nd.offset = r.rd64(); break;
        case BK_PACKET_NODE: case BK_SELECT_MANAGER_NODE: break;
        default: r.bad("plan node type outside the GPU path"); return;
    }
    if (nch < 0 || nch > 2) { r.bad("bad plan num_children"); return; }
    nd.ch.resize((size_t)nch);
    for (int i = 0; i < nch && !r.fail; i++) parse_node(r, nd.ch[(size_t)i], remaining, depth + 1);
}

// ---------------------------------------------------------------- type inference
struct Infer {
    const std::vector<HTuple>* tuples;
    std::string err;
    int code = 0;
    int slot_type(int tuple_id, int slot_id) const {
        for (auto& t : *tuples) if (t.tuple_id == tuple_id) for (auto& s : t.slots) if (s.first == slot_id) return s.second;
        return BK_INVALID_TYPE;
    }
    bool fail(int c, const char* fmt, ...) {
        char buf[256]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        if (!code) { code = c; err = buf; }
        return false;
    }
};
static bool expr_constant(const HExpr& e) {
    if (e.node_type == BK_SLOT_REF || e.node_type == BK_AGG_EXPR) return false;
    for (auto& c : e.ch) if (!expr_constant(c)) return false;
    return true;
}
static void complete(HExpr& e, int nargs, int at, int rt) { e.arg_types.assign((size_t)nargs, at); e.return_type = rt; }

// named builtins of one / two DOUBLE arguments (src/expr/internal_functions.cpp:101-317,336-350) -> MathFn / Math2Fn, -1 = not one
static int math1_fn(const std::string& n) {
    static const std::pair<const char*, int> t[] = {{"sqrt", MF_SQRT}, {"sign", MF_SIGN}, {"sin", MF_SIN}, {"asin", MF_ASIN}, {"cos", MF_COS}, {"acos", MF_ACOS},
                                                    {"tan", MF_TAN}, {"cot", MF_COT}, {"atan", MF_ATAN}, {"ln", MF_LN}, {"bit_count", MF_BIT_COUNT}};
    for (auto& p : t) if (n == p.first) return p.second;
    return -1;
}
static int math2_fn(const std::string& n) {
    static const std::pair<const char*, int> t[] = {{"mod", MF2_FMOD}, {"log", MF2_LOG}, {"pow", MF2_POW}, {"power", MF2_POW}, {"greatest", MF2_GREATEST}, {"least", MF2_LEAST}};
    for (auto& p : t) if (n == p.first) return p.second;
    return -1;
}

// A literal's value as the image of type `to` — ExprValue::cast_to (include/common/expr_value.h:502-611) on the host.
// `via_text`: Literal::cast_to_col_type (include/expr/literal.h:204-210) — a numeric literal that meets a date/time type is first
// written out in decimal and then read as a date ("20240131" -> 2024-01-31); a plain cast_to reinterprets the number as the image.
static bool fold_literal(Infer& in, HExpr& c, int to, bool via_text) {
    if (c.lit_null || to == BK_INVALID_TYPE || to == BK_NULL_TYPE) return true;
    if (c.lit_prim == BK_TIME && dt_is_family(to) && to != BK_TIME)
        return in.fail(BKGPU_EUNSUPPORTED, "a TIME value as DATE/DATETIME/TIMESTAMP depends on the current date: outside the GPU path");
    if (dt_is_family(to) && c.lit_prim != BK_STRING && !dt_is_family(c.lit_prim) && via_text) {
        if (is_double_t(c.lit_prim)) return in.fail(BKGPU_EUNSUPPORTED, "a DOUBLE literal compared as a date/time goes through its text form: outside the GPU path");
        char buf[32];
        if (is_uint_t(c.lit_prim)) snprintf(buf, sizeof buf, "%llu", (unsigned long long)c.lit_bits);
        else snprintf(buf, sizeof buf, "%lld", (long long)c.lit_bits);   // std::to_string, expr_value.h:709-726 (BOOL prints 0 / 1)
        c.lit_str = buf; c.lit_prim = BK_STRING;
    }
    if (c.lit_prim == BK_STRING) {
        if (to == BK_STRING) return true;
        if (!dt_is_family(to)) return in.fail(BKGPU_EUNSUPPORTED, "a STRING literal outside a date/time comparison is outside the GPU path");
        c.lit_bits = parse_literal(c.lit_str.c_str(), c.lit_str.size(), to);
    } else c.lit_bits = host_cast_prim(c.lit_bits, c.lit_prim, to);
    c.lit_prim = to;
    return true;
}

static bool infer_expr(Infer& in, HExpr& e) {
    for (auto& c : e.ch) if (!infer_expr(in, c)) return false;
    e.is_constant = expr_constant(e);
    auto all_int = [&](const std::vector<int>& t) { for (int x : t) if (!is_int_t(x)) return false; return true; };
    auto has = [&](const std::vector<int>& t, bool (*f)(int)) { for (int x : t) if (f(x)) return true; return false; };
    auto has_eq = [&](const std::vector<int>& t, int v) { for (int x : t) if (x == v) return true; return false; };
    switch (e.node_type) {
        case BK_SLOT_REF: {
            int st = in.slot_type(e.tuple_id, e.slot_id);
            if (e.col_type == BK_INVALID_TYPE) e.col_type = st;
            if (e.col_type == BK_INVALID_TYPE) return in.fail(BKGPU_EINVAL, "slot %d_%d has no tuple descriptor", e.tuple_id, e.slot_id);
            if (!is_numeric_path_type(e.col_type) && e.col_type != BK_STRING)
                return in.fail(BKGPU_EUNSUPPORTED, "column %d_%d has type %d, outside the GPU path", e.tuple_id, e.slot_id, e.col_type);
            return true;
        }
        case BK_NULL_LITERAL: if (!e.col_type) e.col_type = BK_NULL_TYPE; return true;
        case BK_BOOL_LITERAL: if (!e.col_type) e.col_type = BK_BOOL; return true;
        case BK_INT_LITERAL: if (!e.col_type) e.col_type = BK_INT64; return true;
        case BK_DOUBLE_LITERAL: if (!e.col_type) e.col_type = BK_DOUBLE; return true;
        case BK_STRING_LITERAL: if (!e.col_type) e.col_type = BK_STRING; return true;
        case BK_DATETIME_LITERAL: if (!e.col_type) e.col_type = BK_DATETIME; return true;
        case BK_TIMESTAMP_LITERAL: if (!e.col_type) e.col_type = BK_TIMESTAMP; return true;
        case BK_DATE_LITERAL: if (!e.col_type) e.col_type = BK_DATE; return true;
        case BK_TIME_LITERAL: if (!e.col_type) e.col_type = BK_TIME; return true;
        case BK_AGG_EXPR: {  // agg_fn_call.cpp:87-122
            int ct = e.ch.empty() ? BK_INVALID_TYPE : e.ch[0].col_type;
            if (e.name == "count_star" || e.name == "count") e.col_type = BK_INT64;
            else if (e.name == "avg") e.col_type = BK_DOUBLE;
            else if (e.name == "sum") {
                if (e.ch.empty()) return in.fail(BKGPU_EINVAL, "sum() without an argument");
                e.col_type = is_double_t(ct) ? BK_DOUBLE : (is_uint_t(ct) ? BK_UINT64 : BK_INT64);
            } else if (e.name == "min" || e.name == "max") {
                if (e.ch.empty()) return in.fail(BKGPU_EINVAL, "%s() without an argument", e.name.c_str());
                e.col_type = ct;
            } else return in.fail(BKGPU_EUNSUPPORTED, "aggregate '%s' is outside the GPU path", e.name.c_str());
            if (e.name != "count_star" && e.name != "count" && !e.ch.empty() && ct == BK_STRING)
                return in.fail(BKGPU_EUNSUPPORTED, "%s over a STRING argument is outside the GPU path", e.name.c_str());
            return true;
        }
        case BK_AND_PREDICATE: case BK_OR_PREDICATE: case BK_XOR_PREDICATE: case BK_NOT_PREDICATE:
        case BK_IS_NULL_PREDICATE: case BK_IS_TRUE_PREDICATE:
            if (!e.col_type) e.col_type = BK_BOOL;
            return true;
        case BK_IN_PREDICATE: {  // InPredicate::singel_open, predicate.cpp:102-148
            if (e.ch.size() < 2) return in.fail(BKGPU_EINVAL, "IN without a value list");
            if (e.ch[0].node_type == BK_SLOT_REF)
                for (size_t i = 1; i < e.ch.size(); i++) if (e.ch[i].is_constant) e.ch[i].col_type = e.ch[0].col_type;
            std::vector<int> types = {e.ch[0].col_type, e.ch[1].col_type};
            int map_type;
            if (all_int(types)) map_type = BK_INT64;
            else if (has_eq(types, BK_DATETIME)) map_type = BK_DATETIME;     // InPredicate::singel_open, predicate.cpp:102-119
            else if (has_eq(types, BK_TIMESTAMP)) map_type = BK_TIMESTAMP;
            else if (has_eq(types, BK_DATE)) map_type = BK_DATE;
            else if (has_eq(types, BK_TIME)) map_type = BK_TIME;
            else if (has(types, is_double_t) || has(types, is_int_t)) map_type = BK_DOUBLE;
            else return in.fail(BKGPU_EUNSUPPORTED, "IN over STRING is outside the GPU path");
            if (e.ch[0].col_type == BK_STRING) return in.fail(BKGPU_EUNSUPPORTED, "IN over a STRING operand is outside the GPU path");
            if (e.ch[0].col_type == BK_TIME && map_type != BK_TIME && dt_is_family(map_type))
                return in.fail(BKGPU_EUNSUPPORTED, "a TIME value as DATE/DATETIME/TIMESTAMP depends on the current date: outside the GPU path");
            for (size_t i = 1; i < e.ch.size(); i++) {
                HExpr& c = e.ch[i];
                if (!is_literal_node(c.node_type)) return in.fail(BKGPU_EUNSUPPORTED, "IN list entries must be literals");
                // Literal::get_value casts to the literal's col_type, then the set takes value.cast_to(_map_type) (predicate.cpp:120-136)
                if (!fold_literal(in, c, c.col_type, false) || !fold_literal(in, c, map_type, false)) return false;
                if (!c.lit_null) c.col_type = map_type;
            }
            e.arg_types.assign(1, map_type);
            if (!e.col_type) e.col_type = BK_BOOL;
            return true;
        }
        default: break;
    }
    // FUNCTION_CALL
    if (!e.arg_types.empty() && e.return_type != BK_INVALID_TYPE) {  // completed by the db already
        if (!e.col_type) e.col_type = e.return_type;
    } else {
        switch (e.fn_op) {  // predicates take the column's type: scalar_fn_call.cpp:57-67
            case BK_FT_EQ: case BK_FT_NE: case BK_FT_GE: case BK_FT_GT: case BK_FT_LE: case BK_FT_LT:
                if (e.ch.size() == 2 && e.ch[0].node_type == BK_SLOT_REF && e.ch[1].is_constant) e.ch[1].col_type = e.ch[0].col_type;
                break;
            default: break;
        }
        std::vector<int> types;
        for (auto& c : e.ch) {
            if (c.col_type == BK_INVALID_TYPE) return in.fail(BKGPU_EINVAL, "child of fn_op %d has INVALID_TYPE", e.fn_op);
            types.push_back(c.col_type);
        }
        switch (e.fn_op) {  // fn_manager.cpp:316-409
            case BK_FT_EQ: case BK_FT_NE: case BK_FT_GE: case BK_FT_GT: case BK_FT_LE: case BK_FT_LT:
                if (types.size() != 2) return in.fail(BKGPU_EINVAL, "comparison needs two operands");
                if (all_int(types)) complete(e, 2, has(types, is_uint_t) ? BK_UINT64 : BK_INT64, BK_BOOL);
                else if (has_eq(types, BK_DATETIME)) complete(e, 2, BK_DATETIME, BK_BOOL);
                else if (has_eq(types, BK_TIMESTAMP)) complete(e, 2, BK_TIMESTAMP, BK_BOOL);
                else if (has_eq(types, BK_DATE)) complete(e, 2, BK_DATE, BK_BOOL);
                else if (has_eq(types, BK_TIME)) complete(e, 2, BK_TIME, BK_BOOL);
                else if (has(types, is_double_t)) complete(e, 2, BK_DOUBLE, BK_BOOL);
                else if (has(types, is_int_t)) complete(e, 2, BK_DOUBLE, BK_BOOL);
                else complete(e, 2, BK_STRING, BK_BOOL);
                break;
            case BK_FT_ADD: case BK_FT_MINUS: case BK_FT_MULTIPLIES:
                if (types.size() != 2) return in.fail(BKGPU_EINVAL, "arithmetic needs two operands");
                if (has(types, is_double_t)) complete(e, 2, BK_DOUBLE, BK_DOUBLE);
                else if (has(types, is_uint_t)) complete(e, 2, BK_UINT64, BK_UINT64);
                else complete(e, 2, BK_INT64, BK_INT64);
                break;
            case BK_FT_DIVIDES: complete(e, 2, BK_DOUBLE, BK_DOUBLE); break;
            case BK_FT_MOD:
                if (has(types, is_uint_t)) complete(e, 2, BK_UINT64, BK_UINT64); else complete(e, 2, BK_INT64, BK_INT64);
                break;
            case BK_FT_BIT_AND: case BK_FT_BIT_OR: case BK_FT_BIT_XOR: case BK_FT_LS: case BK_FT_RS:
                complete(e, 2, BK_UINT64, BK_UINT64); break;
            case BK_FT_BIT_NOT: complete(e, 1, BK_UINT64, BK_UINT64); break;
            case BK_FT_UMINUS:
                if (has(types, is_double_t)) complete(e, 1, BK_DOUBLE, BK_DOUBLE);
                else if (has(types, is_uint_t)) complete(e, 1, BK_UINT64, BK_UINT64);
                else complete(e, 1, BK_INT64, BK_INT64);
                break;
            case BK_FT_LOGIC_NOT: complete(e, 1, BK_BOOL, BK_BOOL); break;
            case BK_FT_COMMON: {   // return_type_map + complete_common_fn (fn_manager.cpp:398-401,466-514); no argument casts
                std::vector<int> merge;
                const size_t n = types.size();
                if (e.name == "if") { if (n != 3) return in.fail(BKGPU_EINVAL, "if() needs three arguments"); merge = {types[1], types[2]}; }
                else if (e.name == "ifnull") { if (n != 2) return in.fail(BKGPU_EINVAL, "ifnull() needs two arguments"); merge = {types[0], types[1]}; }
                else if (e.name == "case_when") {
                    if (n < 2) return in.fail(BKGPU_EINVAL, "case_when needs a WHEN/THEN pair");
                    for (size_t i = 1; i < n; i++) if (i % 2 == 1 || i + 1 == n) merge.push_back(types[i]);
                }
                else if (e.name == "abs" || e.name == "round" || e.name == "cast_to_double" || math1_fn(e.name) >= 0 || math2_fn(e.name) >= 0 || e.name == "pi") e.return_type = BK_DOUBLE;
                else if (e.name == "floor" || e.name == "ceil" || e.name == "ceiling" || e.name == "cast_to_signed") e.return_type = BK_INT64;
                else if (e.name == "cast_to_unsigned") e.return_type = BK_UINT64;
                else return in.fail(BKGPU_EUNSUPPORTED, "function '%s' is outside the GPU path", e.name.c_str());
                if (!merge.empty()) {   // has_merged_type (include/common/type_utils.h:502-560)
                    bool all_null = true, all_equal = true, all_num = true, has_dbl = false, has_u64 = false, has_signed = false;
                    int first = BK_NULL_TYPE;
                    for (int t : merge) {
                        if (t == BK_NULL_TYPE) continue;
                        if (all_null) { first = t; all_null = false; }
                        if (t != first) all_equal = false;
                        if (!(is_double_t(t) || is_int_t(t) || t == BK_BOOL)) all_num = false;
                        if (is_double_t(t)) has_dbl = true;
                        if (t == BK_UINT64) has_u64 = true;
                        if (t == BK_INT8 || t == BK_INT16 || t == BK_INT32 || t == BK_INT64) has_signed = true;
                    }
                    if (all_null) return in.fail(BKGPU_EUNSUPPORTED, "%s: every branch is NULL", e.name.c_str());
                    if (all_equal) e.return_type = first;
                    else if (all_num) e.return_type = has_dbl ? BK_DOUBLE : (has_u64 ? (has_signed ? BK_DOUBLE : BK_UINT64) : BK_INT64);
                    else return in.fail(BKGPU_EUNSUPPORTED, "%s over date/time or STRING branches is outside the GPU path", e.name.c_str());
                    if (e.return_type == BK_STRING || is_datetime_family(e.return_type))
                        return in.fail(BKGPU_EUNSUPPORTED, "%s returning type %d is outside the GPU path", e.name.c_str(), e.return_type);
                }
                if (e.name == "sign" || e.name == "bit_count") e.return_type = BK_INT64;   // return_type_map, fn_manager.cpp:105-128
                if ((e.name == "abs" || e.name == "floor" || e.name == "ceil" || e.name == "ceiling" || e.name.rfind("cast_to_", 0) == 0 || math1_fn(e.name) >= 0) && n != 1)
                    return in.fail(BKGPU_EINVAL, "%s() needs one argument", e.name.c_str());
                if (e.name == "pi" && n != 0) return in.fail(BKGPU_EINVAL, "pi() takes no argument");
                if (math2_fn(e.name) >= 0 && (e.name == "greatest" || e.name == "least" ? n < 1 : n != 2))
                    return in.fail(BKGPU_EINVAL, "%s() needs %s", e.name.c_str(), e.name == "greatest" || e.name == "least" ? "an argument" : "two arguments");
                if (e.name == "round" && (n < 1 || n > 2 || (n == 2 && (!is_literal_node(e.ch[1].node_type) || e.ch[1].lit_null))))
                    return in.fail(BKGPU_EUNSUPPORTED, "round() needs a literal number of decimals");
                for (auto& c : e.ch)
                    if (c.col_type == BK_STRING || is_datetime_family(c.col_type)) return in.fail(BKGPU_EUNSUPPORTED, "%s over STRING / date-time arguments", e.name.c_str());
            } break;
            default: return in.fail(BKGPU_EUNSUPPORTED, "function fn_op=%d name='%s' is outside the GPU path", e.fn_op, e.name.c_str());
        }
        if (!e.col_type) e.col_type = e.return_type;
        // Literal::cast_to_col_type on literal children (scalar_fn_call.cpp:113-117)
        for (size_t i = 0; i < e.arg_types.size() && i < e.ch.size(); i++) {
            HExpr& c = e.ch[i];
            if (is_literal_node(c.node_type) && !c.lit_null) {
                if (!fold_literal(in, c, e.arg_types[i], true)) return false;
                c.col_type = e.arg_types[i];   // value_to_node_type: the literal now is of the argument's type (literal.h:296-310)
            }
        }
    }
    if (e.ch.size() < e.arg_types.size()) return in.fail(BKGPU_EINVAL, "function has fewer children than arg_types");
    for (int at : e.arg_types)
        if (at == BK_STRING) return in.fail(BKGPU_EUNSUPPORTED, "STRING-domain comparison is outside the GPU path");
    for (size_t i = 0; i < e.arg_types.size(); i++) {
        if (e.ch[i].col_type == BK_TIME && e.arg_types[i] != BK_TIME && dt_is_family(e.arg_types[i]))
            return in.fail(BKGPU_EUNSUPPORTED, "a TIME value as DATE/DATETIME/TIMESTAMP depends on the current date: outside the GPU path");
        if (e.ch[i].col_type == BK_STRING) return in.fail(BKGPU_EUNSUPPORTED, "STRING operands are outside the GPU path");
    }
    return true;
}

// ---------------------------------------------------------------- lowering
static bool is_filter(const HNode* n) {
    return n && (n->node_type == BK_WHERE_FILTER_NODE || n->node_type == BK_TABLE_FILTER_NODE);
}

struct Lower {
    Compiled* out;
    Infer* in;
    Program* p;
    std::map<int, int> build_cols;  // (unused unless joining)
    int intern_col(int tuple_id, int slot_id, int prim) {
        auto& cols = out->cols;
        for (size_t i = 0; i < cols.size(); i++) if (cols[i].tuple_id == tuple_id && cols[i].slot_id == slot_id) return (int)i;
        cols.push_back({tuple_id, slot_id, prim});
        out->col_side.push_back(out->build_tuple >= 0 && tuple_id == out->build_tuple ? 1 : 0);
        return (int)cols.size() - 1;
    }
    bool emit(uint8_t op, uint8_t a = 0, uint8_t b = 0, uint8_t c = 0) {
        if (p->n_instr >= MAX_INSTR) return in->fail(BKGPU_EUNSUPPORTED, "expression too long for the device program (%d instructions)", MAX_INSTR);
        p->code[p->n_instr++] = Instr{op, a, b, c};
        return true;
    }
    int add_const(uint64_t bits, bool isnull) {
        if (out->n_const >= MAX_CONST) { in->fail(BKGPU_EUNSUPPORTED, "too many constants (%d)", MAX_CONST); return -1; }
        int i = out->n_const++;
        p->cbits[i] = bits;
        if (isnull) p->cnull |= 1ull << i;
        return i;
    }
    // ExprValue::cast_to between canonical images: a no-op when the image does not change
    bool cast(int from, int to) {
        if (from == to || to == BK_INVALID_TYPE || from == BK_NULL_TYPE) return true;
        if (dt_is_family(from) && dt_is_family(to)) return emit(OP_CAST, (uint8_t)from, (uint8_t)to);   // calendar conversion (datetime.h)
        bool from_int = is_int_t(from) || from == BK_BOOL || is_datetime_family(from);
        if (from_int && (to == BK_INT64 || to == BK_UINT64 || to == BK_DATETIME)) return true;  // sign/zero-extended image reinterpreted
        if (from == BK_FLOAT && to == BK_DOUBLE) return true;                                    // floats travel widened
        return emit(OP_CAST, (uint8_t)from, (uint8_t)to);
    }
    // emits code leaving the value on the stack; returns false on error.  `stack` tracks depth.
    bool expr(const HExpr& e, int& depth) {
        if (depth + 1 > STACK_DEPTH - 1) return in->fail(BKGPU_EUNSUPPORTED, "expression nests deeper than the device stack");
        switch (e.node_type) {
            case BK_SLOT_REF: {
                int st = in->slot_type(e.tuple_id, e.slot_id);
                if (st == BK_INVALID_TYPE) st = e.col_type;
                if (st == BK_STRING) return in->fail(BKGPU_EUNSUPPORTED, "STRING column %d_%d used in an expression", e.tuple_id, e.slot_id);
                int ci = intern_col(e.tuple_id, e.slot_id, st);
                if (ci >= MAX_COLS) return in->fail(BKGPU_EUNSUPPORTED, "more than %d columns referenced", MAX_COLS);
                if (!emit(OP_LOAD_COL, (uint8_t)ci, 0, (uint8_t)out->col_side[(size_t)ci])) return false;  // c = 1: read at the matched build row
                depth++;
                return cast(st, e.col_type);
            }
            case BK_NULL_LITERAL: { int k = add_const(0, true); if (k < 0) return false; depth++; return emit(OP_CONST, (uint8_t)k); }
            case BK_BOOL_LITERAL: case BK_INT_LITERAL: case BK_DOUBLE_LITERAL: case BK_STRING_LITERAL:
            case BK_DATETIME_LITERAL: case BK_TIMESTAMP_LITERAL: case BK_DATE_LITERAL: case BK_TIME_LITERAL: {
                if (e.lit_prim == BK_STRING) return in->fail(BKGPU_EUNSUPPORTED, "a STRING literal outside a date/time comparison is outside the GPU path");
                if (e.lit_prim == BK_TIME && e.col_type != BK_TIME && dt_is_family(e.col_type))
                    return in->fail(BKGPU_EUNSUPPORTED, "a TIME value as DATE/DATETIME/TIMESTAMP depends on the current date: outside the GPU path");
                // Literal::get_value: _value.cast_to(_col_type)
                uint64_t bits = e.col_type ? host_cast_prim(e.lit_bits, e.lit_prim, e.col_type) : e.lit_bits;
                int k = add_const(bits, false); if (k < 0) return false;
                depth++;
                return emit(OP_CONST, (uint8_t)k);
            }
            case BK_AND_PREDICATE: case BK_OR_PREDICATE: {
                if (e.ch.empty() || e.ch.size() > 8) return in->fail(BKGPU_EUNSUPPORTED, "AND/OR with %zu operands", e.ch.size());
                int d0 = depth;
                for (auto& c : e.ch) { if (!expr(c, depth)) return false; if (!to_bool(c)) return false; }
                depth = d0 + 1;
                return emit(e.node_type == BK_AND_PREDICATE ? OP_AND : OP_OR, (uint8_t)e.ch.size());
            }
            case BK_XOR_PREDICATE: {
                if (e.ch.size() != 2) return in->fail(BKGPU_EINVAL, "XOR needs two operands");
                for (auto& c : e.ch) { if (!expr(c, depth)) return false; if (!to_bool(c)) return false; }
                depth--;
                return emit(OP_XOR);
            }
            case BK_NOT_PREDICATE: case BK_IS_TRUE_PREDICATE: {
                if (e.ch.size() != 1) return in->fail(BKGPU_EINVAL, "unary predicate needs one operand");
                if (!expr(e.ch[0], depth) || !to_bool(e.ch[0])) return false;
                return emit(e.node_type == BK_NOT_PREDICATE ? OP_NOT3 : OP_IS_TRUE);
            }
            case BK_IS_NULL_PREDICATE:
                if (e.ch.size() != 1) return in->fail(BKGPU_EINVAL, "IS NULL needs one operand");
                if (!expr(e.ch[0], depth)) return false;
                return emit(OP_IS_NULL);
            case BK_IN_PREDICATE: {
                int map_type = e.arg_types[0];
                if (!expr(e.ch[0], depth) || !cast(e.ch[0].col_type, map_type)) return false;
                int first = out->n_const, cnt = 0; bool has_null = false;
                for (size_t i = 1; i < e.ch.size(); i++) {
                    const HExpr& c = e.ch[i];
                    if (c.lit_null) { has_null = true; continue; }
                    uint64_t bits = c.col_type ? host_cast_prim(c.lit_bits, c.lit_prim, c.col_type) : c.lit_bits;  // Literal::get_value
                    bits = host_cast_prim(bits, c.col_type ? c.col_type : c.lit_prim, map_type);                  // value.cast_to(_map_type)
                    if (add_const(bits, false) < 0) return false;
                    cnt++;
                }
                if (cnt > 255) return in->fail(BKGPU_EUNSUPPORTED, "IN list longer than 255");
                return emit(OP_IN, (uint8_t)first, (uint8_t)cnt, (uint8_t)((has_null ? 16 : 0) | host_prim_class(map_type)));
            }
            case BK_FUNCTION_CALL: {
                if (e.fn_op == BK_FT_COMMON) return common_fn(e, depth);
                int d0 = depth;
                for (size_t i = 0; i < e.ch.size(); i++) {
                    if (!expr(e.ch[i], depth)) return false;
                    if (i < e.arg_types.size() && !cast(e.ch[i].col_type, e.arg_types[i])) return false;
                }
                int at = e.arg_types.empty() ? BK_INT64 : e.arg_types[0];
                uint8_t vc = (uint8_t)host_prim_class(at);
                bool ok;
                switch (e.fn_op) {
                    case BK_FT_EQ: case BK_FT_NE: case BK_FT_GT: case BK_FT_GE: case BK_FT_LT: case BK_FT_LE:
                        ok = emit(OP_CMP, (uint8_t)e.fn_op, vc); depth = d0 + 1; break;
                    case BK_FT_ADD: case BK_FT_MINUS: case BK_FT_MULTIPLIES: ok = emit(OP_ARITH, (uint8_t)e.fn_op, vc); depth = d0 + 1; break;
                    case BK_FT_DIVIDES: ok = emit(OP_DIV_F64); depth = d0 + 1; break;
                    case BK_FT_MOD: ok = emit(OP_MOD, 0, vc); depth = d0 + 1; break;
                    case BK_FT_BIT_AND: case BK_FT_BIT_OR: case BK_FT_BIT_XOR: case BK_FT_LS: case BK_FT_RS:
                        ok = emit(OP_BIT, (uint8_t)e.fn_op); depth = d0 + 1; break;
                    case BK_FT_BIT_NOT: ok = emit(OP_BIT_NOT); break;
                    case BK_FT_UMINUS: ok = emit(OP_NEG, 0, vc); break;   // minus_uint yields the INT64 image too (operators.cpp:31)
                    case BK_FT_LOGIC_NOT: ok = emit(OP_LOGIC_NOT); break;
                    default: return in->fail(BKGPU_EUNSUPPORTED, "fn_op %d", e.fn_op);
                }
                if (!ok) return false;
                int rt = e.return_type;
                if (e.fn_op == BK_FT_UMINUS && at == BK_UINT64) rt = BK_INT64;
                return cast(rt, e.col_type);  // _fn_call(args).cast_to(_col_type)
            }
            default: return in->fail(BKGPU_EUNSUPPORTED, "expr node type %d cannot be lowered", e.node_type);
        }
    }
    // named builtins (FT_COMMON): arguments are NOT cast (arg_types stay empty), the result is cast to the node's col_type
    bool branch(const HExpr& c, int to, int& depth) { return expr(c, depth) && cast(c.col_type, to); }
    bool common_fn(const HExpr& e, int& depth) {
        const int d0 = depth;
        const int ct = e.col_type;
        if (e.name == "if") {
            if (!expr(e.ch[0], depth) || !to_bool(e.ch[0]) || !branch(e.ch[1], ct, depth) || !branch(e.ch[2], ct, depth)) return false;
            depth = d0 + 1;
            return emit(OP_SELECT);
        }
        if (e.name == "ifnull") {
            if (!branch(e.ch[0], ct, depth) || !branch(e.ch[1], ct, depth)) return false;
            depth = d0 + 1;
            return emit(OP_IFNULL);
        }
        if (e.name == "case_when") {   // WHEN c1 THEN t1 ... [ELSE e]  ==  if(c1, t1, if(c2, t2, ... e | NULL))
            const size_t n = e.ch.size(), pairs = n / 2;
            for (size_t i = 0; i < pairs; i++)
                if (!expr(e.ch[2 * i], depth) || !to_bool(e.ch[2 * i]) || !branch(e.ch[2 * i + 1], ct, depth)) return false;
            if (n % 2 == 1) { if (!branch(e.ch[n - 1], ct, depth)) return false; }
            else { int k = add_const(0, true); if (k < 0 || !emit(OP_CONST, (uint8_t)k)) return false; depth++; }
            if (depth > STACK_DEPTH - 1) return in->fail(BKGPU_EUNSUPPORTED, "CASE with %zu branches exceeds the device stack", pairs);
            for (size_t i = 0; i < pairs; i++) if (!emit(OP_SELECT)) return false;
            depth = d0 + 1;
            return true;
        }
        if (e.name == "pi") {
            const double pi = 3.14159265358979323846; uint64_t b; memcpy(&b, &pi, 8);
            int k = add_const(b, false); if (k < 0) return false;
            depth++;
            return emit(OP_CONST, (uint8_t)k) && cast(BK_DOUBLE, ct);
        }
        if (const int m2 = math2_fn(e.name); m2 >= 0) {   // arguments read with get_numberic<double>(); greatest / least fold left to right
            if (!expr(e.ch[0], depth) || !cast(e.ch[0].col_type, BK_DOUBLE)) return false;
            for (size_t i = 1; i < e.ch.size(); i++) {
                if (!expr(e.ch[i], depth) || !cast(e.ch[i].col_type, BK_DOUBLE) || !emit(OP_MATH2, (uint8_t)m2)) return false;
                depth--;
            }
            depth = d0 + 1;
            return cast(BK_DOUBLE, ct);
        }
        if (!expr(e.ch[0], depth)) return false;
        const int at = e.ch[0].col_type;
        if (e.name == "bit_count") return cast(at, BK_UINT64) && emit(OP_MATH, MF_BIT_COUNT) && cast(BK_INT64, ct);
        if (e.name == "cast_to_signed") return cast(at, BK_INT64) && cast(BK_INT64, ct);
        if (e.name == "cast_to_unsigned") return cast(at, BK_UINT64) && cast(BK_UINT64, ct);
        if (e.name == "cast_to_double") return cast(at, BK_DOUBLE) && cast(BK_DOUBLE, ct);
        if (!cast(at, BK_DOUBLE)) return false;   // get_numberic<double>()
        if (e.name == "abs") return emit(OP_MATH, MF_ABS) && cast(BK_DOUBLE, ct);
        if (e.name == "floor") return emit(OP_MATH, MF_FLOOR) && emit(OP_CAST, (uint8_t)BK_DOUBLE, (uint8_t)BK_INT64) && cast(BK_INT64, ct);
        if (e.name == "ceil" || e.name == "ceiling") return emit(OP_MATH, MF_CEIL) && emit(OP_CAST, (uint8_t)BK_DOUBLE, (uint8_t)BK_INT64) && cast(BK_INT64, ct);
        if (e.name == "sign") return emit(OP_MATH, MF_SIGN) && cast(BK_INT64, ct);
        if (const int m1 = math1_fn(e.name); m1 >= 0) return emit(OP_MATH, (uint8_t)m1) && cast(BK_DOUBLE, ct);
        if (e.name == "round") {
            int bits = 0;
            if (e.ch.size() == 2) {   // input[1].get_numberic<int>() of the literal (cast to its col_type first, literal.h:204-206)
                const HExpr& l = e.ch[1];
                uint64_t img = l.col_type ? host_cast_prim(l.lit_bits, l.lit_prim, l.col_type) : l.lit_bits;
                bits = (int)(int64_t)host_cast_prim(img, l.col_type ? l.col_type : l.lit_prim, BK_INT32);
            }
            const double base = std::pow(10.0, bits);
            uint64_t bb; memcpy(&bb, &base, 8);
            int k = add_const(bb, false); if (k < 0) return false;
            return emit(OP_MATH, MF_ROUND, (uint8_t)k) && cast(BK_DOUBLE, ct);
        }
        return in->fail(BKGPU_EUNSUPPORTED, "function '%s'", e.name.c_str());
    }
    // children of logical predicates are read with get_numberic<bool>() (predicate.h:31)
    bool to_bool(const HExpr& c) { return c.col_type == BK_BOOL ? true : emit(OP_CAST, (uint8_t)c.col_type, (uint8_t)BK_BOOL); }
    bool out_reg(int r) { return emit(OP_OUT, (uint8_t)r); }
};

static std::string expr_key(const HExpr& e) {  // structural identity, for lane sharing
    char buf[96];
    snprintf(buf, sizeof buf, "(%d:%d:%d_%d:%d:%llx", e.node_type, e.col_type, e.tuple_id, e.slot_id, e.fn_op, (unsigned long long)e.lit_bits);
    std::string s = buf;
    for (auto& c : e.ch) s += expr_key(c);
    return s + ")";
}
static void expr_cols(const HExpr& e, const Compiled& c, int& mask) {
    if (e.node_type == BK_SLOT_REF)
        for (size_t i = 0; i < c.cols.size(); i++) if (c.cols[i].tuple_id == e.tuple_id && c.cols[i].slot_id == e.slot_id) mask |= 1 << i;
    for (auto& ch : e.ch) expr_cols(ch, c, mask);
}
// can the expression yield NULL although every column it reads is valid?
static bool expr_makes_null(const HExpr& e) {
    if (e.node_type == BK_NULL_LITERAL) return true;
    if (e.node_type == BK_IS_NULL_PREDICATE || e.node_type == BK_IS_TRUE_PREDICATE) return false;
    if (e.node_type == BK_FUNCTION_CALL && (e.fn_op == BK_FT_DIVIDES || e.fn_op == BK_FT_MOD)) return true;
    if (e.node_type == BK_FUNCTION_CALL && e.fn_op == BK_FT_COMMON && e.name == "case_when" && e.ch.size() % 2 == 0) return true;
    if (e.node_type == BK_IN_PREDICATE) for (size_t i = 1; i < e.ch.size(); i++) if (e.ch[i].lit_null) return true;
    for (auto& c : e.ch) if (expr_makes_null(c)) return true;
    return false;
}

static int key_value_bits(int prim) {
    switch (prim) {
        case BK_INT64: case BK_UINT64: case BK_DOUBLE: case BK_DATETIME: case BK_FLOAT: return 64;  // FLOAT keys travel as their double image
        default: return 32;
    }
}

static const HNode* skip_passthrough(const HNode* n, bool* under_packet) {
    while (n && (n->node_type == BK_PACKET_NODE || n->node_type == BK_SELECT_MANAGER_NODE)) {
        if (n->node_type == BK_PACKET_NODE && under_packet) *under_packet = true;
        n = n->ch.empty() ? nullptr : &n->ch[0];
    }
    return n;
}

static bool lower_agg(Infer& in, Compiled& out, const HNode& agg, const std::vector<const HExpr*>& conjuncts, int scan_tuple, bool under_packet, bool allow_direct) {
    out.kind = PK_AGG;
    out.scan_tuple = scan_tuple;
    out.is_merge = agg.node_type == BK_MERGE_AGG_NODE;
    out.emit_default = agg.group_exprs.empty() && (under_packet || out.is_merge);
    out.agg_limit = agg.limit;
    if (agg.group_exprs.size() > MAX_GROUP) return in.fail(BKGPU_EUNSUPPORTED, "more than %d GROUP BY expressions", MAX_GROUP);
    if (agg.agg_fns.size() > MAX_AGG) return in.fail(BKGPU_EUNSUPPORTED, "more than %d aggregate functions", MAX_AGG);
    Program& p = out.prog; memset(&p, 0, sizeof p);
    AggPlan& ap = out.ap; memset(&ap, 0, sizeof ap);
    Lower lw{&out, &in, &p, {}};
    int reg = 0, depth = 0;
    // ---- predicate: all conjuncts non-NULL true (filter_node.cpp:726-734) ----
    ap.pred_out = -1;
    if (out.is_merge && !conjuncts.empty()) return in.fail(BKGPU_EUNSUPPORTED, "filter below MERGE_AGG_NODE");
    bool any_distinct = false;
    for (auto& f : agg.agg_fns) any_distinct = any_distinct || f.distinct;   // (a distinct aggregate is never "initial": is_initialize, agg_fn_call.cpp:322-328)
    if (out.is_merge && agg.group_exprs.empty() && !agg.agg_fns.empty() && !any_distinct) {
        // a scalar merger skips input rows whose aggregates are all still "initial" (AggFnCall::all_is_initialize,
        // agg_node.cpp:519-522): pass = OR_k (count_k <> 0 | value_k IS NOT NULL), three-valued, NULL = skip
        for (size_t k = 0; k < agg.agg_fns.size(); k++) {
            const HExpr& f = agg.agg_fns[k];
            const int st = in.slot_type(f.tuple_id, f.inter_slot);
            if (st == BK_INVALID_TYPE) return in.fail(BKGPU_EINVAL, "MERGE_AGG: slot %d_%d is not declared", f.tuple_id, f.inter_slot);
            const int ci = lw.intern_col(f.tuple_id, f.inter_slot, st);
            if (ci >= MAX_COLS) return in.fail(BKGPU_EUNSUPPORTED, "more than %d columns referenced", MAX_COLS);
            if (!lw.emit(OP_LOAD_COL, (uint8_t)ci, st == BK_STRING ? 1 : 0, 0)) return false;
            if (f.name == "count" || f.name == "count_star") {
                const int z = lw.add_const(0, false); if (z < 0) return false;
                if (!lw.emit(OP_CONST, (uint8_t)z) || !lw.emit(OP_CMP, (uint8_t)BK_FT_NE, (uint8_t)host_prim_class(st))) return false;
            } else if (!lw.emit(OP_IS_NULL) || !lw.emit(OP_NOT3)) return false;
            if (k > 0 && !lw.emit(OP_OR, 2)) return false;
        }
        ap.pred_out = reg;
        if (!lw.out_reg(reg++)) return false;
    }
    if (!conjuncts.empty()) {
        if (conjuncts.size() > 8) return in.fail(BKGPU_EUNSUPPORTED, "more than 8 conjuncts");
        for (auto* c : conjuncts) { depth = 0; if (!lw.expr(*c, depth) || !lw.to_bool(*c)) return false; }
        if (conjuncts.size() > 1 && !lw.emit(OP_AND, (uint8_t)conjuncts.size())) return false;
        ap.pred_out = reg;
        if (!lw.out_reg(reg++)) return false;
    }
    // ---- group keys ----
    ap.n_group = (int)agg.group_exprs.size();
    int n_words = 0, half_word = -1;  // half_word: a word holding one 32-bit field (upper half free)
    for (int g = 0; g < ap.n_group; g++) {
        const HExpr& e = agg.group_exprs[(size_t)g];
        if (e.col_type == BK_STRING) return in.fail(BKGPU_EUNSUPPORTED, "GROUP BY over a STRING key is outside the GPU path");
        depth = 0;
        if (!lw.expr(e, depth)) return false;
        ap.key_out[g] = (uint8_t)reg; ap.key_prim[g] = (uint8_t)e.col_type;
        if (!lw.out_reg(reg++)) return false;
        int bits = key_value_bits(e.col_type);
        ap.key_bits[g] = (uint8_t)bits;
        if (bits == 64) { ap.key_word[g] = (uint8_t)n_words++; ap.key_shift[g] = 0; }
        else if (half_word >= 0) { ap.key_word[g] = (uint8_t)half_word; ap.key_shift[g] = 32; half_word = -1; }
        else { ap.key_word[g] = (uint8_t)n_words; ap.key_shift[g] = 0; half_word = n_words++; }
    }
    if (ap.n_group > 0) {  // null flags: one bit per group expression (encode_exprs_key's null-flag byte)
        int w, sh;
        if (half_word >= 0) { w = half_word; sh = 32; } else { w = n_words++; sh = 0; }
        for (int g = 0; g < ap.n_group; g++) { ap.key_null_word[g] = (uint8_t)w; ap.key_null_shift[g] = (uint8_t)(sh + g); }
    }
    if (n_words > MAX_KEYW) return in.fail(BKGPU_EUNSUPPORTED, "GROUP BY key wider than %d words", MAX_KEYW);
    ap.n_keyw = n_words;
    // ---- aggregates and lanes ----
    ap.n_agg = (int)agg.agg_fns.size();
    const int n_visible = ap.n_agg;
    int n_hidden = 0;
    ap.n_lanes = 1; ap.lane_op[0] = LN_ADD_I64;
    std::map<std::string, int> cnt_lane_of, acc_lane_of, arg_reg_of;
    out.arg_cols_mask.assign((size_t)MAX_AGG, 0);
    out.arg_can_null.assign((size_t)MAX_AGG, false);
    auto new_lane = [&](uint8_t op) -> int {
        if (ap.n_lanes >= MAX_LANES) { in.fail(BKGPU_EUNSUPPORTED, "more than %d accumulator lanes", MAX_LANES); return -1; }
        ap.lane_op[ap.n_lanes] = op; return ap.n_lanes++;
    };
    for (int k = 0; k < ap.n_agg; k++) {
        const HExpr& f = agg.agg_fns[(size_t)k];
        AggSpec& a = ap.agg[k]; memset(&a, 0, sizeof a);
        a.arg_out = 0xFF; a.out_prim = (uint8_t)f.col_type;
        if (out.is_merge && !f.distinct) {
            // MERGE_AGG_NODE: the input rows carry the stores' intermediate slots; AggFnCall::merge (agg_fn_call.cpp:719-822)
            // adds counts and sums, folds MIN/MAX, and adds both halves of an AVG blob.  NULL intermediates are skipped.
            const int st = in.slot_type(f.tuple_id, f.inter_slot);
            if (st == BK_INVALID_TYPE) return in.fail(BKGPU_EINVAL, "MERGE_AGG: slot %d_%d is not declared", f.tuple_id, f.inter_slot);
            const bool blob = f.name == "avg";
            if (blob != (st == BK_STRING)) return in.fail(BKGPU_EUNSUPPORTED, "MERGE_AGG: %s over an intermediate slot of type %d", f.name.c_str(), st);
            const int ci = lw.intern_col(f.tuple_id, f.inter_slot, st);
            if (ci >= MAX_COLS) return in.fail(BKGPU_EUNSUPPORTED, "more than %d columns referenced", MAX_COLS);
            out.arg_cols_mask[(size_t)k] = 1 << ci;
            if (!lw.emit(OP_LOAD_COL, (uint8_t)ci, blob ? 1 : 0, 0)) return false;
            a.arg_out = (uint8_t)reg; a.nullable = 1;
            if (!lw.out_reg(reg++)) return false;
            const uint8_t cls = (uint8_t)(blob ? VC_F64 : host_prim_class(st));
            a.vclass = a.arg_vclass = cls;
            uint8_t op = cls == VC_F64 ? LN_ADD_F64 : LN_ADD_I64;
            if (blob) {
                a.kind = AG_AVG;
                if (n_visible + n_hidden >= MAX_AGG) return in.fail(BKGPU_EUNSUPPORTED, "more than %d aggregate accumulators", MAX_AGG);
                const int hk = n_visible + n_hidden++;
                AggSpec& h = ap.agg[hk]; memset(&h, 0, sizeof h);
                h.kind = AG_COUNT_MERGE; h.hidden = 1; h.vclass = h.arg_vclass = VC_I64; h.nullable = 1; h.acc_owner = 1;
                out.arg_cols_mask[(size_t)hk] = 1 << ci;
                if (!lw.emit(OP_LOAD_COL, (uint8_t)ci, 2, 0)) return false;
                h.arg_out = (uint8_t)reg;
                if (!lw.out_reg(reg++)) return false;
                int hl = new_lane(LN_ADD_I64); if (hl < 0) return false;
                h.acc_lane = (uint8_t)hl;
                a.cnt_lane = (uint8_t)hl; a.cnt_owner = 0;
            } else if (f.name == "count" || f.name == "count_star") {
                a.kind = AG_COUNT_MERGE; a.vclass = VC_I64; op = LN_ADD_I64;
            } else {
                if (f.name == "sum") a.kind = AG_SUM;
                else {
                    const bool mn = f.name == "min";
                    a.kind = mn ? AG_MIN : AG_MAX;
                    op = cls == VC_F64 ? (mn ? LN_MIN_F64 : LN_MAX_F64) : (cls == VC_U64 ? (mn ? LN_MIN_U64 : LN_MAX_U64) : (mn ? LN_MIN_I64 : LN_MAX_I64));
                }
                int cl = new_lane(LN_ADD_I64); if (cl < 0) return false;
                a.cnt_lane = (uint8_t)cl; a.cnt_owner = 1;
            }
            int al = new_lane(op); if (al < 0) return false;
            a.acc_lane = (uint8_t)al; a.acc_owner = 1;
            continue;
        }
        if (f.name == "count_star") { a.kind = AG_COUNT_STAR; continue; }
        if (f.ch.empty()) return in.fail(BKGPU_EINVAL, "%s() without an argument", f.name.c_str());
        if (f.ch.size() > 1) return in.fail(BKGPU_EUNSUPPORTED, "%s() with %zu arguments", f.name.c_str(), f.ch.size());
        const HExpr& arg = f.ch[0];
        // COUNT(non-null literal) is COUNT(*) (agg_fn_call.cpp:246-250)
        if (f.name == "count" && is_literal_node(arg.node_type) && !arg.lit_null) { a.kind = AG_COUNT_STAR; continue; }
        std::string key = expr_key(arg);
        if (!arg_reg_of.count(key)) {
            depth = 0;
            if (!lw.expr(arg, depth)) return false;
            arg_reg_of[key] = reg;
            if (!lw.out_reg(reg++)) return false;
        }
        a.arg_out = (uint8_t)arg_reg_of[key];
        a.arg_vclass = (uint8_t)host_prim_class(arg.col_type);
        a.nullable = 1;
        int m = 0; expr_cols(arg, out, m); out.arg_cols_mask[(size_t)k] = m; out.arg_can_null[(size_t)k] = expr_makes_null(arg);
        if (!cnt_lane_of.count(key)) { int l = new_lane(LN_ADD_I64); if (l < 0) return false; cnt_lane_of[key] = l; a.cnt_owner = 1; }
        a.cnt_lane = (uint8_t)cnt_lane_of[key];
        uint8_t op;
        if (f.name == "count") { a.kind = AG_COUNT; continue; }
        else if (f.name == "sum") { a.kind = AG_SUM; a.vclass = (uint8_t)host_prim_class(f.col_type); op = a.vclass == VC_F64 ? LN_ADD_F64 : LN_ADD_I64; }
        else if (f.name == "avg") { a.kind = AG_AVG; a.vclass = VC_F64; op = LN_ADD_F64; }
        else {
            bool mn = f.name == "min";
            a.kind = mn ? AG_MIN : AG_MAX; a.vclass = a.arg_vclass;
            op = a.vclass == VC_F64 ? (mn ? LN_MIN_F64 : LN_MAX_F64) : (a.vclass == VC_U64 ? (mn ? LN_MIN_U64 : LN_MAX_U64) : (mn ? LN_MIN_I64 : LN_MAX_I64));
        }
        std::string akey = key + "#" + std::to_string((int)op) + "#" + std::to_string((int)a.vclass);
        if (!acc_lane_of.count(akey)) { int l = new_lane(op); if (l < 0) return false; acc_lane_of[akey] = l; a.acc_owner = 1; }
        a.acc_lane = (uint8_t)acc_lane_of[akey];
    }
    ap.n_agg = n_visible + n_hidden;
    out.arg_cols_mask.resize((size_t)ap.n_agg); out.arg_can_null.resize((size_t)ap.n_agg);
    p.n_out = reg;
    if (reg > MAX_GROUP + MAX_AGG + 1) return in.fail(BKGPU_EUNSUPPORTED, "too many program outputs");
    // ---- output schema ----
    for (int g = 0; g < ap.n_group; g++) {
        const HExpr& e = agg.group_exprs[(size_t)g];
        bool sr = e.node_type == BK_SLOT_REF;
        out.out_cols.push_back({sr ? e.tuple_id : -1, sr ? e.slot_id : g, e.col_type, 0});
    }
    for (int k = 0; k < n_visible; k++) {
        const HExpr& f = agg.agg_fns[(size_t)k];
        int ft = in.slot_type(f.tuple_id, f.final_slot);
        if (ft == BK_INVALID_TYPE || ft == BK_STRING) ft = f.col_type;
        out.out_cols.push_back({f.tuple_id, f.final_slot, ft, 0});
        ap.agg[k].out_prim = (uint8_t)ft;
        if (ap.agg[k].kind == AG_AVG) out.out_cols.push_back({f.tuple_id, f.inter_slot != f.final_slot ? f.inter_slot : -1, BK_STRING, 1});
    }
    // ---- "direct" shape detection ----
    out.has_direct = false;
    DirectPlan& d = out.direct; memset(&d, 0, sizeof d); memset(d.agg_val, 0xFF, sizeof d.agg_val);
    do {
        std::vector<int> order;  // cols indices in [terms][key][values] order
        if (!allow_direct || out.is_merge) break;
        if (!conjuncts.empty()) {
            if (conjuncts.size() > 2) break;
            bool ok = true;
            for (auto* cp : conjuncts) {
                const HExpr& c = *cp;
                if (c.node_type != BK_FUNCTION_CALL || c.fn_op < BK_FT_EQ || c.fn_op > BK_FT_LE || c.ch.size() != 2 || c.col_type != BK_BOOL) { ok = false; break; }
                const HExpr *col = &c.ch[0], *lit = &c.ch[1]; int op = c.fn_op;
                if (is_literal_node(col->node_type) && lit->node_type == BK_SLOT_REF) {
                    std::swap(col, lit);
                    op = op == BK_FT_GT ? BK_FT_LT : op == BK_FT_LT ? BK_FT_GT : op == BK_FT_GE ? BK_FT_LE : op == BK_FT_LE ? BK_FT_GE : op;
                }
                if (col->node_type != BK_SLOT_REF || !is_literal_node(lit->node_type) || lit->lit_null) { ok = false; break; }
                int at = c.arg_types[0];
                int st = in.slot_type(col->tuple_id, col->slot_id); if (st == BK_INVALID_TYPE) st = col->col_type;
                if (st != col->col_type) { ok = false; break; }
                int cc = host_prim_class(st), ac = host_prim_class(at);
                if ((cc == VC_F64) != (ac == VC_F64)) { ok = false; break; }  // would need an int<->double conversion per row
                if (is_datetime_family(at) && st != at) { ok = false; break; }   // (a change of type inside the family is a calendar conversion: generic path)
                if (lit->lit_prim == BK_STRING) { ok = false; break; }
                // Literal::get_value casts to its col_type, ScalarFnCall casts that to the arg type
                uint64_t bits = lit->col_type ? host_cast_prim(lit->lit_bits, lit->lit_prim, lit->col_type) : lit->lit_bits;
                bits = host_cast_prim(bits, lit->col_type ? lit->col_type : lit->lit_prim, at);
                DirectTerm& t = d.term[d.n_terms++];
                t.cmp = (uint8_t)op; t.vclass = (uint8_t)ac; t.cbits = bits;
                order.push_back(lw.intern_col(col->tuple_id, col->slot_id, st));
            }
            if (!ok) break;
        }
        if (ap.n_group > 1) break;
        if (ap.n_group == 1) {
            const HExpr& e = agg.group_exprs[0];
            if (e.node_type != BK_SLOT_REF) break;
            int st = in.slot_type(e.tuple_id, e.slot_id); if (st == BK_INVALID_TYPE) st = e.col_type;
            if (st != e.col_type || ap.n_keyw > 2) break;
            d.n_keys = 1;
            order.push_back(lw.intern_col(e.tuple_id, e.slot_id, st));
        }
        bool ok = true;
        std::vector<int> vals;
        for (int k = 0; k < ap.n_agg && ok; k++) {
            const HExpr& f = agg.agg_fns[(size_t)k];
            if (ap.agg[k].kind == AG_COUNT_STAR) continue;
            const HExpr& arg = f.ch[0];
            int st = arg.node_type == BK_SLOT_REF ? in.slot_type(arg.tuple_id, arg.slot_id) : BK_INVALID_TYPE;
            if (arg.node_type != BK_SLOT_REF || (st != BK_INVALID_TYPE && st != arg.col_type)) { ok = false; break; }
            int ci = lw.intern_col(arg.tuple_id, arg.slot_id, arg.col_type);
            size_t pos = 0;
            for (; pos < vals.size(); pos++) if (vals[pos] == ci) break;
            if (pos == vals.size()) vals.push_back(ci);
            d.agg_val[k] = (uint8_t)pos;
        }
        if (!ok || vals.size() > 4) break;
        {   // at most 3 lane operations per value column (ValOps)
            int ops[4] = {0, 0, 0, 0}; bool too_many = false;
            for (int k = 0; k < ap.n_agg; k++)
                if (ap.agg[k].kind != AG_COUNT_STAR && ap.agg[k].kind != AG_COUNT && ap.agg[k].acc_owner && ++ops[d.agg_val[k]] > 3) too_many = true;
            if (too_many) break;
        }
        d.n_vals = (int)vals.size();
        for (int v : vals) order.push_back(v);
        out.direct_cols = order;
        out.has_direct = true;
    } while (0);
    return true;
}

// ---------------------------------------------------------------- explain
static const char* op_name(int op) {
    static const char* n[] = {"END", "LOAD_COL", "CONST", "CAST", "CMP", "ARITH", "DIV_F64", "MOD", "BIT", "BIT_NOT", "NEG",
                              "LOGIC_NOT", "AND", "OR", "XOR", "NOT3", "IS_NULL", "IS_TRUE", "IN", "OUT", "SELECT", "IFNULL", "MATH", "MATH2"};
    return op >= 0 && op <= OP_MATH2 ? n[op] : "?";
}
static void explain(Compiled& c) {
    char buf[256];
    std::string& s = c.explain;
    snprintf(buf, sizeof buf, "kind=%d scan_tuple=%d cols=%zu\n", c.kind, c.scan_tuple, c.cols.size()); s += buf;
    for (size_t i = 0; i < c.cols.size(); i++) { snprintf(buf, sizeof buf, "  col[%zu] = %d_%d prim=%d\n", i, c.cols[i].tuple_id, c.cols[i].slot_id, c.cols[i].prim); s += buf; }
    if (c.kind == PK_AGG || c.kind == PK_JOIN_AGG) {
        const AggPlan& ap = c.ap;
        snprintf(buf, sizeof buf, "agg: n_group=%d n_keyw=%d n_agg=%d n_lanes=%d pred_out=%d direct=%d emit_default=%d\n", ap.n_group, ap.n_keyw,
                 ap.n_agg, ap.n_lanes, ap.pred_out, (int)c.has_direct, (int)c.emit_default); s += buf;
        for (int k = 0; k < ap.n_agg; k++) {
            const AggSpec& a = ap.agg[k];
            snprintf(buf, sizeof buf, "  agg[%d] kind=%d class=%d arg_class=%d acc_lane=%d cnt_lane=%d arg_out=%d out_prim=%d\n", k, a.kind, a.vclass,
                     a.arg_vclass, a.acc_lane, a.cnt_lane, a.arg_out, a.out_prim); s += buf;
        }
        if (c.has_direct) {
            for (int t = 0; t < c.direct.n_terms; t++) {
                snprintf(buf, sizeof buf, "  direct term[%d]: col=%d cmp=%d class=%d const=0x%llx\n", t, c.direct_cols[(size_t)t], c.direct.term[t].cmp,
                         c.direct.term[t].vclass, (unsigned long long)c.direct.term[t].cbits); s += buf;
            }
        }
    }
    if (c.kind == PK_SORT || c.kind == PK_FILTER) {
        snprintf(buf, sizeof buf, "rows: pred_out=%d limit=%lld offset=%lld direct=%d\n", c.ap.pred_out, (long long)c.limit, (long long)c.offset, (int)c.has_direct); s += buf;
        for (auto& k : c.sort_keys) { snprintf(buf, sizeof buf, "  order key: out=%d prim=%d asc=%d null_first=%d\n", k.out_reg, k.prim, (int)k.asc, (int)k.null_first); s += buf; }
    }
    if (c.post) { snprintf(buf, sizeof buf, "post fragment above the aggregate: kind=%d keys=%zu pred_out=%d limit=%lld offset=%lld\n", c.post->kind, c.post->sort_keys.size(),
                           c.post->ap.pred_out, (long long)c.post->limit, (long long)c.post->offset); s += buf; }
    snprintf(buf, sizeof buf, "program: %d instr, %d outputs\n", c.prog.n_instr, c.prog.n_out); s += buf;
    for (int i = 0; i < c.prog.n_instr; i++) {
        const Instr& in = c.prog.code[i];
        snprintf(buf, sizeof buf, "  %02d %-9s a=%d b=%d c=%d", i, op_name(in.op), in.a, in.b, in.c); s += buf;
        if (in.op == OP_CONST) { snprintf(buf, sizeof buf, "   ; 0x%llx%s", (unsigned long long)c.prog.cbits[in.a], ((c.prog.cnull >> in.a) & 1) ? " NULL" : ""); s += buf; }
        s += "\n";
    }
}

// ---------------------------------------------------------------- entry
bool lower_sort(Infer& in, Compiled& out, const HNode& sort, const HNode* filter, const HNode& scan);
bool lower_filter(Infer& in, Compiled& out, const HNode* limit_node, const HNode& filter_or_scan, const HNode& scan);
bool lower_join_agg(Infer& in, Compiled& out, const HNode& agg, const HNode& join, bool under_packet, const HNode* above);
bool lower_post(Infer& in, Compiled& out, const HNode* sort, const HNode* having, const HNode* limit_node);
bool lower_join_rows(Infer& in, Compiled& out, const HNode& join, const HNode* above, const HNode* sort, const HNode* limit_node);

static bool infer_node(Infer& in, HNode& n) {
    for (auto& e : n.conjuncts) if (!infer_expr(in, e)) return false;
    for (auto& e : n.group_exprs) if (!infer_expr(in, e)) return false;
    for (auto& e : n.agg_fns) if (!infer_expr(in, e)) return false;
    for (auto& e : n.order_exprs) if (!infer_expr(in, e)) return false;
    for (auto& c : n.ch) if (!infer_node(in, c)) return false;
    return true;
}

// [SORT ->] [FILTER ->] JOIN at the top of the fragment -> the JOIN node, else nullptr
static const HNode* join_rows_shape(const HNode* t) {
    if (t && t->node_type == BK_SORT_NODE && !t->ch.empty()) t = skip_passthrough(&t->ch[0], nullptr);
    if (t && is_filter(t) && !t->ch.empty()) t = skip_passthrough(&t->ch[0], nullptr);
    return t && t->node_type == BK_JOIN_NODE ? t : nullptr;
}

int compile_plan(const uint8_t* desc, size_t len, Compiled& out, std::string& err) {
    if (!desc || len < 16 || (len & 3)) { err = "plan descriptor is empty or not a multiple of 4 bytes"; return BKGPU_EINVAL; }
    Reader r{}; r.w = (const int32_t*)desc; r.n = len / 4;
    if ((uint32_t)r.rd() != BKGPU_PLAN_MAGIC) { err = "bad plan magic"; return BKGPU_EINVAL; }
    if (r.rd() != BKGPU_PLAN_VERSION) { err = "unsupported plan version"; return BKGPU_EINVAL; }
    int nt = r.rd(), nn = r.rd();
    if (nt < 0 || nt > 16 || nn <= 0 || nn > 64) { err = "bad tuple / node count"; return BKGPU_EINVAL; }
    for (int i = 0; i < nt && !r.fail; i++) {
        HTuple t; t.tuple_id = r.rd(); int ns = r.rd();
        if (ns < 0 || ns > 4096) { r.bad("bad slot count"); break; }
        for (int k = 0; k < ns; k++) { int sid = r.rd(); int pt = r.rd(); t.slots.push_back({sid, pt}); }
        out.tuples.push_back(t);
    }
    HNode root;
    if (!r.fail) parse_node(r, root, nn, 0);
    if (!r.fail && nn != 0) r.bad("plan node count mismatch");
    if (!r.fail && r.pos != r.n) r.bad("trailing words after the plan");
    if (r.fail) { err = r.err; return r.err.find("outside the GPU path") != std::string::npos ? BKGPU_EUNSUPPORTED : BKGPU_EINVAL; }
    Infer in{}; in.tuples = &out.tuples;
    if (!infer_node(in, root)) { err = in.err; return in.code; }

    bool under_packet = false;
    const HNode* top = skip_passthrough(&root, &under_packet);
    if (!top) { err = "empty plan"; return BKGPU_EINVAL; }
    bool ok = false;
    const HNode* limit_node = nullptr;
    if (top->node_type == BK_LIMIT_NODE && !top->ch.empty()) { limit_node = top; top = skip_passthrough(&top->ch[0], &under_packet); }
    // the db-side chain above an aggregate: [LIMIT ->] [SORT ->] [HAVING_FILTER ->] (MERGE_)AGG (separate.cpp:241-260): the operators
    // above the aggregate become a post fragment over the aggregate's output rows
    const HNode *post_sort = nullptr, *post_having = nullptr;
    {
        const HNode* t = top;
        const HNode *ps = nullptr, *ph = nullptr;
        if (t->node_type == BK_SORT_NODE && !t->ch.empty()) { ps = t; t = skip_passthrough(&t->ch[0], nullptr); }
        if (t && (is_filter(t) || t->node_type == BK_HAVING_FILTER_NODE) && !t->ch.empty()) { ph = t; t = skip_passthrough(&t->ch[0], nullptr); }
        if (t && (ps || ph) && (t->node_type == BK_AGG_NODE || t->node_type == BK_MERGE_AGG_NODE)) { post_sort = ps; post_having = ph; top = t; }
    }
    if (top->node_type == BK_AGG_NODE || top->node_type == BK_MERGE_AGG_NODE) {
        if (top->ch.empty()) { err = "AGG node without a child"; return BKGPU_EINVAL; }
        const HNode* c = skip_passthrough(&top->ch[0], nullptr);
        const HNode* filter = nullptr;
        if (is_filter(c)) { filter = c; c = c->ch.empty() ? nullptr : skip_passthrough(&c->ch[0], nullptr); }
        if (c && c->node_type == BK_SCAN_NODE) {
            if (filter && filter->limit != -1) { err = "LIMIT on a filter below an aggregate is order dependent: outside the GPU path"; return BKGPU_EUNSUPPORTED; }
            std::vector<const HExpr*> conj;
            if (filter) for (auto& e : filter->conjuncts) conj.push_back(&e);
            ok = lower_agg(in, out, *top, conj, c->tuple_id, under_packet, true);
            if (ok && limit_node && !post_sort && !post_having) { out.limit = limit_node->limit; out.offset = limit_node->offset; }
        } else if (c && c->node_type == BK_JOIN_NODE) {   // AGG -> [FILTER ->] JOIN: the store-side chain of a filtered join (separate.cpp:241-260)
            if (filter && filter->limit != -1) { err = "LIMIT on a filter below an aggregate is order dependent: outside the GPU path"; return BKGPU_EUNSUPPORTED; }
            ok = lower_join_agg(in, out, *top, *c, under_packet, filter);
            if (ok && limit_node && !post_sort && !post_having) { out.limit = limit_node->limit; out.offset = limit_node->offset; }   // LimitNode over the joined aggregate
        } else { err = "AGG child must be [FILTER ->] SCAN or JOIN"; return BKGPU_EUNSUPPORTED; }
        if (ok && (post_sort || post_having)) ok = lower_post(in, out, post_sort, post_having, limit_node);
    } else if (const HNode* jn = join_rows_shape(top)) {   // [SORT ->] [FILTER ->] JOIN: a join that returns its rows
        const HNode* t = top; const HNode *js = nullptr, *jf = nullptr;
        if (t->node_type == BK_SORT_NODE) { js = t; t = skip_passthrough(&t->ch[0], nullptr); }
        if (is_filter(t)) jf = t;
        ok = lower_join_rows(in, out, *jn, jf, js, limit_node);
    } else if (top->node_type == BK_SORT_NODE) {
        const HNode* c = top->ch.empty() ? nullptr : skip_passthrough(&top->ch[0], nullptr);
        const HNode* filter = nullptr;
        if (is_filter(c)) { filter = c; c = c->ch.empty() ? nullptr : skip_passthrough(&c->ch[0], nullptr); }
        if (!c || c->node_type != BK_SCAN_NODE) { err = "SORT child must be [FILTER ->] SCAN"; return BKGPU_EUNSUPPORTED; }
        ok = lower_sort(in, out, *top, filter, *c);
        if (ok && limit_node) {
            out.offset = limit_node->offset;
            if (limit_node->limit >= 0) {   // (limit -1 = OFFSET only: the sort keeps every row)
                const int64_t lim = limit_node->limit + limit_node->offset;
                if (out.limit < 0 || lim < out.limit) out.limit = lim;
            }
        }
    } else if (is_filter(top) || top->node_type == BK_SCAN_NODE) {
        const HNode* c = top;
        if (is_filter(c)) c = c->ch.empty() ? nullptr : skip_passthrough(&c->ch[0], nullptr);
        if (!c || c->node_type != BK_SCAN_NODE) { err = "FILTER child must be SCAN"; return BKGPU_EUNSUPPORTED; }
        ok = lower_filter(in, out, limit_node, *top, *c);
    } else { err = "plan root is outside the GPU path"; return BKGPU_EUNSUPPORTED; }
    if (!ok) { err = in.err.empty() ? "plan could not be lowered" : in.err; return in.code ? in.code : BKGPU_EUNSUPPORTED; }
    explain(out);
    return BKGPU_OK;
}

// ---------------------------------------------------------------- SORT / filter-only fragments
// Rows of the scan tuple are returned (all slots of its tuple descriptor), ordered / filtered.
static bool intern_scan_tuple(Infer& in, Compiled& out, Lower& lw, int tuple_id) {
    const HTuple* t = nullptr;
    for (auto& x : out.tuples) if (x.tuple_id == tuple_id) t = &x;
    if (!t || t->slots.empty()) return in.fail(BKGPU_EINVAL, "scan tuple %d has no descriptor", tuple_id);
    for (auto& sl : t->slots) {
        if (prim_storage(sl.second) < 0 || sl.second == BK_STRING)
            return in.fail(BKGPU_EUNSUPPORTED, "column %d_%d has type %d: outside the GPU path", tuple_id, sl.first, sl.second);
        int ci = lw.intern_col(tuple_id, sl.first, sl.second);
        if (ci >= MAX_COLS) return in.fail(BKGPU_EUNSUPPORTED, "more than %d columns in the scan tuple", MAX_COLS);
        out.out_cols.push_back({tuple_id, sl.first, sl.second, 0});
    }
    return true;
}
static bool lower_predicate(Infer& in, Lower& lw, const HNode* filter, AggPlan& ap, int& reg) {
    ap.pred_out = -1;
    if (!filter || filter->conjuncts.empty()) return true;
    if (filter->conjuncts.size() > 8) return in.fail(BKGPU_EUNSUPPORTED, "more than 8 conjuncts");
    for (auto& c : filter->conjuncts) { int depth = 0; if (!lw.expr(c, depth) || !lw.to_bool(c)) return false; }
    if (filter->conjuncts.size() > 1 && !lw.emit(OP_AND, (uint8_t)filter->conjuncts.size())) return false;
    ap.pred_out = reg;
    return lw.out_reg(reg++);
}

bool lower_sort(Infer& in, Compiled& out, const HNode& sort, const HNode* filter, const HNode& scan) {
    out.kind = PK_SORT;
    out.scan_tuple = scan.tuple_id;
    out.limit = sort.limit;
    if (filter && filter->limit != -1) return in.fail(BKGPU_EUNSUPPORTED, "LIMIT on a filter below a sort is order dependent");
    Program& p = out.prog; memset(&p, 0, sizeof p);
    memset(&out.ap, 0, sizeof out.ap);
    Lower lw{&out, &in, &p, {}};
    if (!intern_scan_tuple(in, out, lw, scan.tuple_id)) return false;
    int reg = 0;
    if (!lower_predicate(in, lw, filter, out.ap, reg)) return false;
    if (sort.order_exprs.empty()) return in.fail(BKGPU_EINVAL, "SORT node without order expressions");
    if (sort.order_exprs.size() > 4) return in.fail(BKGPU_EUNSUPPORTED, "more than 4 ORDER BY expressions");
    for (size_t i = 0; i < sort.order_exprs.size(); i++) {
        const HExpr& e = sort.order_exprs[i];
        if (e.col_type == BK_STRING) return in.fail(BKGPU_EUNSUPPORTED, "ORDER BY over a STRING key is outside the GPU path");
        int depth = 0;
        if (!lw.expr(e, depth)) return false;
        out.sort_keys.push_back({reg, e.col_type, sort.is_asc[i] != 0, sort.is_null_first[i] != 0});
        if (!lw.out_reg(reg++)) return false;
    }
    p.n_out = reg;
    // direct shape: one plain column key, no filter (config C5)
    out.has_direct = false;
    if (!filter && sort.order_exprs.size() == 1 && sort.order_exprs[0].node_type == BK_SLOT_REF) {
        const HExpr& e = sort.order_exprs[0];
        int st = in.slot_type(e.tuple_id, e.slot_id);
        if (st == e.col_type && (prim_storage(st) == ST_I64 || prim_storage(st) == ST_U64 || prim_storage(st) == ST_F64 ||
                                 prim_storage(st) == ST_I32 || prim_storage(st) == ST_U32)) {
            out.has_direct = true;
            out.direct_cols = {lw.intern_col(e.tuple_id, e.slot_id, st)};
        }
    }
    return true;
}

bool lower_filter(Infer& in, Compiled& out, const HNode* limit_node, const HNode& top, const HNode& scan) {
    out.kind = PK_FILTER;
    out.scan_tuple = scan.tuple_id;
    const HNode* filter = is_filter(&top) ? &top : nullptr;
    out.limit = filter ? filter->limit : scan.limit;   // FilterNode honours its own limit in input order (filter_node.cpp:786-791)
    out.offset = 0;
    if (limit_node) {
        out.offset = limit_node->offset;
        int64_t lim = limit_node->limit < 0 ? -1 : limit_node->limit + limit_node->offset;
        if (lim >= 0 && (out.limit < 0 || lim < out.limit)) out.limit = lim;
    }
    Program& p = out.prog; memset(&p, 0, sizeof p);
    memset(&out.ap, 0, sizeof out.ap);
    Lower lw{&out, &in, &p, {}};
    if (!intern_scan_tuple(in, out, lw, scan.tuple_id)) return false;
    int reg = 0;
    if (!lower_predicate(in, lw, filter, out.ap, reg)) return false;
    p.n_out = reg;
    return true;
}

// [LIMIT ->] [SORT ->] [HAVING ->] over an aggregate: a fragment of kind PK_SORT / PK_FILTER whose input rows are the aggregate's
// output columns (SortNode::open src/exec/sort_node.cpp:278-346, FilterNode::get_next src/exec/filter_node.cpp:726-795 over the
// rows AggNode::get_next emits).  The AVG intermediate blobs ride along as 16-byte payload columns.
bool lower_post(Infer& in, Compiled& out, const HNode* sort, const HNode* having, const HNode* limit_node) {
    auto post = std::make_shared<Compiled>();
    Compiled& pc = *post;
    pc.kind = sort ? PK_SORT : PK_FILTER;
    pc.tuples = out.tuples;
    pc.scan_tuple = -1;
    Program& p = pc.prog; memset(&p, 0, sizeof p);
    memset(&pc.ap, 0, sizeof pc.ap);
    Lower lw{&pc, &in, &p, {}};
    if (out.out_cols.size() > (size_t)MAX_COLS) return in.fail(BKGPU_EUNSUPPORTED, "more than %d aggregate output columns under a SORT / HAVING", MAX_COLS);
    for (const OutCol& oc : out.out_cols) {
        const int prim = oc.kind == 1 ? BK_STRING : oc.prim;
        const int ci = lw.intern_col(oc.tuple_id, oc.slot_id, prim);
        if (ci != (int)pc.out_cols.size()) return in.fail(BKGPU_EUNSUPPORTED, "aggregate output columns %d_%d appear twice", oc.tuple_id, oc.slot_id);
        pc.out_cols.push_back({oc.tuple_id, oc.slot_id, prim, oc.kind});
    }
    const size_t n_payload = pc.cols.size();
    int reg = 0;
    if (having && having->limit != -1 && sort) return in.fail(BKGPU_EUNSUPPORTED, "LIMIT on a HAVING filter below a sort is order dependent");
    if (!lower_predicate(in, lw, having, pc.ap, reg)) return false;
    if (sort) {
        pc.limit = sort->limit;
        if (sort->order_exprs.empty()) return in.fail(BKGPU_EINVAL, "SORT node without order expressions");
        if (sort->order_exprs.size() > 4) return in.fail(BKGPU_EUNSUPPORTED, "more than 4 ORDER BY expressions");
        for (size_t i = 0; i < sort->order_exprs.size(); i++) {
            const HExpr& e = sort->order_exprs[i];
            if (e.col_type == BK_STRING) return in.fail(BKGPU_EUNSUPPORTED, "ORDER BY over a STRING key is outside the GPU path");
            int depth = 0;
            if (!lw.expr(e, depth)) return false;
            pc.sort_keys.push_back({reg, e.col_type, sort->is_asc[i] != 0, sort->is_null_first[i] != 0});
            if (!lw.out_reg(reg++)) return false;
        }
        if (limit_node) {
            pc.offset = limit_node->offset;
            if (limit_node->limit >= 0) { const int64_t lim = limit_node->limit + limit_node->offset; if (pc.limit < 0 || lim < pc.limit) pc.limit = lim; }
        }
    } else {
        pc.limit = having ? having->limit : -1;
        pc.offset = 0;
        if (limit_node) {
            pc.offset = limit_node->offset;
            const int64_t lim = limit_node->limit < 0 ? -1 : limit_node->limit + limit_node->offset;
            if (lim >= 0 && (pc.limit < 0 || lim < pc.limit)) pc.limit = lim;
        }
    }
    p.n_out = reg;
    if (pc.cols.size() != n_payload) return in.fail(BKGPU_EUNSUPPORTED, "SORT / HAVING above an aggregate references a column the aggregate does not output");
    pc.has_direct = false;
    out.post = post;
    return true;
}

// AGG over an equi-join (config C3).  The reference builds its hash map on the OUTER (driver) child and probes
// with the inner child's rows (JoinNode::hash_join, src/exec/join_node.cpp:920-1022; Joiner::construct_hash_map,
// src/exec/joiner.cpp:624-631); the key is the cast equal-slot value (strip_out_equal_slots, joiner.cpp:166-217).
// Filters of both children and the residual join conditions become one predicate over the joined row — for an
// INNER join that is the same set of rows.
bool lower_join_agg(Infer& in, Compiled& out, const HNode& agg, const HNode& join, bool under_packet, const HNode* above) {
    int jt = join.join_type;
    if (jt != BK_INNER_JOIN && jt != BK_LEFT_JOIN && jt != BK_RIGHT_JOIN && jt != BK_SEMI_JOIN && jt != BK_ANTI_SEMI_JOIN)
        return in.fail(BKGPU_EUNSUPPORTED, "join type %d (FULL / NULL) is outside the GPU path", jt);
    if (join.ch.size() != 2) return in.fail(BKGPU_EINVAL, "JOIN node needs two children");
    const HNode* side[2]; const HNode* filt[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) {
        // RIGHT JOIN: the reference swaps the roles — the right child becomes the outer (preserved, driver) table (join_node.cpp:151-156)
        const HNode* c = skip_passthrough(&join.ch[(size_t)(jt == BK_RIGHT_JOIN ? 1 - i : i)], nullptr);
        if (is_filter(c)) { filt[i] = c; c = c->ch.empty() ? nullptr : skip_passthrough(&c->ch[0], nullptr); }
        if (!c || c->node_type != BK_SCAN_NODE) return in.fail(BKGPU_EUNSUPPORTED, "JOIN children must be [FILTER ->] SCAN");
        if (filt[i] && filt[i]->limit != -1) return in.fail(BKGPU_EUNSUPPORTED, "LIMIT below a join is order dependent");
        side[i] = c;
    }
    if (jt == BK_RIGHT_JOIN) jt = BK_LEFT_JOIN;
    // the fused predicate (child filters + residual conditions over the joined row) equals filter-then-join only for INNER joins
    // a filter ABOVE the join sees the joined row: for an INNER join it simply joins the residual conditions; above an outer join it would also
    // have to judge the NULL-extended rows, after the match decision — a second predicate the probe does not carry
    if (jt != BK_INNER_JOIN && above) return in.fail(BKGPU_EUNSUPPORTED, "a filter between the aggregate and a LEFT / SEMI / ANTI join is outside the GPU path");
    if (jt != BK_INNER_JOIN && (filt[0] || filt[1])) return in.fail(BKGPU_EUNSUPPORTED, "LEFT / SEMI / ANTI join over filtered children is outside the GPU path");
    const int build_tuple = side[0]->tuple_id, probe_tuple = side[1]->tuple_id;
    out.build_tuple = build_tuple;
    std::vector<const HExpr*> conj;
    const HExpr *bk = nullptr, *pk = nullptr;
    for (auto& e : join.conjuncts) {
        bool taken = false;
        if (!bk && e.node_type == BK_FUNCTION_CALL && e.fn_op == BK_FT_EQ && e.ch.size() == 2 && e.ch[0].node_type == BK_SLOT_REF && e.ch[1].node_type == BK_SLOT_REF) {
            const HExpr *a = &e.ch[0], *b = &e.ch[1];
            if (a->tuple_id == build_tuple && b->tuple_id == probe_tuple) { bk = a; pk = b; taken = true; }
            else if (b->tuple_id == build_tuple && a->tuple_id == probe_tuple) { bk = b; pk = a; taken = true; }
        }
        if (!taken) conj.push_back(&e);
    }
    if (!bk) return in.fail(BKGPU_EUNSUPPORTED, "join without an equality between the two tables (nested loop) is outside the GPU path");
    for (int i = 0; i < 2; i++) if (filt[i]) for (auto& e : filt[i]->conjuncts) conj.push_back(&e);
    if (above) for (auto& e : above->conjuncts) conj.push_back(&e);
    int ot = bk->col_type, it = pk->col_type, cast;
    auto is_signed_t = [](int t) { return t >= BK_INT8 && t <= BK_INT64; };
    if (ot == it) cast = ot;
    else if (is_signed_t(ot) && is_signed_t(it)) cast = BK_INT64;
    else if (is_uint_t(ot) && is_uint_t(it)) cast = BK_UINT64;
    else return in.fail(BKGPU_EUNSUPPORTED, "join keys of types %d and %d are compared as STRING in the reference: outside the GPU path", ot, it);
    if (cast == BK_STRING || is_double_t(cast)) return in.fail(BKGPU_EUNSUPPORTED, "join key type %d is outside the GPU path", cast);
    if (!lower_agg(in, out, agg, conj, probe_tuple, under_packet, false)) return false;
    out.kind = PK_JOIN_AGG;
    out.join_type = jt; out.join_key_prim = cast;
    Program dummy; memset(&dummy, 0, sizeof dummy);
    Lower lw{&out, &in, &dummy, {}};
    out.build_key_col = lw.intern_col(bk->tuple_id, bk->slot_id, in.slot_type(bk->tuple_id, bk->slot_id));
    out.probe_key_col = lw.intern_col(pk->tuple_id, pk->slot_id, in.slot_type(pk->tuple_id, pk->slot_id));
    if ((int)out.cols.size() > MAX_COLS) return in.fail(BKGPU_EUNSUPPORTED, "more than %d columns referenced", MAX_COLS);
    // fast path: lower the same aggregate again with every column treated as a column of ONE (virtual, joined) tuple
    if (jt == BK_INNER_JOIN) {
        auto jf = std::make_shared<Compiled>();
        jf->tuples = out.tuples; jf->build_tuple = -1;
        Infer in2{}; in2.tuples = &jf->tuples;
        if (lower_agg(in2, *jf, agg, conj, probe_tuple, under_packet, true) && memcmp(&jf->ap.n_keyw, &out.ap.n_keyw, sizeof(int32_t) * 4) == 0 &&
            jf->ap.n_lanes == out.ap.n_lanes) {
            bool ok = true;
            for (auto& c : jf->cols) {
                int m = -1;
                for (size_t i = 0; i < out.cols.size(); i++) if (out.cols[i].tuple_id == c.tuple_id && out.cols[i].slot_id == c.slot_id) m = (int)i;
                if (m < 0) ok = false;
                out.jfast_of_main.push_back(m);
            }
            if (ok) out.jfast = jf; else out.jfast_of_main.clear();
        }
    }
    return true;
}

// A JOIN whose rows are the result ([LIMIT ->] [SORT ->] [FILTER ->] JOIN; JoinNode::get_next, join_node.cpp:1200-1326).  The joined rows that
// pass the conditions are produced on the device batch by batch (pairs + one gather per column) and handed to a sink fragment
// (Compiled::post, kind PK_FILTER without a predicate, or PK_SORT) whose "scan tuple" is the joined row: every slot of both tuples.
// INNER: child filters, the residual conditions and a filter above the join are one predicate.  LEFT / RIGHT: the outer (build) side is
// preserved, unmatched rows come out NULL-extended; filtered children or a filter above are refused (as in lower_join_agg).
bool lower_join_rows(Infer& in, Compiled& out, const HNode& join, const HNode* above, const HNode* sort, const HNode* limit_node) {
    int jt = join.join_type;
    if (jt != BK_INNER_JOIN && jt != BK_LEFT_JOIN && jt != BK_RIGHT_JOIN)
        return in.fail(BKGPU_EUNSUPPORTED, "a join of type %d that returns rows is outside the GPU path (INNER / LEFT / RIGHT are lowered)", jt);
    if (join.ch.size() != 2) return in.fail(BKGPU_EINVAL, "JOIN node needs two children");
    if (above && above->limit != -1) return in.fail(BKGPU_EUNSUPPORTED, "LIMIT on a filter above a join: use a LIMIT node");
    const HNode* side[2]; const HNode* filt[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; i++) {
        const HNode* c = skip_passthrough(&join.ch[(size_t)(jt == BK_RIGHT_JOIN ? 1 - i : i)], nullptr);
        if (is_filter(c)) { filt[i] = c; c = c->ch.empty() ? nullptr : skip_passthrough(&c->ch[0], nullptr); }
        if (!c || c->node_type != BK_SCAN_NODE) return in.fail(BKGPU_EUNSUPPORTED, "JOIN children must be [FILTER ->] SCAN");
        if (filt[i] && filt[i]->limit != -1) return in.fail(BKGPU_EUNSUPPORTED, "LIMIT below a join is order dependent");
        side[i] = c;
    }
    if (jt == BK_RIGHT_JOIN) jt = BK_LEFT_JOIN;
    if (jt != BK_INNER_JOIN && (filt[0] || filt[1] || above)) return in.fail(BKGPU_EUNSUPPORTED, "LEFT join with filtered children or a filter above it is outside the GPU path");
    const int build_tuple = side[0]->tuple_id, probe_tuple = side[1]->tuple_id;
    out.kind = PK_JOIN; out.build_tuple = build_tuple; out.scan_tuple = probe_tuple; out.join_type = jt;
    std::vector<const HExpr*> conj;
    const HExpr *bk = nullptr, *pk = nullptr;
    for (auto& e : join.conjuncts) {
        bool taken = false;
        if (!bk && e.node_type == BK_FUNCTION_CALL && e.fn_op == BK_FT_EQ && e.ch.size() == 2 && e.ch[0].node_type == BK_SLOT_REF && e.ch[1].node_type == BK_SLOT_REF) {
            const HExpr *a = &e.ch[0], *b = &e.ch[1];
            if (a->tuple_id == build_tuple && b->tuple_id == probe_tuple) { bk = a; pk = b; taken = true; }
            else if (b->tuple_id == build_tuple && a->tuple_id == probe_tuple) { bk = b; pk = a; taken = true; }
        }
        if (!taken) conj.push_back(&e);
    }
    if (!bk) return in.fail(BKGPU_EUNSUPPORTED, "join without an equality between the two tables (nested loop) is outside the GPU path");
    for (int i = 0; i < 2; i++) if (filt[i]) for (auto& e : filt[i]->conjuncts) conj.push_back(&e);
    if (above) for (auto& e : above->conjuncts) conj.push_back(&e);
    int ot = bk->col_type, it = pk->col_type, cast;
    auto is_signed_t = [](int t) { return t >= BK_INT8 && t <= BK_INT64; };
    if (ot == it) cast = ot;
    else if (is_signed_t(ot) && is_signed_t(it)) cast = BK_INT64;
    else if (is_uint_t(ot) && is_uint_t(it)) cast = BK_UINT64;
    else return in.fail(BKGPU_EUNSUPPORTED, "join keys of types %d and %d are compared as STRING in the reference: outside the GPU path", ot, it);
    if (cast == BK_STRING || is_double_t(cast)) return in.fail(BKGPU_EUNSUPPORTED, "join key type %d is outside the GPU path", cast);
    out.join_key_prim = cast;
    Program& p = out.prog; memset(&p, 0, sizeof p);
    memset(&out.ap, 0, sizeof out.ap);
    Lower lw{&out, &in, &p, {}};
    // every slot of both tuples is an output column (and a device column): build tuple first, as the joined row lists them
    for (int sd = 0; sd < 2; sd++) {
        const int tid = sd == 0 ? build_tuple : probe_tuple;
        const HTuple* t = nullptr;
        for (auto& x : out.tuples) if (x.tuple_id == tid) t = &x;
        if (!t || t->slots.empty()) return in.fail(BKGPU_EINVAL, "join tuple %d has no descriptor", tid);
        for (auto& sl : t->slots) {
            if (prim_storage(sl.second) < 0 || sl.second == BK_STRING) return in.fail(BKGPU_EUNSUPPORTED, "column %d_%d has type %d: outside the GPU path", tid, sl.first, sl.second);
            if (lw.intern_col(tid, sl.first, sl.second) >= MAX_COLS) return in.fail(BKGPU_EUNSUPPORTED, "more than %d columns in the joined row", MAX_COLS);
            out.out_cols.push_back({tid, sl.first, sl.second, 0});
        }
    }
    const size_t n_payload = out.cols.size();
    out.build_key_col = lw.intern_col(bk->tuple_id, bk->slot_id, in.slot_type(bk->tuple_id, bk->slot_id));
    out.probe_key_col = lw.intern_col(pk->tuple_id, pk->slot_id, in.slot_type(pk->tuple_id, pk->slot_id));
    int reg = 0;
    out.ap.pred_out = -1;
    if (!conj.empty()) {
        if (conj.size() > 8) return in.fail(BKGPU_EUNSUPPORTED, "more than 8 conjuncts");
        for (auto* c : conj) { int depth = 0; if (!lw.expr(*c, depth) || !lw.to_bool(*c)) return false; }
        if (conj.size() > 1 && !lw.emit(OP_AND, (uint8_t)conj.size())) return false;
        out.ap.pred_out = reg;
        if (!lw.out_reg(reg++)) return false;
    }
    p.n_out = reg;
    if (out.cols.size() != n_payload) return in.fail(BKGPU_EINVAL, "join conditions reference a slot that is in neither tuple descriptor");
    // the sink over the joined rows
    auto post = std::make_shared<Compiled>();
    Compiled& pc = *post;
    pc.kind = sort ? PK_SORT : PK_FILTER;
    pc.tuples = out.tuples; pc.scan_tuple = -1;
    Program& pp = pc.prog; memset(&pp, 0, sizeof pp);
    memset(&pc.ap, 0, sizeof pc.ap);
    pc.ap.pred_out = -1;
    Infer in2{}; in2.tuples = &pc.tuples;
    Lower lw2{&pc, &in2, &pp, {}};
    for (const OutCol& oc : out.out_cols) { lw2.intern_col(oc.tuple_id, oc.slot_id, oc.prim); pc.out_cols.push_back(oc); }
    int preg = 0;
    if (sort) {
        pc.limit = sort->limit;
        if (sort->order_exprs.empty()) return in.fail(BKGPU_EINVAL, "SORT node without order expressions");
        if (sort->order_exprs.size() > 4) return in.fail(BKGPU_EUNSUPPORTED, "more than 4 ORDER BY expressions");
        for (size_t i = 0; i < sort->order_exprs.size(); i++) {
            const HExpr& e = sort->order_exprs[i];
            if (e.col_type == BK_STRING) return in.fail(BKGPU_EUNSUPPORTED, "ORDER BY over a STRING key is outside the GPU path");
            int depth = 0;
            if (!lw2.expr(e, depth)) { in.code = in2.code; in.err = in2.err; return false; }
            pc.sort_keys.push_back({preg, e.col_type, sort->is_asc[i] != 0, sort->is_null_first[i] != 0});
            if (!lw2.out_reg(preg++)) return false;
        }
        if (limit_node) {
            pc.offset = limit_node->offset;
            if (limit_node->limit >= 0) { const int64_t lim = limit_node->limit + limit_node->offset; if (pc.limit < 0 || lim < pc.limit) pc.limit = lim; }
        }
    } else {
        pc.limit = join.limit; pc.offset = 0;
        if (limit_node) {
            pc.offset = limit_node->offset;
            const int64_t lim = limit_node->limit < 0 ? -1 : limit_node->limit + limit_node->offset;
            if (lim >= 0 && (pc.limit < 0 || lim < pc.limit)) pc.limit = lim;
        }
    }
    pp.n_out = preg;
    if (pc.cols.size() != out.out_cols.size()) return in.fail(BKGPU_EUNSUPPORTED, "ORDER BY above a join references a column outside the joined row");
    pc.has_direct = false;
    out.post = post;
    out.limit = pc.limit; out.offset = pc.offset;
    return true;
}

}  // namespace bk
