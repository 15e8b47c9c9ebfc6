// bkgpu_dictionary.hpp — STRING columns on the GPU path as ORDER-PRESERVING dictionary codes, on the C++ side of the adapter
// (the Python mirror, with the reasoning and the checks against pyarrow's string kernels: baikaldb_b200/dictionary.py, tests/test_dictionary.py;
// tests/test_host_cpp.py checks that this rewrite produces the SAME plan bytes and the SAME codes).
//
// A fragment whose STRING slots are only compared (= != < <= > >=, IN, IS NULL, LIKE 'literal'), grouped, joined, ordered, counted or MIN / MAX-ed needs
// only the ORDER of the strings: GpuExecNode's child hands the strings over, this builds one sorted dictionary per comparison domain
// (string slots compared with each other share one), replaces every string by its rank (INT32, NULL stays NULL), rewrites the fragment —
// literals become rank thresholds — and maps the codes of the result's key / MIN / MAX columns back.  Byte order is the reference's string
// order (ExprValue::compare -> std::string::compare, include/common/expr_value.h:892-943).  Anything else on a string throws Unsupported:
// refused, never answered differently.
#pragma once
#include <algorithm>
#include <functional>
#include <map>
#include <optional>
#include <stdexcept>
#include "bkgpu_host.hpp"

namespace bkgpu {

struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };

struct StringColumn { int tuple_id, slot_id; std::vector<std::optional<std::string>> values; };

struct EncodedStrings {
    Plan plan;                                                                  // the rewritten fragment
    std::vector<Column> columns;                                                // INT32 code columns, in the order of the string columns given
    std::map<std::pair<int, int>, std::shared_ptr<std::vector<std::string>>> dictionaries;   // scan slot -> its domain's sorted dictionary
    std::map<std::pair<int, int>, std::shared_ptr<std::vector<std::string>>> result_slots;   // result slot -> dictionary to decode it with
    // a result column of codes -> strings (nullopt = NULL); columns that are not string-valued are left to the caller
    bool decodes(const Column& c) const { return result_slots.count({c.tuple_id, c.slot_id}) != 0; }
    std::vector<std::optional<std::string>> decode(const Column& c) const {
        const auto& d = *result_slots.at({c.tuple_id, c.slot_id});
        std::vector<std::optional<std::string>> out((size_t)c.length);
        for (int64_t r = 0; r < c.length; r++)
            if (c.validity.empty() || ((c.validity[(size_t)r >> 3] >> (r & 7)) & 1)) out[(size_t)r] = d[(size_t)c.at<int32_t>(r)];
        return out;
    }
};

// LikePredicate::like<Charset> (include/expr/predicate.h:503-573) restated: `%` any run of characters, `_` one character, the escape character
// takes the next pattern character literally; one backtrack point.  utf8 = false: the Binary charset (one byte per character).
// Returns -1 for a malformed character (like_one then retries as Binary, src/expr/predicate.cpp:509-547), else 0 / 1.
inline int like_code_point(const std::string& s, size_t i, bool utf8) {
    if (!utf8) return 1;
    const uint8_t b = (uint8_t)s[i];
    const int n = b < 0x80 ? 1 : (b >> 5) == 0x6 ? 2 : (b >> 4) == 0xE ? 3 : (b >> 3) == 0x1E ? 4 : 0;
    return n && i + (size_t)n <= s.size() ? n : 0;
}
inline int like_match(const std::string& target, const std::string& pattern, bool utf8, char escape = '\\') {
    size_t tx = 0, px = 0, ntx = 0, npx = 0;
    while (tx < target.size() || px < pattern.size()) {
        if (px < pattern.size()) {
            int pn = like_code_point(pattern, px, utf8);
            if (pn == 0) return -1;
            if (pn == 1 && pattern[px] == '_') {
                if (tx < target.size()) { const int tn = like_code_point(target, tx, utf8); px++; tx += tn > 0 ? (size_t)tn : 1; continue; }
            } else if (pn == 1 && pattern[px] == '%') {
                size_t off = 1;
                if (tx < target.size()) { const int tn = like_code_point(target, tx, utf8); if (tn > 0) off = (size_t)tn; }
                npx = px; ntx = tx + off; px++;
                continue;
            } else {
                if (pn == 1 && pattern[px] == escape && px + 1 < pattern.size()) {
                    px++;
                    pn = like_code_point(pattern, px, utf8);
                    if (pn == 0) return -1;
                }
                if (tx < target.size()) {
                    const int tn = like_code_point(target, tx, utf8);
                    if (tn == 0) return -1;
                    if (tn == pn && target.compare(tx, (size_t)tn, pattern, px, (size_t)pn) == 0) { px += (size_t)pn; tx += (size_t)tn; continue; }
                }
            }
        }
        if (ntx > 0 && ntx <= target.size()) { px = npx; tx = ntx; continue; }
        return 0;
    }
    return 1;
}
inline bool like_one(const std::string& target, const std::string& pattern, bool utf8 = true) {
    int r = like_match(target, pattern, utf8);
    if (r < 0) r = like_match(target, pattern, false);
    return r > 0;
}
constexpr size_t MAX_LIKE_RANGES = 16;   // a LIKE becomes an OR of at most this many code ranges

inline EncodedStrings encode_strings(const Plan& plan, const std::vector<StringColumn>& string_cols) {
    using Key = std::pair<int, int>;
    std::map<Key, const StringColumn*> strings;
    for (auto& c : string_cols) strings[{c.tuple_id, c.slot_id}] = &c;
    auto is_str = [&](const Expr& e) { return e.node_type == BK_SLOT_REF && strings.count({e.tuple_id, e.slot_id}) != 0; };
    auto is_cmp = [](const Expr& e) { return e.node_type == BK_FUNCTION_CALL && e.fn_op >= BK_FT_EQ && e.fn_op <= BK_FT_LE && e.children.size() == 2; };

    // ---- comparison domains (union-find over string slots compared with each other) ----
    std::map<Key, Key> parent;
    for (auto& kv : strings) parent[kv.first] = kv.first;
    std::function<Key(Key)> find = [&](Key k) { while (parent[k] != k) { parent[k] = parent[parent[k]]; k = parent[k]; } return k; };
    std::function<void(const Expr&)> scan_expr = [&](const Expr& e) {
        if (is_cmp(e) && is_str(e.children[0]) && is_str(e.children[1])) {
            const Key a = find({e.children[0].tuple_id, e.children[0].slot_id}), b = find({e.children[1].tuple_id, e.children[1].slot_id});
            parent[a] = b;
        }
        for (auto& c : e.children) scan_expr(c);
    };
    std::function<void(const PlanNode&)> scan_node = [&](const PlanNode& n) {
        for (auto& e : n.conjuncts) scan_expr(e);
        for (auto& e : n.group_exprs) scan_expr(e);
        for (auto& e : n.agg_fn_calls) scan_expr(e);
        for (auto& e : n.order_exprs) scan_expr(e);
        for (auto& c : n.children) scan_node(c);
    };
    scan_node(plan.root);
    EncodedStrings enc;
    std::map<Key, std::shared_ptr<std::vector<std::string>>> of_root;
    for (auto& kv : strings) {
        auto& d = of_root[find(kv.first)];
        if (!d) d = std::make_shared<std::vector<std::string>>();
        for (auto& v : kv.second->values) if (v) d->push_back(*v);
    }
    for (auto& kv : of_root) {
        auto& d = *kv.second;
        std::sort(d.begin(), d.end());                                           // std::string order = unsigned byte order
        d.erase(std::unique(d.begin(), d.end()), d.end());
        if (d.size() >= ((size_t)1 << 31)) throw Unsupported("more than 2^31 distinct strings in one comparison domain");
    }
    for (auto& kv : strings) enc.dictionaries[kv.first] = of_root[find(kv.first)];

    // ---- expressions ----
    auto cmp = [](int op, const char* name, Expr a, Expr b) { return Expr::fn(op, name, {std::move(a), std::move(b)}); };
    std::function<Expr(const Expr&)> rewrite = [&](const Expr& e) -> Expr {
        if (e.node_type == BK_SLOT_REF) return is_str(e) ? Expr::slot_ref(e.tuple_id, e.slot_id, BK_INT32) : e;
        if (is_cmp(e)) {
            const Expr &a = e.children[0], &b = e.children[1];
            if (is_str(a) && is_str(b)) return cmp(e.fn_op, e.name.c_str(), rewrite(a), rewrite(b));
            if (is_str(b) && a.node_type == BK_STRING_LITERAL) {                  // literal on the left: mirror the operator
                switch (e.fn_op) {
                    case BK_FT_LT: return rewrite(cmp(BK_FT_GT, "gt", b, a));
                    case BK_FT_LE: return rewrite(cmp(BK_FT_GE, "ge", b, a));
                    case BK_FT_GT: return rewrite(cmp(BK_FT_LT, "lt", b, a));
                    case BK_FT_GE: return rewrite(cmp(BK_FT_LE, "le", b, a));
                    default: return rewrite(cmp(e.fn_op, e.name.c_str(), b, a));   // EQ / NE are symmetric
                }
            }
            if (is_str(a) && b.node_type == BK_STRING_LITERAL) {
                const auto& d = *enc.dictionaries[{a.tuple_id, a.slot_id}];
                const int64_t lo = std::lower_bound(d.begin(), d.end(), b.name) - d.begin(), hi = std::upper_bound(d.begin(), d.end(), b.name) - d.begin();
                Expr col = rewrite(a);
                switch (e.fn_op) {
                    case BK_FT_EQ: return cmp(BK_FT_EQ, "eq", col, Expr::int_literal(lo != hi ? lo : -1));
                    case BK_FT_NE: return cmp(BK_FT_NE, "ne", col, Expr::int_literal(lo != hi ? lo : -1));
                    case BK_FT_LT: return cmp(BK_FT_LT, "lt", col, Expr::int_literal(lo));
                    case BK_FT_LE: return cmp(BK_FT_LT, "lt", col, Expr::int_literal(hi));
                    case BK_FT_GT: return cmp(BK_FT_GE, "ge", col, Expr::int_literal(hi));
                    default: return cmp(BK_FT_GE, "ge", col, Expr::int_literal(lo));
                }
            }
            if (is_str(a) || is_str(b)) throw Unsupported("'" + e.name + "' between a STRING column and something that is neither a STRING column nor a string literal");
        }
        if (e.node_type == BK_LIKE_PREDICATE && e.children.size() == 2 && is_str(e.children[0])) {
            // the pattern is matched against the DICTIONARY on the host; the ranks that match form ranges (one for a prefix pattern)
            if (e.children[1].node_type != BK_STRING_LITERAL) throw Unsupported("LIKE takes a literal pattern");
            const auto& d = *enc.dictionaries[{e.children[0].tuple_id, e.children[0].slot_id}];
            std::vector<std::pair<int64_t, int64_t>> ranges;
            for (size_t i = 0; i < d.size();) {
                if (!like_one(d[i], e.children[1].name)) { i++; continue; }
                size_t j = i;
                while (j < d.size() && like_one(d[j], e.children[1].name)) j++;
                ranges.push_back({(int64_t)i, (int64_t)j});
                i = j;
            }
            if (ranges.size() > MAX_LIKE_RANGES) throw Unsupported("LIKE '" + e.children[1].name + "' selects " + std::to_string(ranges.size()) + " separate ranges of the dictionary");
            if (ranges.empty()) return cmp(BK_FT_EQ, "eq", rewrite(e.children[0]), Expr::int_literal(-1));
            std::vector<Expr> terms;
            for (auto& r : ranges)
                terms.push_back(Expr::predicate(BK_AND_PREDICATE, BK_FT_LOGIC_AND, "logic_and", {cmp(BK_FT_GE, "ge", rewrite(e.children[0]), Expr::int_literal(r.first)),
                                                                                               cmp(BK_FT_LT, "lt", rewrite(e.children[0]), Expr::int_literal(r.second))}));
            if (terms.size() == 1) return terms[0];
            return Expr::predicate(BK_OR_PREDICATE, BK_FT_LOGIC_OR, "logic_or", std::move(terms));
        }
        if (e.node_type == BK_IN_PREDICATE && !e.children.empty() && is_str(e.children[0])) {
            const auto& d = *enc.dictionaries[{e.children[0].tuple_id, e.children[0].slot_id}];
            std::vector<Expr> args{rewrite(e.children[0])};
            bool any_code = false;
            for (size_t i = 1; i < e.children.size(); i++) {
                const Expr& l = e.children[i];
                if (l.node_type == BK_NULL_LITERAL) { args.push_back(Expr::null_literal()); continue; }
                if (l.node_type != BK_STRING_LITERAL) throw Unsupported("IN over a STRING column takes string literals");
                const auto it = std::lower_bound(d.begin(), d.end(), l.name);
                if (it != d.end() && *it == l.name) { args.push_back(Expr::int_literal(it - d.begin())); any_code = true; }
            }
            if (!any_code) args.push_back(Expr::int_literal(-1));
            return Expr::predicate(BK_IN_PREDICATE, BK_FT_IN, "in", std::move(args));
        }
        const bool touches = std::any_of(e.children.begin(), e.children.end(), is_str);
        if (e.node_type == BK_AGG_EXPR && touches) {
            if (e.name != "count" && e.name != "min" && e.name != "max" && e.name != "count_distinct") throw Unsupported(e.name + "() over a STRING column");
            if (e.name == "min" || e.name == "max") enc.result_slots[{e.tuple_id, e.final_slot_id}] = enc.dictionaries[{e.children[0].tuple_id, e.children[0].slot_id}];
        } else if (touches && !(e.node_type == BK_IS_NULL_PREDICATE || (e.node_type == BK_FUNCTION_CALL && e.fn_op == BK_FT_IS_NULL)))
            throw Unsupported("expression '" + e.name + "' over a STRING column is outside the dictionary-coded path");
        Expr out = e;
        out.children.clear();
        for (auto& c : e.children) out.children.push_back(rewrite(c));
        return out;
    };
    std::function<PlanNode(const PlanNode&)> rewrite_node = [&](const PlanNode& n) {
        PlanNode m = n;
        m.children.clear(); m.conjuncts.clear(); m.group_exprs.clear(); m.agg_fn_calls.clear(); m.order_exprs.clear();
        for (auto& c : n.children) m.children.push_back(rewrite_node(c));
        for (auto& e : n.conjuncts) m.conjuncts.push_back(rewrite(e));
        for (auto& e : n.group_exprs) { m.group_exprs.push_back(rewrite(e)); if (is_str(e)) enc.result_slots[{e.tuple_id, e.slot_id}] = enc.dictionaries[{e.tuple_id, e.slot_id}]; }
        for (auto& e : n.agg_fn_calls) m.agg_fn_calls.push_back(rewrite(e));
        for (auto& e : n.order_exprs) m.order_exprs.push_back(rewrite(e));
        if (n.node_type == BK_SORT_NODE || n.node_type == BK_WHERE_FILTER_NODE || n.node_type == BK_TABLE_FILTER_NODE || n.node_type == BK_JOIN_NODE || n.node_type == BK_SCAN_NODE)
            for (auto& kv : strings) enc.result_slots.insert({kv.first, enc.dictionaries[kv.first]});   // fragments that return rows return the scan slots
        return m;
    };
    enc.plan.root = rewrite_node(plan.root);
    enc.plan.tuples = plan.tuples;
    for (auto& t : enc.plan.tuples)
        for (auto& s : t.slots)
            if (s.second == BK_STRING && (strings.count({t.tuple_id, s.first}) || enc.result_slots.count({t.tuple_id, s.first}))) s.second = BK_INT32;

    // ---- columns: string -> rank in its domain's dictionary ----
    for (auto& c : string_cols) {
        const auto& d = *enc.dictionaries[{c.tuple_id, c.slot_id}];
        std::vector<int32_t> codes(c.values.size(), 0);
        bool any_null = false;
        for (size_t r = 0; r < c.values.size(); r++) {
            if (c.values[r]) codes[r] = (int32_t)(std::lower_bound(d.begin(), d.end(), *c.values[r]) - d.begin());
            else any_null = true;
        }
        Column col = Column::from(c.tuple_id, c.slot_id, BK_INT32, codes);
        if (any_null) {
            col.validity.assign((c.values.size() + 7) / 8 + 1, 0);
            for (size_t r = 0; r < c.values.size(); r++) if (c.values[r]) col.validity[r >> 3] |= (uint8_t)(1u << (r & 7));
        }
        enc.columns.push_back(std::move(col));
    }
    return enc;
}

}  // namespace bkgpu
