// host_main.cpp — drives the GPU path from C++ the way Region::select drives an ExecNode tree
// (src/store/region.cpp:3069-3216), with plans built like the planner's and tables from the
// counter-based generator of baikaldb_b200/datagen.py restated in C++ (bit-identical, so tests/ can check the
// printed rows against the oracle on the same table).
//
//   bkgpu_host plan    <c1|c2|c3|c5>                       hex of the serialized plan (compared with plan.py's bytes)
//   bkgpu_host explain <c1|c2|c3|c5>                       bkgpu_plan_explain of it (no GPU needed)
//   bkgpu_host run     <c1|c2|c3|c5> <rows> [batch_rows]   executes on cuda:0 and prints one result row per line
//   bkgpu_host chunk   - <rows> [capacity]                   CPU only: rows -> Chunk -> column batches -> rows, checked value by value
//   bkgpu_host strings -  <rows>                             CPU only: two fragments over STRING columns rewritten to dictionary codes
//                                                           (bkgpu_dictionary.hpp): plan bytes + a hash of every code column, compared with dictionary.py's
//   bkgpu_host rows    c2 <rows> [capacity]                 the same table fed ROW by row (MemRow-style values with NULLs every
//                                                           17th key) through Chunk -> column batches -> GPU -> Chunk::to_rows
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "bkgpu_host.hpp"
#include "bkgpu_dictionary.hpp"

using namespace bkgpu;

// ------------------------------------------------------------------ generator (datagen.py)
static const uint64_t GOLDEN = 0x9E3779B97F4A7C15ull, C1 = 0xD1B54A32D192ED03ull, C2 = 0x8CB92BA72F3D8DD7ull;
static uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
static uint64_t stream_key(uint64_t seed, uint64_t column_id, uint64_t k) { return mix64(seed + column_id * C1 + k * C2); }
static uint64_t raw64(uint64_t key, int64_t row) { return mix64(key + (uint64_t)(row + 1) * GOLDEN); }
static double u01(uint64_t r) { return (double)(r >> 11) * (1.0 / 9007199254740992.0); }

static std::vector<int32_t> gen_uniform_i32(uint64_t seed, int cid, int64_t row0, int64_t n, int64_t lo, int64_t hi) {
    std::vector<int32_t> v((size_t)n); uint64_t key = stream_key(seed, cid, 0);
    for (int64_t i = 0; i < n; i++) v[(size_t)i] = (int32_t)(lo + (int64_t)(raw64(key, row0 + i) % (uint64_t)(hi - lo)));
    return v;
}
static std::vector<double> gen_u01(uint64_t seed, int cid, int64_t row0, int64_t n) {
    std::vector<double> v((size_t)n); uint64_t key = stream_key(seed, cid, 0);
    for (int64_t i = 0; i < n; i++) v[(size_t)i] = u01(raw64(key, row0 + i));
    return v;
}
static std::vector<double> gen_normal(uint64_t seed, int cid, int64_t row0, int64_t n, double scale) {
    std::vector<double> v((size_t)n); uint64_t k[4];
    for (int j = 0; j < 4; j++) k[j] = stream_key(seed, cid, j);
    for (int64_t i = 0; i < n; i++) {
        double s = u01(raw64(k[0], row0 + i));
        for (int j = 1; j < 4; j++) s = s + u01(raw64(k[j], row0 + i));
        v[(size_t)i] = (s - 2.0) * scale;
    }
    return v;
}
static std::vector<int64_t> gen_i64_full(uint64_t seed, int cid, int64_t row0, int64_t n) {
    std::vector<int64_t> v((size_t)n); uint64_t key = stream_key(seed, cid, 0);
    for (int64_t i = 0; i < n; i++) v[(size_t)i] = (int64_t)raw64(key, row0 + i);
    return v;
}
static std::vector<int32_t> gen_permutation_i32(uint64_t seed, int cid, int64_t row0, int64_t n, int64_t domain) {
    int bits = 0; while (((int64_t)1 << bits) < domain) bits++;     // bit_length(domain - 1)
    if (bits < 2) bits = 2;
    bits += bits & 1;
    const int half = bits / 2; const uint64_t mask = ((uint64_t)1 << half) - 1;
    uint64_t keys[4]; for (int r = 0; r < 4; r++) keys[r] = stream_key(seed, cid, 16 + r);
    std::vector<int32_t> v((size_t)n);
    for (int64_t i = 0; i < n; i++) {
        uint64_t x = (uint64_t)(row0 + i);
        do {
            uint64_t left = x >> half, right = x & mask;
            for (int r = 0; r < 4; r++) { uint64_t f = mix64(right + keys[r]) & mask; uint64_t nl = right; right = left ^ f; left = nl; }
            x = (left << half) | right;
        } while (x >= (uint64_t)domain);
        v[(size_t)i] = (int32_t)x;
    }
    return v;
}

// ------------------------------------------------------------------ plans (queries.py)
static Expr cmp(int op, const char* name, Expr a, Expr b) { return Expr::fn(op, name, {std::move(a), std::move(b)}); }
static PlanNode scan(int tuple) { PlanNode n; n.node_type = BK_SCAN_NODE; n.tuple_id = tuple; return n; }
static PlanNode where(PlanNode child, Expr c) { PlanNode n; n.node_type = BK_WHERE_FILTER_NODE; n.children.push_back(std::move(child)); n.conjuncts.push_back(std::move(c)); return n; }
static PlanNode agg(PlanNode child, int agg_tuple, std::vector<Expr> groups, std::vector<Expr> fns) {
    PlanNode n; n.node_type = BK_AGG_NODE; n.children.push_back(std::move(child)); n.agg_tuple_id = agg_tuple; n.group_exprs = std::move(groups); n.agg_fn_calls = std::move(fns); return n;
}

static Plan plan_c1() {   // SELECT COUNT(*) FROM t WHERE `0_1` < 2^19
    Plan p; PlanNode packet; packet.node_type = BK_PACKET_NODE;
    packet.children.push_back(agg(where(scan(0), cmp(BK_FT_LT, "lt", Expr::slot_ref(0, 1, BK_INT32), Expr::int_literal(1 << 19))), 1, {}, {Expr::agg("count_star", 1, 1, 1, {})}));
    p.root = std::move(packet);
    p.tuples = {{0, {{1, BK_INT32}}}, {1, {{1, BK_INT64}}}};
    return p;
}
static Plan plan_c2() {   // SELECT `0_1`, COUNT(*), SUM(`0_3`), AVG(`0_4`) FROM t WHERE `0_2` < 2^19 GROUP BY `0_1`
    Plan p;
    p.root = agg(where(scan(0), cmp(BK_FT_LT, "lt", Expr::slot_ref(0, 2, BK_INT32), Expr::int_literal(1 << 19))), 1, {Expr::slot_ref(0, 1, BK_INT32)},
                 {Expr::agg("count_star", 1, 1, 1, {}), Expr::agg("sum", 1, 2, 2, {Expr::slot_ref(0, 3, BK_DOUBLE)}), Expr::agg("avg", 1, 3, 4, {Expr::slot_ref(0, 4, BK_DOUBLE)})});
    p.tuples = {{0, {{1, BK_INT32}, {2, BK_INT32}, {3, BK_DOUBLE}, {4, BK_DOUBLE}}}, {1, {{1, BK_INT64}, {2, BK_DOUBLE}, {3, BK_DOUBLE}, {4, BK_STRING}}}};
    return p;
}
static Plan plan_c3() {   // SELECT `1_2`, COUNT(*), SUM(`0_2`) FROM fact JOIN dim ON `0_1` = `1_1` GROUP BY `1_2`; dim = outer (driver) child
    Plan p; PlanNode j; j.node_type = BK_JOIN_NODE; j.join_type = BK_INNER_JOIN;
    j.children.push_back(scan(1)); j.children.push_back(scan(0));
    j.conjuncts.push_back(cmp(BK_FT_EQ, "eq", Expr::slot_ref(1, 1, BK_INT32), Expr::slot_ref(0, 1, BK_INT32)));
    p.root = agg(std::move(j), 2, {Expr::slot_ref(1, 2, BK_INT32)}, {Expr::agg("count_star", 2, 1, 1, {}), Expr::agg("sum", 2, 2, 2, {Expr::slot_ref(0, 2, BK_DOUBLE)})});
    p.tuples = {{0, {{1, BK_INT32}, {2, BK_DOUBLE}}}, {1, {{1, BK_INT32}, {2, BK_INT32}}}, {2, {{1, BK_INT64}, {2, BK_DOUBLE}}}};
    return p;
}
static Plan plan_c5() {   // SELECT `0_1`, `0_2` FROM t ORDER BY `0_1` ASC LIMIT 1000
    Plan p; PlanNode s; s.node_type = BK_SORT_NODE; s.tuple_id = 0; s.limit = 1000; s.children.push_back(scan(0));
    s.order_exprs.push_back(Expr::slot_ref(0, 1, BK_INT64)); s.is_asc.push_back(true); s.is_null_first.push_back(true);
    p.root = std::move(s);
    p.tuples = {{0, {{1, BK_INT64}, {2, BK_INT32}}}};
    return p;
}

// ------------------------------------------------------------------ scans over the synthetic tables, in batches (regions)
static std::unique_ptr<ExecNode> scan_of(const std::string& cfg, int tuple, int64_t rows, int64_t batch_rows) {
    std::vector<RowBatch> batches;
    for (int64_t r0 = 0; r0 < rows; r0 += batch_rows) {
        int64_t n = rows - r0 < batch_rows ? rows - r0 : batch_rows; RowBatch b;
        if (cfg == "c1") b.columns = {Column::from(0, 1, BK_INT32, gen_uniform_i32(1, 1, r0, n, 0, 1 << 20))};
        else if (cfg == "c2") b.columns = {Column::from(0, 1, BK_INT32, gen_uniform_i32(2, 1, r0, n, 0, 1000)), Column::from(0, 2, BK_INT32, gen_uniform_i32(2, 2, r0, n, 0, 1 << 20)),
                                           Column::from(0, 3, BK_DOUBLE, gen_u01(2, 3, r0, n)), Column::from(0, 4, BK_DOUBLE, gen_normal(2, 4, r0, n, 1000.0 * 1.7320508075688772))};
        else if (cfg == "c5") b.columns = {Column::from(0, 1, BK_INT64, gen_i64_full(5, 1, r0, n)), Column::from(0, 2, BK_INT32, gen_uniform_i32(5, 2, r0, n, 0, 1 << 30))};
        else if (cfg == "c3" && tuple == 0) b.columns = {Column::from(0, 1, BK_INT32, gen_uniform_i32(3, 1, r0, n, 0, rows / 10)), Column::from(0, 2, BK_DOUBLE, gen_u01(3, 2, r0, n))};
        else b.columns = {Column::from(1, 1, BK_INT32, gen_permutation_i32(3, 11, r0, n, rows)), Column::from(1, 2, BK_INT32, gen_uniform_i32(3, 12, r0, n, 0, 1000))};
        batches.push_back(std::move(b));
    }
    return std::unique_ptr<ExecNode>(new ColumnScanNode(std::move(batches)));
}

static void print_value(const Column& c, int64_t i) {
    if (c.is_null(i)) { printf("NULL"); return; }
    switch (c.prim_type) {
        case BK_INT32: printf("%d", c.at<int32_t>(i)); break;
        case BK_INT64: printf("%" PRId64, c.at<int64_t>(i)); break;
        case BK_UINT64: printf("%" PRIu64, c.at<uint64_t>(i)); break;
        case BK_DOUBLE: printf("%.17g", c.at<double>(i)); break;
        case BK_STRING: if (c.elem_size == 16) { printf("blob(%.17g,%" PRId64 ")", c.at<double>(2 * i), c.at<int64_t>(2 * i + 1)); break; }   // AVG intermediate {sum,count}
            /* fallthrough */
        default: printf("?"); break;
    }
}

// ------------------------------------------------------------------ strings (tests/test_host_cpp.py builds the same data and plans in Python)
static std::vector<std::optional<std::string>> gen_strings(uint64_t seed, int64_t n, int domain, int null_every) {
    std::vector<std::optional<std::string>> v((size_t)n);
    uint64_t x = seed;
    for (int64_t i = 0; i < n; i++) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        const int idx = (int)((x >> 33) % (uint64_t)domain);
        if (null_every && (x >> 20) % (uint64_t)null_every == 0) continue;
        v[(size_t)i] = "s" + std::to_string((idx * 7) % domain);
    }
    return v;
}
static void print_encoded(const EncodedStrings& enc) {
    for (uint8_t b : enc.plan.serialize()) printf("%02x", b);
    printf("\n");
    for (const auto& c : enc.columns) {
        uint64_t h = 1469598103934665603ull; int64_t nulls = 0;
        for (int64_t r = 0; r < c.length; r++) {
            if (c.is_null(r)) { nulls++; continue; }
            const uint32_t code = (uint32_t)c.at<int32_t>(r);
            for (int b = 0; b < 4; b++) { h ^= (code >> (8 * b)) & 0xFF; h *= 1099511628211ull; }
        }
        printf("%d_%d rows=%" PRId64 " nulls=%" PRId64 " hash=%016" PRIx64 " dict=%zu\n", c.tuple_id, c.slot_id, c.length, nulls, h, enc.dictionaries.at({c.tuple_id, c.slot_id})->size());
    }
}
static int strings_mode(int64_t rows) {
    auto S = [](int t, int s) { return Expr::slot_ref(t, s, BK_STRING); };
    {   // A: SELECT `0_1`, COUNT(*), MIN(`0_2`), MAX(`0_2`), COUNT(`0_2`) WHERE `0_3` >= 's2' AND 's30' > `0_3` AND `0_2` != 'zzz' AND `0_1` IN ('s1', 's5', 'nope') AND `0_2` LIKE 's1%' AND `0_3` LIKE '%2_' GROUP BY `0_1`
        Plan p;
        PlanNode f = where(scan(0), cmp(BK_FT_GE, "ge", S(0, 3), Expr::string_literal("s2")));
        f.conjuncts.push_back(cmp(BK_FT_GT, "gt", Expr::string_literal("s30"), S(0, 3)));
        f.conjuncts.push_back(cmp(BK_FT_NE, "ne", S(0, 2), Expr::string_literal("zzz")));
        f.conjuncts.push_back(Expr::predicate(BK_IN_PREDICATE, BK_FT_IN, "in", {S(0, 1), Expr::string_literal("s1"), Expr::string_literal("s5"), Expr::string_literal("nope")}));
        f.conjuncts.push_back(Expr::predicate(BK_LIKE_PREDICATE, BK_FT_LIKE, "like", {S(0, 2), Expr::string_literal("s1%")}));
        f.conjuncts.push_back(Expr::predicate(BK_LIKE_PREDICATE, BK_FT_LIKE, "like", {S(0, 3), Expr::string_literal("%2_")}));
        p.root = agg(std::move(f), 1, {S(0, 1)}, {Expr::agg("count_star", 1, 1, 1, {}), Expr::agg("min", 1, 2, 2, {S(0, 2)}), Expr::agg("max", 1, 3, 3, {S(0, 2)}), Expr::agg("count", 1, 4, 4, {S(0, 2)})});
        p.tuples = {{0, {{1, BK_STRING}, {2, BK_STRING}, {3, BK_STRING}}}, {1, {{1, BK_INT64}, {2, BK_STRING}, {3, BK_STRING}, {4, BK_INT64}}}};
        std::vector<StringColumn> cols = {{0, 1, gen_strings(11, rows, 37, 0)}, {0, 2, gen_strings(12, rows, 23, 9)}, {0, 3, gen_strings(13, rows, 41, 0)}};
        print_encoded(encode_strings(p, cols));
    }
    {   // B: SELECT `1_2`, COUNT(*) FROM fact JOIN dim ON `1_1` = `0_1` GROUP BY `1_2`  (string join keys: one dictionary for both sides)
        Plan p; PlanNode j; j.node_type = BK_JOIN_NODE; j.join_type = BK_INNER_JOIN;
        j.children.push_back(scan(1)); j.children.push_back(scan(0));
        j.conjuncts.push_back(cmp(BK_FT_EQ, "eq", S(1, 1), S(0, 1)));
        p.root = agg(std::move(j), 2, {Expr::slot_ref(1, 2, BK_INT32)}, {Expr::agg("count_star", 2, 1, 1, {})});
        p.tuples = {{0, {{1, BK_STRING}}}, {1, {{1, BK_STRING}, {2, BK_INT32}}}, {2, {{1, BK_INT64}}}};
        std::vector<StringColumn> cols = {{0, 1, gen_strings(21, rows, 53, 13)}, {1, 1, gen_strings(22, rows / 4 + 1, 61, 0)}};
        print_encoded(encode_strings(p, cols));
    }
    try {   // refused: SUM over a string
        Plan p; p.root = agg(scan(0), 1, {}, {Expr::agg("sum", 1, 1, 1, {S(0, 1)})}); p.tuples = {{0, {{1, BK_STRING}}}, {1, {{1, BK_DOUBLE}}}};
        encode_strings(p, {{0, 1, gen_strings(1, 4, 3, 0)}});
        printf("NOT REFUSED\n"); return 1;
    } catch (const Unsupported& e) { printf("refused: %s\n", e.what()); }
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s plan|explain|run c1|c2|c3|c5 [rows] [batch_rows]\n", argv[0]); return 2; }
    std::string mode = argv[1], cfg = argv[2];
    if (mode == "strings") return strings_mode(argc > 3 ? atoll(argv[3]) : 1000);
    if (mode == "like") {   // bkgpu_host like - <hex target> <hex pattern> <utf8 0|1> ... (triples): the matcher of bkgpu_dictionary.hpp, one result per line
        auto unhex = [](const char* h) { std::string o; for (size_t i = 0; h[i] && h[i + 1]; i += 2) o.push_back((char)strtol(std::string(h + i, 2).c_str(), nullptr, 16)); return o; };
        for (int i = 3; i + 2 < argc; i += 3) printf("%d\n", like_match(unhex(argv[i]), unhex(argv[i + 1]), atoi(argv[i + 2]) != 0));
        return 0;
    }
    Plan plan = cfg == "c1" ? plan_c1() : cfg == "c2" ? plan_c2() : cfg == "c3" ? plan_c3() : plan_c5();
    if (mode == "chunk") {   // CPU-only: rows -> Chunk -> column batches of `capacity` rows -> rows again (f1 adapter round trip)
        const int64_t rows = argc > 3 ? atoll(argv[3]) : 1000, capacity = argc > 4 ? atoll(argv[4]) : 64;
        std::vector<int32_t> k = gen_uniform_i32(2, 1, 0, rows, 0, 1000); std::vector<double> a = gen_u01(2, 3, 0, rows); std::vector<int64_t> w = gen_i64_full(5, 1, 0, rows);
        std::vector<MemRowValues> mem_rows((size_t)rows);
        for (int64_t i = 0; i < rows; i++)
            mem_rows[(size_t)i] = {k[(size_t)i] % 7 == 0 ? Value() : Value::of_int(k[(size_t)i]), i % 5 == 0 ? Value() : Value::of_double(a[(size_t)i]), Value::of_int(w[(size_t)i]),
                                   Value::of_uint((uint64_t)w[(size_t)i] >> 40), Value::of_double((double)(float)a[(size_t)i]), Value::of_int(i & 1)};
        RowScanNode scan({{0, 1, BK_INT32}, {0, 2, BK_DOUBLE}, {0, 3, BK_INT64}, {0, 4, BK_UINT32}, {0, 5, BK_FLOAT}, {0, 6, BK_BOOL}}, mem_rows, capacity);
        RuntimeState st; bool eos = false; int64_t seen = 0, batches = 0, mismatches = 0;
        while (!eos) {
            RowBatch b;
            if (scan.get_next(&st, &b, &eos) < 0) return 1;
            if (b.columns.empty()) continue;
            batches++;
            if (b.size() > capacity) mismatches++;
            std::vector<MemRowValues> back = Chunk::to_rows(b);
            for (auto& r : back) {
                const MemRowValues& o = mem_rows[(size_t)seen++];
                for (size_t c = 0; c < r.size(); c++) {
                    bool same = r[c].is_null == o[c].is_null;
                    if (same && !o[c].is_null) same = (c == 1 || c == 4) ? r[c].f64 == o[c].f64 : (c == 3 ? r[c].u64 == o[c].u64 : r[c].i64 == o[c].i64);
                    if (!same) mismatches++;
                }
            }
        }
        printf("rows=%" PRId64 " batches=%" PRId64 " mismatches=%" PRId64 "\n", seen, batches, mismatches);
        return mismatches == 0 && seen == rows ? 0 : 1;
    }
    if (mode == "plan") { for (uint8_t b : plan.serialize()) printf("%02x", b); printf("\n"); return 0; }
    if (mode == "explain") {
        std::vector<uint8_t> d = plan.serialize(); std::vector<char> text(1 << 16);
        int rc = bkgpu_plan_explain(d.data(), d.size(), text.data(), text.size());
        if (rc != 0) { fprintf(stderr, "explain failed (%d): %s\n", rc, bkgpu_last_error(nullptr)); return 1; }
        fputs(text.data(), stdout); return 0;
    }
    int64_t rows = argc > 3 ? atoll(argv[3]) : 1 << 20, batch_rows = argc > 4 ? atoll(argv[4]) : rows;
    RuntimeState state; state.row_batch_capacity = 4096;
    GpuExecNode root;
    if (root.init(plan) < 0) return 1;
    if (mode == "rows") {   // f1: a row-engine child below the GPU subtree
        std::vector<int32_t> k = gen_uniform_i32(2, 1, 0, rows, 0, 1000), f = gen_uniform_i32(2, 2, 0, rows, 0, 1 << 20);
        std::vector<double> a = gen_u01(2, 3, 0, rows), b = gen_normal(2, 4, 0, rows, 1000.0 * 1.7320508075688772);
        std::vector<MemRowValues> mem_rows((size_t)rows);
        for (int64_t i = 0; i < rows; i++) {
            MemRowValues& r = mem_rows[(size_t)i];
            r = {k[(size_t)i] % 17 == 0 ? Value() : Value::of_int(k[(size_t)i]), Value::of_int(f[(size_t)i]), Value::of_double(a[(size_t)i]), Value::of_double(b[(size_t)i])};
        }
        root.add_child(std::unique_ptr<ExecNode>(new RowScanNode({{0, 1, BK_INT32}, {0, 2, BK_INT32}, {0, 3, BK_DOUBLE}, {0, 4, BK_DOUBLE}}, std::move(mem_rows), argc > 4 ? atoll(argv[4]) : 1024)));
        if (root.open(&state) < 0) { fprintf(stderr, "open failed (%d): %s\n", state.error_code, state.error_msg.c_str()); return 1; }
        bool eos = false; int64_t total = 0;
        while (!eos) {
            RowBatch batch;
            if (root.get_next(&state, &batch, &eos) < 0) { fprintf(stderr, "get_next failed (%d): %s\n", state.error_code, state.error_msg.c_str()); return 1; }
            std::vector<MemRowValues> out = Chunk::to_rows(batch);   // back to rows for a row-engine parent
            for (auto& r : out) {
                for (size_t c = 0; c < r.size(); c++) {
                    const Column& col = batch.columns[c];
                    printf("%s%d_%d=", c ? " " : "", col.tuple_id, col.slot_id);
                    if (r[c].is_null) printf("NULL");
                    else if (col.prim_type == BK_DOUBLE) printf("%.17g", r[c].f64);
                    else if (col.prim_type == BK_STRING) printf("blob(%.17g,%" PRId64 ")", col.at<double>(2 * (int64_t)(&r - &out[0])), col.at<int64_t>(2 * (int64_t)(&r - &out[0]) + 1));
                    else printf("%" PRId64, r[c].i64);
                }
                printf("\n");
            }
            total += batch.size();
        }
        root.close(&state);
        fprintf(stderr, "rows_returned=%" PRId64 " scan_rows=%" PRId64 " filter_rows=%" PRId64 "\n", total, state.num_scan_rows, state.num_filter_rows);
        return 0;
    }
    if (cfg == "c3") { root.add_child(scan_of(cfg, 1, rows / 10, batch_rows)); root.add_child(scan_of(cfg, 0, rows, batch_rows)); }   // driver (dim) first
    else root.add_child(scan_of(cfg, 0, rows, batch_rows));
    if (root.open(&state) < 0) { fprintf(stderr, "open failed (%d): %s\n", state.error_code, state.error_msg.c_str()); return 1; }
    bool eos = false; int64_t total = 0;
    while (!eos) {
        RowBatch batch;
        if (root.get_next(&state, &batch, &eos) < 0) { fprintf(stderr, "get_next failed (%d): %s\n", state.error_code, state.error_msg.c_str()); return 1; }
        for (int64_t i = 0; i < batch.size(); i++) {
            for (size_t c = 0; c < batch.columns.size(); c++) { if (c) printf(" "); printf("%d_%d=", batch.columns[c].tuple_id, batch.columns[c].slot_id); print_value(batch.columns[c], i); }
            printf("\n");
        }
        total += batch.size();
    }
    root.close(&state);
    fprintf(stderr, "rows_returned=%" PRId64 " scan_rows=%" PRId64 " filter_rows=%" PRId64 "\n", total, state.num_scan_rows, state.num_filter_rows);
    return 0;
}
