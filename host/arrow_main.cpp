// arrow_main.cpp — the Arrow bridge of host/bkgpu_arrow.hpp from the command line (tests/test_host_arrow.py):
//   bkgpu_arrow_host describe  <schema.bin> <rows.bin>                       views over the wire batch: name prim width rows nulls hash(valid values)
//   bkgpu_arrow_host roundtrip <schema.bin> <rows.bin> <schema.out> <rows.out>   wire -> bkgpu_columns -> record batch -> wire
//   bkgpu_arrow_host exec      <plan.bin> <schema.bin> <rows.bin> <schema.out> <rows.out>   one fragment on cuda:0, IPC in and out
#include <cinttypes>
#include <cstdio>
#include <fstream>
#include <iterator>
#include "bkgpu_arrow.hpp"

static std::string slurp(const char* path) { std::ifstream f(path, std::ios::binary); return std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>()); }
static void spit(const char* path, const std::string& s) { std::ofstream f(path, std::ios::binary); f.write(s.data(), (std::streamsize)s.size()); }

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s describe|roundtrip|exec ...\n", argv[0]); return 2; }
    const std::string mode = argv[1];
    const int base = mode == "exec" ? 3 : 2;
    const std::string wire_schema = slurp(argv[base]), wire_rows = slurp(argv[base + 1]);   // must outlive the batch: the read is zero-copy
    auto in = bkgpu::from_wire(wire_schema, wire_rows);
    if (!in.ok()) { fprintf(stderr, "read failed: %s\n", in.status().ToString().c_str()); return 1; }
    std::shared_ptr<arrow::RecordBatch> rb = *in;
    std::vector<bkgpu_column> cols;
    std::string err = bkgpu::columns_from_record_batch(*rb, [](int, int) { return 0; }, &cols);
    if (!err.empty()) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    if (mode == "describe") {
        for (const auto& c : cols) {
            int64_t nulls = 0; uint64_t h = 1469598103934665603ull;
            for (int64_t r = 0; r < c.length; r++) {
                if (c.validity && !((c.validity[r >> 3] >> (r & 7)) & 1)) { nulls++; continue; }
                const uint8_t* p = (const uint8_t*)c.values + (size_t)r * (size_t)c.elem_size;
                for (int b = 0; b < c.elem_size; b++) { h ^= p[b]; h *= 1099511628211ull; }
            }
            printf("%d_%d prim=%d width=%d rows=%" PRId64 " nulls=%" PRId64 " hash=%016" PRIx64 "\n", c.tuple_id, c.slot_id, c.prim_type, c.elem_size, c.length, nulls, h);
        }
        return 0;
    }
    std::shared_ptr<arrow::RecordBatch> out;
    if (mode == "roundtrip") {
        auto r = bkgpu::record_batch_from_columns(cols.data(), (int)cols.size(), rb->num_rows());
        if (!r.ok()) { fprintf(stderr, "%s\n", r.status().ToString().c_str()); return 1; }
        out = *r;
    } else {   // exec: the decode -> push -> get_next -> encode loop of a GPU store (no CPU fallback: fails without a device)
        const std::string plan = slurp(argv[2]);
        bkgpu_plan* h = nullptr;
        int rc = bkgpu_init(&h, (const uint8_t*)plan.data(), plan.size(), 0, nullptr);
        if (rc == 0) rc = bkgpu_open(h);
        if (rc == 0) rc = bkgpu_push(h, cols.data(), (int)cols.size(), rb->num_rows(), 0);
        if (rc == 0) rc = bkgpu_finish(h);
        if (rc != 0) { fprintf(stderr, "gpu path failed (%d): %s\n", rc, bkgpu_last_error(h)); return 1; }
        std::vector<std::shared_ptr<arrow::RecordBatch>> batches;
        int eos = 0;
        while (!eos) {
            bkgpu_column oc[64]; int n = 64; int64_t nrows = 0;
            if ((rc = bkgpu_get_next(h, oc, &n, &nrows, &eos)) != 0) { fprintf(stderr, "get_next failed: %s\n", bkgpu_last_error(h)); return 1; }
            auto r = bkgpu::record_batch_from_columns(oc, n, nrows);
            if (!r.ok()) { fprintf(stderr, "%s\n", r.status().ToString().c_str()); return 1; }
            batches.push_back(*r);
        }
        bkgpu_close(h);
        auto t = arrow::Table::FromRecordBatches(batches);
        if (!t.ok()) { fprintf(stderr, "%s\n", t.status().ToString().c_str()); return 1; }
        auto combined = (*t)->CombineChunksToBatch();
        if (!combined.ok()) { fprintf(stderr, "%s\n", combined.status().ToString().c_str()); return 1; }
        out = *combined;
    }
    std::string s, d;
    auto st = bkgpu::to_wire(*out, &s, &d);
    if (!st.ok()) { fprintf(stderr, "%s\n", st.ToString().c_str()); return 1; }
    spit(argv[base + 2], s); spit(argv[base + 3], d);
    return 0;
}
