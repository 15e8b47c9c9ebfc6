// acero_main.cpp — drives host/bkgpu_acero.hpp (the Acero plug-in point) from the command line (tests/test_host_acero.py):
//   bkgpu_acero_host check <plan.bin>                                          registry + factory: does the fragment lower?  (no GPU needed)
//   bkgpu_acero_host exec  <plan.bin> <schema.bin> <rows.bin> <out_schema.bin> <schema.out> <rows.out> [batch_rows]
//        record_batch_source (the wire batch, re-cut into batches of `batch_rows`) -> bkgpu_fragment -> table, IPC in and out
// exit codes: 0 ok, 1 error, 3 = the plan ran into "no CUDA device" (the library has no CPU fallback), 4 = the fragment is refused
#include <arrow/acero/exec_plan.h>
#include <arrow/acero/options.h>
#include <arrow/compute/initialize.h>
#include <arrow/table.h>
#include <cstdio>
#include <fstream>
#include <iterator>
#include "bkgpu_acero.hpp"

namespace ac = arrow::acero;

static std::string slurp(const char* path) { std::ifstream f(path, std::ios::binary); return std::string(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>()); }
static void spit(const char* path, const std::string& s) { std::ofstream f(path, std::ios::binary); f.write(s.data(), (std::streamsize)s.size()); }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s check <plan.bin> | exec <plan.bin> <schema.bin> <rows.bin> <out_schema.bin> <schema.out> <rows.out> [batch_rows]\n", argv[0]); return 2; }
    const std::string mode = argv[1];
    auto st = arrow::compute::Initialize();   // Arrow >= 21 keeps the compute kernels in libarrow_compute
    if (!st.ok()) { fprintf(stderr, "%s\n", st.ToString().c_str()); return 1; }
    st = bkgpu::RegisterAceroNode();
    if (!st.ok()) { fprintf(stderr, "register: %s\n", st.ToString().c_str()); return 1; }
    if (!ac::default_exec_factory_registry()->GetFactory("bkgpu_fragment").ok()) { fprintf(stderr, "factory not found after registration\n"); return 1; }
    if (ac::default_exec_factory_registry()->AddFactory("bkgpu_fragment", bkgpu::BkgpuFragmentNode::Make).ok()) { fprintf(stderr, "a second registration must be refused\n"); return 1; }
    bkgpu::BkgpuFragmentOptions opts;
    opts.plan = slurp(argv[2]);
    std::shared_ptr<arrow::RecordBatch> in;
    std::string wire_schema, wire_rows;
    if (mode == "exec") {
        if (argc < 8) { fprintf(stderr, "exec needs 6 files\n"); return 2; }
        wire_schema = slurp(argv[3]); wire_rows = slurp(argv[4]);
        auto r = bkgpu::from_wire(wire_schema, wire_rows);
        if (!r.ok()) { fprintf(stderr, "read failed: %s\n", r.status().ToString().c_str()); return 1; }
        in = *r;
        const std::string os = slurp(argv[5]);
        arrow::io::BufferReader sr(std::make_shared<arrow::Buffer>((const uint8_t*)os.data(), (int64_t)os.size()));
        auto sch = arrow::ipc::ReadSchema(&sr, nullptr);
        if (!sch.ok()) { fprintf(stderr, "output schema: %s\n", sch.status().ToString().c_str()); return 1; }
        opts.output_schema = *sch;
    } else {   // check: an empty one-column source is enough to reach the factory
        opts.output_schema = arrow::schema({arrow::field("1_1", arrow::int64())});
        auto arr = arrow::MakeArrayOfNull(arrow::int32(), 0).ValueOrDie();
        in = arrow::RecordBatch::Make(arrow::schema({arrow::field("0_1", arrow::int32())}), 0, {arr});
    }
    const int64_t batch_rows = argc > 8 ? std::atoll(argv[8]) : 65536;
    std::vector<std::shared_ptr<arrow::RecordBatch>> batches;
    for (int64_t off = 0; off < in->num_rows() || batches.empty(); off += batch_rows) {
        batches.push_back(in->Slice(off, std::min<int64_t>(batch_rows, in->num_rows() - off)));
        if (in->num_rows() == 0) break;
    }
    ac::Declaration decl = ac::Declaration::Sequence({
        {"record_batch_source", ac::RecordBatchSourceNodeOptions{in->schema(), [batches] { return arrow::MakeVectorIterator(batches); }}},
        {"bkgpu_fragment", opts}});
    if (mode == "check") {   // validate only: the factory runs (and refuses what does not lower) when the plan is built
        auto plan = ac::ExecPlan::Make();
        if (!plan.ok()) { fprintf(stderr, "%s\n", plan.status().ToString().c_str()); return 1; }
        auto node = decl.AddToPlan(plan->get());
        if (!node.ok()) { fprintf(stderr, "%s\n", node.status().ToString().c_str()); return node.status().IsNotImplemented() ? 4 : 1; }
        printf("ok: %s\n", (*node)->kind_name());
        return 0;
    }
    auto table = ac::DeclarationToTable(decl, /*use_threads=*/false);   // the reference runs its Acero plans single-threaded by default (arrow_io_excutor.cpp:266-270)
    if (!table.ok()) {
        const std::string msg = table.status().ToString();
        fprintf(stderr, "%s\n", msg.c_str());
        if (table.status().IsNotImplemented()) return 4;
        return msg.find("no CPU fallback") != std::string::npos || msg.find("CUDA device") != std::string::npos ? 3 : 1;
    }
    auto combined = (*table)->CombineChunksToBatch();
    if (!combined.ok()) { fprintf(stderr, "%s\n", combined.status().ToString().c_str()); return 1; }
    std::string s, d;
    st = bkgpu::to_wire(**combined, &s, &d);
    if (!st.ok()) { fprintf(stderr, "%s\n", st.ToString().c_str()); return 1; }
    spit(argv[6], s); spit(argv[7], d);
    return 0;
}
