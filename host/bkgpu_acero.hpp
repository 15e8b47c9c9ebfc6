// bkgpu_acero.hpp — the SECOND plug-in point of SURVEY.md §8(b): an Acero exec node.
//
// The reference's vectorized engine builds an arrow::acero plan per fragment and extends Acero through its factory registry:
// `default_exec_factory_registry()->AddFactory(name, Node::Make)` for "index_collector", "limit", "exchange_sender",
// "make_default_agg_row_when_no_input", "debug_print", "topk" (src/exec/arrow_exec_node.cpp:444-477, registered at process start,
// src/store/main.cpp:169-174); custom nodes derive from AceroBaseNode (ProcessBatch / Finish, include/exec/acero_base_node.h:33-78).
// `bkgpu_fragment` is such a node: a SINK-LIKE pipeline breaker (as an aggregate or order-by node is) that pushes every input batch into
// a bkgpu plan and, when its input has finished, emits the fragment's result batches downstream.  A store that executes with
// EXEC_ARROW_ACERO can therefore hand the fused `AGG -> FILTER -> scan` / `SORT` / `JOIN` part of its declaration to the GPU by replacing
// those declarations with ONE {"bkgpu_fragment", BkgpuFragmentOptions{plan bytes, output schema}} declaration over the same source.
//
//   bkgpu::RegisterAceroNode();                                     // once, next to ArrowExecNodeManager::RegisterAllArrowExecNode()
//   acero::Declaration::Sequence({{"record_batch_source", ...}, {"bkgpu_fragment", bkgpu::BkgpuFragmentOptions{...}}})
//
// Fields are named "<tuple>_<slot>" on both sides (include/expr/slot_ref.h:72-82).  There is no CPU fallback: without a CUDA device
// StartProducing() fails with the library's message.  Needs Arrow C++ with Acero (built here against pyarrow's bundled libraries).
#pragma once
#include <arrow/acero/exec_plan.h>
#include <arrow/acero/options.h>
#include <arrow/compute/exec.h>
#include <mutex>
#include "bkgpu_arrow.hpp"

namespace bkgpu {

struct BkgpuFragmentOptions : public arrow::acero::ExecNodeOptions {
    std::string plan;                                  // the fragment in the plan word stream (include/bkgpu_plan.h)
    std::shared_ptr<arrow::Schema> output_schema;      // "<tuple>_<slot>" fields of the fragment's result, types per the Chunk map
    int device = 0;
    std::vector<std::pair<std::string, int64_t>> options;   // bkgpu_set_option pairs
    // slot types the plan declares where they differ from the Arrow default of the field (UINT8/16 arrive as uint32, DATE as uint32 ...)
    std::vector<std::tuple<int, int, int>> declared;        // (tuple, slot, pb::PrimitiveType)
};

class BkgpuFragmentNode : public arrow::acero::ExecNode {
 public:
    BkgpuFragmentNode(arrow::acero::ExecPlan* plan, std::vector<arrow::acero::ExecNode*> inputs, BkgpuFragmentOptions opts)
        : arrow::acero::ExecNode(plan, std::move(inputs), {"input"}, opts.output_schema), opts_(std::move(opts)) {}
    ~BkgpuFragmentNode() override { if (h_) bkgpu_close(h_); }

    static arrow::Result<arrow::acero::ExecNode*> Make(arrow::acero::ExecPlan* plan, std::vector<arrow::acero::ExecNode*> inputs,
                                                      const arrow::acero::ExecNodeOptions& options) {
        if (inputs.size() != 1) return arrow::Status::Invalid("bkgpu_fragment takes one input");
        const auto* o = dynamic_cast<const BkgpuFragmentOptions*>(&options);
        if (!o) return arrow::Status::Invalid("bkgpu_fragment needs BkgpuFragmentOptions");
        if (!o->output_schema) return arrow::Status::Invalid("bkgpu_fragment needs the fragment's output schema");
        // the fragment must lower (parse, type inference, bytecode): checked here, on the host, before any batch moves
        char text[256];
        if (bkgpu_plan_explain((const uint8_t*)o->plan.data(), o->plan.size(), text, sizeof text) < 0)
            return arrow::Status::NotImplemented("bkgpu_fragment: the library does not take this fragment: ", text);
        return plan->EmplaceNode<BkgpuFragmentNode>(plan, std::move(inputs), *o);
    }

    const char* kind_name() const override { return "BkgpuFragmentNode"; }

    arrow::Status StartProducing() override {   // ExecNode::open of the GPU node: init + options + open
        int rc = bkgpu_init(&h_, (const uint8_t*)opts_.plan.data(), opts_.plan.size(), opts_.device, nullptr);
        for (size_t i = 0; rc == 0 && i < opts_.options.size(); i++) rc = bkgpu_set_option(h_, opts_.options[i].first.c_str(), opts_.options[i].second);
        if (rc == 0) rc = bkgpu_open(h_);
        if (rc != 0) return arrow::Status::ExecutionError("bkgpu_fragment: ", bkgpu_last_error(h_), " (", rc, ")");
        return arrow::Status::OK();
    }

    // Acero may deliver batches from several threads; a bkgpu plan is one host thread at a time (one CUDA stream): serialised here
    arrow::Status InputReceived(arrow::acero::ExecNode*, arrow::compute::ExecBatch batch) override {
        std::unique_lock<std::mutex> lock(mu_);
        ARROW_ASSIGN_OR_RAISE(auto rb, batch.ToRecordBatch(inputs_[0]->output_schema()));
        std::vector<bkgpu_column> cols;
        const std::string err = columns_from_record_batch(*rb, [this](int t, int s) {
            for (const auto& d : opts_.declared) if (std::get<0>(d) == t && std::get<1>(d) == s) return std::get<2>(d);
            return 0; }, &cols);
        if (!err.empty()) return arrow::Status::TypeError("bkgpu_fragment: ", err);
        if (rb->num_rows() > 0) {
            const int rc = bkgpu_push(h_, cols.data(), (int)cols.size(), rb->num_rows(), 0);   // host buffers are borrowed for the call only
            if (rc != 0) return arrow::Status::ExecutionError("bkgpu_fragment: ", bkgpu_last_error(h_), " (", rc, ")");
        }
        received_++;
        if (total_ >= 0 && received_ == total_) return Finish(std::move(lock));
        return arrow::Status::OK();
    }

    arrow::Status InputFinished(arrow::acero::ExecNode*, int total_batches) override {
        std::unique_lock<std::mutex> lock(mu_);
        total_ = total_batches;
        if (received_ == total_) return Finish(std::move(lock));
        return arrow::Status::OK();
    }

    void PauseProducing(arrow::acero::ExecNode*, int32_t) override {}    // a pipeline breaker: its output is small and produced at the end
    void ResumeProducing(arrow::acero::ExecNode*, int32_t) override {}

 protected:
    arrow::Status StopProducingImpl() override { if (h_) bkgpu_cancel(h_); return arrow::Status::OK(); }

 private:
    // end of input: bkgpu_finish (merge / collective / finalize), then the result batches go downstream
    arrow::Status Finish(std::unique_lock<std::mutex> lock) {
        if (finished_) return arrow::Status::OK();
        finished_ = true;
        int rc = bkgpu_finish(h_);
        if (rc != 0) return arrow::Status::ExecutionError("bkgpu_fragment: ", bkgpu_last_error(h_), " (", rc, ")");
        std::vector<arrow::compute::ExecBatch> out;
        int eos = 0;
        while (!eos) {
            bkgpu_column oc[64]; int n = 64; int64_t nrows = 0;
            if ((rc = bkgpu_get_next(h_, oc, &n, &nrows, &eos)) != 0) return arrow::Status::ExecutionError("bkgpu_fragment: ", bkgpu_last_error(h_));
            if (nrows == 0 && !out.empty()) continue;
            ARROW_ASSIGN_OR_RAISE(auto rb, record_batch_from_columns(oc, n, nrows));
            // the declared output schema decides the column order; the library's names must all be there with the same types
            std::vector<arrow::Datum> vals;
            for (const auto& f : output_schema()->fields()) {
                auto col = rb->GetColumnByName(f->name());
                if (!col) return arrow::Status::Invalid("bkgpu_fragment: the result has no column '", f->name(), "'");
                if (!col->type()->Equals(*f->type())) return arrow::Status::TypeError("bkgpu_fragment: column '", f->name(), "' is ", col->type()->ToString(), ", declared ", f->type()->ToString());
                vals.emplace_back(col);
            }
            arrow::compute::ExecBatch eb(std::move(vals), nrows);
            eb.index = (int)out.size();
            out.push_back(std::move(eb));
        }
        lock.unlock();
        const int n_out = (int)out.size();
        for (auto& eb : out) ARROW_RETURN_NOT_OK(output_->InputReceived(this, std::move(eb)));
        return output_->InputFinished(this, n_out);
    }

    BkgpuFragmentOptions opts_;
    bkgpu_plan* h_ = nullptr;
    std::mutex mu_;
    int received_ = 0, total_ = -1;
    bool finished_ = false;
};

// call once per process (the reference: ArrowExecNodeManager::RegisterAllArrowExecNode, src/exec/arrow_exec_node.cpp:444-477)
inline arrow::Status RegisterAceroNode() {
    return arrow::acero::default_exec_factory_registry()->AddFactory("bkgpu_fragment", BkgpuFragmentNode::Make);
}

}  // namespace bkgpu
