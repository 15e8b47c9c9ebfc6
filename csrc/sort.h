// sort.h — K5 (ORDER BY / top-k) and K1 (filter-only stream compaction): host entry points used by api.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../include/bkgpu.h"
#include "plan.h"

namespace bk {

struct SortState;  // opaque (sort.cu)

struct SortOutCol {
    int tuple_id, slot_id, prim, elem;
    std::vector<uint8_t> values;
    std::vector<uint8_t> validity;  // LSB bitmap, empty = all valid
};

int  sort_open(const Compiled& c, int device, cudaStream_t stream, int64_t region_base, SortState** out, std::string& err);
int  sort_push(SortState* s, const DevCol* cols, int64_t nrows, cudaStream_t stream, bkgpu_stats* stats, std::string& err);
int  sort_finish(SortState* s, void* nccl_comm, int nranks, cudaStream_t stream, bkgpu_stats* stats,
                 std::vector<SortOutCol>& out, int64_t* nrows, std::string& err);
size_t sort_partial_bytes(SortState* s);
int  sort_partial_export(SortState* s, void* dev_dst, cudaStream_t stream, std::string& err);
int  sort_partial_merge(SortState* s, const void* dev_src, int nranks, cudaStream_t stream,
                        std::vector<SortOutCol>& out, int64_t* nrows, std::string& err);
int  sort_reset(SortState* s, cudaStream_t stream, std::string& err);
void sort_close(SortState* s);

}  // namespace bk
