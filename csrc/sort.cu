// sort.cu — placeholder until K5 lands (next commit): every entry point reports "unsupported".
#include "sort.h"
namespace bk {
struct SortState { int dummy; };
int  sort_open(const Compiled&, int, cudaStream_t, int64_t, SortState**, std::string& err) { err = "SORT/FILTER-only plans are not implemented yet"; return BKGPU_EUNSUPPORTED; }
int  sort_push(SortState*, const DevCol*, int64_t, cudaStream_t, bkgpu_stats*, std::string& err) { err = "unsupported"; return BKGPU_EUNSUPPORTED; }
int  sort_finish(SortState*, void*, int, cudaStream_t, bkgpu_stats*, std::vector<SortOutCol>&, int64_t*, std::string& err) { err = "unsupported"; return BKGPU_EUNSUPPORTED; }
size_t sort_partial_bytes(SortState*) { return 0; }
int  sort_partial_export(SortState*, void*, cudaStream_t, std::string& err) { err = "unsupported"; return BKGPU_EUNSUPPORTED; }
int  sort_partial_merge(SortState*, const void*, int, cudaStream_t, std::vector<SortOutCol>&, int64_t*, std::string& err) { err = "unsupported"; return BKGPU_EUNSUPPORTED; }
int  sort_reset(SortState*, cudaStream_t, std::string&) { return 0; }
void sort_close(SortState*) {}
}  // namespace bk
