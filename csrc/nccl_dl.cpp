// nccl_dl.cpp — see nccl_dl.h.  Only the five entry points the merge step needs are bound.
#include "nccl_dl.h"
#include <dlfcn.h>
#include <nccl.h>
#include <mutex>
#include <string>

namespace bk {
namespace {
std::string g_err;
std::once_flag g_once;
void* g_lib = nullptr;
ncclResult_t (*p_GetUniqueId)(ncclUniqueId*) = nullptr;
ncclResult_t (*p_CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
ncclResult_t (*p_CommCount)(const ncclComm_t, int*) = nullptr;
ncclResult_t (*p_CommUserRank)(const ncclComm_t, int*) = nullptr;
ncclResult_t (*p_CommDestroy)(ncclComm_t) = nullptr;
ncclResult_t (*p_AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
const char* (*p_GetErrorString)(ncclResult_t) = nullptr;
ncclResult_t (*p_Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
ncclResult_t (*p_Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
ncclResult_t (*p_GroupStart)() = nullptr;
ncclResult_t (*p_GroupEnd)() = nullptr;

bool load() {
    std::call_once(g_once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { g_lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_lib) break; }
        if (!g_lib) { g_err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
#define BIND(sym) *(void**)(&p_##sym) = dlsym(g_lib, "nccl" #sym)
        BIND(GetUniqueId); BIND(CommInitRank); BIND(CommCount); BIND(CommUserRank); BIND(CommDestroy); BIND(AllGather); BIND(GetErrorString);
        BIND(Send); BIND(Recv); BIND(GroupStart); BIND(GroupEnd);
#undef BIND
        if (!p_GetUniqueId || !p_CommInitRank || !p_CommCount || !p_CommDestroy || !p_AllGather) { g_err = "libnccl lacks a required symbol"; g_lib = nullptr; }
    });
    return g_lib != nullptr;
}
int check(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return 0;
    g_err = std::string(what) + ": " + (p_GetErrorString ? p_GetErrorString(r) : "nccl error");
    return -1;
}
}  // namespace

const char* nccl_last_error() { return g_err.c_str(); }
int nccl_unique_id(uint8_t id_out[128]) {
    if (!load()) return -1;
    ncclUniqueId id;
    if (check(p_GetUniqueId(&id), "ncclGetUniqueId")) return -1;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    __builtin_memcpy(id_out, &id, 128);
    return 0;
}
int nccl_comm_create(void** comm_out, const uint8_t id[128], int nranks, int rank) {
    if (!load()) return -1;
    ncclUniqueId uid; __builtin_memcpy(&uid, id, 128);
    ncclComm_t c = nullptr;
    if (check(p_CommInitRank(&c, nranks, uid, rank), "ncclCommInitRank")) return -1;
    *comm_out = (void*)c;
    return 0;
}
int nccl_comm_count(void* comm, int* nranks) { if (!load()) return -1; return check(p_CommCount((ncclComm_t)comm, nranks), "ncclCommCount"); }
int nccl_comm_rank(void* comm, int* rank) { if (!load() || !p_CommUserRank) return -1; return check(p_CommUserRank((ncclComm_t)comm, rank), "ncclCommUserRank"); }
void nccl_comm_destroy(void* comm) { if (load() && comm) p_CommDestroy((ncclComm_t)comm); }
int nccl_all_gather(void* comm, const void* send, void* recv, size_t words, cudaStream_t stream) {
    if (!load()) return -1;
    return check(p_AllGather(send, recv, words, ncclUint64, (ncclComm_t)comm, stream), "ncclAllGather");
}
int nccl_all_to_all(void* comm, const void* send, void* recv, size_t words, int nranks, cudaStream_t stream) {
    if (!load()) return -1;
    if (!p_Send || !p_Recv || !p_GroupStart || !p_GroupEnd) { g_err = "libnccl lacks ncclSend/ncclRecv"; return -1; }
    if (check(p_GroupStart(), "ncclGroupStart")) return -1;
    for (int r = 0; r < nranks; r++) {
        if (check(p_Send((const uint64_t*)send + (size_t)r * words, words, ncclUint64, r, (ncclComm_t)comm, stream), "ncclSend")) { p_GroupEnd(); return -1; }
        if (check(p_Recv((uint64_t*)recv + (size_t)r * words, words, ncclUint64, r, (ncclComm_t)comm, stream), "ncclRecv")) { p_GroupEnd(); return -1; }
    }
    return check(p_GroupEnd(), "ncclGroupEnd");
}
}  // namespace bk
