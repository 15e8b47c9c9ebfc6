// nccl_dl.h — NCCL reached through dlopen("libnccl.so.2") so the library loads (and the CPU-side
// tests run) on machines without NCCL; in a torch process this resolves to the NCCL torch bundles.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace bk {
const char* nccl_last_error();
int nccl_unique_id(uint8_t id_out[128]);
int nccl_comm_create(void** comm_out, const uint8_t id[128], int nranks, int rank);
int nccl_comm_count(void* comm, int* nranks);
int nccl_comm_rank(void* comm, int* rank);
void nccl_comm_destroy(void* comm);
// all-gather `words` 64-bit words per rank
int nccl_all_gather(void* comm, const void* send, void* recv, size_t words, cudaStream_t stream);
// all-to-all: segment r of `send` (words 64-bit words each) goes to rank r; segment r of `recv` arrives from rank r
// (grouped ncclSend / ncclRecv pairs)
int nccl_all_to_all(void* comm, const void* send, void* recv, size_t words, int nranks, cudaStream_t stream);
}  // namespace bk
