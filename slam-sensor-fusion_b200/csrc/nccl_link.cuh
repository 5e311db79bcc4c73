// nccl_link.cuh -- NCCL behind the C ABI, bound at run time (nccl_link.cu).
#pragma once
#include "common.cuh"

namespace ssf {

struct NcclLink;  // one communicator

// ncclGetUniqueId: 128 bytes that rank 0 creates and every rank passes to nccl_link_create
int nccl_link_unique_id(unsigned char id_out[128]);
int nccl_link_create(const unsigned char id[128], int rank, int world, NcclLink **out);
void nccl_link_destroy(NcclLink *l);
// sum `count` doubles at device pointer buf across the ranks, in place, on `stream` (capturable)
int nccl_link_allreduce(void *link, double *buf, size_t count, cudaStream_t stream);

}  // namespace ssf
