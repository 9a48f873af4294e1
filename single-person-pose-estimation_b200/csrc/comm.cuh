// Communicator handle of the data-parallel path (comm.cu): an NCCL communicator bound at run time.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

struct hgb_comm {
  void* handle = nullptr;   // ncclComm_t
  int nranks = 1, rank = 0;
};

namespace hgb {
// in-place sum over all ranks of `count` floats, asynchronous on `st`
int comm_allreduce_sum_f32(hgb_comm* c, float* buf, int64_t count, cudaStream_t st);
}  // namespace hgb
