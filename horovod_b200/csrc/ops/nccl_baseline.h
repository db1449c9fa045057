// NCCL baseline ("the reference's data path, not the product"): communicator
// bootstrap through the controller transport + dlopen'ed NCCL entry points.
// Parity: horovod/common/ops/nccl_operations.cc:87-131 (InitNCCLComm),
// :133-147 (async error check), :56-85 (destroy / abort).
#pragma once
#include <cuda_runtime.h>
#include <memory>
#include <string>
#include "../common/common.h"
#include "../transport/transport.h"

namespace hvd {

struct NcclComm {
  void* comm = nullptr;  // ncclComm_t
  int nranks = 0, rank = 0;
  ~NcclComm();
};

bool NcclAvailable(std::string* why = nullptr);
bool NcclSupportsDtype(DataType t);
// Collective over `t`.
std::shared_ptr<NcclComm> NcclCreateComm(Transport* t, int device, std::string* why);
// dtype/op follow hvd enums; AVERAGE must already be folded into postscale by the caller.
Status NcclAllReduceCall(NcclComm& c, const void* in, void* out, int64_t count, DataType dtype, ReduceOp op, cudaStream_t s);
Status NcclAsyncError(NcclComm& c);
void NcclAbort(NcclComm& c);

}  // namespace hvd
