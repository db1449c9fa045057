#include "common.h"
#include <chrono>
#include <sstream>

namespace hvd {

const char* const SHUT_DOWN_ERROR_MSG =
    "Horovod has been shut down. This was caused by an exception on one of the ranks or an attempt to "
    "allreduce, allgather or broadcast a tensor after one of the ranks finished execution. If the shutdown "
    "was caused by an exception, you should see the exception in the log before the first shutdown message.";
const char* const NOT_INITIALIZED_ERROR_MSG = "Horovod has not been initialized; use hvd.init().";

std::string DuplicateNameError(const std::string& name) {
  return "Requested to allreduce, allgather, or broadcast a tensor with the same name as another tensor that "
         "is currently being processed.  If you want to request another tensor, use a different tensor name. "
         "(name: " + name + ")";
}

const char* DataTypeName(DataType t) {
  switch (t) {
    case DataType::UINT8: return "uint8";
    case DataType::INT8: return "int8";
    case DataType::UINT16: return "uint16";
    case DataType::INT16: return "int16";
    case DataType::INT32: return "int32";
    case DataType::INT64: return "int64";
    case DataType::FLOAT16: return "float16";
    case DataType::FLOAT32: return "float32";
    case DataType::FLOAT64: return "float64";
    case DataType::BOOL: return "bool";
    case DataType::BFLOAT16: return "bfloat16";
  }
  return "<unknown>";
}

const char* ReduceOpName(ReduceOp op) {
  switch (op) {
    case ReduceOp::AVERAGE: return "average";
    case ReduceOp::SUM: return "sum";
    case ReduceOp::ADASUM: return "adasum";
    case ReduceOp::MIN: return "min";
    case ReduceOp::MAX: return "max";
    case ReduceOp::PRODUCT: return "product";
  }
  return "<unknown>";
}

const char* RequestTypeName(RequestType t) {
  switch (t) {
    case RequestType::ALLREDUCE: return "ALLREDUCE";
    case RequestType::ALLGATHER: return "ALLGATHER";
    case RequestType::BROADCAST: return "BROADCAST";
    case RequestType::JOIN: return "JOIN";
    case RequestType::ADASUM: return "ADASUM";
    case RequestType::ALLTOALL: return "ALLTOALL";
    case RequestType::BARRIER: return "BARRIER";
    case RequestType::REDUCESCATTER: return "REDUCESCATTER";
    case RequestType::PROCESS_SET_ADD: return "PROCESS_SET_ADD";
    case RequestType::PROCESS_SET_REMOVE: return "PROCESS_SET_REMOVE";
    case RequestType::SYMM_ALLOC: return "SYMM_ALLOC";
  }
  return "<unknown>";
}

std::string TensorShape::DebugString() const {
  std::ostringstream os;
  os << "[";
  for (size_t i = 0; i < dims_.size(); ++i) { if (i) os << ", "; os << dims_[i]; }
  os << "]";
  return os.str();
}

uint64_t NowNs() {
  return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace hvd
