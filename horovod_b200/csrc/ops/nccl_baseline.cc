#include "nccl_baseline.h"
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>
#include <vector>
#include "../common/env.h"
#include "../common/logging.h"

namespace hvd {
namespace {
struct Nccl {
  void* lib = nullptr;
  bool ok = false;
  std::string err;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclCommAbort) CommAbort = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclCommGetAsyncError) CommGetAsyncError = nullptr;
  Nccl() {
    std::string path = EnvStr("HVD_NCCL_LIB", "libnccl.so.2");
    lib = dlopen(path.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("dlopen ") + path + ": " + dlerror(); return; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
    AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    CommAbort = (decltype(CommAbort))dlsym(lib, "ncclCommAbort");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    CommGetAsyncError = (decltype(CommGetAsyncError))dlsym(lib, "ncclCommGetAsyncError");
    ok = GetUniqueId && CommInitRank && AllReduce && CommDestroy && CommAbort && GetErrorString;
    if (!ok) err = "missing NCCL symbols";
  }
};
Nccl& N() { static Nccl n; return n; }

bool MapType(DataType t, ncclDataType_t* o) {
  switch (t) {
    case DataType::UINT8: case DataType::BOOL: *o = ncclUint8; return true;
    case DataType::INT8: *o = ncclInt8; return true;
    case DataType::INT32: *o = ncclInt32; return true;
    case DataType::INT64: *o = ncclInt64; return true;
    case DataType::FLOAT16: *o = ncclFloat16; return true;
    case DataType::FLOAT32: *o = ncclFloat32; return true;
    case DataType::FLOAT64: *o = ncclFloat64; return true;
    case DataType::BFLOAT16: *o = ncclBfloat16; return true;
    default: return false;
  }
}
}  // namespace

NcclComm::~NcclComm() {
  if (comm && N().ok) N().CommDestroy((ncclComm_t)comm);
}

bool NcclSupportsDtype(DataType t) { ncclDataType_t o; return MapType(t, &o); }

bool NcclAvailable(std::string* why) {
  if (!N().ok && why) *why = N().err;
  return N().ok;
}

std::shared_ptr<NcclComm> NcclCreateComm(Transport* t, int device, std::string* why) {
  uint64_t okw = N().ok ? 1 : 0;
  t->AllreduceBits(&okw, 1, nullptr, 0);
  if (!okw) { if (why) *why = "NCCL library not loadable: " + N().err; return nullptr; }
  ncclUniqueId id;
  memset(&id, 0, sizeof id);
  if (t->rank() == 0) N().GetUniqueId(&id);
  t->Bcast(&id, sizeof id, 0);
  cudaSetDevice(device);
  auto c = std::make_shared<NcclComm>();
  c->nranks = t->size(); c->rank = t->rank();
  ncclComm_t comm = nullptr;
  ncclResult_t r = N().CommInitRank(&comm, t->size(), id, t->rank());
  uint64_t good = r == ncclSuccess ? 1 : 0;
  t->AllreduceBits(&good, 1, nullptr, 0);
  if (r == ncclSuccess) c->comm = comm;
  if (!good) { if (why) *why = std::string("ncclCommInitRank failed: ") + N().GetErrorString(r); return nullptr; }
  return c;
}

Status NcclAllReduceCall(NcclComm& c, const void* in, void* out, int64_t count, DataType dtype, ReduceOp op, cudaStream_t s) {
  ncclDataType_t dt;
  if (!MapType(dtype, &dt)) return Status::InvalidArgument(std::string("NCCL baseline does not support dtype ") + DataTypeName(dtype));
  ncclRedOp_t ro = ncclSum;
  if (op == ReduceOp::MIN) ro = ncclMin; else if (op == ReduceOp::MAX) ro = ncclMax; else if (op == ReduceOp::PRODUCT) ro = ncclProd;
  ncclResult_t r = N().AllReduce(in, out, (size_t)count, dt, ro, (ncclComm_t)c.comm, s);
  if (r != ncclSuccess) return Status::UnknownError(std::string("ncclAllReduce failed: ") + N().GetErrorString(r));
  return Status::OK();
}

Status NcclAsyncError(NcclComm& c) {
  if (!N().CommGetAsyncError || !c.comm) return Status::OK();
  ncclResult_t e = ncclSuccess;
  N().CommGetAsyncError((ncclComm_t)c.comm, &e);
  if (e != ncclSuccess && e != ncclInProgress) return Status::UnknownError(std::string("NCCL async error: ") + N().GetErrorString(e));
  return Status::OK();
}

void NcclAbort(NcclComm& c) {
  if (c.comm && N().ok) { N().CommAbort((ncclComm_t)c.comm); c.comm = nullptr; }
}

}  // namespace hvd
