// On-the-fly registration of ORDINARY device allocations (cudaMalloc / the PyTorch caching allocator) for the zero-copy
// in-place allreduce: the owner exports its allocation with cudaIpcGetMemHandle, every peer maps it with
// cudaIpcOpenMemHandle, and from then on `hvd.allreduce_(plain_tensor)` is the pure NVLink phase — no pack into the
// symmetric buffer, no unpack — exactly like a tensor from hvd.symm_empty (minus the multicast alias: NVLS needs VMM
// memory bound to the multicast object, which an IPC mapping of a cudaMalloc block cannot be).
//
// Consistency protocol (no extra round in steady state): a request carries a per-rank key derived from the CUDA buffer
// id and address of its tensor; the response cache treats a changed key as INVALID, so any rank whose tensor moved forces
// a fresh negotiation, and a FRESH response is the (collective) moment at which every rank re-exchanges handles.  A
// response replayed from the cache therefore always finds peer mappings that are still current on every rank.
//
// The reference hands the user pointer to ncclAllReduce (ops/nccl_operations.cc:256-261); NCCL does its own (optional)
// user-buffer registration behind that call.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include "../kernels/p2p_kernels.h"
#include "../transport/transport.h"

namespace hvd {

// key a request carries for a plain device tensor (always <= -2; -1 = no key, >= 0 = registered symmetric region)
int64_t IpcKeyFor(const void* ptr);

class IpcRegistry {
 public:
  ~IpcRegistry() { Clear(); }
  struct Entry {
    bool usable = false;                 // false: some rank could not export / import (VMM-backed allocator, no peer access)
    const void* my_ptr = nullptr;
    void* ptr[kern::kMaxPeers] = {};     // the tensor on every rank, as mapped into this process
  };
  // COLLECTIVE over `t` (all ranks execute the same fresh response): (re)exports my tensor's allocation, imports the
  // peers', stores the result under `name`.
  const Entry& Exchange(Transport* t, const std::string& name, const void* ptr, size_t bytes);
  // Entry recorded by the last Exchange for `name` (nullptr if none): valid for cached responses.
  const Entry* Find(const std::string& name) const;
  void Clear();
  size_t opened_allocations() const { return opened_.size(); }

 private:
  using Handle = std::array<char, 64>;
  struct Opened { void* base = nullptr; int refs = 0; };
  struct Stored { Entry e; std::vector<std::pair<int, Handle>> holds; };
  void Release(Stored& s);
  std::map<std::pair<int, Handle>, Opened> opened_;  // (peer, exported handle) -> mapping of that peer's allocation
  std::unordered_map<std::string, Stored> entries_;
};

}  // namespace hvd
