// LRU cache of per-tensor negotiation results.  Once every rank has seen the
// response for a named tensor, later submissions with identical parameters only
// need one bit in the per-cycle bit-vector AND instead of a gather/broadcast
// round through the coordinator.
//
// Parity: horovod/common/response_cache.{h,cc} (ResponseCache; the
// CacheCoordinator's bit packing lives in controller.cc here).  All mutating
// operations (Put / Touch / Erase) happen at globally agreed points of the
// response stream, so slot numbers and LRU order are identical on every rank.
#pragma once
#include <list>
#include <string>
#include <unordered_map>
#include <vector>
#include "message.h"

namespace hvd {

class ResponseCache {
 public:
  enum class State { MISS, HIT, INVALID };
  void set_capacity(uint32_t c);
  uint32_t capacity() const { return capacity_; }
  size_t size() const { return by_name_.size(); }
  void clear();

  State Cached(const Request& r) const;
  // bit of a cached name; UINT32_MAX when absent
  uint32_t PeekBit(const std::string& name) const;
  bool HasBit(uint32_t bit) const { return bit < slots_.size() && slots_[bit].used; }
  // Inserts (or refreshes) the single-tensor response. `local` is this rank's
  // own request for the tensor (nullptr when this rank has joined and holds no
  // entry; such slots never HIT).  Returns the evicted bit or UINT32_MAX.
  uint32_t Put(const Response& single, const Request* local);
  const Response& GetResponse(uint32_t bit);  // also marks most-recently-used
  const Response& PeekResponse(uint32_t bit) const { return slots_[bit].response; }
  void Erase(uint32_t bit);

 private:
  struct Slot {
    bool used = false;
    bool params_valid = false;
    Response response;
    // this rank's request parameters
    RequestType type; DataType dtype; std::vector<int64_t> shape; int32_t device; int32_t root_rank;
    double prescale, postscale; ReduceOp op; int64_t symm_key;
    std::list<uint32_t>::iterator lru_it;
  };
  uint32_t capacity_ = 1024;
  std::vector<Slot> slots_;
  std::vector<uint32_t> free_;
  std::list<uint32_t> lru_;  // front = least recently used
  std::unordered_map<std::string, uint32_t> by_name_;
};

}  // namespace hvd
