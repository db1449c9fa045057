// Process sets: arbitrary rank subgroups, each with its own negotiation state
// (queue, response cache, group table, controller) and — on GPU — its own
// peer-mapped symmetric team.  Set 0 is the global set.
// Parity: horovod/common/process_set.{h,cc} (ProcessSet, ProcessSetTable).
#pragma once
#include <algorithm>
#include <map>
#include <memory>
#include <mutex>
#include <vector>
#include "controller.h"
#include "group_table.h"
#include "response_cache.h"
#include "tensor_queue.h"

namespace hvd {

class SymmTeam;
class IpcRegistry;
struct NcclComm;

struct ProcessSet {
  int32_t id = 0;
  std::vector<int> ranks;                 // global ranks, ascending
  std::shared_ptr<Transport> transport;   // nullptr when this process is not a member
  TensorQueue queue;
  ResponseCache cache;
  GroupTable groups;
  std::unique_ptr<Controller> controller;
  // GPU state, created lazily by the first GPU collective on this set
  std::shared_ptr<SymmTeam> team;
  mutable std::mutex team_mu;  // `team` is published by the background thread, read by enqueueing threads
  bool team_tried = false;
  std::shared_ptr<IpcRegistry> ipc;      // peer mappings of ordinary allocations (ops/ipc_registry.h), cycle thread only
  std::shared_ptr<NcclComm> nccl;
  bool nccl_tried = false;
  // multi-host sets: intra-host peer-mapped team + cross-host communicator of the hierarchical GPU allreduce
  std::shared_ptr<Transport> local_transport, cross_transport;
  std::shared_ptr<SymmTeam> local_team;
  bool hier_tried = false;
  bool member() const { return transport != nullptr; }
  int set_rank() const { return transport ? transport->rank() : -1; }
  int set_size() const { return (int)ranks.size(); }
};

class ProcessSetTable {
 public:
  std::shared_ptr<ProcessSet> Get(int32_t id) const {
    std::lock_guard<std::mutex> l(mu_);
    auto it = sets_.find(id);
    return it == sets_.end() ? nullptr : it->second;
  }
  bool Contains(int32_t id) const { std::lock_guard<std::mutex> l(mu_); return sets_.count(id) > 0; }
  // Returns the id of an existing set with exactly these ranks, or -1.
  int32_t Find(const std::vector<int>& ranks) const {
    std::lock_guard<std::mutex> l(mu_);
    for (auto& kv : sets_) if (kv.second->ranks == ranks) return kv.first;
    return -1;
  }
  int32_t Insert(std::shared_ptr<ProcessSet> ps) {
    std::lock_guard<std::mutex> l(mu_);
    int32_t id;
    if (!free_ids_.empty()) { id = *free_ids_.begin(); free_ids_.erase(free_ids_.begin()); } else { id = next_id_++; }
    ps->id = id;
    sets_[id] = std::move(ps);
    return id;
  }
  void Remove(int32_t id) {
    std::lock_guard<std::mutex> l(mu_);
    if (sets_.erase(id) && id != 0) free_ids_.push_back(id);
    std::sort(free_ids_.begin(), free_ids_.end());
  }
  std::vector<int32_t> Ids() const {
    std::lock_guard<std::mutex> l(mu_);
    std::vector<int32_t> v;
    for (auto& kv : sets_) v.push_back(kv.first);
    return v;
  }
  void Clear() { std::lock_guard<std::mutex> l(mu_); sets_.clear(); free_ids_.clear(); next_id_ = 0; }

 private:
  mutable std::mutex mu_;
  std::map<int32_t, std::shared_ptr<ProcessSet>> sets_;
  std::vector<int32_t> free_ids_;
  int32_t next_id_ = 0;
};

}  // namespace hvd
