// Negotiation: decides, once per cycle, which named tensors are ready on every
// rank of a process set and in which (globally identical) order and grouping
// they execute.
//
// Capability parity with horovod/common/controller.{h,cc}:
//   * cache fast path: one bit-vector AND/OR over the transport, no coordinator
//     round when every queued tensor is a response-cache hit (controller.cc:209-252)
//   * slow path: gather uncached requests at the coordinator (set rank 0), count
//     until every non-joined rank submitted, validate (ConstructResponse,
//     controller.cc:496-843), broadcast, insert into every rank's cache in the
//     same order (controller.cc:456-471)
//   * tensor fusion with look-ahead (FuseResponses, controller.cc:901-1091)
//   * join / barrier / grouped ops / stall inspection
// Differences by design: responses travel unfused and every rank runs the
// (deterministic) fusion planner locally; the AND and OR sections are exchanged
// in ONE transport call; "ready" additionally requires one real (non-joined)
// requester so a fully-joined set never fabricates work; group completeness is
// carried in the request (group_size) instead of being looked up in the
// coordinator's own group table.
#pragma once
#include <deque>
#include <map>
#include <set>
#include <unordered_map>
#include "../transport/transport.h"
#include "common.h"
#include "group_table.h"
#include "message.h"
#include "response_cache.h"
#include "stall_inspector.h"
#include "tensor_queue.h"
#include "timeline.h"

namespace hvd {

struct TunableParams {
  int64_t fusion_threshold_bytes = 128ll << 20;
  double cycle_time_ms = 1.0;
  bool cache_enabled = true;
  int64_t oneshot_max_bytes = 512 << 10;
  int64_t nvls_min_bytes = 1 << 20;
  int32_t comm_ctas = 128;
  uint8_t active = 0;  // autotune still running
};

class Controller {
 public:
  Controller(std::shared_ptr<Transport> transport, TensorQueue* queue, ResponseCache* cache, Timeline* timeline);

  int rank() const { return transport_->rank(); }
  int size() const { return transport_->size(); }
  bool is_coordinator() const { return rank() == 0; }
  Transport* transport() { return transport_.get(); }

  // One negotiation round.  `shutdown_requested`: this rank wants to stop.
  ResponseList ComputeResponseList(bool shutdown_requested);

  // Greedy fusion with look-ahead; public for unit tests.
  static std::deque<Response> FuseResponses(std::deque<Response> responses, int64_t threshold_bytes,
                                            bool disable_group_fusion);
  // Cross-rank validation of one tensor's requests; public for unit tests.
  static Response ConstructResponse(const std::string& name, const std::vector<Request>& requests, int set_size,
                                    const std::vector<int>& joined_ranks);

  void SynchronizeParameters(TunableParams* p);  // coordinator -> all

  void set_fusion_threshold(int64_t b) { fusion_threshold_ = b; }
  void set_cache_enabled(bool e) { cache_enabled_ = e; }
  void set_disable_group_fusion(bool d) { disable_group_fusion_ = d; }
  bool local_joined() const { return local_joined_; }
  StallInspector& stall_inspector() { return stall_; }
  int32_t last_joined_rank() const { return last_joined_rank_; }

 private:
  static bool Cacheable(RequestType t) {
    return t == RequestType::ALLREDUCE || t == RequestType::ADASUM || t == RequestType::ALLGATHER ||
           t == RequestType::BROADCAST || t == RequestType::ALLTOALL || t == RequestType::REDUCESCATTER;
  }
  void CoordinatorHandleRequest(const Request& r, int from_rank);
  void CoordinatorCollectReady(std::vector<Response>* out);

  std::shared_ptr<Transport> transport_;
  TensorQueue* queue_;
  ResponseCache* cache_;
  Timeline* timeline_;
  StallInspector stall_;
  bool stall_shutdown_ = false;  // coordinator: the stall inspector asked for a job-wide shutdown

  int64_t fusion_threshold_ = 128ll << 20;
  bool cache_enabled_ = true;
  bool disable_group_fusion_ = false;
  bool local_joined_ = false;
  int32_t last_joined_rank_ = -1;

  // every rank
  std::map<uint32_t, Request> pending_hits_;                      // locally hit, waiting for the global AND
  std::unordered_map<std::string, Request> inflight_uncached_;    // sent to the coordinator, no response yet
  // coordinator only
  struct PendingTensor { std::vector<Request> requests; std::vector<bool> from; };
  std::unordered_map<std::string, PendingTensor> message_table_;
  std::vector<std::string> table_order_;
  std::vector<int> joined_ranks_;
};

}  // namespace hvd
