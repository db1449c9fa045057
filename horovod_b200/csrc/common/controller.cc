#include "controller.h"
#include <algorithm>
#include <sstream>
#include "logging.h"

namespace hvd {

namespace {
constexpr uint64_t kStatusShutdown = 1, kStatusUncached = 2, kStatusInvalid = 4, kStatusJoined = 8;

inline void SetBit(std::vector<uint64_t>& w, size_t base, uint32_t bit) { w[base + bit / 64] |= (1ull << (bit % 64)); }
inline bool GetBit(const std::vector<uint64_t>& w, size_t base, uint32_t bit) { return (w[base + bit / 64] >> (bit % 64)) & 1; }

int64_t AlignedBytes(const Response& r) {
  if (r.type == ResponseType::ALLGATHER) return r.payload_bytes;  // per-rank first dims x row bytes (ConstructResponse)
  int64_t total = 0;
  for (auto n : r.tensor_sizes) {
    int64_t b = n * (int64_t)DataTypeSize(r.dtype);
    total += (b + FUSION_ALIGN_BYTES - 1) / FUSION_ALIGN_BYTES * FUSION_ALIGN_BYTES;
  }
  return total;
}

// ALLGATHER and REDUCESCATTER fuse like the reference (controller.cc:915-918, :1003-1086); BROADCAST responses of one root
// fuse as well (the reference never fuses them: hvd.broadcast_parameters is one launch per dtype here, not one per tensor).
bool Fusable(ResponseType t) {
  return t == ResponseType::ALLREDUCE || t == ResponseType::ADASUM || t == ResponseType::ALLGATHER ||
         t == ResponseType::REDUCESCATTER || t == ResponseType::BROADCAST;
}

std::string ShapeStr(const std::vector<int64_t>& s) { return TensorShape(s).DebugString(); }
}  // namespace

Controller::Controller(std::shared_ptr<Transport> transport, TensorQueue* queue, ResponseCache* cache, Timeline* timeline)
    : transport_(std::move(transport)), queue_(queue), cache_(cache), timeline_(timeline) {
  stall_.ConfigureFromEnv();
}

void Controller::SynchronizeParameters(TunableParams* p) { transport_->Bcast(p, sizeof(TunableParams), 0); }

// ---------------------------------------------------------------------------

ResponseList Controller::ComputeResponseList(bool shutdown_requested) {
  std::deque<Request> msgs;
  queue_->PopMessagesFromQueue(msgs);

  std::vector<Request> uncached;
  std::set<uint32_t> invalid_bits;
  std::vector<std::pair<uint32_t, Request>> new_hits;
  const bool use_cache = cache_enabled_ && cache_->capacity() > 0;

  for (auto& r : msgs) {
    if (r.type == RequestType::JOIN) { local_joined_ = true; uncached.push_back(std::move(r)); continue; }
    if (use_cache && Cacheable(r.type)) {
      auto st = cache_->Cached(r);
      if (st == ResponseCache::State::HIT) { new_hits.emplace_back(cache_->PeekBit(r.name), std::move(r)); continue; }
      if (st == ResponseCache::State::INVALID) invalid_bits.insert(cache_->PeekBit(r.name));
    }
    uncached.push_back(std::move(r));
  }
  // A group must travel one way: if any member missed, send the hit members
  // through the coordinator as well (their slots are invalidated everywhere).
  if (!new_hits.empty()) {
    std::set<int32_t> missed_groups;
    for (auto& r : uncached) if (r.group_id >= 0) missed_groups.insert(r.group_id);
    for (auto& h : new_hits) {
      if (h.second.group_id >= 0 && missed_groups.count(h.second.group_id)) {
        invalid_bits.insert(h.first);
        uncached.push_back(std::move(h.second));
      } else {
        if (stall_.enabled()) stall_.RecordCachedTensorStart(h.second.name);
        pending_hits_[h.first] = std::move(h.second);
      }
    }
  }
  // Cached tensors that never became globally ready: push them to the
  // coordinator so the stall inspector can name the missing ranks.
  // Decided ONCE per cycle: the cycle blocks in the bit exchange below, so a second look at the clock further down would
  // usually be the one that sees the interval expire and this block would never run.
  const bool stall_check_now = stall_.enabled() && stall_.ShouldPerformCheck();
  if (stall_check_now && !pending_hits_.empty()) {
    std::vector<std::string> stalled;
    stall_.CollectStalledCachedTensors(&stalled);
    for (auto& n : stalled) {
      LOG(DEBUG) << "cached tensor " << n << " is not ready everywhere after the stall-warning time: renegotiating it through the coordinator";
      uint32_t bit = cache_->PeekBit(n);
      auto it = bit == UINT32_MAX ? pending_hits_.end() : pending_hits_.find(bit);
      if (it != pending_hits_.end()) {
        invalid_bits.insert(bit);
        uncached.push_back(std::move(it->second));
        pending_hits_.erase(it);
      }
      stall_.RemoveCachedTensor(n);
    }
  }

  // ---- one combined AND / OR exchange --------------------------------------
  const uint32_t nbits = use_cache ? cache_->capacity() : 0;
  const size_t nw = (nbits + 63) / 64;
  std::vector<uint64_t> and_words(nw, 0), or_words(1 + 2 * nw, 0);  // [status | invalid bits | real-requester bits]
  for (auto& kv : pending_hits_) { SetBit(and_words, 0, kv.first); SetBit(or_words, 1 + nw, kv.first); }
  if (local_joined_) std::fill(and_words.begin(), and_words.end(), ~0ull);
  for (uint32_t b : invalid_bits) SetBit(or_words, 1, b);
  if (shutdown_requested || stall_shutdown_) or_words[0] |= kStatusShutdown;
  if (!uncached.empty()) or_words[0] |= kStatusUncached;
  if (!invalid_bits.empty()) or_words[0] |= kStatusInvalid;
  if (local_joined_) or_words[0] |= kStatusJoined;
  transport_->AllreduceBits(and_words.data(), (int)nw, or_words.data(), (int)or_words.size());
  const uint64_t status = or_words[0];

  ResponseList result;
  result.shutdown = (status & kStatusShutdown) != 0;

  // ---- globally invalidated slots ------------------------------------------
  if (status & kStatusInvalid) {
    for (uint32_t b = 0; b < nbits; ++b) {
      if (!GetBit(or_words, 1, b)) continue;
      auto it = pending_hits_.find(b);
      if (it != pending_hits_.end()) {  // someone else changed this tensor: my hit is void, renegotiate
        stall_.RemoveCachedTensor(it->second.name);
        uncached.push_back(std::move(it->second));
        pending_hits_.erase(it);
      }
      cache_->Erase(b);
      and_words[b / 64] &= ~(1ull << (b % 64));
    }
  }

  // ---- cache hits common to every rank --------------------------------------
  std::deque<Response> responses;
  for (uint32_t b = 0; b < nbits; ++b) {
    if (!GetBit(and_words, 0, b) || !GetBit(or_words, 1 + nw, b) || !cache_->HasBit(b)) continue;
    Response resp = cache_->GetResponse(b);
    // A joined rank has no tensor in the registered region the cached response points at: while any rank is joined every
    // rank (the OR bit is global) runs the cached response through the fusion-buffer kernel, where the joined rank
    // contributes zeros (the slow path does the same in ConstructResponse).
    if (status & kStatusJoined) resp.symm_key = -1;
    resp.from_cache = true;
    auto it = pending_hits_.find(b);
    if (it != pending_hits_.end()) {
      resp.group_id = it->second.group_id;
      stall_.RemoveCachedTensor(it->second.name);
      pending_hits_.erase(it);
    }
    responses.push_back(std::move(resp));
  }

  // ---- slow path: through the coordinator -----------------------------------
  if (status & kStatusUncached) {
    RequestList mine;
    mine.requests = uncached;
    for (auto& r : uncached) {
      if (r.type != RequestType::JOIN) inflight_uncached_[r.name] = r;
      if (timeline_ && !is_coordinator()) { /* coordinator owns the negotiation rows */ }
    }
    std::vector<std::vector<uint8_t>> all;
    transport_->GatherBytes(mine.Serialize(), &all, 0);
    ResponseList fresh;
    if (is_coordinator()) {
      for (int r = 0; r < size(); ++r) {
        RequestList rl = RequestList::Parse(all[r].data(), all[r].size());
        for (auto& q : rl.requests) CoordinatorHandleRequest(q, r);
      }
      CoordinatorCollectReady(&fresh.responses);
      if (stall_check_now) {
        if (stall_.CheckForStalledTensors(size(), joined_ranks_)) fresh.shutdown = true;
        stall_.UpdateCheckTime();
      }
    }
    std::vector<uint8_t> buf;
    if (is_coordinator()) buf = fresh.Serialize();
    transport_->BcastBytes(&buf, 0);
    if (!is_coordinator()) fresh = ResponseList::Parse(buf.data(), buf.size());
    if (fresh.shutdown) result.shutdown = true;

    for (auto& resp : fresh.responses) {
      if (resp.type == ResponseType::JOIN) { local_joined_ = false; last_joined_rank_ = resp.last_joined_rank; }
      for (auto& name : resp.tensor_names) {
        auto it = inflight_uncached_.find(name);
        const Request* local = it == inflight_uncached_.end() ? nullptr : &it->second;
        if (use_cache && resp.type != ResponseType::ERROR && Cacheable((RequestType)resp.type)) {
          uint32_t evicted = cache_->Put(resp, local);
          if (evicted != UINT32_MAX) {
            auto ph = pending_hits_.find(evicted);
            if (ph != pending_hits_.end()) {  // my pending hit lost its slot: renegotiate next cycle
              std::deque<Request> back{std::move(ph->second)};
              pending_hits_.erase(ph);
              queue_->PushMessagesToQueue(back);
            }
          }
        }
        if (local) { resp.group_id = local->group_id; inflight_uncached_.erase(it); }
      }
      responses.push_back(std::move(resp));
    }
  } else if (stall_check_now) {
    // quiet cycle (nobody has new uncached requests): this is exactly when a stalled tensor sits at the coordinator, so
    // the check must run here as well; a shutdown decision reaches the other ranks through next cycle's status bits
    if (is_coordinator() && stall_.CheckForStalledTensors(size(), joined_ranks_)) stall_shutdown_ = true;
    stall_.UpdateCheckTime();
  }

  auto fused = FuseResponses(std::move(responses), fusion_threshold_, disable_group_fusion_);
  result.responses.assign(std::make_move_iterator(fused.begin()), std::make_move_iterator(fused.end()));
  return result;
}

// ---------------------------------------------------------------------------
// coordinator bookkeeping

void Controller::CoordinatorHandleRequest(const Request& r, int from_rank) {
  if (r.type == RequestType::JOIN) {
    if (std::find(joined_ranks_.begin(), joined_ranks_.end(), from_rank) == joined_ranks_.end()) {
      joined_ranks_.push_back(from_rank);
      last_joined_rank_ = from_rank;
    }
    return;
  }
  auto it = message_table_.find(r.name);
  if (it == message_table_.end()) {
    PendingTensor pt;
    pt.from.assign(size(), false);
    it = message_table_.emplace(r.name, std::move(pt)).first;
    table_order_.push_back(r.name);
    if (timeline_) timeline_->NegotiateStart(r.name, r.type);
    if (stall_.enabled()) stall_.RecordUncachedTensorStart(r.name, from_rank, size());
  } else if (stall_.enabled()) {
    stall_.RecordUncachedTensorRank(r.name, from_rank);
  }
  if (it->second.from[from_rank]) {
    LOG(WARNING) << "rank " << from_rank << " submitted tensor " << r.name << " twice in one negotiation";
    return;
  }
  it->second.from[from_rank] = true;
  it->second.requests.push_back(r);
  if (timeline_) timeline_->NegotiateRankReady(r.name, from_rank);
}

void Controller::CoordinatorCollectReady(std::vector<Response>* out) {
  auto is_joined = [&](int r) { return std::find(joined_ranks_.begin(), joined_ranks_.end(), r) != joined_ranks_.end(); };
  // complete = every rank either submitted or has joined
  std::vector<std::string> complete;
  std::unordered_map<int32_t, int> group_complete;  // group_id -> complete members
  for (auto& name : table_order_) {
    auto& pt = message_table_[name];
    bool ok = true;
    for (int r = 0; r < size(); ++r) if (!pt.from[r] && !is_joined(r)) { ok = false; break; }
    if (!ok) continue;
    complete.push_back(name);
    int32_t g = pt.requests[0].group_id;
    if (g >= 0) group_complete[g]++;
  }
  std::set<std::string> emitted;
  for (auto& name : complete) {
    auto& pt = message_table_[name];
    const Request& q = pt.requests[0];
    if (q.group_id >= 0 && q.group_size > 0 && group_complete[q.group_id] < q.group_size) continue;  // hold the group
    Response resp = ConstructResponse(name, pt.requests, size(), joined_ranks_);
    resp.group_id = q.group_id;
    if (timeline_) timeline_->NegotiateEnd(name);
    stall_.RemoveUncachedTensor(name);
    out->push_back(std::move(resp));
    emitted.insert(name);
  }
  if (!emitted.empty()) {
    for (auto& n : emitted) message_table_.erase(n);
    table_order_.erase(std::remove_if(table_order_.begin(), table_order_.end(),
                                      [&](const std::string& n) { return emitted.count(n) > 0; }),
                       table_order_.end());
  }
  if ((int)joined_ranks_.size() == size()) {
    Response j;
    j.type = ResponseType::JOIN;
    j.tensor_names.push_back(JOIN_TENSOR_NAME);
    j.last_joined_rank = last_joined_rank_;
    out->push_back(std::move(j));
    joined_ranks_.clear();
  }
}

// ---------------------------------------------------------------------------

Response Controller::ConstructResponse(const std::string& name, const std::vector<Request>& requests, int set_size,
                                       const std::vector<int>& joined_ranks) {
  Response resp;
  resp.tensor_names.push_back(name);
  const Request& first = requests[0];
  std::ostringstream err;
  bool error = false;

  for (size_t i = 1; i < requests.size() && !error; ++i) {
    if (requests[i].type != first.type) {
      error = true;
      err << "Mismatched collective operations: One rank did an " << RequestTypeName(first.type)
          << ", but another rank did an " << RequestTypeName(requests[i].type) << ".";
    }
  }
  for (size_t i = 1; i < requests.size() && !error; ++i) {
    if (requests[i].dtype != first.dtype) {
      error = true;
      err << "Mismatched data types: One rank had type " << DataTypeName(first.dtype)
          << ", but another rank had type " << DataTypeName(requests[i].dtype) << ".";
    }
  }
  const RequestType t = first.type;
  if (!error && (t == RequestType::ALLREDUCE || t == RequestType::ADASUM || t == RequestType::BROADCAST ||
                 t == RequestType::REDUCESCATTER)) {
    for (size_t i = 1; i < requests.size() && !error; ++i) {
      if (requests[i].shape != first.shape) {
        error = true;
        err << "Mismatched " << RequestTypeName(t) << " tensor shapes: One rank sent a tensor of shape "
            << ShapeStr(first.shape) << ", but another rank sent a tensor of shape " << ShapeStr(requests[i].shape) << ".";
      }
    }
  }
  if (!error && (t == RequestType::ALLREDUCE || t == RequestType::ADASUM || t == RequestType::REDUCESCATTER)) {
    for (size_t i = 1; i < requests.size() && !error; ++i) {
      if (requests[i].prescale != first.prescale || requests[i].postscale != first.postscale) {
        error = true;
        err << "Mismatched prescale and/or postscale factors: One rank sent factors (" << first.prescale << ", "
            << first.postscale << "), but another rank sent factors (" << requests[i].prescale << ", "
            << requests[i].postscale << ").";
      }
      if (!error && requests[i].reduce_op != first.reduce_op) {
        error = true;
        err << "Mismatched reduce operations: One rank sent reduce op " << ReduceOpName(first.reduce_op)
            << ", but another rank sent reduce op " << ReduceOpName(requests[i].reduce_op) << ".";
      }
    }
  }
  if (!error && (t == RequestType::ALLGATHER || t == RequestType::ALLTOALL)) {
    if (first.shape.empty()) {
      error = true;
      err << "Rank zero tried to " << (t == RequestType::ALLGATHER ? "allgather" : "alltoall")
          << " a rank-zero tensor.";
    }
    for (size_t i = 1; i < requests.size() && !error; ++i) {
      if (requests[i].shape.size() != first.shape.size()) {
        error = true;
        err << "Mismatched " << RequestTypeName(t) << " tensor shapes: One rank sent a tensor of rank "
            << first.shape.size() << ", but another rank sent a tensor of rank " << requests[i].shape.size() << ".";
        break;
      }
      for (size_t d = 1; d < first.shape.size(); ++d) {
        if (requests[i].shape[d] != first.shape[d]) {
          error = true;
          err << "Mismatched " << RequestTypeName(t) << " tensor shapes: One rank sent a tensor with dimension " << d
              << " equal to " << first.shape[d] << ", but another rank sent a tensor with dimension " << d
              << " equal to " << requests[i].shape[d] << ".";
          break;
        }
      }
    }
  }
  if (!error && t == RequestType::REDUCESCATTER && first.shape.empty()) {
    error = true;
    err << "Rank zero tried to reducescatter a rank-zero tensor.";
  }
  if (!error && !joined_ranks.empty() &&
      (t == RequestType::ALLGATHER || t == RequestType::BROADCAST || t == RequestType::ALLTOALL ||
       t == RequestType::REDUCESCATTER)) {
    error = true;
    const char* nm = t == RequestType::ALLGATHER ? "Allgather" : t == RequestType::BROADCAST ? "Broadcast"
                     : t == RequestType::ALLTOALL ? "Alltoall" : "Reducescatter";
    err << nm << " is not supported with Join at this time. Specify sparse_to_dense=True if using DistributedOptimizer";
  }
  if (!error && t == RequestType::BROADCAST) {
    for (size_t i = 1; i < requests.size() && !error; ++i) {
      if (requests[i].root_rank != first.root_rank) {
        error = true;
        err << "Mismatched broadcast root ranks: One rank specified root rank " << first.root_rank
            << ", but another rank specified root rank " << requests[i].root_rank << ".";
      }
    }
  }
  if (!error) {
    bool first_cpu = first.device == CPU_DEVICE_ID;
    for (size_t i = 1; i < requests.size() && !error; ++i) {
      if ((requests[i].device == CPU_DEVICE_ID) != first_cpu) {
        error = true;
        err << "Mismatched CPU/GPU device selection: One rank specified device " << (first_cpu ? "CPU" : "GPU")
            << ", but another rank specified device " << (requests[i].device == CPU_DEVICE_ID ? "CPU" : "GPU") << ".";
      }
    }
  }
  if (error) {
    resp.type = ResponseType::ERROR;
    resp.error_message = err.str();
    return resp;
  }

  resp.type = (ResponseType)t;
  resp.dtype = first.dtype;
  resp.prescale = first.prescale;
  resp.postscale = first.postscale;
  resp.reduce_op = first.reduce_op;
  resp.root_rank = first.root_rank;
  if (t == RequestType::ALLTOALL) {
    // uniform only if NO rank gave explicit splits and every rank sends the same number of rows
    bool uniform = true;
    for (auto& q : requests) if (q.root_rank != kUniformSplits || q.shape.empty() || q.shape[0] != first.shape[0]) uniform = false;
    resp.root_rank = uniform ? kUniformSplits : 0;
  }
  resp.symm_key = first.symm_key;
  for (auto& q : requests) if (q.symm_key != first.symm_key) resp.symm_key = -1;  // zero-copy only if EVERY rank registered it identically
  if (resp.symm_key < 0) {
    // every rank offers an in-place plain allocation it can export over CUDA IPC (keys differ per rank by construction)
    bool all_ipc = true;
    for (auto& q : requests) if (q.symm_key > -2) all_ipc = false;
    resp.symm_key = all_ipc ? -2 : -1;
  }
  if ((int)requests.size() < set_size) resp.symm_key = -1;                         // joined ranks have no registered tensor
  resp.devices.assign(set_size, first.device);
  for (auto& q : requests) if (q.request_rank >= 0 && q.request_rank < set_size) resp.devices[q.request_rank] = q.device;
  if (t == RequestType::ALLGATHER) {
    resp.tensor_sizes.assign(set_size, 0);
    for (auto& q : requests) resp.tensor_sizes[q.request_rank] = q.shape[0];
    int64_t row = (int64_t)DataTypeSize(first.dtype);
    for (size_t d = 1; d < first.shape.size(); ++d) row *= first.shape[d];
    for (auto d0 : resp.tensor_sizes) resp.payload_bytes += (d0 * row + FUSION_ALIGN_BYTES - 1) / FUSION_ALIGN_BYTES * FUSION_ALIGN_BYTES;
  } else if (t == RequestType::PROCESS_SET_ADD || t == RequestType::PROCESS_SET_REMOVE || t == RequestType::SYMM_ALLOC) {
    resp.tensor_sizes = first.shape;
  } else if (t != RequestType::BARRIER && t != RequestType::JOIN) {
    resp.tensor_sizes.push_back(TensorShape(first.shape).num_elements());
  }
  return resp;
}

std::deque<Response> Controller::FuseResponses(std::deque<Response> responses, int64_t threshold,
                                               bool disable_group_fusion) {
  std::deque<Response> out;
  while (!responses.empty()) {
    Response r = std::move(responses.front());
    responses.pop_front();
    if (Fusable(r.type) && threshold > 0 && !responses.empty()) {
      int64_t total = AlignedBytes(r);
      std::deque<Response> skipped;
      while (!responses.empty()) {
        Response& n = responses.front();
        bool compatible = n.type == r.type && n.dtype == r.dtype && n.devices == r.devices &&
                          n.prescale == r.prescale && n.postscale == r.postscale && n.reduce_op == r.reduce_op &&
                          n.root_rank == r.root_rank &&
                          (!disable_group_fusion || n.group_id == r.group_id) && n.symm_key == -1 && r.symm_key == -1;
        int64_t nb = compatible ? AlignedBytes(n) : 0;
        if (compatible && total + nb <= threshold) {
          total += nb;
          for (auto& s : n.tensor_names) r.tensor_names.push_back(std::move(s));
          for (auto s : n.tensor_sizes) r.tensor_sizes.push_back(s);
          r.payload_bytes += n.payload_bytes;
        } else {
          skipped.push_back(std::move(n));  // look-ahead: keep scanning, preserve order of what we skip
        }
        responses.pop_front();
      }
      responses = std::move(skipped);
    }
    out.push_back(std::move(r));
  }
  return out;
}

}  // namespace hvd
