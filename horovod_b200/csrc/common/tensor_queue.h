// Thread-safe hand-off between framework threads (enqueue) and the background
// cycle thread: a name-keyed table of in-flight entries plus a FIFO of the
// negotiation requests announcing them.
// Parity: horovod/common/tensor_queue.{h,cc}.
#pragma once
#include <condition_variable>
#include <deque>
#include <mutex>
#include <unordered_map>
#include "common.h"
#include "message.h"

namespace hvd {

class TensorQueue {
 public:
  // Rejects names already in flight (DUPLICATE_NAME error).
  Status AddToTensorQueue(std::shared_ptr<TensorTableEntry> e, Request msg);
  Status AddToTensorQueueMulti(std::vector<std::shared_ptr<TensorTableEntry>>& es, std::vector<Request>& msgs);
  void PopMessagesFromQueue(std::deque<Request>& out);
  void PushMessagesToQueue(std::deque<Request>& msgs);  // re-queue at the front (order preserved)
  // Looks up and removes the entries of a response. Names missing locally
  // (this rank has joined) come back as nullptr.
  void GetTensorEntriesFromResponse(const Response& r, std::vector<std::shared_ptr<TensorTableEntry>>& out);
  std::shared_ptr<TensorTableEntry> GetTensorEntry(const std::string& name) const;
  std::shared_ptr<TensorTableEntry> PopTensorEntry(const std::string& name);
  bool IsTensorPresent(const std::string& name) const;
  // Fails every pending entry with `status` (shutdown / fatal transport error).
  void FinalizeTensorQueue(const Status& status);
  size_t size() const;
  // Wake-up channel for the event-driven cycle loop.
  bool WaitForMessages(double timeout_ms);
  void Notify();

 private:
  mutable std::mutex mu_;
  std::condition_variable cv_;
  std::unordered_map<std::string, std::shared_ptr<TensorTableEntry>> table_;
  std::deque<Request> queue_;
  bool notified_ = false;
};

}  // namespace hvd
