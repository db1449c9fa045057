#include "tensor_queue.h"
#include <chrono>

namespace hvd {

Status TensorQueue::AddToTensorQueue(std::shared_ptr<TensorTableEntry> e, Request msg) {
  std::lock_guard<std::mutex> l(mu_);
  if (table_.count(e->name)) return Status::InvalidArgument(DuplicateNameError(e->name));
  table_.emplace(e->name, std::move(e));
  queue_.push_back(std::move(msg));
  notified_ = true;
  cv_.notify_one();
  return Status::OK();
}

Status TensorQueue::AddToTensorQueueMulti(std::vector<std::shared_ptr<TensorTableEntry>>& es, std::vector<Request>& msgs) {
  std::lock_guard<std::mutex> l(mu_);
  for (auto& e : es) if (table_.count(e->name)) return Status::InvalidArgument(DuplicateNameError(e->name));
  for (size_t i = 0; i < es.size(); ++i) {
    table_.emplace(es[i]->name, es[i]);
    queue_.push_back(std::move(msgs[i]));
  }
  notified_ = true;
  cv_.notify_one();
  return Status::OK();
}

void TensorQueue::PopMessagesFromQueue(std::deque<Request>& out) {
  std::lock_guard<std::mutex> l(mu_);
  while (!queue_.empty()) { out.push_back(std::move(queue_.front())); queue_.pop_front(); }
  notified_ = false;
}

void TensorQueue::PushMessagesToQueue(std::deque<Request>& msgs) {
  std::lock_guard<std::mutex> l(mu_);
  while (!msgs.empty()) { queue_.push_front(std::move(msgs.back())); msgs.pop_back(); }
}

void TensorQueue::GetTensorEntriesFromResponse(const Response& r, std::vector<std::shared_ptr<TensorTableEntry>>& out) {
  std::lock_guard<std::mutex> l(mu_);
  out.reserve(r.tensor_names.size());
  for (auto& n : r.tensor_names) {
    auto it = table_.find(n);
    if (it == table_.end()) { out.push_back(nullptr); continue; }
    out.push_back(std::move(it->second));
    table_.erase(it);
  }
}

std::shared_ptr<TensorTableEntry> TensorQueue::GetTensorEntry(const std::string& name) const {
  std::lock_guard<std::mutex> l(mu_);
  auto it = table_.find(name);
  return it == table_.end() ? nullptr : it->second;
}

std::shared_ptr<TensorTableEntry> TensorQueue::PopTensorEntry(const std::string& name) {
  std::lock_guard<std::mutex> l(mu_);
  auto it = table_.find(name);
  if (it == table_.end()) return nullptr;
  auto e = std::move(it->second);
  table_.erase(it);
  return e;
}

bool TensorQueue::IsTensorPresent(const std::string& name) const {
  std::lock_guard<std::mutex> l(mu_);
  return table_.count(name) > 0;
}

void TensorQueue::FinalizeTensorQueue(const Status& status) {
  std::unordered_map<std::string, std::shared_ptr<TensorTableEntry>> t;
  {
    std::lock_guard<std::mutex> l(mu_);
    t.swap(table_);
    queue_.clear();
  }
  for (auto& kv : t) {
    if (kv.second && kv.second->callback) { Completion c; c.status = status; kv.second->callback(c); }
  }
}

size_t TensorQueue::size() const { std::lock_guard<std::mutex> l(mu_); return table_.size(); }

bool TensorQueue::WaitForMessages(double timeout_ms) {
  std::unique_lock<std::mutex> l(mu_);
  if (notified_ || !queue_.empty()) return true;
  cv_.wait_for(l, std::chrono::duration<double, std::milli>(timeout_ms), [&] { return notified_ || !queue_.empty(); });
  return notified_ || !queue_.empty();
}

void TensorQueue::Notify() {
  std::lock_guard<std::mutex> l(mu_);
  notified_ = true;
  cv_.notify_one();
}

}  // namespace hvd
