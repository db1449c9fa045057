#include "thread_pool.h"
namespace hvd {
void ThreadPool::Create(int n) {
  Reset();
  running_ = true;
  for (int i = 0; i < n; ++i) threads_.emplace_back(&ThreadPool::Loop, this);
}
void ThreadPool::Execute(std::function<void()> f) {
  {
    std::lock_guard<std::mutex> l(mu_);
    if (!running_) { f(); return; }
    work_.push(std::move(f));
  }
  cv_.notify_one();
}
void ThreadPool::Reset() {
  {
    std::lock_guard<std::mutex> l(mu_);
    running_ = false;
  }
  cv_.notify_all();
  for (auto& t : threads_) if (t.joinable()) t.join();
  threads_.clear();
}
void ThreadPool::Loop() {
  while (true) {
    std::function<void()> f;
    {
      std::unique_lock<std::mutex> l(mu_);
      cv_.wait(l, [&] { return !running_ || !work_.empty(); });
      if (work_.empty()) { if (!running_) return; continue; }
      f = std::move(work_.front());
      work_.pop();
    }
    f();
  }
}
}  // namespace hvd
