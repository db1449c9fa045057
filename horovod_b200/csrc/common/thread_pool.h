// Small fixed-size worker pool (GPU completion finalizers).
// Parity: horovod/common/thread_pool.{h,cc}.
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>
namespace hvd {
class ThreadPool {
 public:
  ~ThreadPool() { Reset(); }
  void Create(int n);
  void Execute(std::function<void()> f);
  void Reset();
 private:
  void Loop();
  std::vector<std::thread> threads_;
  std::queue<std::function<void()>> work_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool running_ = false;
};
}  // namespace hvd
