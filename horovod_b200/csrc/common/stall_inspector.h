// Detects tensors that some ranks submitted and others did not.
// The coordinator records which ranks are ready per tensor; after
// HOROVOD_STALL_CHECK_TIME_SECONDS (60) it logs the missing ranks, after
// HOROVOD_STALL_SHUTDOWN_TIME_SECONDS (0 = never) it requests shutdown.
// Locally-cached tensors that never become globally ready are invalidated so
// they re-enter the coordinator path and get reported.
// Parity: horovod/common/stall_inspector.{h,cc}.
#pragma once
#include <chrono>
#include <string>
#include <unordered_map>
#include <vector>

namespace hvd {
class StallInspector {
 public:
  void ConfigureFromEnv();
  bool enabled() const { return enabled_; }
  void set_warning_seconds(double s) { warn_s_ = s; }
  void set_shutdown_seconds(double s) { shutdown_s_ = s; }
  // coordinator side
  void RecordUncachedTensorStart(const std::string& name, int rank, int set_size);
  void RecordUncachedTensorRank(const std::string& name, int rank);
  void RemoveUncachedTensor(const std::string& name);
  // returns true when the job should shut down; `global_ranks` maps set rank -> printable rank
  bool CheckForStalledTensors(int set_size, const std::vector<int>& joined_ranks);
  // every rank
  void RecordCachedTensorStart(const std::string& name);
  void RemoveCachedTensor(const std::string& name);
  void CollectStalledCachedTensors(std::vector<std::string>* names);
  bool ShouldPerformCheck();
  void UpdateCheckTime();

 private:
  using Clock = std::chrono::steady_clock;
  struct Info { Clock::time_point start; std::vector<bool> ready; };
  bool enabled_ = true;
  double warn_s_ = 60.0, shutdown_s_ = 0.0;
  Clock::time_point last_check_ = Clock::now();
  std::unordered_map<std::string, Info> uncached_;
  std::unordered_map<std::string, Clock::time_point> cached_;
};
}  // namespace hvd
