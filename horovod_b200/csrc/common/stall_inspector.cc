#include "stall_inspector.h"
#include <sstream>
#include "common.h"
#include "env.h"
#include "logging.h"

namespace hvd {

void StallInspector::ConfigureFromEnv() {
  enabled_ = !EnvBool(HOROVOD_STALL_CHECK_DISABLE, false);
  warn_s_ = EnvDouble(HOROVOD_STALL_CHECK_TIME_SECONDS, 60.0);
  shutdown_s_ = EnvDouble(HOROVOD_STALL_SHUTDOWN_TIME_SECONDS, 0.0);
  if (shutdown_s_ > 0 && shutdown_s_ < warn_s_) {
    LOG(WARNING) << "HOROVOD_STALL_SHUTDOWN_TIME_SECONDS is less than HOROVOD_STALL_CHECK_TIME_SECONDS; "
                    "stall warnings will not be logged before shutdown.";
  }
}

void StallInspector::RecordUncachedTensorStart(const std::string& name, int rank, int set_size) {
  auto& info = uncached_[name];
  info.start = Clock::now();
  info.ready.assign(set_size, false);
  if (rank >= 0 && rank < set_size) info.ready[rank] = true;
}
void StallInspector::RecordUncachedTensorRank(const std::string& name, int rank) {
  auto it = uncached_.find(name);
  if (it != uncached_.end() && rank >= 0 && rank < (int)it->second.ready.size()) it->second.ready[rank] = true;
}
void StallInspector::RemoveUncachedTensor(const std::string& name) { uncached_.erase(name); }

bool StallInspector::CheckForStalledTensors(int set_size, const std::vector<int>& joined_ranks) {
  bool shutdown = false, header = false;
  auto now = Clock::now();
  std::ostringstream msg;
  for (auto& kv : uncached_) {
    double age = std::chrono::duration<double>(now - kv.second.start).count();
    if (age < warn_s_) continue;
    if (!header) {
      msg << "One or more tensors were submitted to be reduced, gathered or broadcasted by subset of ranks and "
             "are waiting for remainder of ranks for more than " << (int)warn_s_ << " seconds. This may indicate "
             "that different ranks are trying to submit different tensors or that only subset of ranks is "
             "submitting tensors, which will cause deadlock. \nMissing ranks:";
      header = true;
    }
    std::ostringstream missing;
    bool first = true;
    for (int r = 0; r < set_size; ++r) {
      bool joined = false;
      for (int j : joined_ranks) if (j == r) joined = true;
      if (r < (int)kv.second.ready.size() && !kv.second.ready[r] && !joined) { missing << (first ? "" : ", ") << r; first = false; }
    }
    msg << "\n" << kv.first << ": [" << missing.str() << "]";
    if (shutdown_s_ > 0 && age > shutdown_s_) shutdown = true;
  }
  if (header) LOG(WARNING) << msg.str();
  if (shutdown) LOG(ERROR) << "One or more rank (marked by \"!\") is stalled for longer than " << (int)shutdown_s_
                           << " seconds. Will shutdown.";
  return shutdown;
}

void StallInspector::RecordCachedTensorStart(const std::string& name) {
  if (!cached_.count(name)) cached_[name] = Clock::now();
}
void StallInspector::RemoveCachedTensor(const std::string& name) { cached_.erase(name); }
void StallInspector::CollectStalledCachedTensors(std::vector<std::string>* names) {
  auto now = Clock::now();
  for (auto& kv : cached_)
    if (std::chrono::duration<double>(now - kv.second).count() > warn_s_) names->push_back(kv.first);
}
bool StallInspector::ShouldPerformCheck() {
  return enabled_ && std::chrono::duration<double>(Clock::now() - last_check_).count() > std::min(warn_s_, 5.0);
}
void StallInspector::UpdateCheckTime() { last_check_ = Clock::now(); }

}  // namespace hvd
