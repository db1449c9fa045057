#include "timeline.h"
#include <chrono>
#include "logging.h"
#include "nvtx_op_range.h"

namespace hvd {

namespace {
constexpr size_t kRingCapacity = 1 << 20;
std::string JsonEscape(const std::string& s) {
  std::string o;
  for (char c : s) { if (c == '"' || c == '\\') o.push_back('\\'); o.push_back(c); }
  return o;
}
}  // namespace

int64_t Timeline::NowUs() const { return (int64_t)((NowNs() - start_ns_) / 1000); }

void Timeline::Initialize(const std::string& file, int world_size) {
  if (Initialized()) return;
  file_ = fopen(file.c_str(), "w");
  if (!file_) { LOG(ERROR) << "Error opening the Horovod Timeline file " << file << ", will not write a timeline."; return; }
  fputs("[\n", file_);
  first_record_ = true;
  closed_ = false;
  start_ns_ = NowNs();
  ring_.assign(kRingCapacity, Record());
  head_.store(0); tail_.store(0);
  stop_.store(false);
  pids_.clear(); states_.clear(); nvtx_top_.clear(); nvtx_act_.clear();
  writer_ = std::thread(&Timeline::WriterLoop, this);
  initialized_.store(true, std::memory_order_release);
  (void)world_size;
}

void Timeline::Shutdown() {
  if (!Initialized()) return;
  initialized_.store(false, std::memory_order_release);
  { std::lock_guard<std::mutex> l(mu_); }  // a producer that saw initialized_ == true has finished its push
  stop_.store(true, std::memory_order_release);
  if (writer_.joinable()) writer_.join();
  if (file_) { if (!closed_) fputs("\n]\n", file_); fclose(file_); file_ = nullptr; }
  std::lock_guard<std::mutex> l(mu_);
  for (auto& kv : nvtx_act_) NvtxRangeEnd(kv.second, true);
  for (auto& kv : nvtx_top_) NvtxRangeEnd(kv.second, false);
  pids_.clear(); states_.clear(); nvtx_top_.clear(); nvtx_act_.clear();
}

// Producer side (callers hold mu_, so there is one producer at a time): no lock, no syscall.
void Timeline::Push(Record r) {
  const size_t h = head_.load(std::memory_order_relaxed);
  const size_t next = (h + 1) % kRingCapacity;
  if (next == tail_.load(std::memory_order_acquire)) { dropped_.fetch_add(1, std::memory_order_relaxed); return; }  // full: drop, never block
  ring_[h] = std::move(r);
  head_.store(next, std::memory_order_release);
}

void Timeline::WriterLoop() {
  while (true) {
    const size_t t = tail_.load(std::memory_order_relaxed);
    if (t == head_.load(std::memory_order_acquire)) {
      // drained: leave a complete JSON document behind, then nap
      if (!closed_ && !first_record_) { fputs("\n]\n", file_); fflush(file_); closed_ = true; }
      if (stop_.load(std::memory_order_acquire) && t == head_.load(std::memory_order_acquire)) break;
      std::this_thread::sleep_for(std::chrono::microseconds(500));
      continue;
    }
    Record r = std::move(ring_[t]);
    tail_.store((t + 1) % kRingCapacity, std::memory_order_release);
    if (closed_) { fseek(file_, -3, SEEK_END); closed_ = false; }  // overwrite the "\n]\n" written at the last drain
    if (!first_record_) fputs(",\n", file_);
    first_record_ = false;
    if (r.meta) {
      fprintf(file_, "{\"name\": \"process_name\", \"ph\": \"M\", \"pid\": %d, \"args\": {\"name\": \"%s\"}},\n", r.pid,
              JsonEscape(r.name).c_str());
      fprintf(file_, "{\"name\": \"process_sort_index\", \"ph\": \"M\", \"pid\": %d, \"args\": {\"sort_index\": %d}},\n", r.pid, r.pid);
      fprintf(file_, "{\"name\": \"thread_name\", \"ph\": \"M\", \"pid\": %d, \"tid\": 1, \"args\": {\"name\": \"GPU (device-timed)\"}}", r.pid);
    } else {
      fprintf(file_, "{\"ph\": \"%c\"", r.phase);
      if (r.phase != 'E') fprintf(file_, ", \"name\": \"%s\"", JsonEscape(r.name).c_str());
      fprintf(file_, ", \"ts\": %lld, \"pid\": %d", (long long)r.ts_us, r.pid);
      if (r.tid) fprintf(file_, ", \"tid\": %d", r.tid);
      if (r.phase == 'X') fprintf(file_, ", \"dur\": %lld", (long long)r.dur_us);
      if (r.phase == 'i') fputs(", \"s\": \"g\"", file_);
      if (!r.args.empty()) fprintf(file_, ", \"args\": {%s}", r.args.c_str());
      fputs("}", file_);
    }
  }
  fflush(file_);
}

int Timeline::Pid(const std::string& tensor_name) {
  auto it = pids_.find(tensor_name);
  if (it != pids_.end()) return it->second;
  int pid = (int)pids_.size() + 1;
  pids_[tensor_name] = pid;
  Record r; r.meta = true; r.pid = pid; r.name = tensor_name; r.phase = 'M'; r.ts_us = 0;
  Push(std::move(r));
  return pid;
}

void Timeline::NegotiateStart(const std::string& name, RequestType type) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  if (states_[name] != State::UNKNOWN) return;
  Push({'B', Pid(name), std::string("NEGOTIATE_") + RequestTypeName(type), "", NowUs()});
  states_[name] = State::NEGOTIATING;
}
void Timeline::NegotiateRankReady(const std::string& name, int rank) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  if (states_[name] != State::NEGOTIATING) return;
  Push({'X', Pid(name), std::to_string(rank), "", NowUs()});
}
void Timeline::NegotiateEnd(const std::string& name) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  if (states_[name] != State::NEGOTIATING) return;
  Push({'E', Pid(name), "", "", NowUs()});
  states_[name] = State::UNKNOWN;
}
void Timeline::Start(const std::string& name, ResponseType type, size_t bytes) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  if (states_[name] == State::NEGOTIATING) { Push({'E', Pid(name), "", "", NowUs()}); }
  std::string args = bytes ? "\"bytes\": " + std::to_string(bytes) : "";
  Push({'B', Pid(name), ResponseTypeName(type), args, NowUs()});
  states_[name] = State::TOP_LEVEL;
  if (NvtxEnabled()) { NvtxRangeEnd(nvtx_top_[name], false); nvtx_top_[name] = NvtxRangeStart(name + ": " + ResponseTypeName(type), false); }
}
void Timeline::ActivityStart(const std::string& name, const std::string& activity) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  if (states_[name] == State::ACTIVITY) Push({'E', Pid(name), "", "", NowUs()});
  if (states_[name] == State::UNKNOWN) return;
  Push({'B', Pid(name), activity, "", NowUs()});
  states_[name] = State::ACTIVITY;
  if (NvtxEnabled()) { NvtxRangeEnd(nvtx_act_[name], true); nvtx_act_[name] = NvtxRangeStart(name + ": " + activity, true); }
}
void Timeline::ActivityEnd(const std::string& name) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  if (states_[name] != State::ACTIVITY) return;
  Push({'E', Pid(name), "", "", NowUs()});
  states_[name] = State::TOP_LEVEL;
  if (NvtxEnabled()) { NvtxRangeEnd(nvtx_act_[name], true); nvtx_act_[name] = 0; }
}
void Timeline::ActivityStartAll(const std::vector<std::shared_ptr<TensorTableEntry>>& es, const std::string& a) {
  if (!Initialized()) return;
  for (auto& e : es) if (e) ActivityStart(e->name, a);
}
void Timeline::ActivityEndAll(const std::vector<std::shared_ptr<TensorTableEntry>>& es) {
  if (!Initialized()) return;
  for (auto& e : es) if (e) ActivityEnd(e->name);
}
void Timeline::End(const std::string& name, const std::string& args) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  if (states_[name] == State::ACTIVITY) Push({'E', Pid(name), "", "", NowUs()});
  if (states_[name] == State::ACTIVITY || states_[name] == State::TOP_LEVEL) Push({'E', Pid(name), "", args, NowUs()});
  states_[name] = State::UNKNOWN;
  if (NvtxEnabled()) {
    auto a = nvtx_act_.find(name); if (a != nvtx_act_.end()) { NvtxRangeEnd(a->second, true); nvtx_act_.erase(a); }
    auto t = nvtx_top_.find(name); if (t != nvtx_top_.end()) { NvtxRangeEnd(t->second, false); nvtx_top_.erase(t); }
  }
}
void Timeline::DeviceSpan(const std::vector<std::string>& names, const std::string& activity, int64_t start_us, int64_t dur_us) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  for (auto& n : names) {
    Record r{'X', Pid(n), activity, "", start_us};
    r.dur_us = dur_us > 0 ? dur_us : 1;
    r.tid = 1;
    Push(std::move(r));
  }
}
void Timeline::MarkCycleStart() {
  if (!Initialized() || !mark_cycles_) return;
  std::lock_guard<std::mutex> l(mu_);
  Push({'i', 0, "CYCLE_START", "", NowUs()});
}

}  // namespace hvd
