#include "timeline.h"
#include "logging.h"

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
  start_ns_ = NowNs();
  ring_.assign(kRingCapacity, Record());
  head_ = tail_ = 0;
  stop_ = false;
  pids_.clear(); states_.clear();
  writer_ = std::thread(&Timeline::WriterLoop, this);
  initialized_.store(true, std::memory_order_release);
  (void)world_size;
}

void Timeline::Shutdown() {
  if (!Initialized()) return;
  initialized_.store(false, std::memory_order_release);
  {
    std::lock_guard<std::mutex> l(ring_mu_);
    stop_ = true;
    ring_cv_.notify_all();
  }
  if (writer_.joinable()) writer_.join();
  if (file_) { fputs("\n]\n", file_); fclose(file_); file_ = nullptr; }
  std::lock_guard<std::mutex> l(mu_);
  pids_.clear(); states_.clear();
}

void Timeline::Push(Record r) {
  std::lock_guard<std::mutex> l(ring_mu_);
  size_t next = (head_ + 1) % kRingCapacity;
  if (next == tail_) return;  // full: drop (never block the cycle thread)
  ring_[head_] = std::move(r);
  head_ = next;
  ring_cv_.notify_one();
}

void Timeline::WriterLoop() {
  while (true) {
    Record r;
    {
      std::unique_lock<std::mutex> l(ring_mu_);
      ring_cv_.wait(l, [&] { return stop_ || head_ != tail_; });
      if (head_ == tail_) { if (stop_) break; continue; }
      r = std::move(ring_[tail_]);
      tail_ = (tail_ + 1) % kRingCapacity;
    }
    if (!first_record_) fputs(",\n", file_);
    first_record_ = false;
    if (r.meta) {
      fprintf(file_, "{\"name\": \"process_name\", \"ph\": \"M\", \"pid\": %d, \"args\": {\"name\": \"%s\"}},\n", r.pid,
              JsonEscape(r.name).c_str());
      fprintf(file_, "{\"name\": \"process_sort_index\", \"ph\": \"M\", \"pid\": %d, \"args\": {\"sort_index\": %d}}", r.pid, r.pid);
    } else {
      fprintf(file_, "{\"ph\": \"%c\"", r.phase);
      if (r.phase != 'E') fprintf(file_, ", \"name\": \"%s\"", JsonEscape(r.name).c_str());
      fprintf(file_, ", \"ts\": %lld, \"pid\": %d", (long long)r.ts_us, r.pid);
      if (r.phase == 'X') fputs(", \"dur\": 0", file_);
      if (r.phase == 'i') fputs(", \"s\": \"g\"", file_);
      if (!r.args.empty()) fprintf(file_, ", \"args\": {%s}", r.args.c_str());
      fputs("}", file_);
    }
    // keep the file loadable even if the process dies: flush periodically
    static int n = 0;
    if ((++n & 0xff) == 0) fflush(file_);
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
}
void Timeline::ActivityStart(const std::string& name, const std::string& activity) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  if (states_[name] == State::ACTIVITY) Push({'E', Pid(name), "", "", NowUs()});
  if (states_[name] == State::UNKNOWN) return;
  Push({'B', Pid(name), activity, "", NowUs()});
  states_[name] = State::ACTIVITY;
}
void Timeline::ActivityEnd(const std::string& name) {
  if (!Initialized()) return;
  std::lock_guard<std::mutex> l(mu_);
  if (states_[name] != State::ACTIVITY) return;
  Push({'E', Pid(name), "", "", NowUs()});
  states_[name] = State::TOP_LEVEL;
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
}
void Timeline::MarkCycleStart() {
  if (!Initialized() || !mark_cycles_) return;
  std::lock_guard<std::mutex> l(mu_);
  Push({'i', 0, "CYCLE_START", "", NowUs()});
}

}  // namespace hvd
