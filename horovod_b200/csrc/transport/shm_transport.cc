// Shared-memory control-plane overlay for ranks on one host.
//
// The reference synchronises its response-cache bit vector with an
// MPI_Allreduce / gloo allreduce over TCP every cycle
// (response_cache.cc:428,465; mpi_controller.cc:117-127).  On one NVSwitch box
// all ranks share a host, so the AND/OR and the barrier run on a POSIX shm
// segment with per-rank sequence-stamped slots: ~1-2 us instead of a socket
// round trip, which matters because a small NVLink allreduce is ~10 us.
#include <fcntl.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include "../common/logging.h"
#include "transport.h"

namespace hvd {
namespace {

constexpr int kMaxWords = 512;  // 32768 bits; larger vectors fall back to the base transport
constexpr int kMaxRanks = 64;

struct alignas(64) Slot {
  std::atomic<uint64_t> seq;
  int32_t pid;
  int32_t pad;
  uint64_t data[2][kMaxWords];
};

// Process sets (sub-communicators) of a single-host job negotiate through their own, smaller slots: one "channel" per
// Split() call.  Channels are handed out by a counter that advances on EVERY rank for EVERY Split (members or not), so the
// members of a set agree on theirs without talking; they are not recycled — when they run out, a set uses the base transport.
constexpr int kSubChannels = 8;
constexpr int kSubWords = 64;   // 4096 cache bits per process set

struct alignas(64) SubSlot {
  std::atomic<uint64_t> seq;
  uint64_t data[2][kSubWords];
};

struct Segment {
  std::atomic<uint32_t> magic;
  uint32_t nranks;
  Slot slots[kMaxRanks];
  SubSlot sub[kSubChannels][kMaxRanks];
};

// A process that exited but has not been reaped yet (a zombie) still answers kill(pid, 0): look at its state as well.
bool ProcessAlive(int pid) {
  if (kill(pid, 0) != 0 && errno == ESRCH) return false;
  char path[64];
  snprintf(path, sizeof path, "/proc/%d/stat", pid);
  FILE* f = fopen(path, "r");
  if (!f) return errno != ENOENT;          // no procfs: trust kill()
  char buf[512];
  size_t n = fread(buf, 1, sizeof buf - 1, f);
  fclose(f);
  buf[n] = 0;
  const char* close_paren = strrchr(buf, ')');   // "pid (comm) S ...": comm may contain spaces and parentheses
  return !(close_paren && close_paren[1] == ' ' && (close_paren[2] == 'Z' || close_paren[2] == 'X'));
}

// Spin (pause -> yield -> 50 us naps) until the slot's sequence reaches `k`; once a second make sure its owner still lives.
void WaitSeqReaches(std::atomic<uint64_t>& seq, const int32_t& owner_pid, uint64_t k, int r);
void WaitSlot(Slot& s, uint64_t k, int r) { WaitSeqReaches(s.seq, s.pid, k, r); }
// Waiting policy: `pauses` busy polls (lowest latency, burns a core), then `yields` sched_yield polls (lets a runnable peer
// on the same core in), then 50 us naps (a job whose ranks have nothing to agree on must not spin).  Containers with a CPU
// quota may prefer fewer yields: HVD_SHM_SPIN_PAUSES / HVD_SHM_SPIN_YIELDS.
struct SpinPolicy {
  uint64_t pauses = 2000, yields = 18000;
  double timeout_s = 0;     // HVD_SHM_TIMEOUT_SECONDS: give up on a peer that is alive but does not arrive (0 = wait forever)
  SpinPolicy() {
    if (const char* e = getenv("HVD_SHM_SPIN_PAUSES")) pauses = (uint64_t)std::max(0LL, atoll(e));
    if (const char* e = getenv("HVD_SHM_SPIN_YIELDS")) yields = (uint64_t)std::max(0LL, atoll(e));
    if (const char* e = getenv("HVD_SHM_TIMEOUT_SECONDS")) timeout_s = std::max(0.0, atof(e));
  }
};

// Peers' pids are only meaningful inside one pid namespace.  Ranks that share /dev/shm but not the pid namespace (one
// container per rank with host IPC) would look dead to each other: the wrappers verify once, right after start-up, that
// every peer's pid is visible and otherwise switch the liveness check off (failures then surface through the base transport).
std::atomic<bool> g_trust_pids{true};

void VerifyPeerPids(Segment* seg, int nslots, int my_slot) {
  for (int i = 0; i < nslots; ++i) {
    const int pid = seg->slots[i].pid;
    if (i != my_slot && pid > 0 && !ProcessAlive(pid)) {
      if (g_trust_pids.exchange(false))
        LOG(WARNING) << "peer pids are not visible from this process (separate pid namespaces?); shared-memory waits will not detect dead peers";
      return;
    }
  }
}

void WaitSeqReaches(std::atomic<uint64_t>& seq, const int32_t& owner_pid, uint64_t k, int r) {
  static const SpinPolicy policy;
  uint64_t spins = 0;
  auto last_check = std::chrono::steady_clock::now();
  const auto started = last_check;
  while (seq.load(std::memory_order_acquire) < k) {
    ++spins;
    if (spins < policy.pauses) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    } else if (spins < policy.pauses + policy.yields) {
      std::this_thread::yield();
    } else {
      std::this_thread::sleep_for(std::chrono::microseconds(50));
      auto now = std::chrono::steady_clock::now();
      if (now - last_check > std::chrono::seconds(1)) {
        last_check = now;
        int pid = g_trust_pids.load(std::memory_order_relaxed) ? (int)owner_pid : 0;
        if (pid > 0 && !ProcessAlive(pid))
          throw TransportError("rank " + std::to_string(r) + " (pid " + std::to_string(pid) + ") died");
        if (policy.timeout_s > 0 && std::chrono::duration<double>(now - started).count() > policy.timeout_s)
          throw TransportError("rank " + std::to_string(r) + " did not reach the shared-memory exchange within " +
                               std::to_string((int)policy.timeout_s) + " s (HVD_SHM_TIMEOUT_SECONDS)");
      }
    }
  }
}

char* MapNamed(const std::string& name, size_t bytes, bool create);
size_t DataSlotBytes();

// Variable-length blobs to a root slot / from a root slot through per-member data slots: every member writes [length][payload]
// into its slot, one barrier, the root reads.  A payload that does not fit a slot is announced in the header and EVERY member
// (all of them read all headers) returns false together — the caller then takes the socket path.  `me` / `root` are slot
// indices, `barrier` synchronises exactly the `nslots` members, `half` comes from the communicator's piece counter.
template <typename BarrierFn>
bool SlotGather(int me, int nslots, int root, int half, BarrierFn barrier, char* data, size_t slot_bytes, const std::vector<uint8_t>& mine,
                std::vector<std::vector<uint8_t>>* all) {
  const size_t cap = slot_bytes - 8;
  auto slot_of = [&](int r) { return data + ((size_t)r * 2 + (size_t)half) * slot_bytes; };
  const int64_t len = (int64_t)mine.size();
  memcpy(slot_of(me), &len, 8);
  if (len > 0 && (size_t)len <= cap) memcpy(slot_of(me) + 8, mine.data(), (size_t)len);
  barrier();
  for (int r = 0; r < nslots; ++r) {
    int64_t l = 0;
    memcpy(&l, slot_of(r), 8);
    if ((size_t)l > cap) return false;
  }
  if (me != root) return true;
  all->assign((size_t)nslots, {});
  for (int r = 0; r < nslots; ++r) {
    int64_t l = 0;
    memcpy(&l, slot_of(r), 8);
    (*all)[(size_t)r].assign((const uint8_t*)slot_of(r) + 8, (const uint8_t*)slot_of(r) + 8 + l);
  }
  return true;
}

template <typename BarrierFn>
bool SlotBcast(int me, int root, int half, BarrierFn barrier, char* data, size_t slot_bytes, std::vector<uint8_t>* buf) {
  const size_t cap = slot_bytes - 8;
  char* slot = data + ((size_t)root * 2 + (size_t)half) * slot_bytes;
  if (me == root) {
    const int64_t len = (int64_t)buf->size();
    memcpy(slot, &len, 8);
    if (len > 0 && (size_t)len <= cap) memcpy(slot + 8, buf->data(), (size_t)len);
  }
  barrier();
  int64_t len = 0;
  memcpy(&len, slot, 8);
  if ((size_t)len > cap) return false;
  if (me != root) buf->assign((const uint8_t*)slot + 8, (const uint8_t*)slot + 8 + len);
  return true;
}

// the whole communicator shares the slots (rank i owns slot i)
bool SlotGatherBytes(Transport* t, char* data, size_t slot_bytes, const std::vector<uint8_t>& mine, std::vector<std::vector<uint8_t>>* all,
                     int root) {
  const int half = (int)(t->ShmNextPiece() & 1);
  return SlotGather(t->rank(), t->size(), root, half, [t] { t->Barrier(); }, data, slot_bytes, mine, all);
}
bool SlotBcastBytes(Transport* t, char* data, size_t slot_bytes, std::vector<uint8_t>* buf, int root) {
  const int half = (int)(t->ShmNextPiece() & 1);
  return SlotBcast(t->rank(), root, half, [t] { t->Barrier(); }, data, slot_bytes, buf);
}

// A process set of a single-host job: point-to-point traffic goes through the parent, the per-cycle bit exchange and the
// barrier run on the set's own channel of the parent's segment.
class ShmSubTransport : public SubTransport {
 public:
  // `slots[i]`: index of member i in the segment's slot arrays (default: its rank in the parent — single-host jobs; the
  // two-level plane passes the members' LOCAL indices, its segment being per host)
  ShmSubTransport(Transport* parent, std::vector<int> ranks, int my_index, Segment* seg, int channel, std::vector<int> slots = {})
      : SubTransport(parent, std::move(ranks), my_index), seg_(seg), channel_(channel), slots_(std::move(slots)) {
    if (slots_.empty()) slots_ = ranks_;
  }
  void AllreduceBits(uint64_t* and_words, int n_and, uint64_t* or_words, int n_or) override {
    const int n = n_and + n_or;
    if (size() == 1) return;
    if (n > kSubWords) { Transport::AllreduceBits(and_words, n_and, or_words, n_or); return; }
    const uint64_t k = ++round_;
    const int buf = (int)(k & 1);
    SubSlot* row = seg_->sub[channel_];
    SubSlot& me = row[slots_[(size_t)my_]];
    if (n_and) memcpy(me.data[buf], and_words, (size_t)n_and * 8);
    if (n_or) memcpy(me.data[buf] + n_and, or_words, (size_t)n_or * 8);
    me.seq.store(k, std::memory_order_release);
    for (int i = 0; i < size(); ++i) {
      if (i == my_) continue;
      SubSlot& s = row[slots_[(size_t)i]];
      WaitSeqReaches(s.seq, seg_->slots[slots_[(size_t)i]].pid, k, ranks_[i]);
      for (int w = 0; w < n_and; ++w) and_words[w] &= s.data[buf][w];
      for (int w = 0; w < n_or; ++w) or_words[w] |= s.data[buf][n_and + w];
    }
  }
  void Barrier() override { AllreduceBits(nullptr, 0, nullptr, 0); }
  void GatherBytes(const std::vector<uint8_t>& mine, std::vector<std::vector<uint8_t>>* all, int root) override {
    if (!data_ || size() == 1 || !SlotGatherBytes(this, data_, slot_bytes_, mine, all, root)) Transport::GatherBytes(mine, all, root);
  }
  void BcastBytes(std::vector<uint8_t>* buf, int root) override {
    if (!data_ || size() == 1 || !SlotBcastBytes(this, data_, slot_bytes_, buf, root)) Transport::BcastBytes(buf, root);
  }
  std::string Describe() const override {
    return "control: shared memory channel " + std::to_string(channel_) + " (" + std::to_string(size()) + " of the host's ranks); host data: " +
           (data_ ? "shared-memory slots of " + std::to_string(slot_bytes_) + " bytes" : std::string("ring over the base transport"));
  }
  ~ShmSubTransport() override { if (data_) munmap(data_, data_bytes_); }
  bool ShmDataPlane(ShmData* out) override {
    if (!data_) return false;
    out->base = data_; out->slot_bytes = slot_bytes_;
    return true;
  }
  uint64_t ShmNextPiece() override { return piece_++; }

  // Collective among the members (they are all inside Split() at the same time): the first member creates the set's own
  // data segment, the others map it; the set's fresh control channel carries the two agreement rounds.
  void CreateDataPlane(const std::string& name) {
    if (size() < 2) return;
    const char* dp = getenv("HVD_SHM_DATA_PLANE");
    if (dp && atoi(dp) == 0) return;
    const size_t slot = DataSlotBytes(), total = slot * 2 * (size_t)size();
    char* data = nullptr;
    if (my_ == 0) { shm_unlink(name.c_str()); data = MapNamed(name, total, true); }
    uint64_t ok = my_ == 0 ? (uint64_t)(data != nullptr) : 1;
    AllreduceBits(&ok, 1, nullptr, 0);
    if (ok && my_ != 0) data = MapNamed(name, total, false);
    uint64_t ok2 = ok ? (uint64_t)(data != nullptr) : 0;
    AllreduceBits(&ok2, 1, nullptr, 0);
    if (my_ == 0) shm_unlink(name.c_str());
    if (!ok2) { if (data) munmap(data, total); return; }
    data_ = data; data_bytes_ = total; slot_bytes_ = slot;
  }

 private:
  Segment* seg_;
  int channel_;
  std::vector<int> slots_;
  uint64_t round_ = 0;
  char* data_ = nullptr;
  size_t data_bytes_ = 0, slot_bytes_ = 0;
  uint64_t piece_ = 0;
};

class ShmControlTransport : public Transport {
 public:
  ShmControlTransport(std::shared_ptr<Transport> base, Segment* seg) : base_(std::move(base)), seg_(seg) {
    seg_->slots[base_->rank()].pid = (int32_t)getpid();
  }
  ~ShmControlTransport() override {
    munmap(seg_, sizeof(Segment));
    if (data_) munmap(data_, data_bytes_);
  }
  void AttachData(char* data, size_t total_bytes, size_t slot_bytes) { data_ = data; data_bytes_ = total_bytes; slot_bytes_ = slot_bytes; }
  bool ShmDataPlane(ShmData* out) override {
    if (!data_) return false;
    out->base = data_; out->slot_bytes = slot_bytes_;
    return true;
  }
  uint64_t ShmNextPiece() override { return piece_++; }
  std::shared_ptr<Transport> Split(const std::vector<int>& ranks) override {
    const int channel = next_channel_++;                         // advances on members and non-members alike
    auto it = std::find(ranks.begin(), ranks.end(), rank());
    if (it == ranks.end()) return nullptr;
    const int idx = (int)(it - ranks.begin());
    if (channel >= kSubChannels) return std::make_shared<SubTransport>(this, ranks, idx);
    auto sub = std::make_shared<ShmSubTransport>(this, ranks, idx, seg_, channel);
    if (data_) sub->CreateDataPlane(name_ + "-c" + std::to_string(channel) + "-d");
    return sub;
  }
  void set_name(const std::string& n) { name_ = n; }
  std::string Describe() const override {
    return std::string("control: shared memory (") + std::to_string(size()) + " ranks, one host); host data: " +
           (data_ ? "shared-memory slots of " + std::to_string(slot_bytes_) + " bytes" : std::string("ring over the base transport"));
  }
  int host_id(int i) const override { return base_->host_id(i); }
  int rank() const override { return base_->rank(); }
  int size() const override { return base_->size(); }
  int global_rank(int i) const override { return base_->global_rank(i); }
  bool single_host() const override { return true; }
  void Send(int p, const void* b, size_t n) override { base_->Send(p, b, n); }
  void Recv(int p, void* b, size_t n) override { base_->Recv(p, b, n); }
  void SendRecv(int sp, const void* sb, size_t sn, int rp, void* rb, size_t rn) override {
    base_->SendRecv(sp, sb, sn, rp, rb, rn);
  }
  void AlltoallvBytes(const char* in, const int64_t* sd, char* out, const int64_t* rd, const uint8_t* skip) override {
    base_->AlltoallvBytes(in, sd, out, rd, skip);
  }

  void AllreduceBits(uint64_t* and_words, int n_and, uint64_t* or_words, int n_or) override {
    const int n = n_and + n_or;
    if (size() == 1) return;
    if (n > kMaxWords) { Transport::AllreduceBits(and_words, n_and, or_words, n_or); return; }
    const uint64_t k = ++round_;
    const int buf = (int)(k & 1);
    Slot& me = seg_->slots[rank()];
    if (n_and) memcpy(me.data[buf], and_words, (size_t)n_and * 8);
    if (n_or) memcpy(me.data[buf] + n_and, or_words, (size_t)n_or * 8);
    me.seq.store(k, std::memory_order_release);
    for (int r = 0; r < size(); ++r) {
      if (r == rank()) continue;
      Slot& s = seg_->slots[r];
      WaitSeq(s, k, r);
      for (int i = 0; i < n_and; ++i) and_words[i] &= s.data[buf][i];
      for (int i = 0; i < n_or; ++i) or_words[i] |= s.data[buf][n_and + i];
    }
  }
  void Barrier() override { AllreduceBits(nullptr, 0, nullptr, 0); }
  // The coordinator round of uncached requests (names, shapes -> responses) through the data slots (SlotGatherBytes).
  void GatherBytes(const std::vector<uint8_t>& mine, std::vector<std::vector<uint8_t>>* all, int root) override {
    if (!data_ || size() == 1 || !SlotGatherBytes(this, data_, slot_bytes_, mine, all, root)) Transport::GatherBytes(mine, all, root);
  }
  void BcastBytes(std::vector<uint8_t>* buf, int root) override {
    if (!data_ || size() == 1 || !SlotBcastBytes(this, data_, slot_bytes_, buf, root)) Transport::BcastBytes(buf, root);
  }
  // Small integer tables (alltoall split matrices, IPC handle records, topology at init) through the same slots and round
  // counter as the bit vectors: one publication + n - 1 reads instead of a star over sockets.
  void AllgatherInts(const int64_t* mine, int n, int64_t* out) override {
    if (size() == 1 || n > kMaxWords || n <= 0) { Transport::AllgatherInts(mine, n, out); return; }
    const uint64_t k = ++round_;
    const int buf = (int)(k & 1);
    Slot& me = seg_->slots[rank()];
    memcpy(me.data[buf], mine, (size_t)n * 8);
    me.seq.store(k, std::memory_order_release);
    memcpy(out + (size_t)rank() * (size_t)n, mine, (size_t)n * 8);
    for (int r = 0; r < size(); ++r) {
      if (r == rank()) continue;
      Slot& s = seg_->slots[r];
      WaitSeq(s, k, r);
      memcpy(out + (size_t)r * (size_t)n, s.data[buf], (size_t)n * 8);
    }
  }

 private:
  static void WaitSeq(Slot& s, uint64_t k, int r) { WaitSlot(s, k, r); }
  std::shared_ptr<Transport> base_;
  Segment* seg_;
  uint64_t round_ = 0;
  char* data_ = nullptr;
  size_t data_bytes_ = 0, slot_bytes_ = 0;
  uint64_t piece_ = 0;
  int next_channel_ = 0;
  std::string name_;
};
size_t DataSlotBytes() {
  size_t slot = 1u << 20;
  if (const char* sb = getenv("HVD_SHM_SLOT_BYTES")) slot = (size_t)std::max(4096LL, atoll(sb));
  return (slot + 4095) & ~(size_t)4095;
}

// Maps `bytes` of a named segment; the creator reserves the pages up front (posix_fallocate) so that a too-small /dev/shm
// shows up here as an error instead of a SIGBUS in the middle of a collective.
char* MapNamed(const std::string& name, size_t bytes, bool create) {
  int fd = create ? shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600) : shm_open(name.c_str(), O_RDWR, 0600);
  if (fd < 0) return nullptr;
  if (create && (ftruncate(fd, (off_t)bytes) != 0 || posix_fallocate(fd, 0, (off_t)bytes) != 0)) {
    close(fd);
    shm_unlink(name.c_str());
    return nullptr;
  }
  void* p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  return p == MAP_FAILED ? nullptr : (char*)p;
}


// Two-level control plane of a MULTI-host job: the ranks of one host fold their bit vectors in a per-host shm segment, the
// host leaders exchange the folded vectors over the base (TCP) transport, and the leaders publish the result back through
// shm.  Per negotiation cycle the job sends 2 (H - 1) small TCP messages instead of 2 (N - 1) — with 8 GPUs per host an
// 8x lighter load on rank 0, and no socket at all between the ranks of a host.
class HierShmControlTransport : public Transport {
 public:
  HierShmControlTransport(std::shared_ptr<Transport> base, Segment* seg, std::vector<int> local, int local_index,
                          std::vector<int> leaders)
      : base_(std::move(base)), seg_(seg), local_(std::move(local)), li_(local_index), leaders_(std::move(leaders)) {
    seg_->slots[li_].pid = (int32_t)getpid();
  }
  ~HierShmControlTransport() override {
    munmap(seg_, sizeof(Segment));
    if (data_) munmap(data_, data_bytes_);
  }
  // ---- two-level data plane for host tensors ----
  void AttachData(char* data, size_t total, size_t slot, std::vector<std::vector<int>> column) {
    data_ = data; data_bytes_ = total; slot_bytes_ = slot; column_ = std::move(column);
    cross_ = base_->Split(column_[li_]);
  }
  bool HierDataPlane(HierData* out) override {
    if (!data_ || !cross_) return false;
    out->local.base = data_; out->local.slot_bytes = slot_bytes_;
    out->local_rank = li_; out->local_size = (int)local_.size();
    out->cross = cross_.get(); out->column = &column_;
    return true;
  }
  uint64_t ShmNextPiece() override { return piece_++; }
  // A process set whose members all live on ONE host negotiates and moves host tensors through that host's segment (own
  // channel, own data slots), like a process set of a single-host job; sets that span hosts stay on the sockets.
  std::shared_ptr<Transport> Split(const std::vector<int>& ranks) override {
    const int channel = next_channel_++;                         // advances on every rank for every set
    auto it = std::find(ranks.begin(), ranks.end(), rank());
    if (it == ranks.end()) return nullptr;
    const int idx = (int)(it - ranks.begin());
    std::vector<int> slots;
    for (int r : ranks) {
      auto at = std::find(local_.begin(), local_.end(), r);
      if (at == local_.end()) break;
      slots.push_back((int)(at - local_.begin()));
    }
    if (slots.size() != ranks.size() || channel >= kSubChannels || ranks.size() < 2) return std::make_shared<SubTransport>(this, ranks, idx);
    auto sub = std::make_shared<ShmSubTransport>(this, ranks, idx, seg_, channel, slots);
    if (data_) sub->CreateDataPlane(name_ + "-c" + std::to_string(channel) + "-d");
    return sub;
  }
  void set_name(const std::string& n) { name_ = n; }
  void LocalBarrier() override {
    const uint64_t k = ++local_round_;
    SubSlot* row = seg_->sub[0];
    row[li_].seq.store(k, std::memory_order_release);
    for (int j = 0; j < (int)local_.size(); ++j)
      if (j != li_) WaitSeqReaches(row[j].seq, seg_->slots[j].pid, k, local_[j]);
  }
  int rank() const override { return base_->rank(); }
  int size() const override { return base_->size(); }
  int global_rank(int i) const override { return base_->global_rank(i); }
  bool single_host() const override { return false; }
  int host_id(int i) const override { return base_->host_id(i); }
  void Send(int p, const void* b, size_t n) override { base_->Send(p, b, n); }
  void Recv(int p, void* b, size_t n) override { base_->Recv(p, b, n); }
  void SendRecv(int sp, const void* sb, size_t sn, int rp, void* rb, size_t rn) override {
    base_->SendRecv(sp, sb, sn, rp, rb, rn);
  }
  void AlltoallvBytes(const char* in, const int64_t* sd, char* out, const int64_t* rd, const uint8_t* skip) override {
    base_->AlltoallvBytes(in, sd, out, rd, skip);
  }

  void AllreduceBits(uint64_t* and_words, int n_and, uint64_t* or_words, int n_or) override {
    const int n = n_and + n_or;
    if (n > kMaxWords) { Transport::AllreduceBits(and_words, n_and, or_words, n_or); return; }
    const uint64_t k = ++round_;
    const int buf = (int)(k & 1);
    const int nlocal = (int)local_.size();
    Slot& result = seg_->slots[nlocal];          // one slot past the members: the leader's answer
    if (li_ != 0) {
      Slot& me = seg_->slots[li_];
      if (n_and) memcpy(me.data[buf], and_words, (size_t)n_and * 8);
      if (n_or) memcpy(me.data[buf] + n_and, or_words, (size_t)n_or * 8);
      me.seq.store(k, std::memory_order_release);
      WaitSlot(result, k, local_[0]);
      if (n_and) memcpy(and_words, result.data[buf], (size_t)n_and * 8);
      if (n_or) memcpy(or_words, result.data[buf] + n_and, (size_t)n_or * 8);
      return;
    }
    // ---- host leader ----
    for (int j = 1; j < nlocal; ++j) {
      Slot& s = seg_->slots[j];
      WaitSlot(s, k, local_[j]);
      for (int i = 0; i < n_and; ++i) and_words[i] &= s.data[buf][i];
      for (int i = 0; i < n_or; ++i) or_words[i] |= s.data[buf][n_and + i];
    }
    if (leaders_.size() > 1) {
      std::vector<uint64_t> mine((size_t)std::max(n, 1));
      if (n_and) memcpy(mine.data(), and_words, (size_t)n_and * 8);
      if (n_or) memcpy(mine.data() + n_and, or_words, (size_t)n_or * 8);
      int me = 0;
      while (leaders_[me] != rank()) ++me;
      base_->AllreduceBitsAmong(leaders_, me, mine.data(), n_and, n);
      if (n_and) memcpy(and_words, mine.data(), (size_t)n_and * 8);
      if (n_or) memcpy(or_words, mine.data() + n_and, (size_t)n_or * 8);
    }
    if (n_and) memcpy(result.data[buf], and_words, (size_t)n_and * 8);
    if (n_or) memcpy(result.data[buf] + n_and, or_words, (size_t)n_or * 8);
    result.seq.store(k, std::memory_order_release);
  }
  void Barrier() override { AllreduceBits(nullptr, 0, nullptr, 0); }

  // Coordinator round, two-level: the ranks of a host hand their blobs to the host leader through the host's data slots, the
  // leaders send ONE message each to the root (H - 1 messages at the root instead of N - 1); responses travel the other way.
  // Needs the root to be a host leader (the coordinator, rank 0, is) and the host data plane; otherwise the plain star.
  void GatherBytes(const std::vector<uint8_t>& mine, std::vector<std::vector<uint8_t>>* all, int root) override {
    if (!data_ || std::find(leaders_.begin(), leaders_.end(), root) == leaders_.end()) { Transport::GatherBytes(mine, all, root); return; }
    const int L = (int)local_.size();
    std::vector<std::vector<uint8_t>> loc;
    const int half = (int)(ShmNextPiece() & 1);
    if (!SlotGather(li_, L, 0, half, [this] { LocalBarrier(); }, data_, slot_bytes_, mine, &loc)) {
      // a blob of this host does not fit a slot: the host's ranks (they all saw the same headers) use their sockets
      if (li_ == 0) {
        loc.assign((size_t)L, {});
        loc[0] = mine;
        for (int j = 1; j < L; ++j) {
          int64_t n = 0;
          base_->Recv(local_[(size_t)j], &n, sizeof n);
          loc[(size_t)j].resize((size_t)n);
          if (n) base_->Recv(local_[(size_t)j], loc[(size_t)j].data(), (size_t)n);
        }
      } else {
        int64_t n = (int64_t)mine.size();
        base_->Send(local_[0], &n, sizeof n);
        if (n) base_->Send(local_[0], mine.data(), mine.size());
      }
    }
    if (li_ != 0) return;
    if (rank() == root) {
      all->assign((size_t)size(), {});
      for (int j = 0; j < L; ++j) (*all)[(size_t)local_[(size_t)j]] = std::move(loc[(size_t)j]);
      for (int leader : leaders_) {
        if (leader == root) continue;
        int64_t n = 0;
        base_->Recv(leader, &n, sizeof n);
        std::vector<uint8_t> pack((size_t)n);
        if (n) base_->Recv(leader, pack.data(), (size_t)n);
        size_t pos = 0;                                  // [rank:int32][len:int64][payload] ...
        while (pos + 12 <= pack.size()) {
          int32_t rk = 0; int64_t len = 0;
          memcpy(&rk, pack.data() + pos, 4); memcpy(&len, pack.data() + pos + 4, 8);
          pos += 12;
          if (rk < 0 || rk >= size() || pos + (size_t)len > pack.size()) throw TransportError("malformed gather pack from host leader " + std::to_string(leader));
          (*all)[(size_t)rk].assign(pack.begin() + (long)pos, pack.begin() + (long)(pos + (size_t)len));
          pos += (size_t)len;
        }
      }
    } else {
      std::vector<uint8_t> pack;
      for (int j = 0; j < L; ++j) {
        const int32_t rk = (int32_t)local_[(size_t)j];
        const int64_t len = (int64_t)loc[(size_t)j].size();
        const size_t at = pack.size();
        pack.resize(at + 12 + (size_t)len);
        memcpy(pack.data() + at, &rk, 4); memcpy(pack.data() + at + 4, &len, 8);
        if (len) memcpy(pack.data() + at + 12, loc[(size_t)j].data(), (size_t)len);
      }
      int64_t n = (int64_t)pack.size();
      base_->Send(root, &n, sizeof n);
      if (n) base_->Send(root, pack.data(), pack.size());
    }
  }
  void BcastBytes(std::vector<uint8_t>* buf, int root) override {
    if (!data_ || std::find(leaders_.begin(), leaders_.end(), root) == leaders_.end()) { Transport::BcastBytes(buf, root); return; }
    if (li_ == 0) {
      if (rank() == root) {
        int64_t n = (int64_t)buf->size();
        for (int leader : leaders_) {
          if (leader == root) continue;
          base_->Send(leader, &n, sizeof n);
          if (n) base_->Send(leader, buf->data(), buf->size());
        }
      } else {
        int64_t n = 0;
        base_->Recv(root, &n, sizeof n);
        buf->resize((size_t)n);
        if (n) base_->Recv(root, buf->data(), (size_t)n);
      }
    }
    const int L = (int)local_.size();
    const int half = (int)(ShmNextPiece() & 1);
    if (SlotBcast(li_, 0, half, [this] { LocalBarrier(); }, data_, slot_bytes_, buf)) return;
    if (li_ == 0) {
      int64_t n = (int64_t)buf->size();
      for (int j = 1; j < L; ++j) { base_->Send(local_[(size_t)j], &n, sizeof n); if (n) base_->Send(local_[(size_t)j], buf->data(), buf->size()); }
    } else {
      int64_t n = 0;
      base_->Recv(local_[0], &n, sizeof n);
      buf->resize((size_t)n);
      if (n) base_->Recv(local_[0], buf->data(), (size_t)n);
    }
  }
  // Integer tables, two-level: host slots -> leader, leaders exchange whole host rows through the first leader, full table back
  // through the host slots (needs the homogeneous layout the data plane is built for; otherwise the star of the base class).
  void AllgatherInts(const int64_t* mine, int n, int64_t* out) override {
    const int L = (int)local_.size(), H = (int)leaders_.size(), N = size();
    if (!data_ || column_.empty() || n <= 0 || (size_t)N * (size_t)n * 8 > slot_bytes_) { Transport::AllgatherInts(mine, n, out); return; }
    auto slot_of = [&](int j, int half) { return (int64_t*)(data_ + ((size_t)j * 2 + (size_t)half) * slot_bytes_); };
    int x = 0;                                           // my host's position in the column tables
    while (column_[(size_t)li_][(size_t)x] != rank()) ++x;
    int half = (int)(ShmNextPiece() & 1);
    memcpy(slot_of(li_, half), mine, (size_t)n * 8);
    LocalBarrier();
    std::vector<int64_t> table((size_t)N * (size_t)n);
    if (li_ == 0) {
      std::vector<int64_t> row((size_t)L * (size_t)n);
      for (int j = 0; j < L; ++j) memcpy(row.data() + (size_t)j * (size_t)n, slot_of(j, half), (size_t)n * 8);
      auto place = [&](int host, const int64_t* r) {
        for (int j = 0; j < L; ++j) memcpy(table.data() + (size_t)column_[(size_t)j][(size_t)host] * (size_t)n, r + (size_t)j * (size_t)n, (size_t)n * 8);
      };
      if (rank() == leaders_[0]) {
        place(x, row.data());
        std::vector<int64_t> other((size_t)L * (size_t)n);
        for (int h = 1; h < H; ++h) {
          base_->Recv(leaders_[(size_t)h], other.data(), other.size() * 8);
          int hx = 0;                                    // which column position that leader's host has
          while (column_[0][(size_t)hx] != leaders_[(size_t)h]) ++hx;
          place(hx, other.data());
        }
        for (int h = 1; h < H; ++h) base_->Send(leaders_[(size_t)h], table.data(), table.size() * 8);
      } else {
        base_->Send(leaders_[0], row.data(), row.size() * 8);
        base_->Recv(leaders_[0], table.data(), table.size() * 8);
      }
    }
    half = (int)(ShmNextPiece() & 1);
    if (li_ == 0) memcpy(slot_of(0, half), table.data(), table.size() * 8);
    LocalBarrier();
    memcpy(out, slot_of(0, half), (size_t)N * (size_t)n * 8);
  }
  std::string Describe() const override {
    return "control: two-level (shared memory among the " + std::to_string(local_.size()) + " ranks of this host, " +
           std::to_string(leaders_.size()) + " host leaders over the base transport); host data: " +
           (data_ && cross_ ? "two-level (shared-memory slots of " + std::to_string(slot_bytes_) + " bytes inside a host, " +
                                  std::to_string(local_.size()) + " parallel cross-host rings)"
                            : std::string("ring over the base transport"));
  }

 private:
  std::shared_ptr<Transport> base_;
  Segment* seg_;
  std::vector<int> local_;     // ranks of my host, ascending; local_[0] is the host leader
  int li_;                     // my index in local_
  std::vector<int> leaders_;   // leader of every host, ascending
  uint64_t round_ = 0;
  char* data_ = nullptr;
  size_t data_bytes_ = 0, slot_bytes_ = 0;
  uint64_t piece_ = 0, local_round_ = 0;
  std::vector<std::vector<int>> column_;
  std::shared_ptr<Transport> cross_;
  int next_channel_ = 1;       // row 0 of the segment's sub-slots carries LocalBarrier()
  std::string name_;
};

}  // namespace

std::shared_ptr<Transport> WrapWithHierarchicalControl(std::shared_ptr<Transport> base, const std::string& segment_name) {
  const int n = base->size(), me = base->rank();
  if (n == 1 || base->single_host()) return base;
  std::vector<int> local, leaders;
  std::vector<int> seen_hosts;
  for (int r = 0; r < n; ++r) {
    const int h = base->host_id(r);
    if (h == base->host_id(me)) local.push_back(r);
    if (std::find(seen_hosts.begin(), seen_hosts.end(), h) == seen_hosts.end()) { seen_hosts.push_back(h); leaders.push_back(r); }
  }
  // worth it only if some host has several ranks; every rank evaluates the same table, so the decision is collective
  bool any_shared = (int)leaders.size() < n;
  // the size limit is checked on the LARGEST host of the global table, not on this rank's own host: with uneven hosts a
  // per-host test would send some ranks into the collective below and let the others return
  int max_local = 0;
  for (int h : seen_hosts) {
    int c = 0;
    for (int r = 0; r < n; ++r) if (base->host_id(r) == h) ++c;
    max_local = std::max(max_local, c);
  }
  if (!any_shared || max_local + 1 > kMaxRanks) return base;
  const int li = (int)(std::find(local.begin(), local.end(), me) - local.begin());
  const std::string name = "/" + segment_name + "-h" + std::to_string(base->host_id(me));
  Segment* seg = nullptr;
  int ok = 1;
  if (li == 0) {
    shm_unlink(name.c_str());
    seg = (Segment*)MapNamed(name, sizeof(Segment), true);
    if (!seg) ok = 0; else { memset((void*)seg, 0, sizeof(Segment)); seg->nranks = (uint32_t)local.size(); seg->magic.store(0x48564448); }
  }
  uint64_t okw = (uint64_t)ok;
  base->AllreduceBits(&okw, 1, nullptr, 0);            // every leader created its segment
  if (okw && li != 0) { seg = (Segment*)MapNamed(name, sizeof(Segment), false); if (!seg) ok = 0; }
  uint64_t ok2 = okw ? (uint64_t)ok : 0;
  base->AllreduceBits(&ok2, 1, nullptr, 0);            // everybody mapped it
  if (li == 0) shm_unlink(name.c_str());
  if (!ok2) {
    if (seg) munmap(seg, sizeof(Segment));
    LOG(DEBUG) << "two-level control plane unavailable; negotiation stays on the base transport";
    return base;
  }
  const int nlocal = (int)local.size();
  // column[l] = ranks with local index l, host by host; only defined when every host runs the same number of ranks
  std::vector<std::vector<int>> column;
  bool homogeneous = n % (int)leaders.size() == 0 && nlocal == n / (int)leaders.size();
  if (homogeneous) {
    column.assign((size_t)nlocal, {});
    for (int h : seen_hosts) {
      int l = 0;
      for (int r = 0; r < n; ++r) if (base->host_id(r) == h) { if (l < nlocal) column[(size_t)l].push_back(r); ++l; }
      if (l != nlocal) homogeneous = false;
    }
  }
  auto raw = base;
  auto hier = std::make_shared<HierShmControlTransport>(std::move(base), seg, std::move(local), li, std::move(leaders));
  hier->set_name(name);
  hier->Barrier();                       // every rank of this host has published its pid
  VerifyPeerPids(seg, nlocal, li);

  // ---- data plane: per-host slots; collective over ALL ranks (every host takes the same decision) ----
  const char* dp = getenv("HVD_SHM_DATA_PLANE");
  uint64_t want = homogeneous && nlocal > 1 && !(dp && atoi(dp) == 0);
  hier->AllreduceBits(&want, 1, nullptr, 0);
  if (!want) return hier;
  const size_t slot = DataSlotBytes(), total = slot * 2 * (size_t)nlocal;
  const std::string dname = name + "-d";
  char* data = nullptr;
  if (li == 0) { shm_unlink(dname.c_str()); data = MapNamed(dname, total, true); }
  uint64_t okd = li == 0 ? (uint64_t)(data != nullptr) : 1;
  hier->AllreduceBits(&okd, 1, nullptr, 0);
  if (okd && li != 0) data = MapNamed(dname, total, false);
  uint64_t okd2 = okd ? (uint64_t)(data != nullptr) : 0;
  hier->AllreduceBits(&okd2, 1, nullptr, 0);
  if (li == 0) shm_unlink(dname.c_str());
  if (!okd2) { if (data) munmap(data, total); return hier; }
  hier->AttachData(data, total, slot, std::move(column));
  (void)raw;
  return hier;
}

namespace {
}  // namespace

std::shared_ptr<Transport> WrapWithShmControl(std::shared_ptr<Transport> base, const std::string& segment_name) {
  if (base->size() == 1 || base->size() > kMaxRanks || !base->single_host()) return base;
  // rank 0 creates + zero-fills, then a base barrier, others open, barrier, rank 0 unlinks.
  Segment* seg = nullptr;
  int ok = 1;
  const std::string name = "/" + segment_name;
  if (base->rank() == 0) {
    shm_unlink(name.c_str());
    int fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0 || ftruncate(fd, sizeof(Segment)) != 0) ok = 0;
    if (ok) {
      void* p = mmap(nullptr, sizeof(Segment), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (p == MAP_FAILED) ok = 0; else { seg = (Segment*)p; memset((void*)seg, 0, sizeof(Segment)); seg->nranks = base->size(); seg->magic.store(0x48564442); }
    }
    if (fd >= 0) close(fd);
  }
  uint64_t okw = ok;
  base->AllreduceBits(&okw, 1, nullptr, 0);  // doubles as the "segment exists" barrier
  if (okw && base->rank() != 0) {
    int fd = shm_open(name.c_str(), O_RDWR, 0600);
    if (fd < 0) ok = 0;
    else {
      void* p = mmap(nullptr, sizeof(Segment), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (p == MAP_FAILED) ok = 0; else seg = (Segment*)p;
      close(fd);
    }
  }
  uint64_t ok2 = okw ? (uint64_t)ok : 0;
  base->AllreduceBits(&ok2, 1, nullptr, 0);
  if (base->rank() == 0) shm_unlink(name.c_str());
  if (!ok2) {
    if (seg) munmap(seg, sizeof(Segment));
    LOG(DEBUG) << "shared-memory control plane unavailable; using the base transport";
    return base;
  }
  auto shm = std::make_shared<ShmControlTransport>(base, seg);
  shm->set_name(name);
  shm->Barrier();                        // every rank has published its pid
  VerifyPeerPids(seg, base->size(), base->rank());

  // ---- data plane (host tensors of a single-host job never touch a socket) ----
  const char* dp = getenv("HVD_SHM_DATA_PLANE");
  if (dp && atoi(dp) == 0) return shm;
  const size_t slot = DataSlotBytes();
  const size_t total = slot * 2 * (size_t)base->size();
  const std::string dname = name + "-d";
  char* data = nullptr;
  if (base->rank() == 0) { shm_unlink(dname.c_str()); data = MapNamed(dname, total, true); }
  uint64_t okd = base->rank() == 0 ? (data != nullptr) : 1;
  shm->AllreduceBits(&okd, 1, nullptr, 0);
  if (okd && base->rank() != 0) data = MapNamed(dname, total, false);
  uint64_t okd2 = okd ? (uint64_t)(data != nullptr) : 0;
  shm->AllreduceBits(&okd2, 1, nullptr, 0);
  if (base->rank() == 0) shm_unlink(dname.c_str());
  if (!okd2) {
    if (data) munmap(data, total);
    LOG(DEBUG) << "shared-memory data plane unavailable (" << total << " bytes of /dev/shm); host tensors use the base transport";
    return shm;
  }
  shm->AttachData(data, total, slot);
  return shm;
}

}  // namespace hvd
