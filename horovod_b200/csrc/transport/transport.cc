#include "transport.h"
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>

namespace hvd {

// ---------------------------------------------------------------------------
// Default star-shaped collectives built from point-to-point primitives.

void Transport::GatherBytes(const std::vector<uint8_t>& mine, std::vector<std::vector<uint8_t>>* all, int root) {
  if (rank() == root) {
    all->assign(size(), {});
    (*all)[root] = mine;
    for (int r = 0; r < size(); ++r) {
      if (r == root) continue;
      int64_t n = 0;
      Recv(r, &n, sizeof n);
      (*all)[r].resize((size_t)n);
      if (n) Recv(r, (*all)[r].data(), (size_t)n);
    }
  } else {
    int64_t n = (int64_t)mine.size();
    Send(root, &n, sizeof n);
    if (n) Send(root, mine.data(), mine.size());
  }
}

void Transport::BcastBytes(std::vector<uint8_t>* buf, int root) {
  if (rank() == root) {
    int64_t n = (int64_t)buf->size();
    for (int r = 0; r < size(); ++r) {
      if (r == root) continue;
      Send(r, &n, sizeof n);
      if (n) Send(r, buf->data(), buf->size());
    }
  } else {
    int64_t n = 0;
    Recv(root, &n, sizeof n);
    buf->resize((size_t)n);
    if (n) Recv(root, buf->data(), (size_t)n);
  }
}

void Transport::Bcast(void* buf, size_t n, int root) {
  if (size() == 1 || n == 0) return;
  if (rank() == root) {
    for (int r = 0; r < size(); ++r) if (r != root) Send(r, buf, n);
  } else {
    Recv(root, buf, n);
  }
}

void Transport::AlltoallvBytes(const char* in, const int64_t* sd, char* out, const int64_t* rd, const uint8_t* skip) {
  const int n = size(), r = rank();
  for (int k = 1; k < n; ++k) {
    const int to = (r + k) % n, from = (r - k + n) % n;
    const size_t sn = skip && skip[to] ? 0 : (size_t)(sd[to + 1] - sd[to]), rn = skip && skip[from] ? 0 : (size_t)(rd[from + 1] - rd[from]);
    if (sn && rn) SendRecv(to, in + sd[to], sn, from, out + rd[from], rn);
    else if (sn) Send(to, in + sd[to], sn);
    else if (rn) Recv(from, out + rd[from], rn);
  }
}

void Transport::AllreduceBits(uint64_t* and_words, int n_and, uint64_t* or_words, int n_or) {
  if (size() == 1) return;
  const int n = n_and + n_or;
  std::vector<uint64_t> mine((size_t)std::max(n, 1));
  if (n_and) memcpy(mine.data(), and_words, n_and * 8);
  if (n_or) memcpy(mine.data() + n_and, or_words, n_or * 8);
  std::vector<int> all(size());
  for (int i = 0; i < size(); ++i) all[i] = i;
  AllreduceBitsAmong(all, rank(), mine.data(), n_and, n);
  if (n_and) memcpy(and_words, mine.data(), n_and * 8);
  if (n_or) memcpy(or_words, mine.data() + n_and, n_or * 8);
}

static int BitsTreeMinRanks() {
  static const int v = [] {
    const char* e = getenv("HVD_BITS_TREE_MIN_RANKS");
    return e ? std::max(2, atoi(e)) : 9;
  }();
  return v;
}

void Transport::AllreduceBitsAmong(const std::vector<int>& peers, int me, uint64_t* words, int n_and, int n) {
  const int m = (int)peers.size();
  if (m <= 1) return;
  const size_t wire = (size_t)std::max(n, 1) * 8;          // an empty vector (barrier) still travels as one word
  std::vector<uint64_t> tmp((size_t)std::max(n, 1));
  auto fold = [&] {
    for (int i = 0; i < n_and; ++i) words[i] &= tmp[i];
    for (int i = n_and; i < n; ++i) words[i] |= tmp[i];
  };
  if (m < BitsTreeMinRanks()) {
    // star: the replies of m - 1 members arrive concurrently, the root pays one receive each
    if (me == 0) {
      for (int i = 1; i < m; ++i) { Recv(peers[i], tmp.data(), wire); fold(); }
      for (int i = 1; i < m; ++i) Send(peers[i], words, wire);
    } else {
      Send(peers[0], words, wire);
      Recv(peers[0], words, wire);
    }
    return;
  }
  // recursive doubling over the largest power of two p <= m; the m - p extra members hand their words to a partner first
  // and get the result back at the end: ceil(log2 m) + 2 message times instead of 2 (m - 1) at the root
  int p = 1;
  while (p * 2 <= m) p *= 2;
  const int extra = m - p;
  if (me >= p) {
    Send(peers[me - p], words, wire);
    Recv(peers[me - p], words, wire);
    return;
  }
  if (me < extra) { Recv(peers[me + p], tmp.data(), wire); fold(); }
  for (int mask = 1; mask < p; mask <<= 1) {
    const int partner = peers[me ^ mask];
    SendRecv(partner, words, wire, partner, tmp.data(), wire);
    fold();
  }
  if (me < extra) Send(peers[me + p], words, wire);
}

void Transport::Barrier() {
  uint64_t dummy = 0;
  AllreduceBits(nullptr, 0, &dummy, 1);
}

void Transport::AllgatherInts(const int64_t* mine, int n, int64_t* out) {
  const int sz = size();
  memcpy(out + (size_t)rank() * n, mine, (size_t)n * 8);
  if (sz == 1) return;
  if (rank() == 0) {
    for (int r = 1; r < sz; ++r) Recv(r, out + (size_t)r * n, (size_t)n * 8);
    for (int r = 1; r < sz; ++r) Send(r, out, (size_t)n * 8 * sz);
  } else {
    Send(0, mine, (size_t)n * 8);
    Recv(0, out, (size_t)n * 8 * sz);
  }
}

std::shared_ptr<Transport> Transport::Split(const std::vector<int>& ranks) {
  auto it = std::find(ranks.begin(), ranks.end(), rank());
  if (it == ranks.end()) return nullptr;
  return std::make_shared<SubTransport>(this, ranks, (int)(it - ranks.begin()));
}

// ---------------------------------------------------------------------------
// Loopback hub: per ordered (src,dst) pair a byte queue.

class LoopbackHub {
 public:
  explicit LoopbackHub(int n) : n_(n), q_((size_t)n * n) {}
  int size() const { return n_; }
  void Push(int src, int dst, const void* p, size_t n) {
    auto& q = q_[(size_t)src * n_ + dst];
    std::lock_guard<std::mutex> l(q.m);
    auto* c = (const uint8_t*)p;
    q.bytes.insert(q.bytes.end(), c, c + n);
    q.cv.notify_all();
  }
  void Pop(int src, int dst, void* p, size_t n) {
    auto& q = q_[(size_t)src * n_ + dst];
    std::unique_lock<std::mutex> l(q.m);
    q.cv.wait(l, [&] { return q.bytes.size() >= n || closed_; });
    if (q.bytes.size() < n) throw TransportError("loopback hub closed");
    std::copy(q.bytes.begin(), q.bytes.begin() + n, (uint8_t*)p);
    q.bytes.erase(q.bytes.begin(), q.bytes.begin() + n);
  }
  void Close() {
    closed_ = true;
    for (auto& q : q_) { std::lock_guard<std::mutex> l(q.m); q.cv.notify_all(); }
  }

 private:
  struct Q { std::mutex m; std::condition_variable cv; std::deque<uint8_t> bytes; };
  int n_;
  std::vector<Q> q_;
  volatile bool closed_ = false;
};

namespace {
class LoopbackTransport : public Transport {
 public:
  LoopbackTransport(std::shared_ptr<LoopbackHub> hub, int rank, std::vector<int> hosts = {})
      : hub_(std::move(hub)), rank_(rank), hosts_(std::move(hosts)) {}
  int rank() const override { return rank_; }
  int size() const override { return hub_->size(); }
  // optional host table: lets unit tests present the ranks of one process as several hosts (two-level planes)
  bool single_host() const override {
    for (int h : hosts_) if (h != hosts_[0]) return false;
    return true;
  }
  int host_id(int i) const override { return i >= 0 && i < (int)hosts_.size() ? hosts_[(size_t)i] : 0; }
  void Send(int peer, const void* b, size_t n) override { hub_->Push(rank_, peer, b, n); }
  void Recv(int peer, void* b, size_t n) override { hub_->Pop(peer, rank_, b, n); }
  void SendRecv(int sp, const void* sb, size_t sn, int rp, void* rb, size_t rn) override {
    if (sn) hub_->Push(rank_, sp, sb, sn);  // queues are unbounded: send never blocks
    if (rn) hub_->Pop(rp, rank_, rb, rn);
  }

 private:
  std::shared_ptr<LoopbackHub> hub_;
  int rank_;
  std::vector<int> hosts_;
};
}  // namespace

std::shared_ptr<LoopbackHub> CreateLoopbackHub(int size) { return std::make_shared<LoopbackHub>(size); }
std::shared_ptr<Transport> LoopbackEndpoint(std::shared_ptr<LoopbackHub> hub, int rank, std::vector<int> hosts) {
  return std::make_shared<LoopbackTransport>(std::move(hub), rank, std::move(hosts));
}

}  // namespace hvd
