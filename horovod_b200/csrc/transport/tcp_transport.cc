// Full-mesh TCP transport + HTTP rendezvous client.
//
// Parity: horovod/common/gloo/gloo_context.cc:67-94 (Rendezvous → full mesh
// through an HTTP store) and gloo/http_store.cc (PUT/GET/DELETE, 404 polling).
// No gloo / HTTPRequest dependency: plain BSD sockets.
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>
#include <chrono>
#include <cstring>
#include <sstream>
#include <thread>
#include "../common/logging.h"
#include "transport.h"
#include <algorithm>

namespace hvd {
namespace {

double Now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int ConnectTo(const std::string& host, int port, double timeout_s) {
  struct addrinfo hints {}, *res = nullptr;
  hints.ai_family = AF_INET;
  hints.ai_socktype = SOCK_STREAM;
  std::string ports = std::to_string(port);
  double deadline = Now() + timeout_s;
  while (true) {
    if (getaddrinfo(host.c_str(), ports.c_str(), &hints, &res) == 0 && res) {
      int fd = socket(res->ai_family, res->ai_socktype, res->ai_protocol);
      if (fd >= 0) {
        if (connect(fd, res->ai_addr, res->ai_addrlen) == 0) {
          freeaddrinfo(res);
          int one = 1;
          setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
          return fd;
        }
        close(fd);
      }
      freeaddrinfo(res);
      res = nullptr;
    }
    if (Now() > deadline) throw TransportError("connect to " + host + ":" + ports + " timed out");
    std::this_thread::sleep_for(std::chrono::milliseconds(20));
  }
}

void WriteAll(int fd, const void* buf, size_t n) {
  auto* p = (const char*)buf;
  while (n) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k < 0) {
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK) { struct pollfd pf {fd, POLLOUT, 0}; poll(&pf, 1, 1000); continue; }
      throw TransportError(std::string("send failed: ") + strerror(errno));
    }
    p += k; n -= (size_t)k;
  }
}

// A peer that is alive but frozen (SIGSTOP, a hung node) never closes its socket: with HVD_TCP_TIMEOUT_SECONDS (default: the
// value of HVD_SHM_TIMEOUT_SECONDS, else 0 = wait forever) a receive that makes no progress for that long fails instead of
// blocking the cycle thread forever (the reference's Gloo transport gives up after HOROVOD_GLOO_TIMEOUT_SECONDS).
double RecvTimeoutSeconds() {
  static const double t = [] {
    const char* e = getenv("HVD_TCP_TIMEOUT_SECONDS");
    if (!e) e = getenv("HVD_SHM_TIMEOUT_SECONDS");
    if (!e) e = getenv("HOROVOD_GLOO_TIMEOUT_SECONDS");    // hvdrun --gloo-timeout-seconds: what the reference's Gloo ops honour
    return e ? std::max(0.0, atof(e)) : 0.0;
  }();
  return t;
}

// Negotiation messages and small payloads are answered within tens of microseconds; a blocking recv() pays a sleep / wake-up
// round trip through the scheduler that is longer than that.  So every receive first polls the socket without blocking for
// HVD_TCP_SPIN_US microseconds (default 50; 0 = block immediately), then falls back to the blocking path.
double SpinSeconds() {
  static const double t = [] {
    const char* e = getenv("HVD_TCP_SPIN_US");
    return (e ? std::max(0.0, atof(e)) : 50.0) * 1e-6;
  }();
  return t;
}

void ReadAll(int fd, void* buf, size_t n) {
  auto* p = (char*)buf;
  const double timeout_s = RecvTimeoutSeconds();
  const double spin_s = SpinSeconds();
  if (spin_s > 0) {
    const double until = Now() + spin_s;
    while (n) {
      ssize_t k = ::recv(fd, p, n, MSG_DONTWAIT);
      if (k > 0) { p += k; n -= (size_t)k; continue; }
      if (k == 0) throw TransportError("peer closed connection");
      if (errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR) throw TransportError(std::string("recv failed: ") + strerror(errno));
      if (Now() > until) break;
      __builtin_ia32_pause();
    }
  }
  while (n) {
    if (timeout_s > 0) {
      struct pollfd pf {fd, POLLIN, 0};
      int rc = poll(&pf, 1, (int)(timeout_s * 1000));
      if (rc == 0) throw TransportError("a peer sent nothing for " + std::to_string((int)timeout_s) + " s (HVD_TCP_TIMEOUT_SECONDS): frozen or unreachable");
      if (rc < 0) { if (errno == EINTR) continue; throw TransportError(std::string("poll failed: ") + strerror(errno)); }
    }
    ssize_t k = ::recv(fd, p, n, 0);
    if (k == 0) throw TransportError("peer closed connection");
    if (k < 0) {
      if (errno == EINTR) continue;
      if (errno == EAGAIN || errno == EWOULDBLOCK) { struct pollfd pf {fd, POLLIN, 0}; poll(&pf, 1, 1000); continue; }
      throw TransportError(std::string("recv failed: ") + strerror(errno));
    }
    p += k; n -= (size_t)k;
  }
}

class TcpTransport : public Transport {
 public:
  TcpTransport(int rank, int size) : fds_(size, -1), rank_(rank), size_(size) {}
  ~TcpTransport() override {
    for (int fd : fds_) if (fd >= 0) { shutdown(fd, SHUT_RDWR); close(fd); }
  }
  int rank() const override { return rank_; }
  int size() const override { return size_; }
  bool single_host() const override { return single_host_; }
  int host_id(int i) const override { return i >= 0 && i < (int)host_ids_.size() ? host_ids_[i] : 0; }
  void Send(int peer, const void* b, size_t n) override { WriteAll(fd(peer), b, n); }
  void Recv(int peer, void* b, size_t n) override { ReadAll(fd(peer), b, n); }
  void SendRecv(int sp, const void* sbuf, size_t sn, int rp, void* rbuf, size_t rn) override {
    if (sp == rank_ && rp == rank_) { if (rn) memcpy(rbuf, sbuf, std::min(sn, rn)); return; }
    if (sn == 0) { if (rn) Recv(rp, rbuf, rn); return; }
    if (rn == 0) { Send(sp, sbuf, sn); return; }
    int sfd = fd(sp), rfd = fd(rp);
    auto* s = (const char*)sbuf; auto* r = (char*)rbuf;
    const double timeout_s = RecvTimeoutSeconds();
    double last_progress = Now();
    double spin_until = SpinSeconds() > 0 ? last_progress + SpinSeconds() : 0;
    while (sn || rn) {
      struct pollfd pf[2]; int np = 0, si = -1, ri = -1;
      if (sn) { pf[np] = {sfd, POLLOUT, 0}; si = np++; }
      if (rn) { pf[np] = {rfd, POLLIN, 0}; ri = np++; }
      // same reasoning as in ReadAll: the first look at the sockets does not sleep
      int rc = poll(pf, np, 0);
      if (rc == 0 && spin_until > 0 && Now() < spin_until) { __builtin_ia32_pause(); continue; }
      if (rc == 0) rc = poll(pf, np, timeout_s > 0 ? (int)std::min(5000.0, timeout_s * 1000) : 5000);
      if (rc < 0) { if (errno == EINTR) continue; throw TransportError("poll failed"); }
      if (rc == 0) {
        if (timeout_s > 0 && Now() - last_progress > timeout_s)
          throw TransportError("no progress on a send/receive pair for " + std::to_string((int)timeout_s) + " s (HVD_TCP_TIMEOUT_SECONDS): peer frozen or unreachable");
        continue;
      }
      last_progress = Now();
      if (spin_until > 0) spin_until = last_progress + SpinSeconds();
      if (ri >= 0 && (pf[ri].revents & (POLLIN | POLLHUP | POLLERR))) {
        ssize_t k = ::recv(rfd, r, rn, MSG_DONTWAIT);
        if (k == 0) throw TransportError("peer closed connection");
        if (k < 0) { if (errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR) throw TransportError(std::string("recv failed: ") + strerror(errno)); }
        else { r += k; rn -= (size_t)k; }
      }
      if (si >= 0 && (pf[si].revents & (POLLOUT | POLLHUP | POLLERR))) {
        ssize_t k = ::send(sfd, s, sn, MSG_DONTWAIT | MSG_NOSIGNAL);
        if (k < 0) { if (errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR) throw TransportError(std::string("send failed: ") + strerror(errno)); }
        else { s += k; sn -= (size_t)k; }
      }
    }
  }
  // All peers at once: every socket with bytes left to send or receive sits in ONE poll set, so a small exchange costs one
  // message time instead of n - 1 and large blocks stream over all connections concurrently.
  void AlltoallvBytes(const char* in, const int64_t* sd, char* out, const int64_t* rd, const uint8_t* skip) override {
    static const bool concurrent = [] { const char* e = getenv("HVD_TCP_ALLTOALL_CONCURRENT"); return !e || atoi(e) != 0; }();
    if (!concurrent) { Transport::AlltoallvBytes(in, sd, out, rd, skip); return; }
    struct Leg { int fd; const char* s; size_t sn; char* r; size_t rn; };
    std::vector<Leg> legs;
    for (int k = 1; k < size_; ++k) {              // rotation order: the first sockets polled differ from rank to rank
      const int p = (rank_ + k) % size_;
      if (skip && skip[p]) continue;
      Leg l{fd(p), in + sd[p], (size_t)(sd[p + 1] - sd[p]), out + rd[p], (size_t)(rd[p + 1] - rd[p])};
      if (l.sn || l.rn) legs.push_back(l);
    }
    const double timeout_s = RecvTimeoutSeconds();
    double last_progress = Now();
    double spin_until = SpinSeconds() > 0 ? last_progress + SpinSeconds() : 0;
    std::vector<struct pollfd> pf;
    while (true) {
      pf.clear();
      for (auto& l : legs) if (l.sn || l.rn) pf.push_back({l.fd, (short)((l.sn ? POLLOUT : 0) | (l.rn ? POLLIN : 0)), 0});
      if (pf.empty()) return;
      int rc = poll(pf.data(), (nfds_t)pf.size(), 0);
      if (rc == 0 && spin_until > 0 && Now() < spin_until) { __builtin_ia32_pause(); continue; }
      if (rc == 0) rc = poll(pf.data(), (nfds_t)pf.size(), timeout_s > 0 ? (int)std::min(5000.0, timeout_s * 1000) : 5000);
      if (rc < 0) { if (errno == EINTR) continue; throw TransportError("poll failed"); }
      if (rc == 0) {
        if (timeout_s > 0 && Now() - last_progress > timeout_s)
          throw TransportError("no progress on an all-to-all exchange for " + std::to_string((int)timeout_s) + " s (HVD_TCP_TIMEOUT_SECONDS): peer frozen or unreachable");
        continue;
      }
      size_t i = 0;
      for (auto& l : legs) {
        if (!(l.sn || l.rn)) continue;
        const short ev = pf[i++].revents;
        if (l.rn && (ev & (POLLIN | POLLHUP | POLLERR))) {
          ssize_t k = ::recv(l.fd, l.r, l.rn, MSG_DONTWAIT);
          if (k == 0) throw TransportError("peer closed connection");
          if (k < 0) { if (errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR) throw TransportError(std::string("recv failed: ") + strerror(errno)); }
          else { l.r += k; l.rn -= (size_t)k; }
        }
        if (l.sn && (ev & (POLLOUT | POLLHUP | POLLERR))) {
          ssize_t k = ::send(l.fd, l.s, l.sn, MSG_DONTWAIT | MSG_NOSIGNAL);
          if (k < 0) { if (errno != EAGAIN && errno != EWOULDBLOCK && errno != EINTR) throw TransportError(std::string("send failed: ") + strerror(errno)); }
          else { l.s += k; l.sn -= (size_t)k; }
        }
      }
      last_progress = Now();
      if (spin_until > 0) spin_until = last_progress + SpinSeconds();
    }
  }
  std::vector<int> fds_;
  bool single_host_ = true;
  std::vector<int> host_ids_;

 private:
  int fd(int peer) const {
    if (peer < 0 || peer >= size_ || fds_[peer] < 0) throw TransportError("no connection to rank " + std::to_string(peer));
    return fds_[peer];
  }
  int rank_, size_;
};

}  // namespace

std::shared_ptr<Transport> CreateTcpTransport(int rank, int size, KVStore* store, const std::string& scope,
                                              const std::string& advertise_addr, double timeout_s,
                                              const std::vector<std::string>& hostnames) {
  auto t = std::make_shared<TcpTransport>(rank, size);
  if (size == 1) return t;
  // listen on an ephemeral port
  int lfd = socket(AF_INET, SOCK_STREAM, 0);
  if (lfd < 0) throw TransportError("socket() failed");
  int one = 1;
  setsockopt(lfd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
  struct sockaddr_in addr {};
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_ANY);
  addr.sin_port = 0;
  if (bind(lfd, (struct sockaddr*)&addr, sizeof addr) != 0 || listen(lfd, size + 8) != 0) {
    close(lfd);
    throw TransportError("bind/listen failed");
  }
  socklen_t alen = sizeof addr;
  getsockname(lfd, (struct sockaddr*)&addr, &alen);
  int port = ntohs(addr.sin_port);
  store->Set(scope, "addr." + std::to_string(rank), advertise_addr + ":" + std::to_string(port));

  // connect to every lower rank, accept from every higher rank
  std::vector<std::string> hosts(size);
  hosts[rank] = advertise_addr;
  for (int p = 0; p < rank; ++p) {
    std::string a = store->Get(scope, "addr." + std::to_string(p), timeout_s);
    auto c = a.rfind(':');
    std::string host = a.substr(0, c);
    hosts[p] = host;
    int pport = atoi(a.substr(c + 1).c_str());
    int fd = ConnectTo(host, pport, timeout_s);
    int32_t me = rank;
    WriteAll(fd, &me, 4);
    t->fds_[p] = fd;
  }
  double deadline = Now() + timeout_s;
  for (int k = rank + 1; k < size; ++k) {
    struct pollfd pf {lfd, POLLIN, 0};
    while (true) {
      int rc = poll(&pf, 1, 200);
      if (rc > 0) break;
      if (Now() > deadline) { close(lfd); throw TransportError("timed out waiting for peers to connect (rendezvous)"); }
    }
    int fd = accept(lfd, nullptr, nullptr);
    if (fd < 0) { close(lfd); throw TransportError("accept failed"); }
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    int32_t who = -1;
    ReadAll(fd, &who, 4);
    if (who <= rank || who >= size || t->fds_[who] >= 0) { close(fd); close(lfd); throw TransportError("bad rendezvous handshake"); }
    t->fds_[who] = fd;
  }
  close(lfd);
  // big socket buffers for the CPU data plane
  for (int fd : t->fds_) {
    if (fd < 0) continue;
    int sz = 4 << 20;
    setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &sz, sizeof sz);
    setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &sz, sizeof sz);
  }
  if (!hostnames.empty()) {
    for (auto& h : hostnames) if (h != hostnames[rank]) t->single_host_ = false;
    std::vector<std::string> uniq;
    for (auto& h : hostnames) {
      auto it = std::find(uniq.begin(), uniq.end(), h);
      t->host_ids_.push_back((int)(it - uniq.begin()));
      if (it == uniq.end()) uniq.push_back(h);
    }
  }
  return t;
}

// ---------------------------------------------------------------------------
// HTTP KV client (HTTP/1.0, Connection: close, one request per socket)

int HttpKVStore::Request(const std::string& method, const std::string& path, const std::string& body,
                         std::string* resp) {
  int fd = ConnectTo(host_, port_, 30.0);
  std::ostringstream os;
  os << method << " " << path << " HTTP/1.0\r\nHost: " << host_ << "\r\nContent-Length: " << body.size()
     << "\r\nConnection: close\r\n\r\n";
  std::string head = os.str();
  try {
    WriteAll(fd, head.data(), head.size());
    if (!body.empty()) WriteAll(fd, body.data(), body.size());
  } catch (...) { close(fd); throw; }
  std::string all;
  char buf[4096];
  while (true) {
    ssize_t k = ::recv(fd, buf, sizeof buf, 0);
    if (k < 0 && errno == EINTR) continue;
    if (k <= 0) break;
    all.append(buf, (size_t)k);
  }
  close(fd);
  int status = 0;
  if (all.size() > 12) status = atoi(all.c_str() + 9);
  auto pos = all.find("\r\n\r\n");
  if (resp) *resp = pos == std::string::npos ? "" : all.substr(pos + 4);
  return status;
}

void HttpKVStore::Set(const std::string& scope, const std::string& key, const std::string& value) {
  for (int attempt = 0; attempt < 3; ++attempt) {
    int st = Request("PUT", "/" + scope + "/" + key, value, nullptr);
    if (st == 200) return;
    std::this_thread::sleep_for(std::chrono::milliseconds(500));
  }
  throw TransportError("rendezvous PUT failed for " + scope + "/" + key);
}

std::string HttpKVStore::Get(const std::string& scope, const std::string& key, double timeout_s) {
  double deadline = Now() + timeout_s;
  std::string out;
  while (true) {
    int st = Request("GET", "/" + scope + "/" + key, "", &out);
    if (st == 200) return out;
    if (Now() > deadline) throw TransportError("rendezvous GET timed out for " + scope + "/" + key);
    std::this_thread::sleep_for(std::chrono::milliseconds(10));
  }
}

void HttpKVStore::Finalize(const std::string& scope) {
  try { Request("DELETE", "/" + scope + "/", "", nullptr); } catch (...) {}
}

std::string HttpKVStore::LocalAddress() {
  try {
    int fd = ConnectTo(host_, port_, 10.0);
    struct sockaddr_in a {};
    socklen_t l = sizeof a;
    getsockname(fd, (struct sockaddr*)&a, &l);
    close(fd);
    char buf[64];
    inet_ntop(AF_INET, &a.sin_addr, buf, sizeof buf);
    return buf;
  } catch (...) { return "127.0.0.1"; }
}

}  // namespace hvd
