// Control-plane + CPU data-plane transport abstraction.
//
// A Transport is a communicator over an ordered set of ranks.  Implementations
// provide blocking point-to-point byte moves; the collectives the controller
// needs (gather/bcast of serialized lists, bit-vector AND/OR, barrier,
// fixed-size allgather) have default star-shaped implementations on top and can
// be overridden (the shared-memory transport overrides the bit-vector
// allreduce and the barrier with CPU atomics).
//
// Parity: the virtual hooks of horovod/common/controller.h:143-156 and the
// MPI/Gloo controllers (mpi/mpi_controller.cc, gloo/gloo_controller.cc); the
// reference depends on MPI or Gloo for these, this is self-contained.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace hvd {

class TransportError : public std::runtime_error {
 public:
  explicit TransportError(const std::string& m) : std::runtime_error(m) {}
};

// Shared-memory data plane of a single-host communicator: every rank owns two slots ("halves") of `slot_bytes` in one
// POSIX shm segment that all ranks map.  Collectives move a message through it in pieces; piece k uses half k & 1, and
// because every piece contains a barrier after its writes, a half is never overwritten while a peer still reads it.
struct ShmData {
  char* base = nullptr;
  size_t slot_bytes = 0;
  char* slot(int rank, int half) const { return base + ((size_t)rank * 2 + (size_t)half) * slot_bytes; }
};

class Transport;
// Two-level data plane of a multi-host communicator whose hosts all run the same number of ranks: `local` are the shm slots
// of the ranks of THIS host (indexed by local rank), `cross` connects the ranks with my local index on every host, and
// `column[l]` lists, host by host, the communicator ranks with local index l.
struct HierData {
  ShmData local;
  int local_rank = 0, local_size = 1;
  Transport* cross = nullptr;
  const std::vector<std::vector<int>>* column = nullptr;
};

class Transport {
 public:
  virtual ~Transport() = default;
  virtual int rank() const = 0;
  virtual int size() const = 0;
  // Global rank (in the root communicator) of local index i.
  virtual int global_rank(int i) const { return i; }

  // ---- point to point (blocking) ----
  virtual void Send(int peer, const void* buf, size_t n) = 0;
  virtual void Recv(int peer, void* buf, size_t n) = 0;
  // Full-duplex exchange that cannot deadlock when both sides send first.
  virtual void SendRecv(int send_peer, const void* sbuf, size_t sn, int recv_peer, void* rbuf, size_t rn) = 0;

  // Personalised exchange with every other rank: byte range [sd[p], sd[p+1]) of `in` goes to rank p, [rd[p], rd[p+1]) of `out`
  // comes from rank p (the caller copies its own block).  Default: n - 1 rounds of SendRecv (to r + k, from r - k); the TCP
  // mesh progresses all peers at once.
  // `skip[p] != 0` leaves rank p out (its blocks travel some other way).
  virtual void AlltoallvBytes(const char* in, const int64_t* sd, char* out, const int64_t* rd, const uint8_t* skip = nullptr);

  // ---- collectives (default: star through `root`) ----
  virtual void GatherBytes(const std::vector<uint8_t>& mine, std::vector<std::vector<uint8_t>>* all, int root = 0);
  virtual void BcastBytes(std::vector<uint8_t>* buf, int root = 0);
  // In-place: and_words <- AND over ranks, or_words <- OR over ranks.
  virtual void AllreduceBits(uint64_t* and_words, int n_and, uint64_t* or_words, int n_or);
  // The same reduction among a subset: `peers` are ranks of THIS communicator (identical list on every member), `me` is the
  // caller's position in it, `words` holds n_and AND-words followed by n - n_and OR-words.  Star through peers[0] for small
  // groups, recursive doubling (log2 rounds of pairwise exchanges) from HVD_BITS_TREE_MIN_RANKS members on.
  void AllreduceBitsAmong(const std::vector<int>& peers, int me, uint64_t* words, int n_and, int n);
  virtual void Barrier();
  virtual void AllgatherInts(const int64_t* mine, int n, int64_t* out);
  virtual void Bcast(void* buf, size_t n, int root);

  // Sub-communicator over `ranks` (indices in *this* communicator). Every
  // member must call with the same list; non-members get nullptr.
  virtual std::shared_ptr<Transport> Split(const std::vector<int>& ranks);

  // true when every rank of this communicator lives on the same host
  virtual bool single_host() const { return true; }
  // Small integer naming the host of local index i (equal ids <=> same host); used to derive the intra-host and
  // cross-host sub-communicators of the hierarchical GPU collectives.
  virtual int host_id(int /*i*/) const { return 0; }

  // Shared-memory data plane (see ShmData); false when this communicator has none.  ShmNextPiece() returns the running
  // piece number (identical on every rank because collectives are issued in the same order everywhere).
  // one line for logs / hvd.control_plane_info(): what the negotiation and the host data path run on
  virtual std::string Describe() const { return "control: point-to-point star over the base transport; host data: ring over the base transport"; }
  virtual bool HierDataPlane(HierData* /*out*/) { return false; }
  virtual void LocalBarrier() {}     // among the ranks of this host only (HierDataPlane users)
  virtual bool ShmDataPlane(ShmData* /*out*/) { return false; }
  virtual uint64_t ShmNextPiece() { return 0; }
};

// View of a parent transport restricted to a rank subset.
class SubTransport : public Transport {
 public:
  SubTransport(Transport* parent, std::vector<int> ranks, int my_index)
      : parent_(parent), ranks_(std::move(ranks)), my_(my_index) {}
  int rank() const override { return my_; }
  int size() const override { return (int)ranks_.size(); }
  int global_rank(int i) const override { return parent_->global_rank(ranks_[i]); }
  void Send(int peer, const void* b, size_t n) override { parent_->Send(ranks_[peer], b, n); }
  void Recv(int peer, void* b, size_t n) override { parent_->Recv(ranks_[peer], b, n); }
  void SendRecv(int sp, const void* sb, size_t sn, int rp, void* rb, size_t rn) override {
    parent_->SendRecv(ranks_[sp], sb, sn, ranks_[rp], rb, rn);
  }
  bool single_host() const override {
    for (int r : ranks_) if (parent_->host_id(r) != parent_->host_id(ranks_[0])) return false;
    return true;
  }
  int host_id(int i) const override { return parent_->host_id(ranks_[i]); }

 protected:
  Transport* parent_;
  std::vector<int> ranks_;
  int my_;
};

// ---------------------------------------------------------------------------
// Rendezvous key-value store (bootstrap only).
class KVStore {
 public:
  virtual ~KVStore() = default;
  virtual void Set(const std::string& scope, const std::string& key, const std::string& value) = 0;
  // Blocks until the key exists or timeout; throws TransportError on timeout.
  virtual std::string Get(const std::string& scope, const std::string& key, double timeout_s) = 0;
  virtual void Finalize(const std::string& scope) {}
};

// Client for the launcher's HTTP KV server (PUT/GET /scope/key; GET polls on
// 404).  Parity: horovod/common/gloo/http_store.{h,cc}.
class HttpKVStore : public KVStore {
 public:
  HttpKVStore(std::string host, int port) : host_(std::move(host)), port_(port) {}
  void Set(const std::string& scope, const std::string& key, const std::string& value) override;
  std::string Get(const std::string& scope, const std::string& key, double timeout_s) override;
  void Finalize(const std::string& scope) override;
  // returns the local address of a socket connected to the server (the NIC
  // that routes to the launcher) — published as this rank's mesh address.
  std::string LocalAddress();

 private:
  // returns HTTP status, fills body
  int Request(const std::string& method, const std::string& path, const std::string& body, std::string* resp);
  std::string host_;
  int port_;
};

// Full-mesh TCP transport.  One socket per peer pair, created at init through
// the KV store (lower rank listens, higher rank connects).
std::shared_ptr<Transport> CreateTcpTransport(int rank, int size, KVStore* store, const std::string& scope,
                                              const std::string& advertise_addr, double timeout_s,
                                              const std::vector<std::string>& hostnames = {});

// In-process loopback hub for single-process multi-"rank" unit tests: N
// transports exchanging through in-memory queues (something the reference has
// no equivalent of — its multi-rank tests always need real MPI/Gloo processes).
class LoopbackHub;
std::shared_ptr<LoopbackHub> CreateLoopbackHub(int size);
// `hosts`: optional host id per rank (unit tests of the two-level planes present one process as several hosts)
std::shared_ptr<Transport> LoopbackEndpoint(std::shared_ptr<LoopbackHub> hub, int rank, std::vector<int> hosts = {});

// Shared-memory control plane overlay: wraps a base transport whose ranks are
// all on one host and replaces AllreduceBits/Barrier with atomics on a POSIX
// shm segment (microseconds instead of a TCP round trip).
std::shared_ptr<Transport> WrapWithShmControl(std::shared_ptr<Transport> base, const std::string& segment_name);
// Multi-host variant: bit vectors are folded per host in shared memory and only the host leaders talk over the base
// transport (2 (H-1) messages per negotiation cycle instead of 2 (N-1)).  Returns `base` for single-host jobs, jobs with
// one rank per host, or when shared memory is unavailable.
std::shared_ptr<Transport> WrapWithHierarchicalControl(std::shared_ptr<Transport> base, const std::string& segment_name);

}  // namespace hvd
