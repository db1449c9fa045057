// Chrome-tracing timeline (chrome://tracing / Perfetto JSON).  Each tensor is a
// "process" row; its lifetime goes NEGOTIATING -> TOP_LEVEL -> ACTIVITY.
// Records go through a bounded lock-free single-producer / single-consumer ring
// (producers are serialised by the state-machine mutex, exactly like the
// reference's TimelineWriter behind Timeline's mutex) to a writer thread, so the
// cycle thread never blocks on file IO or on the writer.  The file is a complete
// JSON document after every drain (the closing bracket is rewritten in place), so
// a trace of a crashed or still running job loads.  Every activity is mirrored as
// an NVTX range in the "hvd" domain, and GPU collectives additionally get a
// device-timed row ("GPU" thread of the tensor) from CUDA events recorded around
// their kernels.  Can be started/stopped at runtime (hvd.start_timeline /
// stop_timeline).
// Parity: horovod/common/timeline.{h,cc} (boost::lockfree spsc_queue + NVTX mirror
// :332-425 there); activity names follow common.h:80-114.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include "common.h"
#include "message.h"

namespace hvd {

// activity names
#define HVD_ACT_WAIT_FOR_DATA "WAIT_FOR_DATA"
#define HVD_ACT_QUEUE "QUEUE"
#define HVD_ACT_MEMCPY_IN_FUSION_BUFFER "MEMCPY_IN_FUSION_BUFFER"
#define HVD_ACT_MEMCPY_OUT_FUSION_BUFFER "MEMCPY_OUT_FUSION_BUFFER"
#define HVD_ACT_P2P_ALLREDUCE_ONESHOT "P2P_ALLREDUCE_ONESHOT"
#define HVD_ACT_P2P_ALLREDUCE_TWOSHOT "P2P_ALLREDUCE_TWOSHOT"
#define HVD_ACT_P2P_ALLREDUCE_NVLS "P2P_ALLREDUCE_NVLS"
#define HVD_ACT_HIER_ALLREDUCE "HIERARCHICAL_ALLREDUCE"
#define HVD_ACT_P2P_ALLGATHER "P2P_ALLGATHER"
#define HVD_ACT_P2P_BROADCAST "P2P_BROADCAST"
#define HVD_ACT_P2P_ALLTOALL "P2P_ALLTOALL"
#define HVD_ACT_P2P_REDUCESCATTER "P2P_REDUCESCATTER"
#define HVD_ACT_P2P_ADASUM "P2P_ADASUM"
#define HVD_ACT_NCCL_ALLREDUCE "NCCL_ALLREDUCE"
#define HVD_ACT_CPU_ALLREDUCE "CPU_ALLREDUCE"
#define HVD_ACT_CPU_ALLGATHER "CPU_ALLGATHER"
#define HVD_ACT_CPU_BROADCAST "CPU_BROADCAST"
#define HVD_ACT_CPU_ALLTOALL "CPU_ALLTOALL"
#define HVD_ACT_CPU_REDUCESCATTER "CPU_REDUCESCATTER"
#define HVD_ACT_CPU_ADASUM "CPU_ADASUM"

class Timeline {
 public:
  ~Timeline() { Shutdown(); }
  void Initialize(const std::string& file, int world_size);
  void Shutdown();
  bool Initialized() const { return initialized_.load(std::memory_order_acquire); }
  void SetMarkCycles(bool v) { mark_cycles_ = v; }

  void NegotiateStart(const std::string& name, RequestType type);
  void NegotiateRankReady(const std::string& name, int rank);
  void NegotiateEnd(const std::string& name);
  void Start(const std::string& name, ResponseType type, size_t bytes = 0);
  void ActivityStart(const std::string& name, const std::string& activity);
  void ActivityEnd(const std::string& name);
  void ActivityStartAll(const std::vector<std::shared_ptr<TensorTableEntry>>& es, const std::string& activity);
  void ActivityEndAll(const std::vector<std::shared_ptr<TensorTableEntry>>& es);
  void End(const std::string& name, const std::string& args = "");
  void MarkCycleStart();
  // Device-timed span of a GPU collective: [start_us, start_us + dur_us) on the timeline clock, measured with CUDA events
  // on the op's stream; shown on the "GPU" thread of every tensor of the response.
  void DeviceSpan(const std::vector<std::string>& names, const std::string& activity, int64_t start_us, int64_t dur_us);
  // timeline clock (microseconds since Initialize) for a host timestamp taken with NowNs()
  int64_t ToTimelineUs(uint64_t host_ns) const { return (int64_t)((host_ns - start_ns_) / 1000); }
  uint64_t session_start_ns() const { return start_ns_; }
  uint64_t dropped_records() const { return dropped_.load(std::memory_order_relaxed); }

 private:
  enum class State { UNKNOWN, NEGOTIATING, TOP_LEVEL, ACTIVITY };
  struct Record { char phase; int pid; std::string name; std::string args; int64_t ts_us; bool meta = false; int64_t dur_us = 0; int tid = 0; };
  void Push(Record r);
  int Pid(const std::string& tensor_name);  // allocates + emits process_name metadata
  void WriterLoop();
  int64_t NowUs() const;

  std::atomic<bool> initialized_{false};
  bool mark_cycles_ = false;
  std::mutex mu_;  // producers (cycle thread + finalizer threads)
  std::unordered_map<std::string, int> pids_;
  std::unordered_map<std::string, State> states_;
  std::unordered_map<std::string, uint64_t> nvtx_top_, nvtx_act_;  // NVTX mirror: open range ids per tensor
  uint64_t start_ns_ = 0;
  // bounded lock-free SPSC ring: head_ is written by the (mutex-serialised) producers, tail_ by the writer thread
  std::vector<Record> ring_;
  std::atomic<size_t> head_{0}, tail_{0};
  std::atomic<uint64_t> dropped_{0};
  std::thread writer_;
  std::atomic<bool> stop_{false};
  FILE* file_ = nullptr;
  bool first_record_ = true;
  bool closed_ = false;  // the file currently ends with the closing bracket
};

}  // namespace hvd
