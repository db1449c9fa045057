// NVTX ranges around every collective, from enqueue to completion, in an "hvd" domain with registered strings, plus
// a mirror of timeline activities.  Disabled with HOROVOD_DISABLE_NVTX_RANGES=1.
// Parity: horovod/common/nvtx_op_range.{h,cc} + the NVTX mirror in timeline.cc:332-425.
#pragma once
#include <cstdint>
#include <string>

namespace hvd {

enum class NvtxOp : int { ALLREDUCE = 0, GROUPED_ALLREDUCE, ALLGATHER, GROUPED_ALLGATHER, BROADCAST, ALLTOALL, REDUCESCATTER,
                          GROUPED_REDUCESCATTER, JOIN, BARRIER, ADASUM, COUNT };

class NvtxOpRange {
 public:
  NvtxOpRange() = default;
  // starts a range `<op name>` carrying the payload size; no-op when NVTX is disabled or no tool is attached
  void Start(NvtxOp op, int64_t payload_bytes);
  void End();
  ~NvtxOpRange() { End(); }
  NvtxOpRange(const NvtxOpRange&) = delete;
  NvtxOpRange& operator=(const NvtxOpRange&) = delete;

 private:
  uint64_t id_ = 0;
  bool active_ = false;
};

// Free-form ranges of the timeline mirror ("<tensor>: <op or activity>") in the reference's domains: top-level ops in
// "HorovodTimeline", nested activities in "HorovodTimelineActivities"; 0 = not started (NVTX off).
uint64_t NvtxRangeStart(const std::string& message, bool activity);
void NvtxRangeEnd(uint64_t id, bool activity);

// instant marker in the hvd domain (cycle starts, autotune changes)
void NvtxMark(const char* message);
bool NvtxEnabled();

}  // namespace hvd
