#include "nvtx_op_range.h"
#include <nvtx3/nvToolsExt.h>
#include "env.h"

namespace hvd {
namespace {
struct Domain {
  bool enabled = false;
  nvtxDomainHandle_t dom = nullptr;
  // timeline mirror: same two domains as the reference (timeline.cc:332-425) so existing Nsight filters keep working
  nvtxDomainHandle_t tl_dom = nullptr, tl_act_dom = nullptr;
  nvtxStringHandle_t names[(int)NvtxOp::COUNT] = {};
  Domain() {
    enabled = !EnvBool("HOROVOD_DISABLE_NVTX_RANGES", false);
    if (!enabled) return;
    dom = nvtxDomainCreateA("hvd");
    tl_dom = nvtxDomainCreateA("HorovodTimeline");
    tl_act_dom = nvtxDomainCreateA("HorovodTimelineActivities");
    static const char* kNames[] = {"HorovodAllreduce", "HorovodGroupedAllreduce", "HorovodAllgather", "HorovodGroupedAllgather",
                                   "HorovodBroadcast", "HorovodAlltoall", "HorovodReducescatter", "HorovodGroupedReducescatter",
                                   "HorovodJoin", "HorovodBarrier", "HorovodAdasum"};
    for (int i = 0; i < (int)NvtxOp::COUNT; ++i) names[i] = nvtxDomainRegisterStringA(dom, kNames[i]);
  }
};
Domain& D() { static Domain d; return d; }
}  // namespace

bool NvtxEnabled() { return D().enabled; }

void NvtxOpRange::Start(NvtxOp op, int64_t payload_bytes) {
  Domain& d = D();
  if (!d.enabled || active_) return;
  nvtxEventAttributes_t a = {};
  a.version = NVTX_VERSION;
  a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
  a.messageType = NVTX_MESSAGE_TYPE_REGISTERED;
  a.message.registered = d.names[(int)op];
  a.payloadType = NVTX_PAYLOAD_TYPE_INT64;
  a.payload.llValue = payload_bytes;
  id_ = nvtxDomainRangeStartEx(d.dom, &a);
  active_ = true;
}

void NvtxOpRange::End() {
  if (!active_) return;
  nvtxDomainRangeEnd(D().dom, id_);
  active_ = false;
}

uint64_t NvtxRangeStart(const std::string& message, bool activity) {
  Domain& d = D();
  if (!d.enabled) return 0;
  nvtxDomainHandle_t dom = activity ? d.tl_act_dom : d.tl_dom;
  nvtxEventAttributes_t a = {};
  a.version = NVTX_VERSION;
  a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
  a.messageType = NVTX_MESSAGE_TYPE_ASCII;
  a.message.ascii = message.c_str();
  return (uint64_t)nvtxDomainRangeStartEx(dom, &a) + 1;  // +1: 0 stays "no range"
}
void NvtxRangeEnd(uint64_t id, bool activity) {
  if (id == 0) return;
  nvtxDomainRangeEnd(activity ? D().tl_act_dom : D().tl_dom, (nvtxRangeId_t)(id - 1));
}

void NvtxMark(const char* message) {
  Domain& d = D();
  if (!d.enabled) return;
  nvtxEventAttributes_t a = {};
  a.version = NVTX_VERSION;
  a.size = NVTX_EVENT_ATTRIB_STRUCT_SIZE;
  a.messageType = NVTX_MESSAGE_TYPE_ASCII;
  a.message.ascii = message;
  nvtxDomainMarkEx(d.dom, &a);
}

}  // namespace hvd
