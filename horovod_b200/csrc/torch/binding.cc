// PyTorch binding: wraps at::Tensor arguments into TensorTableEntry objects,
// records a ready event on the caller's current CUDA stream, enqueues into the
// engine and hands back an integer handle.  Completion is event-chained: for
// GPU collectives `wait_and_clear` makes the caller's stream wait on the
// engine's completion event instead of blocking the host.
//
// Capability parity: horovod/torch/mpi_ops_v2.cc (Do{Allreduce,...}, PollHandle,
// WaitAndClear), adapter_v2.cc (TorchTensor / TorchOpContext allocation),
// ready_event.cc (pooled events on the current stream), handle_manager.cc,
// cuda_util.cc (device guard).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <unordered_map>
#include "../common/engine.h"
#include "../kernels/p2p_kernels.h"

namespace py = pybind11;
using namespace hvd;

namespace {

// ---- handle manager --------------------------------------------------------
struct HandleState {
  std::atomic<bool> done{false};
  Status status;
  std::vector<SharedEvent*> events;  // one per fused response that carried members of this handle
  int device = CPU_DEVICE_ID;
  std::vector<int32_t> received_splits;
  int32_t last_joined_rank = -1;
  std::vector<at::Tensor> keep_alive;
  at::Tensor received_splits_out;
  int pending = 1;  // grouped ops complete when every member did
};

class HandleManager {
 public:
  int Allocate(int pending = 1) {
    std::lock_guard<std::mutex> l(mu_);
    int h = next_++;
    auto st = std::make_shared<HandleState>();
    st->pending = pending;
    map_[h] = st;
    return h;
  }
  std::shared_ptr<HandleState> Get(int h) {
    std::lock_guard<std::mutex> l(mu_);
    auto it = map_.find(h);
    return it == map_.end() ? nullptr : it->second;
  }
  void MarkDone(int h, const Completion& c) {
    std::lock_guard<std::mutex> l(mu_);
    auto it = map_.find(h);
    if (it == map_.end()) { if (c.done_event) GpuContext::Get().Release((SharedEvent*)c.done_event); return; }
    auto& st = *it->second;
    if (!c.status.ok() && st.status.ok()) st.status = c.status;
    if (c.done_event) {
      auto* ev = (SharedEvent*)c.done_event;
      bool dup = false;
      for (auto* x : st.events) if (x == ev) dup = true;
      if (dup) GpuContext::Get().Release(ev);  // grouped members fused into one response share the event
      else st.events.push_back(ev);
    }
    if (!c.received_splits.empty()) st.received_splits = c.received_splits;
    if (c.last_joined_rank >= 0) st.last_joined_rank = c.last_joined_rank;
    if (--st.pending <= 0) { st.done = true; cv_.notify_all(); }
  }
  void Wait(const std::shared_ptr<HandleState>& st) {
    // a fused NVLink allreduce completes in tens of microseconds: spin briefly before paying a futex sleep + wake-up
    const auto spin_end = std::chrono::steady_clock::now() + std::chrono::microseconds(200);
    while (!st->done.load(std::memory_order_acquire) && std::chrono::steady_clock::now() < spin_end) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    if (st->done.load(std::memory_order_acquire)) { std::lock_guard<std::mutex> l(mu_); return; }
    std::unique_lock<std::mutex> l(mu_);
    cv_.wait(l, [&] { return st->done.load(); });
  }
  bool Done(const std::shared_ptr<HandleState>& st) { return st->done.load(std::memory_order_acquire); }
  void Release(int h) { std::lock_guard<std::mutex> l(mu_); map_.erase(h); }
  void Reset() {
    std::lock_guard<std::mutex> l(mu_);
    for (auto& kv : map_) for (auto* ev : kv.second->events) GpuContext::Get().Release(ev);
    map_.clear();
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  std::unordered_map<int, std::shared_ptr<HandleState>> map_;
  int next_ = 0;
};
HandleManager g_handles;

// ---- ready events ----------------------------------------------------------
class EventPool {
 public:
  cudaEvent_t Get(int device) {
    {
      std::lock_guard<std::mutex> l(mu_);
      auto& v = pool_[device];
      if (!v.empty()) { auto e = v.back(); v.pop_back(); return e; }
    }
    cudaEvent_t e;
    C10_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return e;
  }
  void Put(int device, cudaEvent_t e) { std::lock_guard<std::mutex> l(mu_); pool_[device].push_back(e); }

 private:
  std::mutex mu_;
  std::unordered_map<int, std::vector<cudaEvent_t>> pool_;
};
EventPool g_events;

DataType MapDtype(at::ScalarType t) {
  switch (t) {
    case at::kByte: return DataType::UINT8;
    case at::kChar: return DataType::INT8;
    case at::kShort: return DataType::INT16;
    case at::kInt: return DataType::INT32;
    case at::kLong: return DataType::INT64;
    case at::kHalf: return DataType::FLOAT16;
    case at::kBFloat16: return DataType::BFLOAT16;
    case at::kFloat: return DataType::FLOAT32;
    case at::kDouble: return DataType::FLOAT64;
    case at::kBool: return DataType::BOOL;
    default: throw std::invalid_argument(std::string("Horovod: unsupported tensor dtype ") + c10::toString(t));
  }
}

int DeviceOf(const at::Tensor& t) { return t.is_cuda() ? (int)t.get_device() : CPU_DEVICE_ID; }

// Unnamed ops are numbered per (op type, process set) — not by handle: handle counters diverge between ranks as soon as
// one rank issues an extra op (around hvd.join(), or inside a process set the other rank is not a member of, e.g. the
// unnamed allreduce in an autograd backward), and the names must match across the ranks of the set.
std::mutex g_noname_mu;
std::unordered_map<std::string, int> g_noname_counters;
std::string OpName(const char* op, const std::string& name, int process_set_id) {
  if (!name.empty()) return std::string(op) + "." + name;
  std::lock_guard<std::mutex> l(g_noname_mu);
  const std::string key = std::string(op) + (process_set_id ? ".ps" + std::to_string(process_set_id) : "");
  return key + ".noname." + std::to_string(g_noname_counters[key]++);
}

// Process-set ids are recycled: a new set must start numbering from zero on every member.
void ResetNonameCounters(int process_set_id) {
  std::lock_guard<std::mutex> l(g_noname_mu);
  const std::string tag = ".ps" + std::to_string(process_set_id);
  for (auto it = g_noname_counters.begin(); it != g_noname_counters.end();) {
    const std::string& k = it->first;
    if (k.size() >= tag.size() && k.compare(k.size() - tag.size(), tag.size(), tag) == 0) it = g_noname_counters.erase(it);
    else ++it;
  }
}

void ThrowIfError(const Status& st) {
  if (st.ok()) return;
  switch (st.type()) {
    case StatusType::INVALID_ARGUMENT: throw std::invalid_argument(st.reason());
    case StatusType::PRECONDITION_ERROR: throw std::logic_error(st.reason());
    default: throw std::runtime_error(st.reason());
  }
}

std::shared_ptr<TensorTableEntry> MakeEntry(const at::Tensor& in, const std::string& name) {
  TORCH_CHECK(in.is_non_overlapping_and_dense(), "Horovod: tensor must be dense (contiguous in some memory format)");
  auto e = std::make_shared<TensorTableEntry>();
  e->name = name;
  e->input = in.data_ptr();
  e->dtype = MapDtype(in.scalar_type());
  e->device = DeviceOf(in);
  std::vector<int64_t> dims(in.sizes().begin(), in.sizes().end());
  e->shape = TensorShape(dims);
  if (in.is_cuda()) {
    c10::cuda::CUDAGuard guard(in.device());
    cudaEvent_t ev = g_events.Get(e->device);
    C10_CUDA_CHECK(cudaEventRecord(ev, c10::cuda::getCurrentCUDAStream(e->device)));
    e->ready_event = ev;
  }
  return e;
}

CompletionCallback MakeCallback(int handle, int device, void* ready_event) {
  return [handle, device, ready_event](const Completion& c) {
    if (ready_event) g_events.Put(device, (cudaEvent_t)ready_event);
    g_handles.MarkDone(handle, c);
  };
}

// Output allocation for collectives whose size is negotiated: runs on the
// background thread; for CUDA the freshly allocated block is ordered before
// the hvd stream touches it.
OutputAllocator MakeAllocator(at::Tensor output) {
  return [output](const std::vector<int64_t>& shape) mutable -> void* {
    if (output.is_cuda()) {
      c10::cuda::CUDAGuard guard(output.device());
      output.resize_(shape);
      int dev = (int)output.get_device();
      cudaEvent_t ev = g_events.Get(dev);
      cudaEventRecord(ev, c10::cuda::getCurrentCUDAStream(dev));
      cudaStreamWaitEvent(GpuContext::Get().Stream(dev), ev, 0);
      g_events.Put(dev, ev);
    } else {
      output.resize_(shape);
    }
    return output.data_ptr();
  };
}

// ---- ops --------------------------------------------------------------------

int DoAllreduce(at::Tensor tensor, at::Tensor output, const std::string& name, int op, double prescale, double postscale,
                int process_set_id) {
  int h = g_handles.Allocate();
  auto e = MakeEntry(tensor, OpName("allreduce", name, process_set_id));
  e->output = output.data_ptr();
  e->reduce_op = (ReduceOp)op; e->prescale = prescale; e->postscale = postscale;
  e->callback = MakeCallback(h, e->device, e->ready_event);
  auto st = g_handles.Get(h);
  st->device = e->device; st->keep_alive = {tensor, output};
  std::vector<std::shared_ptr<TensorTableEntry>> es{e};
  Status s = Engine::Get().EnqueueAllreduces(es, process_set_id);
  if (!s.ok()) { if (e->ready_event) g_events.Put(e->device, (cudaEvent_t)e->ready_event); g_handles.Release(h); ThrowIfError(s); }
  return h;
}

int DoGroupedAllreduce(std::vector<at::Tensor> tensors, std::vector<at::Tensor> outputs, const std::string& name, int op,
                       double prescale, double postscale, int process_set_id) {
  TORCH_CHECK(tensors.size() == outputs.size() && !tensors.empty(), "grouped_allreduce: bad tensor lists");
  int h = g_handles.Allocate((int)tensors.size());
  auto st = g_handles.Get(h);
  std::vector<std::shared_ptr<TensorTableEntry>> es;
  std::string base = OpName("grouped_allreduce", name, process_set_id);
  for (size_t i = 0; i < tensors.size(); ++i) {
    auto e = MakeEntry(tensors[i], base + "_" + std::to_string(i + 1) + "of" + std::to_string(tensors.size()));
    e->output = outputs[i].data_ptr();
    e->reduce_op = (ReduceOp)op; e->prescale = prescale; e->postscale = postscale;
    e->callback = MakeCallback(h, e->device, e->ready_event);
    st->device = e->device;
    st->keep_alive.push_back(tensors[i]); st->keep_alive.push_back(outputs[i]);
    es.push_back(e);
  }
  Status s = Engine::Get().EnqueueAllreduces(es, process_set_id);
  if (!s.ok()) { for (auto& e : es) if (e->ready_event) g_events.Put(e->device, (cudaEvent_t)e->ready_event); g_handles.Release(h); ThrowIfError(s); }
  return h;
}

int DoAllgather(at::Tensor tensor, at::Tensor output, const std::string& name, int process_set_id) {
  int h = g_handles.Allocate();
  auto e = MakeEntry(tensor, OpName("allgather", name, process_set_id));
  e->alloc_output = MakeAllocator(output);
  e->callback = MakeCallback(h, e->device, e->ready_event);
  auto st = g_handles.Get(h);
  st->device = e->device; st->keep_alive = {tensor, output};
  std::vector<std::shared_ptr<TensorTableEntry>> es{e};
  Status s = Engine::Get().EnqueueAllgathers(es, process_set_id);
  if (!s.ok()) { if (e->ready_event) g_events.Put(e->device, (cudaEvent_t)e->ready_event); g_handles.Release(h); ThrowIfError(s); }
  return h;
}

int DoGroupedAllgather(std::vector<at::Tensor> tensors, std::vector<at::Tensor> outputs, const std::string& name,
                       int process_set_id) {
  TORCH_CHECK(tensors.size() == outputs.size() && !tensors.empty(), "grouped_allgather: bad tensor lists");
  int h = g_handles.Allocate((int)tensors.size());
  auto st = g_handles.Get(h);
  std::vector<std::shared_ptr<TensorTableEntry>> es;
  std::string base = OpName("grouped_allgather", name, process_set_id);
  for (size_t i = 0; i < tensors.size(); ++i) {
    auto e = MakeEntry(tensors[i], base + "_" + std::to_string(i + 1) + "of" + std::to_string(tensors.size()));
    e->alloc_output = MakeAllocator(outputs[i]);
    e->callback = MakeCallback(h, e->device, e->ready_event);
    st->device = e->device;
    st->keep_alive.push_back(tensors[i]); st->keep_alive.push_back(outputs[i]);
    es.push_back(e);
  }
  Status s = Engine::Get().EnqueueAllgathers(es, process_set_id);
  if (!s.ok()) { for (auto& e : es) if (e->ready_event) g_events.Put(e->device, (cudaEvent_t)e->ready_event); g_handles.Release(h); ThrowIfError(s); }
  return h;
}

int DoBroadcast(at::Tensor tensor, at::Tensor output, int root_rank, const std::string& name, int process_set_id) {
  int h = g_handles.Allocate();
  auto e = MakeEntry(tensor, OpName("broadcast", name, process_set_id));
  e->output = output.data_ptr();
  e->root_rank = root_rank;
  e->callback = MakeCallback(h, e->device, e->ready_event);
  auto st = g_handles.Get(h);
  st->device = e->device; st->keep_alive = {tensor, output};
  Status s = Engine::Get().EnqueueBroadcast(e, process_set_id);
  if (!s.ok()) { if (e->ready_event) g_events.Put(e->device, (cudaEvent_t)e->ready_event); g_handles.Release(h); ThrowIfError(s); }
  return h;
}

int DoAlltoall(at::Tensor tensor, at::Tensor splits, at::Tensor output, at::Tensor received_splits, const std::string& name,
               int process_set_id) {
  int h = g_handles.Allocate();
  auto e = MakeEntry(tensor, OpName("alltoall", name, process_set_id));
  if (splits.defined() && splits.numel() > 0) {
    at::Tensor cpu_splits = splits.to(at::kCPU, at::kInt).contiguous();  // synchronous D2H when splits live on the GPU (mpi_ops_v2.cc:603-650)
    e->splits.assign(cpu_splits.data_ptr<int32_t>(), cpu_splits.data_ptr<int32_t>() + cpu_splits.numel());
  }
  e->alloc_output = MakeAllocator(output);
  e->callback = MakeCallback(h, e->device, e->ready_event);
  auto st = g_handles.Get(h);
  st->device = e->device; st->keep_alive = {tensor, output};
  st->received_splits_out = received_splits;
  Status s = Engine::Get().EnqueueAlltoall(e, process_set_id);
  if (!s.ok()) { if (e->ready_event) g_events.Put(e->device, (cudaEvent_t)e->ready_event); g_handles.Release(h); ThrowIfError(s); }
  return h;
}

int DoReducescatter(at::Tensor tensor, at::Tensor output, const std::string& name, int op, double prescale, double postscale,
                    int process_set_id) {
  int h = g_handles.Allocate();
  auto e = MakeEntry(tensor, OpName("reducescatter", name, process_set_id));
  e->alloc_output = MakeAllocator(output);
  e->reduce_op = (ReduceOp)op; e->prescale = prescale; e->postscale = postscale;
  e->callback = MakeCallback(h, e->device, e->ready_event);
  auto st = g_handles.Get(h);
  st->device = e->device; st->keep_alive = {tensor, output};
  std::vector<std::shared_ptr<TensorTableEntry>> es{e};
  Status s = Engine::Get().EnqueueReducescatters(es, process_set_id);
  if (!s.ok()) { if (e->ready_event) g_events.Put(e->device, (cudaEvent_t)e->ready_event); g_handles.Release(h); ThrowIfError(s); }
  return h;
}

int DoGroupedReducescatter(std::vector<at::Tensor> tensors, std::vector<at::Tensor> outputs, const std::string& name, int op,
                           double prescale, double postscale, int process_set_id) {
  TORCH_CHECK(tensors.size() == outputs.size() && !tensors.empty(), "grouped_reducescatter: bad tensor lists");
  int h = g_handles.Allocate((int)tensors.size());
  auto st = g_handles.Get(h);
  std::vector<std::shared_ptr<TensorTableEntry>> es;
  std::string base = OpName("grouped_reducescatter", name, process_set_id);
  for (size_t i = 0; i < tensors.size(); ++i) {
    auto e = MakeEntry(tensors[i], base + "_" + std::to_string(i + 1) + "of" + std::to_string(tensors.size()));
    e->alloc_output = MakeAllocator(outputs[i]);
    e->reduce_op = (ReduceOp)op; e->prescale = prescale; e->postscale = postscale;
    e->callback = MakeCallback(h, e->device, e->ready_event);
    st->device = e->device;
    st->keep_alive.push_back(tensors[i]); st->keep_alive.push_back(outputs[i]);
    es.push_back(e);
  }
  Status s = Engine::Get().EnqueueReducescatters(es, process_set_id);
  if (!s.ok()) { for (auto& e : es) if (e->ready_event) g_events.Put(e->device, (cudaEvent_t)e->ready_event); g_handles.Release(h); ThrowIfError(s); }
  return h;
}

int DoJoin(int device, int process_set_id) {
  int h = g_handles.Allocate();
  auto e = std::make_shared<TensorTableEntry>();
  e->device = device;
  e->callback = MakeCallback(h, device, nullptr);
  Status s = Engine::Get().EnqueueJoin(e, process_set_id);
  if (!s.ok()) { g_handles.Release(h); ThrowIfError(s); }
  return h;
}

int DoBarrier(int process_set_id) {
  int h = g_handles.Allocate();
  auto e = std::make_shared<TensorTableEntry>();
  e->callback = MakeCallback(h, CPU_DEVICE_ID, nullptr);
  Status s = Engine::Get().EnqueueBarrier(e, process_set_id);
  if (!s.ok()) { g_handles.Release(h); ThrowIfError(s); }
  return h;
}

bool PollHandle(int h) {
  auto st = g_handles.Get(h);
  if (!st) throw std::invalid_argument("Handle " + std::to_string(h) + " was not created or has been cleared.");
  if (!g_handles.Done(st)) { Engine::Get().RequestFlush(); return false; }
  if (st->status.ok()) {
    for (auto* ev : st->events) {
      cudaError_t q = cudaEventQuery(ev->ev);
      if (q == cudaErrorNotReady) { cudaGetLastError(); return false; }
    }
  }
  return true;
}

// Returns last_joined_rank (join) or -1.
int WaitAndClear(int h) {
  auto st = g_handles.Get(h);
  if (!st) throw std::invalid_argument("Handle " + std::to_string(h) + " was not created or has been cleared.");
  if (!g_handles.Done(st)) {
    py::gil_scoped_release release;
    Engine::Get().BeginWait();
    g_handles.Wait(st);
    Engine::Get().EndWait();
  }
  Status status = st->status;
  for (auto* ev : st->events) {
    if (status.ok()) {
      // chain instead of block: the caller's current stream waits for the collective
      cudaStream_t cur = c10::cuda::getCurrentCUDAStream(ev->device);
      cudaStreamWaitEvent(cur, ev->ev, 0);
    }
    GpuContext::Get().Release(ev);
  }
  st->events.clear();
  if (st->received_splits_out.defined() && !st->received_splits.empty()) {
    at::Tensor cpu = at::from_blob(st->received_splits.data(), {(int64_t)st->received_splits.size()}, at::kInt).clone();
    st->received_splits_out.resize_({(int64_t)st->received_splits.size()});
    st->received_splits_out.copy_(cpu);
  }
  int joined = st->last_joined_rank;
  g_handles.Release(h);
  ThrowIfError(status);
  return joined;
}

void Reset() {
  g_handles.Reset();
  std::lock_guard<std::mutex> l(g_noname_mu);
  g_noname_counters.clear();
}

// ---- registered symmetric memory -------------------------------------------------------------
// Collective: returns a uint8 CUDA tensor of `nbytes` that lives in peer-mapped memory. In-place allreduces on (views
// of) it skip the fusion-buffer pack/unpack and run the zero-copy NVLink kernel. The memory stays valid until
// hvd.shutdown().
at::Tensor SymmEmpty(int64_t nbytes, int device, int process_set_id) {
  TORCH_CHECK(nbytes > 0, "symm_empty: size must be positive");
  std::string err;
  void* p = nullptr;
  std::shared_ptr<void> keep;
  {
    py::gil_scoped_release release;
    p = Engine::Get().AllocSymmetric((size_t)nbytes, device, process_set_id, &err, &keep);
  }
  if (!p) throw std::runtime_error("symm_empty failed: " + err);
  auto opts = at::TensorOptions().dtype(at::kByte).device(at::kCUDA, device);
  // the tensor co-owns the team's mappings: it stays valid memory after hvd.shutdown() / an elastic reset (it then simply
  // is no longer registered with the new team and takes the packed path)
  return at::from_blob(p, {nbytes}, [keep](void*) mutable { keep.reset(); }, opts);
}

// In-place allreduce of a tensor that lives in registered symmetric memory, launched as ONE kernel on the current CUDA
// stream of the calling thread — which may be capturing into a CUDA graph.  No handle, no negotiation, no host
// synchronisation: ordering is the stream's.  reduce_op follows hvd.ReduceOp (0 Average, 1 Sum, 3 Min, 4 Max, 5 Product).
void CapturedAllreduce(at::Tensor tensor, int reduce_op, double prescale, double postscale, int process_set_id, int max_ctas) {
  TORCH_CHECK(tensor.is_cuda(), "captured_allreduce_: CUDA tensor expected");
  TORCH_CHECK(tensor.is_non_overlapping_and_dense(), "captured_allreduce_: tensor must be dense");
  c10::cuda::CUDAGuard guard(tensor.device());
  cudaStream_t stream = c10::cuda::getCurrentCUDAStream(tensor.get_device());
  const int64_t bytes = (int64_t)tensor.numel() * (int64_t)tensor.element_size();
  ThrowIfError(Engine::Get().CapturedAllreduce(tensor.data_ptr(), bytes, MapDtype(tensor.scalar_type()), (ReduceOp)reduce_op, prescale,
                                               postscale, process_set_id, max_ctas, (void*)stream));
}

// ---- fused optimizer kernels (B200-native extra; see kernels/optim_kernels.cu) --------------
void FusedSgdStep(std::vector<at::Tensor> params, std::vector<at::Tensor> grads, std::vector<at::Tensor> momenta, double lr,
                  double momentum, double dampening, double weight_decay, bool nesterov, double grad_scale, bool first_step) {
  TORCH_CHECK(params.size() == grads.size(), "fused_sgd: list sizes differ");
  if (params.empty()) return;
  const bool has_mom = momentum != 0.0;
  TORCH_CHECK(!has_mom || momenta.size() == params.size(), "fused_sgd: momentum buffers missing");
  int device = (int)params[0].get_device();
  c10::cuda::CUDAGuard guard(params[0].device());
  cudaStream_t s = c10::cuda::getCurrentCUDAStream(device);
  std::vector<kern::SgdTensor> table(params.size());
  int64_t maxc = 0;
  for (size_t i = 0; i < params.size(); ++i) {
    TORCH_CHECK(params[i].is_cuda() && params[i].is_non_overlapping_and_dense() && grads[i].strides() == params[i].strides(), "fused_sgd: params/grads must be dense CUDA tensors with identical layout");
    table[i].param = params[i].data_ptr(); table[i].grad = grads[i].data_ptr();
    table[i].momentum = has_mom ? momenta[i].data_ptr() : nullptr;
    table[i].count = params[i].numel();
    maxc = std::max(maxc, table[i].count);
  }
  const auto* dt = (const kern::SgdTensor*)GpuContext::Get().Stage(device, table.data(), table.size() * sizeof(kern::SgdTensor), s);
  TORCH_CHECK(dt != nullptr, "fused_sgd: table too large");
  cudaError_t e = kern::LaunchFusedSgd(dt, (int)table.size(), maxc, (float)lr, (float)momentum, (float)dampening, (float)weight_decay,
                                       nesterov ? 1 : 0, (float)grad_scale, first_step ? 1 : 0,
                                       (int)MapDtype(params[0].scalar_type()), (int)MapDtype(grads[0].scalar_type()), s);
  TORCH_CHECK(e == cudaSuccess, "fused_sgd launch failed: ", cudaGetErrorString(e));
}

void FusedAdamStep(std::vector<at::Tensor> params, std::vector<at::Tensor> grads, std::vector<at::Tensor> exp_avg,
                   std::vector<at::Tensor> exp_avg_sq, double lr, double beta1, double beta2, double eps, double weight_decay,
                   int64_t step, double grad_scale, bool adamw) {
  TORCH_CHECK(params.size() == grads.size() && params.size() == exp_avg.size() && params.size() == exp_avg_sq.size(), "fused_adam: list sizes differ");
  if (params.empty()) return;
  int device = (int)params[0].get_device();
  c10::cuda::CUDAGuard guard(params[0].device());
  cudaStream_t s = c10::cuda::getCurrentCUDAStream(device);
  std::vector<kern::AdamTensor> table(params.size());
  int64_t maxc = 0;
  for (size_t i = 0; i < params.size(); ++i) {
    TORCH_CHECK(params[i].is_cuda() && params[i].is_non_overlapping_and_dense() && grads[i].strides() == params[i].strides(), "fused_adam: params/grads must be dense CUDA tensors with identical layout");
    table[i].param = params[i].data_ptr(); table[i].grad = grads[i].data_ptr();
    table[i].exp_avg = exp_avg[i].data_ptr(); table[i].exp_avg_sq = exp_avg_sq[i].data_ptr();
    table[i].count = params[i].numel();
    maxc = std::max(maxc, table[i].count);
  }
  const auto* dt = (const kern::AdamTensor*)GpuContext::Get().Stage(device, table.data(), table.size() * sizeof(kern::AdamTensor), s);
  TORCH_CHECK(dt != nullptr, "fused_adam: table too large");
  const float c1 = 1.0f - (float)std::pow(beta1, (double)step), c2 = 1.0f - (float)std::pow(beta2, (double)step);
  cudaError_t e = kern::LaunchFusedAdamW(dt, (int)table.size(), maxc, (float)lr, (float)beta1, (float)beta2, (float)eps,
                                         (float)weight_decay, c1, c2, (float)grad_scale, adamw ? 1 : 0,
                                         (int)MapDtype(params[0].scalar_type()), (int)MapDtype(grads[0].scalar_type()), s);
  TORCH_CHECK(e == cudaSuccess, "fused_adam launch failed: ", cudaGetErrorString(e));
}

}  // namespace

PYBIND11_MODULE(_hvd_torch, m) {
  m.doc() = "horovod_b200 PyTorch binding";
  m.def("allreduce_async", &DoAllreduce);
  m.def("grouped_allreduce_async", &DoGroupedAllreduce);
  m.def("allgather_async", &DoAllgather);
  m.def("grouped_allgather_async", &DoGroupedAllgather);
  m.def("broadcast_async", &DoBroadcast);
  m.def("alltoall_async", &DoAlltoall);
  m.def("reducescatter_async", &DoReducescatter);
  m.def("grouped_reducescatter_async", &DoGroupedReducescatter);
  m.def("join", &DoJoin);
  m.def("barrier", &DoBarrier);
  m.def("poll", &PollHandle);
  m.def("wait_and_clear", &WaitAndClear);
  m.def("reset", &Reset);
  m.def("reset_noname_counters", &ResetNonameCounters);
  m.def("symm_empty", &SymmEmpty);
  m.def("captured_allreduce_", &CapturedAllreduce);
  m.def("fused_sgd_step", &FusedSgdStep);
  m.def("fused_adam_step", &FusedAdamStep);
}
