#include "cpu_ops.h"
#include <algorithm>
#include <atomic>
#include <mutex>
#include <deque>
#include <condition_variable>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "../common/half.h"
#include "../common/thread_pool.h"
#include <thread>

namespace hvd {
namespace cpu {

namespace {

template <typename T> inline T Combine(T a, T b, ReduceOp op) {
  switch (op) {
    case ReduceOp::MIN: return b < a ? b : a;
    case ReduceOp::MAX: return b > a ? b : a;
    case ReduceOp::PRODUCT: return a * b;
    default: return a + b;
  }
}
// dst and src never overlap (a slot / receive buffer vs the caller's buffer): __restrict__ lets the loops vectorise
// without runtime alias checks; the operator is selected once, outside the loop.
template <typename T> void ReduceT(T* __restrict__ d, const T* __restrict__ s, int64_t n, ReduceOp op) {
  switch (op) {
    case ReduceOp::MIN: for (int64_t i = 0; i < n; ++i) d[i] = s[i] < d[i] ? s[i] : d[i]; return;
    case ReduceOp::MAX: for (int64_t i = 0; i < n; ++i) d[i] = s[i] > d[i] ? s[i] : d[i]; return;
    case ReduceOp::PRODUCT: for (int64_t i = 0; i < n; ++i) d[i] = d[i] * s[i]; return;
    default: for (int64_t i = 0; i < n; ++i) d[i] = d[i] + s[i]; return;
  }
}
template <float (*ToF)(uint16_t), uint16_t (*FromF)(float)> void Reduce16(uint16_t* d, const uint16_t* s, int64_t n, ReduceOp op) {
  for (int64_t i = 0; i < n; ++i) d[i] = FromF(Combine<float>(ToF(d[i]), ToF(s[i]), op));
}
#if defined(__x86_64__)
// fp16 through the F16C unit (8 conversions per instruction) when the CPU has it; the function carries its own target
// attribute, so the rest of the library still runs on CPUs without AVX.  Operand order of min / max keeps the scalar
// path's NaN behaviour (an unordered compare keeps the destination value).
__attribute__((target("avx,f16c"))) void ReduceHalfF16C(uint16_t* __restrict__ d, const uint16_t* __restrict__ s, int64_t n, ReduceOp op) {
  int64_t i = 0;
  for (; i + 8 <= n; i += 8) {
    const __m256 a = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i*)(d + i)));
    const __m256 b = _mm256_cvtph_ps(_mm_loadu_si128((const __m128i*)(s + i)));
    __m256 r;
    switch (op) {
      case ReduceOp::MIN: r = _mm256_min_ps(b, a); break;
      case ReduceOp::MAX: r = _mm256_max_ps(b, a); break;
      case ReduceOp::PRODUCT: r = _mm256_mul_ps(a, b); break;
      default: r = _mm256_add_ps(a, b); break;
    }
    _mm_storeu_si128((__m128i*)(d + i), _mm256_cvtps_ph(r, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
  }
  for (; i < n; ++i) d[i] = FloatToHalfBits(Combine<float>(HalfBitsToFloat(d[i]), HalfBitsToFloat(s[i]), op));
}
__attribute__((target("avx,f16c"))) void ScaleHalfF16C(uint16_t* b, int64_t n, float f) {
  const __m256 scale = _mm256_set1_ps(f);
  int64_t i = 0;
  for (; i + 8 <= n; i += 8)
    _mm_storeu_si128((__m128i*)(b + i), _mm256_cvtps_ph(_mm256_mul_ps(_mm256_cvtph_ps(_mm_loadu_si128((const __m128i*)(b + i))), scale),
                                                        _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
  for (; i < n; ++i) b[i] = FloatToHalfBits(HalfBitsToFloat(b[i]) * f);
}
bool HaveF16C() { static const bool have = __builtin_cpu_supports("f16c") && __builtin_cpu_supports("avx"); return have; }
#else
bool HaveF16C() { return false; }
#endif

// bf16: widening is a shift, narrowing is round-to-nearest-even with the NaN case as a select (no branch), and the operator is
// chosen outside the loop — the whole body vectorises.
inline uint16_t NarrowBF16(float v) {
  uint32_t x; memcpy(&x, &v, 4);
  const uint32_t rounded = (x + 0x7fffu + ((x >> 16) & 1u)) >> 16;
  const uint32_t quiet_nan = (x >> 16) | 0x40u;
  return (uint16_t)(((x & 0x7fffffffu) > 0x7f800000u) ? quiet_nan : rounded);
}
template <typename F> void ReduceBF16With(uint16_t* __restrict__ d, const uint16_t* __restrict__ s, int64_t n, F f) {
  for (int64_t i = 0; i < n; ++i) d[i] = NarrowBF16(f(BF16BitsToFloat(d[i]), BF16BitsToFloat(s[i])));
}
void ReduceBF16(uint16_t* d, const uint16_t* s, int64_t n, ReduceOp op) {
  switch (op) {
    case ReduceOp::MIN: ReduceBF16With(d, s, n, [](float a, float b) { return b < a ? b : a; }); return;
    case ReduceOp::MAX: ReduceBF16With(d, s, n, [](float a, float b) { return b > a ? b : a; }); return;
    case ReduceOp::PRODUCT: ReduceBF16With(d, s, n, [](float a, float b) { return a * b; }); return;
    default: ReduceBF16With(d, s, n, [](float a, float b) { return a + b; }); return;
  }
}

template <typename T> void ScaleT(T* b, int64_t n, double s) { for (int64_t i = 0; i < n; ++i) b[i] = (T)(b[i] * s); }

double LoadAsDouble(const void* p, int64_t i, DataType t) {
  switch (t) {
    case DataType::FLOAT16: return HalfBitsToFloat(((const uint16_t*)p)[i]);
    case DataType::BFLOAT16: return BF16BitsToFloat(((const uint16_t*)p)[i]);
    case DataType::FLOAT32: return ((const float*)p)[i];
    default: return ((const double*)p)[i];
  }
}
void StoreFromDouble(void* p, int64_t i, DataType t, double v) {
  switch (t) {
    case DataType::FLOAT16: ((uint16_t*)p)[i] = FloatToHalfBits((float)v); break;
    case DataType::BFLOAT16: ((uint16_t*)p)[i] = FloatToBF16Bits((float)v); break;
    case DataType::FLOAT32: ((float*)p)[i] = (float)v; break;
    default: ((double*)p)[i] = v; break;
  }
}

}  // namespace

void ReduceInto(void* dst, const void* src, int64_t n, DataType dtype, ReduceOp op) {
  switch (dtype) {
    case DataType::UINT8: case DataType::BOOL: ReduceT((uint8_t*)dst, (const uint8_t*)src, n, op); break;
    case DataType::INT8: ReduceT((int8_t*)dst, (const int8_t*)src, n, op); break;
    case DataType::UINT16: ReduceT((uint16_t*)dst, (const uint16_t*)src, n, op); break;
    case DataType::INT16: ReduceT((int16_t*)dst, (const int16_t*)src, n, op); break;
    case DataType::INT32: ReduceT((int32_t*)dst, (const int32_t*)src, n, op); break;
    case DataType::INT64: ReduceT((int64_t*)dst, (const int64_t*)src, n, op); break;
    case DataType::FLOAT32: ReduceT((float*)dst, (const float*)src, n, op); break;
    case DataType::FLOAT64: ReduceT((double*)dst, (const double*)src, n, op); break;
    case DataType::FLOAT16:
#if defined(__x86_64__)
      if (HaveF16C()) { ReduceHalfF16C((uint16_t*)dst, (const uint16_t*)src, n, op); break; }
#endif
      Reduce16<HalfBitsToFloat, FloatToHalfBits>((uint16_t*)dst, (const uint16_t*)src, n, op);
      break;
    case DataType::BFLOAT16: ReduceBF16((uint16_t*)dst, (const uint16_t*)src, n, op); break;
  }
}

void ScaleBuffer(void* buf, int64_t n, DataType dtype, double s) {
  if (s == 1.0) return;
  switch (dtype) {
    case DataType::UINT8: case DataType::BOOL: ScaleT((uint8_t*)buf, n, s); break;
    case DataType::INT8: ScaleT((int8_t*)buf, n, s); break;
    case DataType::UINT16: ScaleT((uint16_t*)buf, n, s); break;
    case DataType::INT16: ScaleT((int16_t*)buf, n, s); break;
    case DataType::INT32: ScaleT((int32_t*)buf, n, s); break;
    case DataType::INT64: ScaleT((int64_t*)buf, n, s); break;
    case DataType::FLOAT32: { float* b = (float*)buf; float f = (float)s; for (int64_t i = 0; i < n; ++i) b[i] *= f; break; }
    case DataType::FLOAT64: ScaleT((double*)buf, n, s); break;
    case DataType::FLOAT16: {
      uint16_t* b = (uint16_t*)buf;
#if defined(__x86_64__)
      if (HaveF16C()) { ScaleHalfF16C(b, n, (float)s); break; }
#endif
      for (int64_t i = 0; i < n; ++i) b[i] = FloatToHalfBits(HalfBitsToFloat(b[i]) * (float)s);
      break;
    }
    case DataType::BFLOAT16: { uint16_t* b = (uint16_t*)buf; const float f = (float)s; for (int64_t i = 0; i < n; ++i) b[i] = NarrowBF16(BF16BitsToFloat(b[i]) * f); break; }
  }
}

// ---------------------------------------------------------------------------
// A few helper threads for the bulk phases (copy-in / reduce / copy-out) of big host messages: one core moves ~14 GB/s, the
// memory system of a training host several times that.  HVD_CPU_THREADS sets the team size per rank.  Off by default (1): with
// the default 1 MiB slots a piece is too small to amortise the hand-off; HVD_SHM_SLOT_BYTES=8388608 HVD_CPU_THREADS=4 took a
// 64 MiB allreduce at np=2 from 15 ms to 8 ms on the build container.  HVD_CPU_THREADS=0 picks min(4, cores / local ranks / 2).
class BulkTeam {
 public:
  static BulkTeam& Get() { static BulkTeam t; return t; }
  int size() const { return k_; }
  // fn(lo, hi) over [0, n) split into size() contiguous ranges; returns when all ranges are done
  template <typename F> void For(int64_t n, int64_t min_per_thread, F fn) {
    int k = (int)std::min<int64_t>(k_, std::max<int64_t>(1, n / std::max<int64_t>(1, min_per_thread)));
    if (k <= 1) { fn((int64_t)0, n); return; }
    std::atomic<int> left{k - 1};
    for (int i = 1; i < k; ++i) {
      const int64_t lo = n * i / k, hi = n * (i + 1) / k;
      pool_.Execute([&fn, &left, lo, hi] { fn(lo, hi); left.fetch_sub(1, std::memory_order_release); });
    }
    fn((int64_t)0, n / k);
    while (left.load(std::memory_order_acquire) > 0) std::this_thread::yield();
  }

 private:
  BulkTeam() {
    int k = 1;
    if (const char* e = getenv("HVD_CPU_THREADS")) k = atoi(e);
    if (k <= 0) {
      int local = 1;
      if (const char* e = getenv("HOROVOD_LOCAL_SIZE")) local = std::max(1, atoi(e));
      const int cores = (int)std::max(1u, std::thread::hardware_concurrency());
      k = std::min(4, std::max(1, cores / local / 2));
    }
    k_ = std::min(k, 16);
    if (k_ > 1) pool_.Create(k_ - 1);
  }
  int k_ = 1;
  ThreadPool pool_;
};

inline void BulkCopy(void* dst, const void* src, size_t bytes) {
  if (bytes < (1u << 20) || BulkTeam::Get().size() == 1) { memcpy(dst, src, bytes); return; }
  BulkTeam::Get().For((int64_t)bytes, 256 << 10, [&](int64_t lo, int64_t hi) { memcpy((char*)dst + lo, (const char*)src + lo, (size_t)(hi - lo)); });
}
inline void BulkReduce(void* dst, const void* src, int64_t count, DataType dtype, ReduceOp op) {
  const size_t es = DataTypeSize(dtype);
  if ((size_t)count * es < (1u << 20) || BulkTeam::Get().size() == 1) { ReduceInto(dst, src, count, dtype, op); return; }
  BulkTeam::Get().For(count, (int64_t)((256 << 10) / es), [&](int64_t lo, int64_t hi) {
    ReduceInto((char*)dst + lo * es, (const char*)src + lo * es, hi - lo, dtype, op);
  });
}

// One thread per process that executes ReduceInto jobs in submission order: the cross-host ring hands it the chunk it just
// received and goes back to the sockets (RingAllreduce).  Created on first use; HVD_RING_CHUNK_BYTES (default 1 MiB, 0 = no
// pipelining) is the chunk size.
class RingReducer {
 public:
  static RingReducer& Get() { static RingReducer r; return r; }
  uint64_t Submit(void* dst, const void* src, int64_t count, DataType dtype, ReduceOp op) {
    std::lock_guard<std::mutex> g(mu_);
    if (!started_) { started_ = true; thread_ = std::thread([this] { Run(); }); }
    jobs_.push_back({dst, src, count, dtype, op});
    cv_.notify_one();
    return ++submitted_;
  }
  void Wait(uint64_t ticket) {
    if (ticket == 0) return;
    for (int spin = 0; done_.load(std::memory_order_acquire) < ticket; ++spin)
      if (spin < 2000) __builtin_ia32_pause(); else std::this_thread::yield();
  }
  ~RingReducer() {
    { std::lock_guard<std::mutex> g(mu_); stop_ = true; cv_.notify_one(); }
    if (thread_.joinable()) thread_.join();
  }

 private:
  struct Job { void* dst; const void* src; int64_t count; DataType dtype; ReduceOp op; };
  void Run() {
    std::unique_lock<std::mutex> lk(mu_);
    while (true) {
      cv_.wait(lk, [this] { return stop_ || !jobs_.empty(); });
      if (jobs_.empty()) return;
      Job j = jobs_.front();
      jobs_.pop_front();
      lk.unlock();
      ReduceInto(j.dst, j.src, j.count, j.dtype, j.op);
      done_.fetch_add(1, std::memory_order_release);
      lk.lock();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<Job> jobs_;
  std::thread thread_;
  bool started_ = false, stop_ = false;
  uint64_t submitted_ = 0;
  std::atomic<uint64_t> done_{0};
};

static int64_t RingChunkElems(size_t es) {
  static const int64_t bytes = [] {
    const char* e = getenv("HVD_RING_CHUNK_BYTES");
    return e ? std::max<int64_t>(0, atoll(e)) : (int64_t)1 << 20;
  }();
  return bytes / (int64_t)es;
}

// One reduce-scatter step of a ring: send `slen` elements from `send`, receive `rlen` elements from `prev` and fold them into
// `acc`.  Big steps move in chunks through two staging buffers: the reducer thread folds chunk k while this thread already
// exchanges chunk k + 1, so the wire and the adds overlap.  `tmp` is scratch owned by the caller (sized on demand).
// `longest` is the longest segment of the WHOLE ring (known to every rank): it alone decides whether and into how many pieces
// the steps are cut, so both ends of every connection agree on the message boundaries (message-based transports need that).
static void RingReduceStep(Transport* t, int next, int prev, const char* send, int64_t slen, char* acc, int64_t rlen, int64_t longest,
                           DataType dtype, ReduceOp op, std::vector<char>* tmp) {
  const size_t es = DataTypeSize(dtype);
  const int64_t chunk = RingChunkElems(es);
  if (chunk <= 0 || longest < 2 * chunk) {
    if (tmp->size() < (size_t)rlen * es) tmp->resize((size_t)rlen * es);
    t->SendRecv(next, send, (size_t)slen * es, prev, tmp->data(), (size_t)rlen * es);
    ReduceInto(acc, tmp->data(), rlen, dtype, op);
    return;
  }
  if (tmp->size() < (size_t)(2 * chunk) * es) tmp->resize((size_t)(2 * chunk) * es);
  RingReducer& red = RingReducer::Get();
  const int64_t pieces = (longest + chunk - 1) / chunk;
  uint64_t ticket[2] = {0, 0};
  for (int64_t k = 0; k < pieces; ++k) {
    const int64_t so = std::min(slen, k * chunk), sc = std::min(chunk, slen - so);
    const int64_t ro = std::min(rlen, k * chunk), rc = std::min(chunk, rlen - ro);
    char* st = tmp->data() + (size_t)(k & 1) * (size_t)chunk * es;
    red.Wait(ticket[k & 1]);                       // the reduce that read this staging half two pieces ago
    t->SendRecv(next, send + so * es, (size_t)sc * es, prev, st, (size_t)rc * es);
    ticket[k & 1] = rc > 0 ? red.Submit(acc + ro * es, st, rc, dtype, op) : 0;
  }
  red.Wait(ticket[0]);                             // the next ring step sends what this one reduced
  red.Wait(ticket[1]);
}

static std::atomic<unsigned long long> g_path_count[3];
unsigned long long HostPathCount(int which) { return which >= 0 && which < 3 ? g_path_count[which].load(std::memory_order_relaxed) : 0; }
static inline bool Took(int path, bool taken) { if (taken) g_path_count[path].fetch_add(1, std::memory_order_relaxed); return taken; }

void RingAllreduce(Transport* t, char* b, int64_t count, DataType dtype, ReduceOp op);
void RingAllgatherv(Transport* t, char* o, const std::vector<int64_t>& bytes, const std::vector<int64_t>& displ);

// ---------------------------------------------------------------------------
// Shared-memory data plane (single-host communicators; see transport.h:ShmData).  The host analogue of the two-shot GPU
// kernel: every rank publishes a piece of its buffer in its slot, reduces the 1/N of the piece it owns straight out of
// the peers' slots, and everybody copies the reduced chunks back.  No socket, no intermediate copies besides the slot.

namespace {

bool ShmAllreduce(Transport* t, char* b, int64_t count, DataType dtype, ReduceOp op) {
  ShmData d;
  if (!t->ShmDataPlane(&d)) return false;
  const int n = t->size(), r = t->rank();
  const size_t es = DataTypeSize(dtype);
  const int64_t per_piece = (int64_t)(d.slot_bytes / es);
  static const bool prof = getenv("HVD_SHM_PROFILE") != nullptr;
  double tt[5] = {0, 0, 0, 0, 0};
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  for (int64_t done = 0; done < count; done += per_piece) {
    const int64_t m = std::min(per_piece, count - done);
    const int half = (int)(t->ShmNextPiece() & 1);
    char* mine = d.slot(r, half);
    double t0 = prof ? now() : 0;
    BulkCopy(mine, b + done * es, (size_t)m * es);
    double t1 = prof ? now() : 0;
    t->Barrier();
    double t2 = prof ? now() : 0;
    const int64_t lo = m * r / n, hi = m * (r + 1) / n;
    for (int p = 1; p < n && hi > lo; ++p) BulkReduce(mine + lo * es, d.slot((r + p) % n, half) + lo * es, hi - lo, dtype, op);
    double t3 = prof ? now() : 0;
    t->Barrier();
    double t4 = prof ? now() : 0;
    for (int q = 0; q < n; ++q) {
      const int64_t ql = m * q / n, qh = m * (q + 1) / n;
      if (qh > ql) BulkCopy(b + (done + ql) * es, d.slot(q, half) + ql * es, (size_t)(qh - ql) * es);
    }
    if (prof) { double t5 = now(); tt[0] += t1 - t0; tt[1] += t2 - t1; tt[2] += t3 - t2; tt[3] += t4 - t3; tt[4] += t5 - t4; }
  }
  if (prof && count * (int64_t)es >= (1 << 20))
    fprintf(stderr, "[shm allreduce %lld B rank %d] in %.2f ms, barrier %.2f, reduce %.2f, barrier %.2f, out %.2f\n", (long long)(count * es), r,
            tt[0] * 1e3, tt[1] * 1e3, tt[2] * 1e3, tt[3] * 1e3, tt[4] * 1e3);
  return true;
}

// Every step each rank publishes, for EVERY destination q, the next `sub` elements of q's segment (destination-major layout
// inside its slot); after the barrier rank q folds sub-chunk q of all n slots straight into its output.  All ranks copy and
// reduce the same amount per step (walking the buffer front to back instead keeps one owner busy reducing n slots while the
// others wait: 64 MiB of input took 46 ms at np=4, an ALLREDUCE of the same buffer 25 ms).
bool ShmReducescatter(Transport* t, const char* b, const std::vector<int64_t>& off, char* out, DataType dtype, ReduceOp op) {
  ShmData d;
  if (!t->ShmDataPlane(&d)) return false;
  const int n = t->size(), r = t->rank();
  const size_t es = DataTypeSize(dtype);
  const int64_t sub = (int64_t)(d.slot_bytes / es) / n;
  if (sub < 16) return false;
  int64_t longest = 0;
  for (int q = 0; q < n; ++q) longest = std::max(longest, off[q + 1] - off[q]);
  const int64_t mine_len = off[r + 1] - off[r];
  for (int64_t done = 0; done < longest; done += sub) {
    const int half = (int)(t->ShmNextPiece() & 1);
    char* slot = d.slot(r, half);
    for (int q = 0; q < n; ++q) {
      const int64_t c = std::min(sub, off[q + 1] - off[q] - done);
      if (c > 0) memcpy(slot + (size_t)q * (size_t)sub * es, b + (off[q] + done) * es, (size_t)c * es);
    }
    t->Barrier();
    const int64_t c = std::min(sub, mine_len - done);
    if (c > 0) {
      char* dst = out + done * es;
      memcpy(dst, slot + (size_t)r * (size_t)sub * es, (size_t)c * es);
      for (int p = 1; p < n; ++p) ReduceInto(dst, d.slot((r + p) % n, half) + (size_t)r * (size_t)sub * es, c, dtype, op);
    }
  }
  return true;
}

bool ShmAllgatherv(Transport* t, const char* in, char* o, const std::vector<int64_t>& bytes, const std::vector<int64_t>& displ) {
  ShmData d;
  if (!t->ShmDataPlane(&d)) return false;
  const int n = t->size(), r = t->rank();
  const int64_t longest = *std::max_element(bytes.begin(), bytes.end());
  const int64_t S = (int64_t)d.slot_bytes;
  for (int64_t done = 0; done < longest; done += S) {
    const int half = (int)(t->ShmNextPiece() & 1);
    const int64_t mine = std::min(S, bytes[r] - done);
    if (mine > 0) memcpy(d.slot(r, half), in + done, (size_t)mine);
    t->Barrier();
    for (int q = 0; q < n; ++q) {
      const int64_t theirs = std::min(S, bytes[q] - done);
      if (q != r && theirs > 0) memcpy(o + displ[q] + done, d.slot(q, half), (size_t)theirs);
    }
  }
  return true;
}

bool ShmBroadcast(Transport* t, char* buf, int64_t bytes, int root) {
  ShmData d;
  if (!t->ShmDataPlane(&d)) return false;
  const int r = t->rank();
  const int64_t S = (int64_t)d.slot_bytes;
  for (int64_t done = 0; done < bytes; done += S) {
    const int half = (int)(t->ShmNextPiece() & 1);
    const int64_t m = std::min(S, bytes - done);
    if (r == root) memcpy(d.slot(root, half), buf + done, (size_t)m);
    t->Barrier();
    if (r != root) memcpy(buf + done, d.slot(root, half), (size_t)m);
  }
  return true;
}

// Every rank publishes, in front of each piece of its send buffer, where the block for each destination starts (a
// receiver knows how much it gets from a peer, not where that block sits in the peer's buffer) and how long the buffer is
// (so that all ranks agree on the number of pieces after the first barrier).
// Personalised exchange among the ranks that share one set of shm slots.  `group[i]` is the communicator rank that owns slot
// i (the whole communicator on one host; the ranks of this host in a multi-host job), `me` the caller's slot index; `barrier`
// synchronises exactly that group.  Blocks for / from ranks outside the group are not touched.
template <typename BarrierFn>
bool SlotAlltoallv(Transport* t, const ShmData& d, const std::vector<int>& group, int me, BarrierFn barrier, const char* in,
                   const std::vector<int64_t>& sd, char* out, const std::vector<int64_t>& rd) {
  const int g = (int)group.size();
  const int64_t hdr = (int64_t)(g + 2) * 8;
  const int64_t S = (int64_t)d.slot_bytes - hdr;
  if (S < 4096) return false;
  // what I send inside the group, packed in group order
  std::vector<int64_t> gd((size_t)g + 1, 0);
  int64_t my_longest_block = 0;
  for (int i = 0; i < g; ++i) {
    const int64_t len = i == me ? 0 : sd[(size_t)group[(size_t)i] + 1] - sd[(size_t)group[(size_t)i]];
    gd[(size_t)i + 1] = gd[(size_t)i] + len;
    my_longest_block = std::max(my_longest_block, len);
  }
  const int64_t my_total = gd[(size_t)g];
  // ---- round 0: header (packed displacements, total, longest block) + all my blocks when they fit into one slot ----
  int half = (int)(t->ShmNextPiece() & 1);
  {
    char* mine = d.slot(me, half);
    int64_t* h = (int64_t*)mine;
    for (int i = 0; i < g; ++i) h[i] = gd[(size_t)i];
    h[g] = my_total;
    h[g + 1] = my_longest_block;
    if (my_total > 0 && my_total <= S)
      for (int i = 0; i < g; ++i)
        if (gd[(size_t)i + 1] > gd[(size_t)i]) memcpy(mine + hdr + gd[(size_t)i], in + sd[(size_t)group[(size_t)i]], (size_t)(gd[(size_t)i + 1] - gd[(size_t)i]));
  }
  barrier();
  int64_t longest_total = 0, longest_block = 0;
  for (int q = 0; q < g; ++q) {
    const int64_t* h = (const int64_t*)d.slot(q, half);
    longest_total = std::max(longest_total, h[g]);
    longest_block = std::max(longest_block, h[g + 1]);
  }
  auto rlen_of = [&](int q) { return rd[(size_t)group[(size_t)q] + 1] - rd[(size_t)group[(size_t)q]]; };
  if (longest_total <= S) {
    // small exchange: everything is already published, every rank pulls its blocks (one barrier in total)
    for (int q = 0; q < g; ++q) {
      if (q == me || rlen_of(q) == 0) continue;
      const char* theirs = d.slot(q, half);
      memcpy(out + rd[(size_t)group[(size_t)q]], theirs + hdr + ((const int64_t*)theirs)[me], (size_t)rlen_of(q));
    }
    return true;
  }
  // ---- large exchange: g - 1 rounds; in round k every rank publishes (a piece of) its block for slot me + k and pulls from
  // slot me - k, so each slot has exactly one reader and every rank copies the same amount per piece (publishing the send
  // buffer front to back instead makes all ranks target rank 0 first, then rank 1, ...: one busy receiver, g - 1 idle) ----
  const int64_t pieces = (longest_block + S - 1) / S;
  for (int k = 1; k < g; ++k) {
    const int to = (me + k) % g, from = (me - k + g) % g;
    const int64_t slen = sd[(size_t)group[(size_t)to] + 1] - sd[(size_t)group[(size_t)to]], rlen = rlen_of(from);
    for (int64_t p = 0; p < pieces; ++p) {
      half = (int)(t->ShmNextPiece() & 1);
      const int64_t so = std::min(slen, p * S), sc = std::min(S, slen - so);
      if (sc > 0) memcpy(d.slot(me, half) + hdr, in + sd[(size_t)group[(size_t)to]] + so, (size_t)sc);
      barrier();
      const int64_t ro = std::min(rlen, p * S), rc = std::min(S, rlen - ro);
      if (rc > 0) memcpy(out + rd[(size_t)group[(size_t)from]] + ro, d.slot(from, half) + hdr, (size_t)rc);
    }
  }
  return true;
}

bool ShmAlltoallv(Transport* t, const char* in, const std::vector<int64_t>& sd, char* out, const std::vector<int64_t>& rd) {
  ShmData d;
  if (!t->ShmDataPlane(&d)) return false;
  std::vector<int> all((size_t)t->size());
  for (int i = 0; i < t->size(); ++i) all[(size_t)i] = i;
  return SlotAlltoallv(t, d, all, t->rank(), [t] { t->Barrier(); }, in, sd, out, rd);
}

// Multi-host, same number of ranks on every host: the blocks for the ranks of my host travel through the host's shm slots, the
// blocks for every other rank over the TCP mesh (all of them in one poll set).
bool HierAlltoallv(Transport* t, const char* in, const std::vector<int64_t>& sd, char* out, const std::vector<int64_t>& rd) {
  HierData h;
  if (!t->HierDataPlane(&h)) return false;
  const auto& column = *h.column;
  const int L = h.local_size, l = h.local_rank, n = t->size();
  int x = -1;
  for (size_t i = 0; i < column[(size_t)l].size(); ++i) if (column[(size_t)l][i] == t->rank()) x = (int)i;
  if (x < 0) return false;
  std::vector<int> host((size_t)L);
  std::vector<uint8_t> local((size_t)n, 0);
  for (int c = 0; c < L; ++c) { host[(size_t)c] = column[(size_t)c][(size_t)x]; local[(size_t)host[(size_t)c]] = 1; }
  if (L > 1 && !SlotAlltoallv(t, h.local, host, l, [t] { t->LocalBarrier(); }, in, sd, out, rd)) return false;
  t->AlltoallvBytes(in, sd.data(), out, rd.data(), local.data());
  return true;
}

// Multi-host, same number of ranks on every host.  Per piece: publish in the host's shm slots -> every local rank reduces the
// 1/L chunk it owns over the ranks of its host -> the L chunk owners run L cross-host ring allreduces in parallel, each with
// the owners of the same chunk on the other hosts (1/L of the bytes per TCP stream) -> everybody copies the chunks back.
bool HierAllreduce(Transport* t, char* b, int64_t count, DataType dtype, ReduceOp op) {
  HierData h;
  if (!t->HierDataPlane(&h)) return false;
  const int L = h.local_size, l = h.local_rank;
  const size_t es = DataTypeSize(dtype);
  const int64_t per_piece = (int64_t)(h.local.slot_bytes / es);
  for (int64_t done = 0; done < count; done += per_piece) {
    const int64_t m = std::min(per_piece, count - done);
    const int half = (int)(t->ShmNextPiece() & 1);
    char* mine = h.local.slot(l, half);
    memcpy(mine, b + done * es, (size_t)m * es);
    t->LocalBarrier();
    const int64_t lo = m * l / L, hi = m * (l + 1) / L;
    if (hi > lo) {
      for (int p = 1; p < L; ++p) ReduceInto(mine + lo * es, h.local.slot((l + p) % L, half) + lo * es, hi - lo, dtype, op);
      RingAllreduce(h.cross, mine + lo * es, hi - lo, dtype, op);
    }
    t->LocalBarrier();
    for (int q = 0; q < L; ++q) {
      const int64_t ql = m * q / L, qh = m * (q + 1) / L;
      if (qh > ql) memcpy(b + (done + ql) * es, h.local.slot(q, half) + ql * es, (size_t)(qh - ql) * es);
    }
  }
  return true;
}

// Two-level allgather: the ranks with my local index exchange their blocks across hosts (L parallel rings), then the ranks
// of a host hand each other the columns they collected through shm.
bool HierAllgatherv(Transport* t, char* o, const std::vector<int64_t>& bytes, const std::vector<int64_t>& displ) {
  HierData h;
  if (!t->HierDataPlane(&h)) return false;
  const int L = h.local_size, l = h.local_rank;
  const auto& column = *h.column;
  const int H = (int)column[0].size();
  // (1) cross-host ring over my column: afterwards `o` holds the blocks of every rank with my local index
  {
    std::vector<int64_t> cb((size_t)H), cd((size_t)H);
    for (int x = 0; x < H; ++x) { cb[(size_t)x] = bytes[(size_t)column[(size_t)l][(size_t)x]]; cd[(size_t)x] = displ[(size_t)column[(size_t)l][(size_t)x]]; }
    // the cross ring addresses blocks by their position in `o`: pass absolute displacements
    std::vector<int64_t> d2(cd.begin(), cd.end());
    d2.push_back(0);
    RingAllgatherv(h.cross, o, cb, d2);
  }
  // (2) inside the host: column c travels through the slot of the local rank c, one piece (of the concatenated column) at a time
  int64_t longest = 0;
  std::vector<int64_t> col_bytes((size_t)L, 0);
  for (int c = 0; c < L; ++c) { for (int r : column[(size_t)c]) col_bytes[(size_t)c] += bytes[(size_t)r]; longest = std::max(longest, col_bytes[(size_t)c]); }
  const int64_t S = (int64_t)h.local.slot_bytes;
  auto walk = [&](int c, int64_t piece_lo, int64_t piece_hi, char* slot, bool into_slot) {
    int64_t pos = 0;                                  // running offset inside the concatenated column c
    for (int r : column[(size_t)c]) {
      const int64_t blo = pos, bhi = pos + bytes[(size_t)r];
      const int64_t lo = std::max(blo, piece_lo), hi = std::min(bhi, piece_hi);
      if (hi > lo) {
        char* user = o + displ[(size_t)r] + (lo - blo);
        char* shm = slot + (lo - piece_lo);
        if (into_slot) memcpy(shm, user, (size_t)(hi - lo)); else memcpy(user, shm, (size_t)(hi - lo));
      }
      pos = bhi;
    }
  };
  for (int64_t done = 0; done < longest; done += S) {
    const int half = (int)(t->ShmNextPiece() & 1);
    walk(l, done, done + S, h.local.slot(l, half), true);
    t->LocalBarrier();
    for (int c = 0; c < L; ++c) if (c != l) walk(c, done, done + S, h.local.slot(c, half), false);
  }
  return true;
}

// Ring reduce-scatter in place: afterwards segment r of `b` (counts / off in elements) holds the reduction over all ranks.
void RingReducescatter(Transport* t, char* b, const std::vector<int64_t>& counts, const std::vector<int64_t>& off, DataType dtype, ReduceOp op) {
  const int n = t->size(), r = t->rank();
  if (n == 1) return;
  const size_t es = DataTypeSize(dtype);
  std::vector<char> tmp;
  const int64_t maxseg = *std::max_element(counts.begin(), counts.end());
  const int next = (r + 1) % n, prev = (r - 1 + n) % n;
  for (int s = 0; s < n - 1; ++s) {
    int si = (r - s - 1 + 2 * n) % n, ri = (r - s - 2 + 2 * n) % n;
    RingReduceStep(t, next, prev, b + off[si] * es, counts[si], b + off[ri] * es, counts[ri], maxseg, dtype, op, &tmp);
  }
}

// Multi-host, same number of ranks on every host.  (1) Inside a host, destination-major steps through the shm slots: local rank
// c ends up with the host's partial sums of every segment that belongs to a rank with local index c (on any host).  (2) The H
// ranks with my local index reduce-scatter those partial sums over their cross-host ring: 1/L of the bytes per TCP stream, and
// nothing that stays inside a host ever touches a socket.
bool HierReducescatter(Transport* t, const char* b, const std::vector<int64_t>& off, char* out, DataType dtype, ReduceOp op) {
  HierData h;
  if (!t->HierDataPlane(&h)) return false;
  const auto& column = *h.column;
  const int L = h.local_size, l = h.local_rank;
  const int H = (int)column[0].size();
  const size_t es = DataTypeSize(dtype);
  const int64_t sub = (int64_t)(h.local.slot_bytes / es) / L;
  if (sub < 16) return false;
  int x_me = -1;
  for (int x = 0; x < H; ++x) if (column[(size_t)l][(size_t)x] == t->rank()) x_me = x;
  if (x_me < 0 || h.cross->rank() != x_me) return false;
  auto seg_len = [&](int rank) { return off[(size_t)rank + 1] - off[(size_t)rank]; };
  std::vector<int64_t> col_len((size_t)L, 0);
  int64_t longest = 0;
  for (int c = 0; c < L; ++c) { for (int rk : column[(size_t)c]) col_len[(size_t)c] += seg_len(rk); longest = std::max(longest, col_len[(size_t)c]); }
  // copies elements [lo, lo + cnt) of the virtual vector "segments of column c, host by host" out of the user buffer
  auto gather = [&](int c, int64_t lo, int64_t cnt, char* dst) {
    int64_t pos = 0;
    for (int rk : column[(size_t)c]) {
      const int64_t blo = pos, bhi = pos + seg_len(rk);
      const int64_t a = std::max(blo, lo), z = std::min(bhi, lo + cnt);
      if (z > a) memcpy(dst + (a - lo) * es, b + (off[(size_t)rk] + (a - blo)) * es, (size_t)(z - a) * es);
      pos = bhi;
    }
  };
  std::vector<char> partial((size_t)std::max<int64_t>(col_len[(size_t)l], 1) * es);
  for (int64_t done = 0; done < longest; done += sub) {
    const int half = (int)(t->ShmNextPiece() & 1);
    char* slot = h.local.slot(l, half);
    for (int c = 0; c < L; ++c) {
      const int64_t cnt = std::min(sub, col_len[(size_t)c] - done);
      if (cnt > 0) gather(c, done, cnt, slot + (size_t)c * (size_t)sub * es);
    }
    t->LocalBarrier();
    const int64_t cnt = std::min(sub, col_len[(size_t)l] - done);
    if (cnt > 0) {
      char* dst = partial.data() + done * es;
      memcpy(dst, slot + (size_t)l * (size_t)sub * es, (size_t)cnt * es);
      for (int p = 1; p < L; ++p) ReduceInto(dst, h.local.slot((l + p) % L, half) + (size_t)l * (size_t)sub * es, cnt, dtype, op);
    }
  }
  std::vector<int64_t> cc((size_t)H), coff((size_t)H + 1, 0);
  for (int x = 0; x < H; ++x) { cc[(size_t)x] = seg_len(column[(size_t)l][(size_t)x]); coff[(size_t)x + 1] = coff[(size_t)x] + cc[(size_t)x]; }
  RingReducescatter(h.cross, partial.data(), cc, coff, dtype, op);
  if (cc[(size_t)x_me]) memcpy(out, partial.data() + coff[(size_t)x_me] * es, (size_t)cc[(size_t)x_me] * es);
  return true;
}

void TreeBroadcast(Transport* t, void* buf, int64_t bytes, int root);

// Two-level broadcast: across hosts only the ranks with the root's local index talk (binomial tree over H ranks instead of
// N), then every host hands the data to its other ranks through the shm slot of that local index.
bool HierBroadcast(Transport* t, char* buf, int64_t bytes, int root) {
  HierData h;
  if (!t->HierDataPlane(&h)) return false;
  const auto& column = *h.column;
  int root_l = -1, root_h = -1;
  for (int l = 0; l < h.local_size && root_l < 0; ++l)
    for (int x = 0; x < (int)column[(size_t)l].size(); ++x)
      if (column[(size_t)l][(size_t)x] == root) { root_l = l; root_h = x; break; }
  if (root_l < 0) return false;
  if (h.local_rank == root_l) TreeBroadcast(h.cross, buf, bytes, root_h);
  const int64_t S = (int64_t)h.local.slot_bytes;
  for (int64_t done = 0; done < bytes; done += S) {
    const int half = (int)(t->ShmNextPiece() & 1);
    const int64_t m = std::min(S, bytes - done);
    if (h.local_rank == root_l) memcpy(h.local.slot(root_l, half), buf + done, (size_t)m);
    t->LocalBarrier();
    if (h.local_rank != root_l) memcpy(buf + done, h.local.slot(root_l, half), (size_t)m);
  }
  return true;
}

}  // namespace

// ---------------------------------------------------------------------------

void Allreduce(Transport* t, void* buf, int64_t count, DataType dtype, ReduceOp op) {
  const int n = t->size();
  if (n == 1 || count == 0) return;
  char* b = (char*)buf;
  if (Took(0, ShmAllreduce(t, b, count, dtype, op))) return;
  if (Took(1, HierAllreduce(t, b, count, dtype, op))) return;
  Took(2, true);
  RingAllreduce(t, b, count, dtype, op);
}

void RingAllreduce(Transport* t, char* b, int64_t count, DataType dtype, ReduceOp op) {
  const int n = t->size(), r = t->rank();
  if (n == 1 || count == 0) return;
  const size_t es = DataTypeSize(dtype);
  if ((size_t)count * es < 32768 || count < n) {
    // latency regime: reduce at rank 0, broadcast
    if (r == 0) {
      std::vector<char> tmp((size_t)count * es);
      for (int p = 1; p < n; ++p) { t->Recv(p, tmp.data(), tmp.size()); ReduceInto(b, tmp.data(), count, dtype, op); }
      for (int p = 1; p < n; ++p) t->Send(p, b, (size_t)count * es);
    } else {
      t->Send(0, b, (size_t)count * es);
      t->Recv(0, b, (size_t)count * es);
    }
    return;
  }
  // bandwidth regime: ring reduce-scatter + ring allgather
  std::vector<int64_t> off(n + 1);
  for (int i = 0; i <= n; ++i) off[i] = count * i / n;
  int64_t maxseg = 0;
  for (int i = 0; i < n; ++i) maxseg = std::max(maxseg, off[i + 1] - off[i]);
  const int next = (r + 1) % n, prev = (r - 1 + n) % n;
  std::vector<char> tmp;
  for (int s = 0; s < n - 1; ++s) {
    const int si = (r - s + n) % n, ri = (r - s - 1 + n) % n;
    RingReduceStep(t, next, prev, b + off[si] * es, off[si + 1] - off[si], b + off[ri] * es, off[ri + 1] - off[ri], maxseg, dtype, op, &tmp);
  }
  for (int s = 0; s < n - 1; ++s) {
    int si = (r + 1 - s + n) % n, ri = (r - s + n) % n;
    t->SendRecv(next, b + off[si] * es, (size_t)(off[si + 1] - off[si]) * es, prev, b + off[ri] * es, (size_t)(off[ri + 1] - off[ri]) * es);
  }
}

void Allgatherv(Transport* t, const void* in, void* out, const std::vector<int64_t>& bytes) {
  const int n = t->size(), r = t->rank();
  std::vector<int64_t> displ(n + 1, 0);
  for (int i = 0; i < n; ++i) displ[i + 1] = displ[i] + bytes[i];
  char* o = (char*)out;
  if (in != o + displ[r] && bytes[r]) memcpy(o + displ[r], in, (size_t)bytes[r]);
  if (n == 1) return;
  if (Took(0, ShmAllgatherv(t, o + displ[r], o, bytes, displ))) return;
  if (Took(1, HierAllgatherv(t, o, bytes, displ))) return;
  Took(2, true);
  RingAllgatherv(t, o, bytes, displ);
}

void RingAllgatherv(Transport* t, char* o, const std::vector<int64_t>& bytes, const std::vector<int64_t>& displ) {
  const int n = t->size(), r = t->rank();
  if (n == 1) return;
  const int next = (r + 1) % n, prev = (r - 1 + n) % n;
  for (int s = 0; s < n - 1; ++s) {
    int si = (r - s + n) % n, ri = (r - s - 1 + n) % n;
    t->SendRecv(next, o + displ[si], (size_t)bytes[si], prev, o + displ[ri], (size_t)bytes[ri]);
  }
}

void Broadcast(Transport* t, void* buf, int64_t bytes, int root) {
  const int n = t->size();
  if (n == 1 || bytes == 0) return;
  if (Took(0, ShmBroadcast(t, (char*)buf, bytes, root))) return;
  if (Took(1, HierBroadcast(t, (char*)buf, bytes, root))) return;
  Took(2, true);
  TreeBroadcast(t, buf, bytes, root);
}

namespace {
void TreeBroadcast(Transport* t, void* buf, int64_t bytes, int root) {
  const int n = t->size(), r = t->rank();
  if (n == 1 || bytes == 0) return;
  int vr = (r - root + n) % n;
  const int64_t chunk = RingChunkElems(1);
  if (n >= 3 && chunk > 0 && bytes >= 4 * chunk) {
    // big message, three or more ranks: a chain in rank order, chunk by chunk — every link carries the message once and the
    // links run concurrently ((chunks + n - 2) chunk times), where the binomial tree makes the root send the WHOLE message
    // log2(n) times in a row
    char* b = (char*)buf;
    const int64_t pieces = (bytes + chunk - 1) / chunk;
    auto len = [&](int64_t k) { return (size_t)std::min(chunk, bytes - k * chunk); };
    const int prev = ((vr - 1 + n) % n + root) % n, next = ((vr + 1) % n + root) % n;
    if (vr == 0) {
      for (int64_t k = 0; k < pieces; ++k) t->Send(next, b + k * chunk, len(k));
    } else if (vr == n - 1) {
      for (int64_t k = 0; k < pieces; ++k) t->Recv(prev, b + k * chunk, len(k));
    } else {
      t->Recv(prev, b, len(0));
      for (int64_t k = 0; k < pieces; ++k) {
        if (k + 1 < pieces) t->SendRecv(next, b + k * chunk, len(k), prev, b + (k + 1) * chunk, len(k + 1));
        else t->Send(next, b + k * chunk, len(k));
      }
    }
    return;
  }
  // binomial tree rooted at `root`
  int mask = 1;
  while (mask < n) {
    if (vr & mask) { t->Recv(((vr - mask) + root) % n, buf, (size_t)bytes); break; }
    mask <<= 1;
  }
  mask >>= 1;
  while (mask > 0) {
    if (vr + mask < n) t->Send(((vr + mask) + root) % n, buf, (size_t)bytes);
    mask >>= 1;
  }
}
}  // namespace

void Alltoallv(Transport* t, const void* in, const std::vector<int64_t>& sb, void* out, const std::vector<int64_t>& rb) {
  const int n = t->size(), r = t->rank();
  std::vector<int64_t> sd(n + 1, 0), rd(n + 1, 0);
  for (int i = 0; i < n; ++i) { sd[i + 1] = sd[i] + sb[i]; rd[i + 1] = rd[i] + rb[i]; }
  const char* i8 = (const char*)in; char* o8 = (char*)out;
  if (sb[r]) memcpy(o8 + rd[r], i8 + sd[r], (size_t)sb[r]);
  if (n > 1 && Took(0, ShmAlltoallv(t, i8, sd, o8, rd))) return;
  if (n > 1 && Took(1, HierAlltoallv(t, i8, sd, o8, rd))) return;
  if (n > 1) { Took(2, true); t->AlltoallvBytes(i8, sd.data(), o8, rd.data()); }
}

void Reducescatter(Transport* t, void* buf, const std::vector<int64_t>& counts, void* out, DataType dtype, ReduceOp op) {
  const int n = t->size(), r = t->rank();
  const size_t es = DataTypeSize(dtype);
  std::vector<int64_t> off(n + 1, 0);
  for (int i = 0; i < n; ++i) off[i + 1] = off[i] + counts[i];
  char* b = (char*)buf;
  if (n > 1 && Took(0, ShmReducescatter(t, b, off, (char*)out, dtype, op))) return;
  if (n > 1 && Took(1, HierReducescatter(t, b, off, (char*)out, dtype, op))) return;
  if (n > 1) { Took(2, true); RingReducescatter(t, b, counts, off, dtype, op); }
  if (counts[r]) memcpy(out, b + off[r] * es, (size_t)counts[r] * es);
}

// ---------------------------------------------------------------------------
// Adasum VHDD

Status AdasumAllreduce(Transport* t, void* buf, const std::vector<int64_t>& tensor_counts, DataType dtype) {
  const int n = t->size(), r = t->rank();
  if (n == 1) return Status::OK();
  if (n & (n - 1)) return Status::PreconditionError("Running Adasum with non-power-of-2 ranks is not supported yet.");
  if (!(dtype == DataType::FLOAT16 || dtype == DataType::BFLOAT16 || dtype == DataType::FLOAT32 || dtype == DataType::FLOAT64))
    return Status::PreconditionError("Adasum supports only floating point tensors.");
  const size_t es = DataTypeSize(dtype);
  const int nt = (int)tensor_counts.size();
  std::vector<int64_t> toff(nt + 1, 0);
  for (int i = 0; i < nt; ++i) toff[i + 1] = toff[i] + tensor_counts[i];
  const int64_t total = toff[nt];
  char* b = (char*)buf;

  struct Lvl { int64_t start, len, kept_start, kept_len, sent_start, sent_len; };
  std::vector<Lvl> levels;
  int64_t start = 0, len = total;
  std::vector<char> recvbuf;
  std::vector<double> dots(3 * (size_t)nt), dots_tmp(3 * (size_t)nt);

  for (int level = 1; level < n; level <<= 1) {
    const int partner = r ^ level;
    const bool lower = (r & level) == 0;
    const int64_t h0 = len / 2, h1 = len - h0;
    Lvl L;
    L.start = start; L.len = len;
    L.kept_start = lower ? start : start + h0; L.kept_len = lower ? h0 : h1;
    L.sent_start = lower ? start + h0 : start; L.sent_len = lower ? h1 : h0;
    levels.push_back(L);
    recvbuf.resize((size_t)L.kept_len * es);
    t->SendRecv(partner, b + L.sent_start * es, (size_t)L.sent_len * es, partner, recvbuf.data(), (size_t)L.kept_len * es);
    // own kept piece belongs to A on the lower side, to B on the upper side
    const char* mine = b + L.kept_start * es;
    const char* theirs = recvbuf.data();
    std::fill(dots.begin(), dots.end(), 0.0);
    for (int ti = 0; ti < nt; ++ti) {
      int64_t lo = std::max(toff[ti], L.kept_start), hi = std::min(toff[ti + 1], L.kept_start + L.kept_len);
      double dab = 0, daa = 0, dbb = 0;
      for (int64_t e = lo; e < hi; ++e) {
        double m = LoadAsDouble(mine, e - L.kept_start, dtype), o = LoadAsDouble(theirs, e - L.kept_start, dtype);
        double a = lower ? m : o, bb = lower ? o : m;
        dab += a * bb; daa += a * a; dbb += bb * bb;
      }
      dots[3 * ti] = dab; dots[3 * ti + 1] = daa; dots[3 * ti + 2] = dbb;
    }
    // sum the partial dots over the 2*level ranks that jointly hold this vector pair
    for (int d = 1; d < 2 * level; d <<= 1) {
      int p = r ^ d;
      t->SendRecv(p, dots.data(), dots.size() * 8, p, dots_tmp.data(), dots_tmp.size() * 8);
      for (size_t i = 0; i < dots.size(); ++i) dots[i] += dots_tmp[i];
    }
    for (int ti = 0; ti < nt; ++ti) {
      int64_t lo = std::max(toff[ti], L.kept_start), hi = std::min(toff[ti + 1], L.kept_start + L.kept_len);
      if (lo >= hi) continue;
      const double dab = dots[3 * ti], daa = dots[3 * ti + 1], dbb = dots[3 * ti + 2];
      const double tiny = std::sqrt(DBL_MIN);
      double ac = 1.0, bc = 1.0;
      if (daa >= tiny) ac = 1.0 - dab / (2.0 * daa);
      if (dbb >= tiny) bc = 1.0 - dab / (2.0 * dbb);
      for (int64_t e = lo; e < hi; ++e) {
        double m = LoadAsDouble(mine, e - L.kept_start, dtype), o = LoadAsDouble(theirs, e - L.kept_start, dtype);
        double a = lower ? m : o, bb = lower ? o : m;
        StoreFromDouble(b + L.kept_start * es, e - L.kept_start, dtype, ac * a + bc * bb);
      }
    }
    start = L.kept_start; len = L.kept_len;
  }
  // distance-halving allgather: undo the splits in reverse order
  for (int li = (int)levels.size() - 1, level = n >> 1; li >= 0; --li, level >>= 1) {
    const Lvl& L = levels[li];
    const int partner = r ^ level;
    t->SendRecv(partner, b + L.kept_start * es, (size_t)L.kept_len * es, partner, b + L.sent_start * es, (size_t)L.sent_len * es);
  }
  return Status::OK();
}

}  // namespace cpu
}  // namespace hvd
