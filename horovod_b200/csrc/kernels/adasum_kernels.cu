// Adasum (adaptive summation) on GPUs over peer-mapped symmetric buffers, sm_100a.
//
// The reference only has a CPU implementation of the pairwise Adasum reduction
// (ops/adasum/adasum.h:195-435, vector-halving distance-doubling over MPI); its
// "GPU Adasum" inside one node is an NCCL sum divided by local_size
// (ops/adasum_gpu_operations.cc).  Here all log2(N) VHDD levels run on the GPUs:
//   pack      tensors -> fp32 fused vector in my symmetric buffer (prescale)
//   level l   partner = rank ^ 2^l; I keep one half K of my current segment.
//     dots    for K: a = lower group's copy, b = upper group's copy, where the
//             partner's copy is read DIRECTLY from its buffer over NVLink;
//             per-tensor partial <a,b>, |a|^2, |b|^2 accumulate in fp64 into a
//             scratch table that lives in symmetric memory
//     combine every rank sums the partials of the 2^(l+1) ranks that jointly hold
//             the vector pair (peer loads of their scratch tables), derives
//             acoeff = 1 - <a,b>/(2|a|^2), bcoeff = 1 - <a,b>/(2|b|^2) per tensor
//             and overwrites K with acoeff*a + bcoeff*b (again pulling b / a
//             from the partner's buffer)
//   gather    every rank pulls every other rank's final 1/N slice, applies
//             postscale, casts and scatters into the output tensors.
// Every kernel begins with the CTA-index flag barrier; together with stream
// order on each GPU this orders all cross-GPU reads after the writes they need.
#include "p2p_common.cuh"

namespace hvd {
namespace kern {
namespace {

constexpr int kRow = kThreads * 16;
constexpr int kAChunk = 16384;

template <typename F>
__device__ __forceinline__ void for_my_chunks(int64_t lo_b, int64_t hi_b, int cta, int grid, F f) {
  for (int64_t c = lo_b / kAChunk; c * kAChunk < hi_b; ++c) {
    if ((int)(c % grid) != cta) continue;
    int64_t lo = c * kAChunk, hi = lo + kAChunk;
    if (lo < lo_b) lo = lo_b;
    if (hi > hi_b) hi = hi_b;
    if (lo < hi) f(lo, hi);
  }
}

__device__ __forceinline__ float4 ld_f4(const void* p) {
  uint4 v = ld_stream(p);
  return make_float4(__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w));
}
__device__ __forceinline__ void st_f4(void* p, float4 f) {
  st_stream(p, make_uint4(__float_as_uint(f.x), __float_as_uint(f.y), __float_as_uint(f.z), __float_as_uint(f.w)));
}

// ---- pack: T -> fp32 fused vector -------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
adasum_pack_kernel(const __grid_constant__ CommParams cp, const TensorDesc* __restrict__ descs, int nd, int64_t total, float prescale) {
  const int cta = blockIdx.x, grid = gridDim.x;
  char* buf = reinterpret_cast<char*>(cp.buf[cp.rank]);
  DescCursor cur;
  cur.init(descs, nd, total);
  for_my_chunks(0, total, cta, grid, [&](int64_t lo, int64_t hi) {
    for (int64_t o = lo + (int64_t)threadIdx.x * 16; o < hi; o += kRow) {
      cur.seek(o);
      const TensorDesc& d = cur.d[cur.i];
      const int64_t e = (o - cur.lo) / 4;
      float a[4];
      load_elems<T, 4, float>(reinterpret_cast<const T*>(d.in) + e, d.count - e, a);
      st_f4(buf + o, make_float4(a[0] * prescale, a[1] * prescale, a[2] * prescale, a[3] * prescale));
    }
  });
}

// ---- level phase 1: partial dot products ---------------------------------------------
__global__ void __launch_bounds__(kThreads, 2)
adasum_dots_kernel(const __grid_constant__ CommParams cp, const TensorDesc* __restrict__ descs, int nd, int64_t total,
                   int64_t keep_lo, int64_t keep_hi, int partner, int lower, double* __restrict__ scratch) {
  extern __shared__ double s_acc[];  // [3 * nd]
  const int cta = blockIdx.x, grid = gridDim.x;
  uint32_t epoch = cp.epochs[cta];
  for (int i = threadIdx.x; i < 3 * nd; i += kThreads) s_acc[i] = 0.0;
  const bool alive = peer_barrier(cp, epoch, cta);  // also orders the smem zeroing (it ends with __syncthreads)
  const char* mine = reinterpret_cast<const char*>(cp.buf[cp.rank]);
  const char* theirs = reinterpret_cast<const char*>(cp.buf[partner]);
  DescCursor cur;
  cur.init(descs, nd, total);
  int t_cur = -1;
  double dab = 0, daa = 0, dbb = 0;
  auto flush = [&]() {
    if (t_cur >= 0 && (dab != 0 || daa != 0 || dbb != 0)) {
      atomicAdd(&s_acc[3 * t_cur], dab); atomicAdd(&s_acc[3 * t_cur + 1], daa); atomicAdd(&s_acc[3 * t_cur + 2], dbb);
    }
    dab = daa = dbb = 0;
  };
  if (alive) {
    for_my_chunks(keep_lo, keep_hi, cta, grid, [&](int64_t lo, int64_t hi) {
      for (int64_t o = lo + (int64_t)threadIdx.x * 16; o < hi; o += kRow) {
        const float4 m = ld_f4(mine + o), p = ld_f4(theirs + o);
        cur.seek(o);
        if (cur.i != t_cur) { flush(); t_cur = cur.i; }
        const float4 a = lower ? m : p, b = lower ? p : m;
        dab += (double)a.x * b.x + (double)a.y * b.y + (double)a.z * b.z + (double)a.w * b.w;
        daa += (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z + (double)a.w * a.w;
        dbb += (double)b.x * b.x + (double)b.y * b.y + (double)b.z * b.z + (double)b.w * b.w;
      }
    });
  }
  flush();
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * nd; i += kThreads) if (s_acc[i] != 0.0) atomicAdd(&scratch[i], s_acc[i]);
  if (threadIdx.x == 0) cp.epochs[cta] = epoch;
}

// ---- level phase 2: coefficients + in-place combine of my half ----------------------------
__global__ void __launch_bounds__(kThreads, 2)
adasum_combine_kernel(const __grid_constant__ CommParams cp, const TensorDesc* __restrict__ descs, int nd, int64_t total,
                      int64_t keep_lo, int64_t keep_hi, int partner, int lower, int group_base, int group_size,
                      int64_t scratch_byte_off) {
  extern __shared__ float s_coef[];  // [2 * nd]: acoeff, bcoeff
  const int cta = blockIdx.x, grid = gridDim.x;
  uint32_t epoch = cp.epochs[cta];
  const bool alive = peer_barrier(cp, epoch, cta);
  if (alive) {
    for (int t = threadIdx.x; t < nd; t += kThreads) {
      double dab = 0, daa = 0, dbb = 0;
      for (int g = 0; g < group_size; ++g) {
        const double* s = reinterpret_cast<const double*>(reinterpret_cast<const char*>(cp.flags[group_base + g]) + scratch_byte_off);
        dab += s[3 * t]; daa += s[3 * t + 1]; dbb += s[3 * t + 2];
      }
      const double tiny = 1.4916681462400413e-154;  // sqrt(DBL_MIN), as in the reference (adasum.h:397-404)
      s_coef[2 * t] = daa >= tiny ? (float)(1.0 - dab / (2.0 * daa)) : 1.f;
      s_coef[2 * t + 1] = dbb >= tiny ? (float)(1.0 - dab / (2.0 * dbb)) : 1.f;
    }
  }
  __syncthreads();
  char* mine = reinterpret_cast<char*>(cp.buf[cp.rank]);
  const char* theirs = reinterpret_cast<const char*>(cp.buf[partner]);
  DescCursor cur;
  cur.init(descs, nd, total);
  if (alive) {
    for_my_chunks(keep_lo, keep_hi, cta, grid, [&](int64_t lo, int64_t hi) {
      for (int64_t o = lo + (int64_t)threadIdx.x * 16; o < hi; o += kRow) {
        const float4 m = ld_f4(mine + o), p = ld_f4(theirs + o);
        cur.seek(o);
        const float ac = s_coef[2 * cur.i], bc = s_coef[2 * cur.i + 1];
        const float4 a = lower ? m : p, b = lower ? p : m;
        st_f4(mine + o, make_float4(ac * a.x + bc * b.x, ac * a.y + bc * b.y, ac * a.z + bc * b.z, ac * a.w + bc * b.w));
      }
    });
  }
  if (threadIdx.x == 0) cp.epochs[cta] = epoch;
}

// ---- final: pull every rank's slice, postscale, cast, scatter --------------------------------
struct GatherRanges { int64_t lo[kMaxPeers]; int64_t hi[kMaxPeers]; };

template <typename T>
__global__ void __launch_bounds__(kThreads, 2)
adasum_gather_kernel(const __grid_constant__ CommParams cp, const TensorDesc* __restrict__ descs, int nd, int64_t total,
                     const __grid_constant__ GatherRanges rg, float postscale) {
  const int cta = blockIdx.x, grid = gridDim.x;
  uint32_t epoch = cp.epochs[cta];
  const bool alive = peer_barrier(cp, epoch, cta);
  DescCursor cur;
  cur.init(descs, nd, total);
  for (int k = 0; k < cp.nranks && alive; ++k) {
    int p = cp.rank + k; if (p >= cp.nranks) p -= cp.nranks;
    const char* src = reinterpret_cast<const char*>(cp.buf[p]);
    cur.i = -1;
    for_my_chunks(rg.lo[p], rg.hi[p], cta, grid, [&](int64_t lo, int64_t hi) {
      for (int64_t o = lo + (int64_t)threadIdx.x * 16; o < hi; o += kRow) {
        const float4 v = ld_f4(src + o);
        cur.seek(o);
        const TensorDesc& d = cur.d[cur.i];
        const int64_t e = (o - cur.lo) / 4;
        if (e >= d.count) continue;
        float a[4] = {v.x * postscale, v.y * postscale, v.z * postscale, v.w * postscale};
        store_elems<T, 4, float>(reinterpret_cast<T*>(d.out) + e, d.count - e, a);
      }
    });
  }
  if (threadIdx.x == 0) cp.epochs[cta] = epoch;
}

// `only` < 0: the whole sequence (production: one GPU per process).  `only` = k: just the k-th launch of the sequence
// (pack = 0, dots/combine of level l = 1 + 2l / 2 + 2l, gather last) — the single-GPU simulation issues launch k of
// EVERY rank before launch k+1 of any rank, otherwise streams that share a hardware queue deadlock (a kernel spinning in
// its barrier sits in front of the peer kernel it is waiting for).
template <typename T>
cudaError_t run_adasum(const CommParams& cp, const AdasumArgs& a, double prescale, double postscale, cudaStream_t s, int only) {
  const int n = cp.nranks, rank = cp.rank;
  const int64_t total = a.total_bytes;
  int step = 0;
  if (only < 0 || only == step) {
    adasum_pack_kernel<T><<<a.ctas, kThreads, 0, s>>>(cp, a.descs, a.ndesc, total, (float)prescale);
    CountKernelLaunch();
  }
  ++step;
  // replay the halving for every rank (element granularity: 4 floats = one 16 B vector)
  GatherRanges rg;
  for (int p = 0; p < n; ++p) {
    int64_t lo = 0, hi = total;
    for (int d = 1; d < n; d <<= 1) {
      int64_t mid = lo + (((hi - lo) / 2) & ~(int64_t)15);
      if ((p & d) == 0) hi = mid; else lo = mid;
    }
    rg.lo[p] = lo; rg.hi[p] = hi;
  }
  int64_t lo = 0, hi = total;
  int level = 0;
  const size_t dots_smem = (size_t)3 * a.ndesc * sizeof(double), coef_smem = (size_t)2 * a.ndesc * sizeof(float);
  for (int d = 1; d < n; d <<= 1, ++level) {
    const int partner = rank ^ d, lower = (rank & d) == 0;
    const int64_t mid = lo + (((hi - lo) / 2) & ~(int64_t)15);
    const int64_t klo = lower ? lo : mid, khi = lower ? mid : hi;
    const int64_t off = kFlagWords * 4 + (int64_t)(level & 1) * a.scratch_stride_bytes;
    double* scratch = reinterpret_cast<double*>(reinterpret_cast<char*>(cp.flags[rank]) + off);
    if (only < 0 || only == step) {
      cudaError_t e = cudaMemsetAsync(scratch, 0, (size_t)3 * a.ndesc * sizeof(double), s);
      if (e != cudaSuccess) return e;
      adasum_dots_kernel<<<a.ctas, kThreads, dots_smem, s>>>(cp, a.descs, a.ndesc, total, klo, khi, partner, lower, scratch);
      CountKernelLaunch();
    }
    ++step;
    if (only < 0 || only == step) {
      adasum_combine_kernel<<<a.ctas, kThreads, coef_smem, s>>>(cp, a.descs, a.ndesc, total, klo, khi, partner, lower,
                                                                 rank & ~(2 * d - 1), 2 * d, off);
      CountKernelLaunch();
    }
    ++step;
    lo = klo; hi = khi;
  }
  if (only < 0 || only == step) {
    adasum_gather_kernel<T><<<a.ctas, kThreads, 0, s>>>(cp, a.descs, a.ndesc, total, rg, (float)postscale);
    CountKernelLaunch();
  }
  return cudaGetLastError();
}

}  // namespace

int AdasumNumLaunches(int nranks) {
  int levels = 0;
  for (int d = 1; d < nranks; d <<= 1) ++levels;
  return 2 + 2 * levels;
}

cudaError_t LaunchAdasum(const CommParams& cp, const AdasumArgs& args, double prescale, double postscale, cudaStream_t stream) {
  return LaunchAdasumStep(cp, args, prescale, postscale, stream, -1);
}

cudaError_t LaunchAdasumStep(const CommParams& cp, const AdasumArgs& args, double prescale, double postscale, cudaStream_t stream,
                             int only) {
  if (args.ctas < 1 || args.ctas > kMaxCtas || (cp.nranks & (cp.nranks - 1))) return cudaErrorInvalidValue;
  if (args.ndesc > kAdasumMaxTensors) return cudaErrorInvalidValue;  // smem / scratch table bound
  switch (args.dtype) {
    case 7: return run_adasum<float>(cp, args, prescale, postscale, stream, only);
    case 6: return run_adasum<__half>(cp, args, prescale, postscale, stream, only);
    case 10: return run_adasum<__nv_bfloat16>(cp, args, prescale, postscale, stream, only);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace kern
}  // namespace hvd
