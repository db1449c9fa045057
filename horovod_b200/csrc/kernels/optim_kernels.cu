// Multi-tensor fused optimizer updates (SGD+momentum, Adam/AdamW) for sm_100a.
// One launch updates every parameter of the model: blockIdx.y selects the
// tensor from a device table, blockIdx.x strides over it with 16 B vectors.
// The reference leaves the optimizer to the framework (one or several kernels
// per parameter); SURVEY 2.5b lists "the optimizer update" as un-fused headroom.
// `grad_scale` lets DistributedOptimizer fold gradient post-division into the
// update instead of a separate elementwise pass.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "p2p_kernels.h"

namespace hvd {
namespace kern {
namespace {

constexpr int kOptThreads = 256;
constexpr int kOptVec = 4;
constexpr int kOptIters = 4;
constexpr int kOptPerBlock = kOptThreads * kOptVec * kOptIters;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

template <typename T> struct alignas(sizeof(T) * kOptVec) Vec4 { T v[kOptVec]; };

template <typename T> __device__ __forceinline__ void load4(const T* p, int64_t i, int64_t n, bool aligned, float* o) {
  if (aligned && i + kOptVec <= n) {
    Vec4<T> v = *reinterpret_cast<const Vec4<T>*>(p + i);
#pragma unroll
    for (int k = 0; k < kOptVec; ++k) o[k] = to_f<T>(v.v[k]);
  } else {
#pragma unroll
    for (int k = 0; k < kOptVec; ++k) o[k] = i + k < n ? to_f<T>(p[i + k]) : 0.f;
  }
}
template <typename T> __device__ __forceinline__ void store4(T* p, int64_t i, int64_t n, bool aligned, const float* o) {
  if (aligned && i + kOptVec <= n) {
    Vec4<T> v;
#pragma unroll
    for (int k = 0; k < kOptVec; ++k) v.v[k] = from_f<T>(o[k]);
    *reinterpret_cast<Vec4<T>*>(p + i) = v;
  } else {
#pragma unroll
    for (int k = 0; k < kOptVec; ++k) if (i + k < n) p[i + k] = from_f<T>(o[k]);
  }
}
template <typename T> __device__ __forceinline__ bool is_aligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) % (sizeof(T) * kOptVec)) == 0; }

template <typename P, typename G>
__global__ void __launch_bounds__(kOptThreads)
fused_sgd_kernel(const SgdTensor* __restrict__ table, float lr, float momentum, float dampening, float wd, int nesterov,
                 float grad_scale, int first_step) {
  const SgdTensor t = table[blockIdx.y];
  const int64_t n = t.count;
  const int64_t base = (int64_t)blockIdx.x * kOptPerBlock;
  if (base >= n) return;
  P* p = reinterpret_cast<P*>(t.param);
  const G* g = reinterpret_cast<const G*>(t.grad);
  P* mbuf = reinterpret_cast<P*>(t.momentum);
  const bool pa = is_aligned<P>(p) && (!mbuf || is_aligned<P>(mbuf)), ga = is_aligned<G>(g);
#pragma unroll
  for (int it = 0; it < kOptIters; ++it) {
    const int64_t i = base + ((int64_t)it * kOptThreads + threadIdx.x) * kOptVec;
    if (i >= n) break;
    float pv[kOptVec], gv[kOptVec], mv[kOptVec];
    load4<P>(p, i, n, pa, pv);
    load4<G>(g, i, n, ga, gv);
    if (mbuf && !first_step) load4<P>(mbuf, i, n, pa, mv);
#pragma unroll
    for (int k = 0; k < kOptVec; ++k) {
      float d = gv[k] * grad_scale;
      if (wd != 0.f) d = fmaf(wd, pv[k], d);
      if (mbuf) {
        float b = first_step ? d : fmaf(momentum, mv[k], (1.f - dampening) * d);
        mv[k] = b;
        d = nesterov ? fmaf(momentum, b, d) : b;
      }
      pv[k] = fmaf(-lr, d, pv[k]);
    }
    store4<P>(p, i, n, pa, pv);
    if (mbuf) store4<P>(mbuf, i, n, pa, mv);
  }
}

template <typename P, typename G>
__global__ void __launch_bounds__(kOptThreads)
fused_adam_kernel(const AdamTensor* __restrict__ table, float lr, float b1, float b2, float eps, float wd, float c1, float c2,
                  float grad_scale, int adamw) {
  const AdamTensor t = table[blockIdx.y];
  const int64_t n = t.count;
  const int64_t base = (int64_t)blockIdx.x * kOptPerBlock;
  if (base >= n) return;
  P* p = reinterpret_cast<P*>(t.param);
  const G* g = reinterpret_cast<const G*>(t.grad);
  float* m = reinterpret_cast<float*>(t.exp_avg);
  float* v = reinterpret_cast<float*>(t.exp_avg_sq);
  const bool pa = is_aligned<P>(p), ga = is_aligned<G>(g), sa = is_aligned<float>(m) && is_aligned<float>(v);
  const float rsc2 = rsqrtf(c2), step = lr / c1;
#pragma unroll
  for (int it = 0; it < kOptIters; ++it) {
    const int64_t i = base + ((int64_t)it * kOptThreads + threadIdx.x) * kOptVec;
    if (i >= n) break;
    float pv[kOptVec], gv[kOptVec], mv[kOptVec], vv[kOptVec];
    load4<P>(p, i, n, pa, pv);
    load4<G>(g, i, n, ga, gv);
    load4<float>(m, i, n, sa, mv);
    load4<float>(v, i, n, sa, vv);
#pragma unroll
    for (int k = 0; k < kOptVec; ++k) {
      float d = gv[k] * grad_scale;
      if (adamw) pv[k] *= (1.f - lr * wd); else if (wd != 0.f) d = fmaf(wd, pv[k], d);
      mv[k] = fmaf(b1, mv[k], (1.f - b1) * d);
      vv[k] = fmaf(b2, vv[k], (1.f - b2) * d * d);
      const float denom = sqrtf(vv[k]) * rsc2 + eps;
      pv[k] -= step * mv[k] / denom;
    }
    store4<P>(p, i, n, pa, pv);
    store4<float>(m, i, n, sa, mv);
    store4<float>(v, i, n, sa, vv);
  }
}

}  // namespace

cudaError_t LaunchFusedSgd(const SgdTensor* table, int n, int64_t max_count, float lr, float momentum, float dampening,
                           float weight_decay, int nesterov, float grad_scale, int first_step, int param_dtype,
                           int grad_dtype, cudaStream_t stream) {
  if (n <= 0 || max_count <= 0) return cudaSuccess;
  dim3 grid((unsigned)((max_count + kOptPerBlock - 1) / kOptPerBlock), (unsigned)n);
  if (param_dtype == 7 && grad_dtype == 7) fused_sgd_kernel<float, float><<<grid, kOptThreads, 0, stream>>>(table, lr, momentum, dampening, weight_decay, nesterov, grad_scale, first_step);
  else if (param_dtype == 7 && grad_dtype == 10) fused_sgd_kernel<float, __nv_bfloat16><<<grid, kOptThreads, 0, stream>>>(table, lr, momentum, dampening, weight_decay, nesterov, grad_scale, first_step);
  else if (param_dtype == 10 && grad_dtype == 10) fused_sgd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kOptThreads, 0, stream>>>(table, lr, momentum, dampening, weight_decay, nesterov, grad_scale, first_step);
  else if (param_dtype == 6 && grad_dtype == 6) fused_sgd_kernel<__half, __half><<<grid, kOptThreads, 0, stream>>>(table, lr, momentum, dampening, weight_decay, nesterov, grad_scale, first_step);
  else return cudaErrorInvalidValue;
  CountKernelLaunch();
  return cudaGetLastError();
}

cudaError_t LaunchFusedAdamW(const AdamTensor* table, int n, int64_t max_count, float lr, float beta1, float beta2,
                             float eps, float weight_decay, float bias_c1, float bias_c2, float grad_scale, int adamw,
                             int param_dtype, int grad_dtype, cudaStream_t stream) {
  if (n <= 0 || max_count <= 0) return cudaSuccess;
  dim3 grid((unsigned)((max_count + kOptPerBlock - 1) / kOptPerBlock), (unsigned)n);
  if (param_dtype == 7 && grad_dtype == 7) fused_adam_kernel<float, float><<<grid, kOptThreads, 0, stream>>>(table, lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2, grad_scale, adamw);
  else if (param_dtype == 7 && grad_dtype == 10) fused_adam_kernel<float, __nv_bfloat16><<<grid, kOptThreads, 0, stream>>>(table, lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2, grad_scale, adamw);
  else if (param_dtype == 10 && grad_dtype == 10) fused_adam_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, kOptThreads, 0, stream>>>(table, lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2, grad_scale, adamw);
  else return cudaErrorInvalidValue;
  CountKernelLaunch();
  return cudaGetLastError();
}

}  // namespace kern
}  // namespace hvd
