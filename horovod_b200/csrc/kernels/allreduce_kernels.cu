// Fused allreduce / reducescatter over peer-mapped symmetric buffers (sm_100a).
//
// One launch per fused response:
//   pack     every tensor -> my symmetric buffer   (prescale, cast T -> wire W)
//   barrier  CTA b of all ranks (release/acquire flags written into peer memory)
//   one-shot : each rank loads the same 16 B vector from all N peers over
//              NVLink, reduces in fp32/int32/fp64 registers, applies
//              postscale (Average), casts W -> T and stores straight into the
//              output tensors.                                  [(N-1)·S in]
//   two-shot : rank r reduces the chunks it owns and stores the result into
//              every peer's buffer, barrier, every rank unpacks its own buffer
//              into the output tensors.                         [2(N-1)/N·S]
//   NVLS     : same as two-shot but the reduction is one
//              multimem.ld_reduce (the NVSwitch adds) and the broadcast one
//              multimem.st on the multicast mapping.            [~S/N in+out]
// The chunk -> CTA mapping is the same in all phases (chunk c belongs to CTA
// c % grid on every rank), so CTA b only ever needs to synchronise with CTA b
// of its peers: no grid-wide barrier, no cooperative launch.
//
// Replaces: batched_scaled_memcpy_k + ncclAllReduce + batched_scaled_memcpy_k
// (reference ops/cuda/cuda_kernels.cu:259-324, ops/nccl_operations.cc:231-283).
#include <cstdlib>
#include "p2p_common.cuh"

namespace hvd {
namespace kern {

static unsigned long long g_launches = 0;
unsigned long long KernelLaunchCount() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
void CountKernelLaunch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

namespace {

constexpr int kRowBytes = kThreads * 16;  // one 16 B vector per thread

template <typename W> struct Nvls { static constexpr bool ok = false; };
template <> struct Nvls<float> {
  static constexpr bool ok = true;
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
template <> struct Nvls<__nv_bfloat16> {
  static constexpr bool ok = true;
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
template <> struct Nvls<__half> {
  static constexpr bool ok = true;
  static __device__ __forceinline__ uint4 ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.f16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
__device__ __forceinline__ void multimem_st(void* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- phase bodies ----------------------------------------------------------
// Every loop handles U rows (U * 8 KiB) per trip: all loads of a trip are issued
// before the first dependent store so each thread keeps U (x N peers) 16 B
// requests in flight — NVLink round trips are ~2 us, HBM ~0.7 us.

template <typename T, typename W, int U>
__device__ __forceinline__ void pack_range(const TensorDesc* descs, int nd, int64_t total, char* buf, int64_t lo, int64_t hi,
                                           typename ScaleOf<typename Traits<W>::Acc>::type scale) {
  using A = typename Traits<W>::Acc;
  constexpr int NW = 16 / (int)sizeof(W);
  constexpr int64_t kSrcRow = (int64_t)kRowBytes / (int)sizeof(W) * (int)sizeof(T);  // source bytes per wire row
  DescCursor cur;
  cur.init(descs, nd, total);
  for (int64_t o0 = lo + (int64_t)threadIdx.x * 16; o0 < hi; o0 += (int64_t)U * kRowBytes) {
    A a[U][NW];
    cur.seek(o0);
    const int64_t olast = o0 + (int64_t)(U - 1) * kRowBytes;
    const int64_t e0 = (o0 - cur.lo) / (int64_t)sizeof(W);
    const char* src = cur.in + e0 * (int64_t)sizeof(T);
    // fast path: all U vectors of this trip are full vectors of ONE tensor and the source is 16 B aligned (the row
    // stride is a multiple of 16 B, so one check covers the trip): no cursor work, no bounds checks
    if (olast < hi && olast + 16 <= cur.lo + cur.count * (int64_t)sizeof(W) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < U; ++j) load_elems_fast<T, NW, A>(reinterpret_cast<const T*>(src + j * kSrcRow), a[j]);
#pragma unroll
      for (int j = 0; j < U; ++j) {
#pragma unroll
        for (int i = 0; i < NW; ++i) a[j][i] = apply_scale<A>(a[j][i], scale);
        st_stream(buf + o0 + (int64_t)j * kRowBytes, pack_vec<W, NW>(a[j]));
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t o = o0 + (int64_t)j * kRowBytes;
      if (o < hi) {
        cur.seek(o);
        const int64_t e = (o - cur.lo) / (int64_t)sizeof(W);
        load_elems<T, NW, A>(reinterpret_cast<const T*>(cur.in) + e, cur.count - e, a[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t o = o0 + (int64_t)j * kRowBytes;
      if (o < hi) {
#pragma unroll
        for (int i = 0; i < NW; ++i) a[j][i] = apply_scale<A>(a[j][i], scale);
        st_stream(buf + o, pack_vec<W, NW>(a[j]));
      }
    }
  }
}

template <typename T, typename W, int U>
__device__ __forceinline__ void unpack_range(const TensorDesc* descs, int nd, int64_t total, const char* buf, int64_t lo,
                                             int64_t hi) {
  using A = typename Traits<W>::Acc;
  constexpr int NW = 16 / (int)sizeof(W);
  constexpr int64_t kDstRow = (int64_t)kRowBytes / (int)sizeof(W) * (int)sizeof(T);
  DescCursor cur;
  cur.init(descs, nd, total);
  for (int64_t o0 = lo + (int64_t)threadIdx.x * 16; o0 < hi; o0 += (int64_t)U * kRowBytes) {
    uint4 v[U];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t o = o0 + (int64_t)j * kRowBytes;
      if (o < hi) v[j] = ld_stream(buf + o);
    }
    cur.seek(o0);
    const int64_t olast = o0 + (int64_t)(U - 1) * kRowBytes;
    const int64_t e0 = (o0 - cur.lo) / (int64_t)sizeof(W);
    char* dst = cur.out + e0 * (int64_t)sizeof(T);
    if (olast < hi && olast + 16 <= cur.lo + cur.count * (int64_t)sizeof(W) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
      for (int j = 0; j < U; ++j) {
        A a[NW];
        unpack_vec<W, NW>(v[j], a);
        store_elems_fast<T, NW, A>(reinterpret_cast<T*>(dst + j * kDstRow), a);
      }
      continue;
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t o = o0 + (int64_t)j * kRowBytes;
      if (o < hi) {
        cur.seek(o);
        const int64_t e = (o - cur.lo) / (int64_t)sizeof(W);
        if (e < cur.count) {
          A a[NW];
          unpack_vec<W, NW>(v[j], a);
          store_elems<T, NW, A>(reinterpret_cast<T*>(cur.out) + e, cur.count - e, a);
        }
      }
    }
  }
}

// Issues the loads of the vector at byte offset `o` from all peers.
template <int NR> __device__ __forceinline__ void peer_loads(const CommParams& cp, int64_t o, uint4* v) {
#pragma unroll
  for (int k = 0; k < NR; ++k) {
    // start at my own buffer and rotate so that at any instant the N ranks pull from N different peers
    int p = cp.rank + k; if (p >= cp.nranks) p -= cp.nranks;
    if (k < cp.nranks) v[k] = ld_stream(reinterpret_cast<const char*>(cp.buf[p]) + o);
  }
}
template <typename W, int NR>
__device__ __forceinline__ void peer_combine(const CommParams& cp, const uint4* v, int op, typename Traits<W>::Acc* acc) {
  using A = typename Traits<W>::Acc;
  constexpr int NW = 16 / (int)sizeof(W);
  unpack_vec<W, NW>(v[0], acc);
#pragma unroll
  for (int k = 1; k < NR; ++k) {
    if (k < cp.nranks) {
      A b[NW];
      unpack_vec<W, NW>(v[k], b);
#pragma unroll
      for (int i = 0; i < NW; ++i) acc[i] = combine<A>(acc[i], b[i], op);
    }
  }
}

template <typename T, typename W, int NR>
__global__ void __launch_bounds__(kThreads, 2)
allreduce_kernel(const __grid_constant__ CommParams cp, const __grid_constant__ AllreduceArgs a, const int chunk_bytes) {
  using A = typename Traits<W>::Acc;
  using S = typename ScaleOf<A>::type;
  constexpr int NW = 16 / (int)sizeof(W);
  constexpr int U = 8 / NR;  // rows per trip in the peer phases
  constexpr int UP = NW >= 16 ? 4 : 8;  // rows per trip in the local pack / unpack phases (8 x 16 B loads in flight per thread)
  const int cta = blockIdx.x, grid = gridDim.x;
  const TensorDesc* descs = a.descs ? a.descs : a.inline_descs;
  const TensorDesc* odescs = a.out_descs ? a.out_descs : descs;
  const int nout = a.out_descs ? a.nout : a.ndesc;
  uint32_t epoch = cp.epochs[cta];
  char* mybuf = reinterpret_cast<char*>(cp.buf[cp.rank]);
  const int64_t total = a.total_bytes;
  const int64_t nchunks = (total + chunk_bytes - 1) / chunk_bytes;
  const S prescale = (S)a.prescale, postscale = (S)a.postscale;
  bool alive = true;

  // ---- pack ----
  for (int64_t c = cta; c < nchunks; c += grid) {
    const int64_t lo = c * chunk_bytes;
    const int64_t hi = lo + chunk_bytes < total ? lo + chunk_bytes : total;
    pack_range<T, W, UP>(descs, a.ndesc, total, mybuf, lo, hi, prescale);
  }
  if (cp.nranks > 1) alive = peer_barrier(cp, epoch, cta); else __syncthreads();

  if (a.variant == kOneShot) {
    // ---- reduce straight into the outputs ----
    DescCursor cur;
    cur.init(odescs, nout, total);
    for (int64_t c = cta; c < nchunks && alive; c += grid) {
      int64_t lo = c * chunk_bytes;
      int64_t hi = lo + chunk_bytes < total ? lo + chunk_bytes : total;
      if (lo < a.reduce_lo) lo = a.reduce_lo;
      if (hi > a.reduce_hi) hi = a.reduce_hi;
      if (a.oneshot_nvls) {
        // reduce-scatter through the switch: one multimem.ld_reduce per vector instead of N peer loads
        if constexpr (Nvls<W>::ok) {
          constexpr int UN = 4;
          for (int64_t o0 = lo + (int64_t)threadIdx.x * 16; o0 < hi; o0 += (int64_t)UN * kRowBytes) {
            uint4 v[UN];
#pragma unroll
            for (int j = 0; j < UN; ++j) {
              const int64_t o = o0 + (int64_t)j * kRowBytes;
              if (o < hi) v[j] = Nvls<W>::ld_reduce(reinterpret_cast<const char*>(cp.mc_buf) + o);
            }
#pragma unroll
            for (int j = 0; j < UN; ++j) {
              const int64_t o = o0 + (int64_t)j * kRowBytes;
              if (o < hi) {
                A acc[NW];
                unpack_vec<W, NW>(v[j], acc);
                cur.seek(o);
                const int64_t e = (o - cur.lo) / (int64_t)sizeof(W);
                if (e < cur.count) {
#pragma unroll
                  for (int i = 0; i < NW; ++i) acc[i] = apply_scale<A>(acc[i], postscale);
                  store_elems<T, NW, A>(reinterpret_cast<T*>(cur.out) + e, cur.count - e, acc);
                }
              }
            }
          }
        }
        continue;
      }
      for (int64_t o0 = lo + (int64_t)threadIdx.x * 16; o0 < hi; o0 += (int64_t)U * kRowBytes) {
        uint4 v[U][NR];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int64_t o = o0 + (int64_t)j * kRowBytes;
          if (o < hi) peer_loads<NR>(cp, o, v[j]);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int64_t o = o0 + (int64_t)j * kRowBytes;
          if (o < hi) {
            A acc[NW];
            peer_combine<W, NR>(cp, v[j], a.op, acc);
            cur.seek(o);
            const int64_t e = (o - cur.lo) / (int64_t)sizeof(W);
            if (e < cur.count) {
#pragma unroll
              for (int i = 0; i < NW; ++i) acc[i] = apply_scale<A>(acc[i], postscale);
              store_elems<T, NW, A>(reinterpret_cast<T*>(cur.out) + e, cur.count - e, acc);
            }
          }
        }
      }
    }
  } else {
    // ---- reduce the chunks I own, publish to every peer ----
    for (int64_t c = cta, k = 0; c < nchunks && alive; c += grid, ++k) {
      int owner = (int)((k + cta) % cp.nranks);
      if (owner != cp.rank) continue;
      const int64_t lo = c * chunk_bytes;
      const int64_t hi = lo + chunk_bytes < total ? lo + chunk_bytes : total;
      if (a.variant == kNvls) {
        if constexpr (Nvls<W>::ok) {
          constexpr int UN = 4;
          for (int64_t o0 = lo + (int64_t)threadIdx.x * 16; o0 < hi; o0 += (int64_t)UN * kRowBytes) {
            uint4 v[UN];
#pragma unroll
            for (int j = 0; j < UN; ++j) {
              const int64_t o = o0 + (int64_t)j * kRowBytes;
              if (o < hi) v[j] = Nvls<W>::ld_reduce(reinterpret_cast<const char*>(cp.mc_buf) + o);
            }
#pragma unroll
            for (int j = 0; j < UN; ++j) {
              const int64_t o = o0 + (int64_t)j * kRowBytes;
              if (o < hi) {
                if (postscale != (S)1) {
                  A acc[NW];
                  unpack_vec<W, NW>(v[j], acc);
#pragma unroll
                  for (int i = 0; i < NW; ++i) acc[i] = apply_scale<A>(acc[i], postscale);
                  v[j] = pack_vec<W, NW>(acc);
                }
                multimem_st(reinterpret_cast<char*>(cp.mc_buf) + o, v[j]);
              }
            }
          }
        }
      } else {
        for (int64_t o0 = lo + (int64_t)threadIdx.x * 16; o0 < hi; o0 += (int64_t)U * kRowBytes) {
          uint4 v[U][NR];
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const int64_t o = o0 + (int64_t)j * kRowBytes;
            if (o < hi) peer_loads<NR>(cp, o, v[j]);
          }
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const int64_t o = o0 + (int64_t)j * kRowBytes;
            if (o < hi) {
              A acc[NW];
              peer_combine<W, NR>(cp, v[j], a.op, acc);
#pragma unroll
              for (int i = 0; i < NW; ++i) acc[i] = apply_scale<A>(acc[i], postscale);
              const uint4 r = pack_vec<W, NW>(acc);
#pragma unroll
              for (int p = 0; p < NR; ++p) {
                int q = cp.rank + p; if (q >= cp.nranks) q -= cp.nranks;
                if (p < cp.nranks) st_stream(reinterpret_cast<char*>(cp.buf[q]) + o, r);
              }
            }
          }
        }
      }
    }
    if (alive) alive = peer_barrier(cp, epoch, cta);
    // ---- unpack my buffer ----
    for (int64_t c = cta; c < nchunks && alive; c += grid) {
      int64_t lo = c * chunk_bytes;
      int64_t hi = lo + chunk_bytes < total ? lo + chunk_bytes : total;
      if (lo < a.reduce_lo) lo = a.reduce_lo;
      if (hi > a.reduce_hi) hi = a.reduce_hi;
      if (lo < hi) unpack_range<T, W, UP>(odescs, nout, total, mybuf, lo, hi);
    }
  }
  if (threadIdx.x == 0) cp.epochs[cta] = epoch;
}

// One chunk of the in-switch allreduce: UN independent multimem.ld_reduce per thread are issued before the first
// multimem.st so UN x 16 B per thread are in flight through the NVSwitch.
template <typename W, int UN>
__device__ __forceinline__ void nvls_inplace_rows(char* mc, int64_t lo, int64_t hi, typename ScaleOf<typename Traits<W>::Acc>::type scale) {
  using A = typename Traits<W>::Acc;
  using S = typename ScaleOf<A>::type;
  constexpr int NW = 16 / (int)sizeof(W);
  for (int64_t o0 = lo + (int64_t)threadIdx.x * 16; o0 < hi; o0 += (int64_t)UN * kRowBytes) {
    uint4 v[UN];
#pragma unroll
    for (int j = 0; j < UN; ++j) { const int64_t o = o0 + (int64_t)j * kRowBytes; if (o < hi) v[j] = Nvls<W>::ld_reduce(mc + o); }
#pragma unroll
    for (int j = 0; j < UN; ++j) {
      const int64_t o = o0 + (int64_t)j * kRowBytes;
      if (o < hi) {
        if (scale != (S)1) {
          A acc[NW];
          unpack_vec<W, NW>(v[j], acc);
#pragma unroll
          for (int i = 0; i < NW; ++i) acc[i] = apply_scale<A>(acc[i], scale);
          v[j] = pack_vec<W, NW>(acc);
        }
        multimem_st(mc + o, v[j]);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Zero-copy allreduce on REGISTERED tensors (memory allocated with hvd.symm_empty / the bucketed DistributedOptimizer):
// the user tensor itself is peer-mapped, so there is no pack and no unpack — the kernel is only the NVLink phase.
//   barrier A   every rank's data is final (the kernel was ordered after the producer's ready event)
//   reduce      rank r sums the chunks it owns out of all N tensors and stores the result into all N tensors
//               (multimem.ld_reduce + multimem.st when the region is multicast-bound: the NVSwitch does the add
//               and the broadcast); a chunk has exactly one reader — its owner — so writing in place is safe
//   barrier B   every result landed
template <typename W, int NR>
__global__ void __launch_bounds__(kThreads, 2)
inplace_allreduce_kernel(const __grid_constant__ CommParams cp, const __grid_constant__ InplaceArgs a) {
  using A = typename Traits<W>::Acc;
  using S = typename ScaleOf<A>::type;
  constexpr int NW = 16 / (int)sizeof(W);
  constexpr int U = 8 / NR;
  const int cta = blockIdx.x, grid = gridDim.x;
  uint32_t epoch = cp.epochs[cta];
  const int64_t total = a.bytes;  // multiple of 16
  const int chunk_bytes = a.chunk_bytes;
  const int64_t nchunks = (total + chunk_bytes - 1) / chunk_bytes;
  const S scale = (S)a.scale;
  bool alive = peer_barrier(cp, epoch, cta);
  for (int64_t c = cta, k = 0; c < nchunks && alive; c += grid, ++k) {
    if ((int)((k + cta) % cp.nranks) != cp.rank) continue;
    const int64_t lo = c * chunk_bytes;
    const int64_t hi = lo + chunk_bytes < total ? lo + chunk_bytes : total;
    if (a.use_multicast) {
      if constexpr (Nvls<W>::ok) {
        char* mc = reinterpret_cast<char*>(a.mc);
        if (a.nvls_unroll == 8) nvls_inplace_rows<W, 8>(mc, lo, hi, scale);
        else nvls_inplace_rows<W, 4>(mc, lo, hi, scale);
      }
    } else {
      for (int64_t o0 = lo + (int64_t)threadIdx.x * 16; o0 < hi; o0 += (int64_t)U * kRowBytes) {
        uint4 v[U][NR];
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int64_t o = o0 + (int64_t)j * kRowBytes;
          if (o < hi) {
#pragma unroll
            for (int p = 0; p < NR; ++p) {
              int q = cp.rank + p; if (q >= cp.nranks) q -= cp.nranks;
              if (p < cp.nranks) v[j][p] = ld_stream(reinterpret_cast<const char*>(a.ptr[q]) + o);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int64_t o = o0 + (int64_t)j * kRowBytes;
          if (o < hi) {
            A acc[NW];
            peer_combine<W, NR>(cp, v[j], a.op, acc);
#pragma unroll
            for (int i = 0; i < NW; ++i) acc[i] = apply_scale<A>(acc[i], scale);
            const uint4 r = pack_vec<W, NW>(acc);
#pragma unroll
            for (int p = 0; p < NR; ++p) {
              int q = cp.rank + p; if (q >= cp.nranks) q -= cp.nranks;
              if (p < cp.nranks) st_stream(reinterpret_cast<char*>(a.ptr[q]) + o, r);
            }
          }
        }
      }
    }
  }
  if (alive) peer_barrier(cp, epoch, cta);
  if (threadIdx.x == 0) cp.epochs[cta] = epoch;
}


// ---------------------------------------------------------------------------
// Software-pipelined allreduce for LARGE fused responses of ordinary (unregistered) tensors (opt-in,
// HVD_PIPELINED_ALLREDUCE=1; variant kPipelined).
//
// The three-phase kernel above runs pack, reduce and unpack back to back: NVLink idles while HBM is busy and vice versa
// (ncu + sweeps: 451 GB/s busBW at 1 GiB vs 834 GB/s for the zero-copy kernel).  Here the message is cut into chunks that
// travel through a ring of slots in the symmetric buffer, and the CTAs are specialised by blockIdx & 3:
//     0     pack    chunk c   : tensors -> my slot (prescale, cast)
//     1, 2  reduce  chunk c-1 : my 1/N slice of the slot on ALL ranks: multimem.ld_reduce + multimem.st (or P2P loads /
//                               stores), postscale folded in
//     3     unpack  chunk c-2 : my slot -> output tensors (cast)
// so HBM traffic of chunks c and c-2 overlaps the NVLink traffic of chunk c-1.  Hand-offs (all monotonic counters,
// `pipe_base` + chunk index + 1, never reset):
//     packed[r]   written by rank r's LAST pack CTA of a chunk into every rank's flag region (st.release.sys)
//                 reduce waits for packed[q] of every rank q before touching chunk c
//     reduced[r]  written by rank r's LAST reduce CTA of a chunk into every rank's flag region
//                 unpack waits for reduced[q] of every rank q (all slices of my slot are final)
//     unpack_done local: the slot of chunk c may be re-packed with chunk c + slots once chunk c is unpacked here (which
//                 implies that every peer finished reading it)
// "last CTA of a chunk" is found with a per-slot arrival counter (threadfence + atomicAdd, reset by the last arriver
// BEFORE it publishes).  Roles are interleaved over blockIdx so that any resident prefix of the grid contains all of
// them; the grid must nevertheless fit on the device (every role waits for the others).

__device__ __forceinline__ uint32_t* pipe_area(const CommParams& cp, int r) {
  return reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(cp.flags[r]) + kPipeAreaOffset);
}

// Threads [0, n) each spin on words[i] until it reaches `target` (wrap-safe), with the abort / timeout protocol of
// peer_barrier.  Returns false when the wait was abandoned.  Ends with __syncthreads().
__device__ __forceinline__ bool pipe_wait(const CommParams& cp, const uint32_t* words, int n, uint32_t target) {
  __shared__ int s_abort;
  if (threadIdx.x == 0) s_abort = 0;
  __syncthreads();
  if ((int)threadIdx.x < n) {
    const uint32_t* w = words + threadIdx.x;
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while ((int32_t)(ld_relaxed_sys(w) - target) < 0) {
      if ((++spins & 0x3fff) == 0 && cp.abort_flag) {
        if (*reinterpret_cast<volatile int*>(cp.abort_flag)) { s_abort = 1; break; }
        if (cp.timeout_ns) {
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          else if (now - t0 > cp.timeout_ns) { *reinterpret_cast<volatile int*>(cp.abort_flag) = 2; s_abort = 1; break; }
        }
      }
    }
    (void)ld_acquire_sys(w);
  }
  __syncthreads();
  const int aborted = s_abort;
  __syncthreads();  // everyone has read the verdict before the next call may reset it
  return aborted == 0;
}

// All threads of the CTA call it after their last store of a chunk.  Returns true in exactly one CTA per chunk (the
// one that arrives last); that CTA has already reset the counter for the slot's next use.
__device__ __forceinline__ bool pipe_last_cta(uint32_t* counter, uint32_t nctas) {
  __shared__ int s_last;
  __threadfence_system();  // my stores (to local HBM, peers, or the multicast alias) before my arrival
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t old = atomicAdd(counter, 1u);
    s_last = old == nctas - 1;
    if (s_last) { atomicExch(counter, 0u); __threadfence_system(); }
  }
  __syncthreads();
  return s_last != 0;
}

template <typename W, int NR>
__device__ __forceinline__ void pipe_p2p_reduce(const CommParams& cp, int64_t delta, int64_t lo, int64_t hi, int op,
                                                typename ScaleOf<typename Traits<W>::Acc>::type postscale) {
  // [lo, hi) are fused-buffer offsets; the data lives at buf[q] + delta + offset on every rank q
  using A = typename Traits<W>::Acc;
  constexpr int NW = 16 / (int)sizeof(W);
  constexpr int U = 8 / NR;
  for (int64_t o0 = lo + (int64_t)threadIdx.x * 16; o0 < hi; o0 += (int64_t)U * kRowBytes) {
    uint4 v[U][NR];
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t o = o0 + (int64_t)j * kRowBytes;
      if (o < hi) peer_loads<NR>(cp, delta + o, v[j]);
    }
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const int64_t o = o0 + (int64_t)j * kRowBytes;
      if (o < hi) {
        A acc[NW];
        peer_combine<W, NR>(cp, v[j], op, acc);
#pragma unroll
        for (int i = 0; i < NW; ++i) acc[i] = apply_scale<A>(acc[i], postscale);
        const uint4 r = pack_vec<W, NW>(acc);
#pragma unroll
        for (int p = 0; p < NR; ++p) {
          int q = cp.rank + p; if (q >= cp.nranks) q -= cp.nranks;
          if (p < cp.nranks) st_stream(reinterpret_cast<char*>(cp.buf[q]) + delta + o, r);
        }
      }
    }
  }
}

template <typename T, typename W, int NR>
__global__ void __launch_bounds__(kThreads, 2)
pipelined_allreduce_kernel(const __grid_constant__ CommParams cp, const __grid_constant__ AllreduceArgs a) {
  using A = typename Traits<W>::Acc;
  using S = typename ScaleOf<A>::type;
  constexpr int NW = 16 / (int)sizeof(W);
  constexpr int UP = NW >= 16 ? 4 : 8;
  const int role = blockIdx.x & 3, quad = blockIdx.x >> 2, nquad = gridDim.x >> 2;
  const TensorDesc* descs = a.descs ? a.descs : a.inline_descs;
  const int64_t total = a.total_bytes, C = a.pipe_chunk_bytes;
  const int K = a.pipe_slots;
  const int64_t nchunks = (total + C - 1) / C;
  const uint32_t base = a.pipe_base;
  uint32_t* mine = pipe_area(cp, cp.rank);
  char* mybuf = reinterpret_cast<char*>(cp.buf[cp.rank]);
  constexpr int64_t kBlock = 8 * (int64_t)kRowBytes;  // 32 KiB work items inside a chunk

  if (role == 0) {
    // ---------------- pack ----------------
    const S prescale = (S)a.prescale;
    for (int64_t c = 0; c < nchunks; ++c) {
      if (c >= K && !pipe_wait(cp, mine + kPipeUnpackDone, 1, base + (uint32_t)(c - K) + 1u)) return;
      const int64_t clo = c * C, chi = clo + C < total ? clo + C : total;
      const int64_t delta = (c % K) * C - clo;  // fused offset -> slot offset
      for (int64_t o = clo + quad * kBlock; o < chi; o += (int64_t)nquad * kBlock)
        pack_range<T, W, UP>(descs, a.ndesc, total, mybuf + delta, o, o + kBlock < chi ? o + kBlock : chi, prescale);
      if (pipe_last_cta(mine + kPipePackCnt + (c % K), (uint32_t)nquad)) {
        if ((int)threadIdx.x < cp.nranks) st_release_sys(pipe_area(cp, threadIdx.x) + kPipePacked + cp.rank, base + (uint32_t)c + 1u);
      }
    }
  } else if (role == 3) {
    // ---------------- unpack ----------------
    for (int64_t c = 0; c < nchunks; ++c) {
      if (!pipe_wait(cp, mine + kPipeReduced, cp.nranks, base + (uint32_t)c + 1u)) return;
      const int64_t clo = c * C, chi = clo + C < total ? clo + C : total;
      const int64_t delta = (c % K) * C - clo;
      for (int64_t o = clo + quad * kBlock; o < chi; o += (int64_t)nquad * kBlock)
        unpack_range<T, W, UP>(descs, a.ndesc, total, mybuf + delta, o, o + kBlock < chi ? o + kBlock : chi);
      if (pipe_last_cta(mine + kPipeUnpackCnt + (c % K), (uint32_t)nquad)) {
        if (threadIdx.x == 0) st_release_sys(mine + kPipeUnpackDone, base + (uint32_t)c + 1u);
      }
    }
  } else {
    // ---------------- reduce + broadcast of my slice ----------------
    const S postscale = (S)a.postscale;
    const int ri = quad * 2 + (role - 1), nred = nquad * 2;
    const int64_t kRBlock = a.pipe_rblock_bytes;  // work item of one reduce CTA (multiple of the 4 KiB row)
    for (int64_t c = 0; c < nchunks; ++c) {
      if (!pipe_wait(cp, mine + kPipePacked, cp.nranks, base + (uint32_t)c + 1u)) return;
      const int64_t clo = c * C, chi = clo + C < total ? clo + C : total;
      const int64_t delta = (c % K) * C - clo;
      const int64_t units = (chi - clo) / 16;  // chunk length is a multiple of 128
      const int64_t slo = clo + 16 * (units * cp.rank / cp.nranks), shi = clo + 16 * (units * (cp.rank + 1) / cp.nranks);
      for (int64_t o = slo + ri * kRBlock; o < shi; o += (int64_t)nred * kRBlock) {
        const int64_t e = o + kRBlock < shi ? o + kRBlock : shi;
        if (a.pipe_use_nvls) {
          if constexpr (Nvls<W>::ok) nvls_inplace_rows<W, 4>(reinterpret_cast<char*>(cp.mc_buf) + delta, o, e, postscale);
        } else {
          pipe_p2p_reduce<W, NR>(cp, delta, o, e, a.op, postscale);
        }
      }
      if (pipe_last_cta(mine + kPipeRedCnt + (c % K), (uint32_t)nred)) {
        if ((int)threadIdx.x < cp.nranks) st_release_sys(pipe_area(cp, threadIdx.x) + kPipeReduced + cp.rank, base + (uint32_t)c + 1u);
      }
    }
  }
}

template <typename T, typename W>
cudaError_t launch_pipelined(const CommParams& cp, const AllreduceArgs& a, cudaStream_t s) {
  const int grid = (a.ctas / 4) * 4;
  if (grid < 4) return cudaErrorInvalidValue;
  if (cp.nranks <= 2) pipelined_allreduce_kernel<T, W, 2><<<grid, kThreads, 0, s>>>(cp, a);
  else if (cp.nranks <= 4) pipelined_allreduce_kernel<T, W, 4><<<grid, kThreads, 0, s>>>(cp, a);
  else pipelined_allreduce_kernel<T, W, 8><<<grid, kThreads, 0, s>>>(cp, a);
  CountKernelLaunch();
  return cudaGetLastError();
}

template <typename W>
cudaError_t launch_inplace(const CommParams& cp, const InplaceArgs& a, cudaStream_t s) {
  if (cp.nranks <= 2) inplace_allreduce_kernel<W, 2><<<a.ctas, kThreads, 0, s>>>(cp, a);
  else if (cp.nranks <= 4) inplace_allreduce_kernel<W, 4><<<a.ctas, kThreads, 0, s>>>(cp, a);
  else inplace_allreduce_kernel<W, 8><<<a.ctas, kThreads, 0, s>>>(cp, a);
  CountKernelLaunch();
  return cudaGetLastError();
}

// Stand-alone pack / unpack (NCCL baseline path): whole grid strides over rows.
template <typename T, typename W>
__global__ void __launch_bounds__(kThreads)
pack_unpack_kernel(char* buffer, const TensorDesc* descs, int nd, int64_t total, double scale, int direction) {
  using A = typename Traits<W>::Acc;
  using S = typename ScaleOf<A>::type;
  constexpr int NW = 16 / (int)sizeof(W);
  DescCursor cur;
  cur.init(descs, nd, total);
  for (int64_t o = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * 16; o < total; o += (int64_t)gridDim.x * kRowBytes) {
    cur.seek(o);
    const int64_t e = (o - cur.lo) / (int64_t)sizeof(W);
    A acc[NW];
    if (direction == 0) {
      load_elems<T, NW, A>(reinterpret_cast<const T*>(cur.in) + e, cur.count - e, acc);
#pragma unroll
      for (int i = 0; i < NW; ++i) acc[i] = apply_scale<A>(acc[i], (S)scale);
      st_stream(buffer + o, pack_vec<W, NW>(acc));
    } else {
      if (e >= cur.count) continue;
      unpack_vec<W, NW>(ld_stream(buffer + o), acc);
#pragma unroll
      for (int i = 0; i < NW; ++i) acc[i] = apply_scale<A>(acc[i], (S)scale);
      store_elems<T, NW, A>(reinterpret_cast<T*>(cur.out) + e, cur.count - e, acc);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) scale_kernel(const T* in, T* out, int64_t n, double scale) {
  using A = typename Traits<T>::Acc;
  using S = typename ScaleOf<A>::type;
  constexpr int N = 16 / (int)sizeof(T);
  for (int64_t e = ((int64_t)blockIdx.x * kThreads + threadIdx.x) * N; e < n; e += (int64_t)gridDim.x * kThreads * N) {
    A a[N];
    load_elems<T, N, A>(in + e, n - e, a);
#pragma unroll
    for (int i = 0; i < N; ++i) a[i] = apply_scale<A>(a[i], (S)scale);
    store_elems<T, N, A>(out + e, n - e, a);
  }
}

template <typename T, typename W>
cudaError_t launch_tw(const CommParams& cp, const AllreduceArgs& a, int chunk_bytes, cudaStream_t s) {
  dim3 grid(a.ctas), block(kThreads);
  if (cp.nranks <= 2) allreduce_kernel<T, W, 2><<<grid, block, 0, s>>>(cp, a, chunk_bytes);
  else if (cp.nranks <= 4) allreduce_kernel<T, W, 4><<<grid, block, 0, s>>>(cp, a, chunk_bytes);
  else allreduce_kernel<T, W, 8><<<grid, block, 0, s>>>(cp, a, chunk_bytes);
  CountKernelLaunch();
  return cudaGetLastError();
}

template <typename T, typename W>
cudaError_t launch_pu(char* buffer, const TensorDesc* descs, int nd, int64_t total, double scale, int dir, int ctas,
                      cudaStream_t s) {
  pack_unpack_kernel<T, W><<<ctas, kThreads, 0, s>>>(buffer, descs, nd, total, scale, dir);
  CountKernelLaunch();
  return cudaGetLastError();
}

}  // namespace

// dtype codes: hvd::DataType
#define HVD_DISPATCH(dtype, wire, FN, ...)                                                         \
  switch (dtype) {                                                                                 \
    case 0: return FN<uint8_t, uint8_t>(__VA_ARGS__);                                               \
    case 9: return FN<uint8_t, uint8_t>(__VA_ARGS__);                                               \
    case 1: return FN<int8_t, int8_t>(__VA_ARGS__);                                                 \
    case 2: case 3: return FN<int16_t, int16_t>(__VA_ARGS__);                                       \
    case 4: return FN<int32_t, int32_t>(__VA_ARGS__);                                               \
    case 5: return FN<int64_t, int64_t>(__VA_ARGS__);                                               \
    case 6: return FN<__half, __half>(__VA_ARGS__);                                                 \
    case 7:                                                                                        \
      if (wire == 10) return FN<float, __nv_bfloat16>(__VA_ARGS__);                                 \
      if (wire == 6) return FN<float, __half>(__VA_ARGS__);                                         \
      return FN<float, float>(__VA_ARGS__);                                                         \
    case 8: return FN<double, double>(__VA_ARGS__);                                                 \
    case 10: return FN<__nv_bfloat16, __nv_bfloat16>(__VA_ARGS__);                                  \
    default: return cudaErrorInvalidValue;                                                         \
  }

cudaError_t LaunchAllreduce(const CommParams& cp, const AllreduceArgs& args, cudaStream_t stream) {
  if (args.ctas < 1 || args.ctas > kMaxCtas || cp.nranks < 1 || cp.nranks > kMaxPeers) return cudaErrorInvalidValue;
  if (args.total_bytes <= 0) return cudaSuccess;
  if (args.variant == kNvls) {
    const int w = args.dtype == 7 ? (args.wire_dtype == 10 || args.wire_dtype == 6 ? args.wire_dtype : 7) : args.dtype;
    if (!(w == 7 || w == 6 || w == 10) || args.op != 1 || cp.mc_buf == nullptr) return cudaErrorInvalidValue;
  }
  if (args.oneshot_nvls) {
    const int w = args.dtype == 7 ? (args.wire_dtype == 10 || args.wire_dtype == 6 ? args.wire_dtype : 7) : args.dtype;
    if (args.variant != kOneShot || !(w == 7 || w == 6 || w == 10) || args.op != 1 || cp.mc_buf == nullptr) return cudaErrorInvalidValue;
  }
  if (args.variant == kPipelined) {
    const int w = args.dtype == 7 ? (args.wire_dtype == 10 || args.wire_dtype == 6 ? args.wire_dtype : 7) : args.dtype;
    if (args.pipe_chunk_bytes <= 0 || (args.pipe_chunk_bytes % 4096) || args.pipe_slots < 2 || args.pipe_slots > kPipeMaxSlots ||
        args.pipe_rblock_bytes < kRowBytes || (args.pipe_rblock_bytes % kRowBytes) ||
        args.out_descs != nullptr || args.reduce_lo != 0 || args.reduce_hi != args.total_bytes)
      return cudaErrorInvalidValue;
    if (args.pipe_use_nvls && (!(w == 7 || w == 6 || w == 10) || args.op != 1 || cp.mc_buf == nullptr)) return cudaErrorInvalidValue;
    HVD_DISPATCH(args.dtype, args.wire_dtype, launch_pipelined, cp, args, stream)
  }
  // chunk size: enough chunks that every (rank, CTA) pair owns work, bounded by [8 KiB, 32 KiB]
  int64_t want = args.total_bytes / ((int64_t)args.ctas * cp.nranks);
  int chunk = (int)((want / kRowBytes) * kRowBytes);
  if (chunk < kRowBytes) chunk = kRowBytes;
  if (chunk > kChunkBytes) chunk = kChunkBytes;
  HVD_DISPATCH(args.dtype, args.wire_dtype, launch_tw, cp, args, chunk, stream)
}

cudaError_t LaunchInplaceAllreduce(const CommParams& cp, const InplaceArgs& args, cudaStream_t stream) {
  if (args.ctas < 1 || args.ctas > kMaxCtas || args.bytes <= 0 || (args.bytes & 15)) return cudaErrorInvalidValue;
  InplaceArgs a = args;
  int64_t want = a.bytes / ((int64_t)a.ctas * cp.nranks);
  int chunk = (int)((want / kRowBytes) * kRowBytes);
  if (chunk < kRowBytes) chunk = kRowBytes;
  const int max_chunk = [] { const char* e = getenv("HVD_INPLACE_CHUNK_BYTES"); int v = e ? atoi(e) : kChunkBytes;
                                    return v < kRowBytes ? kRowBytes : (v / kRowBytes) * kRowBytes; }();
  const int unroll = [] { const char* e = getenv("HVD_NVLS_UNROLL"); return e && atoi(e) == 8 ? 8 : 4; }();
  if (chunk > max_chunk) chunk = max_chunk;
  a.chunk_bytes = chunk;
  a.nvls_unroll = unroll;
  switch (a.dtype) {
    case 7: return launch_inplace<float>(cp, a, stream);
    case 6: return launch_inplace<__half>(cp, a, stream);
    case 10: return launch_inplace<__nv_bfloat16>(cp, a, stream);
    case 8: if (a.use_multicast) return cudaErrorInvalidValue; return launch_inplace<double>(cp, a, stream);
    case 4: if (a.use_multicast) return cudaErrorInvalidValue; return launch_inplace<int32_t>(cp, a, stream);
    case 5: if (a.use_multicast) return cudaErrorInvalidValue; return launch_inplace<int64_t>(cp, a, stream);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t LaunchPackUnpack(void* buffer, const TensorDesc* descs, int ndesc, int64_t total_bytes, int dtype,
                             int wire_dtype, double scale, int direction, int ctas, cudaStream_t stream) {
  if (total_bytes <= 0) return cudaSuccess;
  HVD_DISPATCH(dtype, wire_dtype, launch_pu, (char*)buffer, descs, ndesc, total_bytes, scale, direction, ctas, stream)
}

namespace {
template <typename T, typename W> cudaError_t launch_scale(const void* in, void* out, int64_t n, double scale, cudaStream_t s) {
  int64_t per = (int64_t)kThreads * (16 / sizeof(T));
  int64_t blocks = (n + per - 1) / per;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  scale_kernel<T><<<(int)blocks, kThreads, 0, s>>>((const T*)in, (T*)out, n, scale);
  CountKernelLaunch();
  return cudaGetLastError();
}
}  // namespace

cudaError_t LaunchScale(const void* in, void* out, int64_t count, int dtype, double scale, cudaStream_t stream) {
  if (count <= 0) return cudaSuccess;
  const int wire = dtype;
  HVD_DISPATCH(dtype, wire, launch_scale, in, out, count, scale, stream)
}

}  // namespace kern
}  // namespace hvd
