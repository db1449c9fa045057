// CPU data plane: collectives on host tensors over the Transport
// (TCP full mesh between hosts / loopback in unit tests).
//
// Parity: the Gloo ops (horovod/common/ops/gloo_operations.cc:33-494 — ring /
// halving-doubling allreduce, allgatherv, broadcast, alltoallv, reducescatter)
// and MPI ops (ops/mpi_operations.cc).  Self-contained: no gloo, no MPI.
#pragma once
#include <cstdint>
#include <vector>
#include "../common/common.h"
#include "../transport/transport.h"

namespace hvd {
namespace cpu {

// Elementwise helpers -------------------------------------------------------
// dst[i] = op(dst[i], src[i])
void ReduceInto(void* dst, const void* src, int64_t count, DataType dtype, ReduceOp op);
// buf[i] *= scale (float types and, like the reference's ScaleBufferCPUImpl, integer types too)
void ScaleBuffer(void* buf, int64_t count, DataType dtype, double scale);

// Collectives (blocking; run on the background thread) ----------------------
// In-place allreduce of `count` elements.
void Allreduce(Transport* t, void* buf, int64_t count, DataType dtype, ReduceOp op);
// Variable-size allgather of byte blocks: block r has bytes[r] bytes and lands at out + displ[r].
// `in` may alias out + displ[rank].
void Allgatherv(Transport* t, const void* in, void* out, const std::vector<int64_t>& bytes);
void Broadcast(Transport* t, void* buf, int64_t bytes, int root);
// send_bytes[p] bytes go to rank p (packed consecutively in `in`); recv_bytes[p] arrive from p (packed in `out`).
void Alltoallv(Transport* t, const void* in, const std::vector<int64_t>& send_bytes, void* out,
               const std::vector<int64_t>& recv_bytes);
// In: `buf` holds all segments (counts[r] elements for rank r, consecutive; modified in place).
// Out: `out` receives this rank's reduced segment.
void Reducescatter(Transport* t, void* buf, const std::vector<int64_t>& counts, void* out, DataType dtype, ReduceOp op);

// Adasum (vector-halving distance-doubling, reference ops/adasum/adasum.h:195-435).
// `buf` holds the fused tensors back to back; tensor_counts[i] elements each.
// Requires power-of-two size. dtype: FLOAT16/BFLOAT16/FLOAT32/FLOAT64.
Status AdasumAllreduce(Transport* t, void* buf, const std::vector<int64_t>& tensor_counts, DataType dtype);

// How many host collectives took which data path since start-up: 0 = shared-memory slots, 1 = two-level (shm + cross-host
// rings), 2 = ring / tree / star over the base transport.
unsigned long long HostPathCount(int which);

}  // namespace cpu
}  // namespace hvd
