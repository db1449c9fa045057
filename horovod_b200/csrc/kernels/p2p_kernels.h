// Host-side interface of the sm_100a peer-to-peer collective kernels.
//
// These kernels replace the reference's three-launch GPU data path
// (batched_scaled_memcpy_k -> ncclAllReduce -> batched_scaled_memcpy_k,
// horovod/common/ops/cuda/cuda_kernels.cu:259-324 + ops/nccl_operations.cc:185-287)
// with ONE launch that packs (prescale + cast) gradients into a symmetric
// buffer, synchronises with peer GPUs through release/acquire flags in peer
// memory, reduces by loading/storing peer buffers directly over NVLink (or
// through the NVSwitch with multimem.ld_reduce / multimem.st), applies
// postscale/average + cast and scatters the result to the output tensors.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hvd {
namespace kern {

// number of kernels of this library launched by this process (bench.py reports it as gpu_launches)
unsigned long long KernelLaunchCount();
void CountKernelLaunch();

constexpr int kMaxPeers = 8;          // one NVSwitch domain of the HGX B200 box
constexpr int kMaxCtas = 256;         // flag slots per team
// 256-thread CTAs capped at 128 registers: two fit on an SM and one fits NEXT TO resident compute CTAs of an
// overlapping backward pass, so a communication kernel does not have to wait for whole SMs to drain
// (512-thread x 128-register CTAs needed an empty SM each; measured 2x step-time inflation at 2 GPUs).
constexpr int kThreads = 256;
constexpr int kChunkBytes = 32768;    // kThreads * 16 B * 8
constexpr int kInlineDescs = 6;
constexpr int kFlagWords = kMaxCtas * kMaxPeers;
// Independent barrier channels (flag words + epochs) of one team: see SymmTeam::Params
constexpr int kNumChannels = 8;
constexpr int kGraphChannel = 1;              // collectives captured into CUDA graphs (framework streams)
constexpr int kLatencyChannel = 3;            // the latency lane of small allreduce responses
constexpr int kAuxChannel = 2;                // lane 1 of the dual-lane large-message allreduce (auxiliary hvd stream)
constexpr int kChannelFlagsOffset = 128 * 1024;  // bytes from the start of the flag region; channel c >= 1 at + (c-1) * kFlagWords * 4

// Everything a kernel needs to talk to its peers. Passed by value
// (__grid_constant__).
struct CommParams {
  int nranks;
  int rank;
  void* buf[kMaxPeers];         // op data buffer of every rank, mapped into this process
  uint32_t* flags[kMaxPeers];   // flag words of every rank: [cta][src_rank]
  void* mc_buf;                 // multicast (NVLS) mapping of the same buffer, or nullptr
  uint32_t* epochs;             // local: per-CTA barrier epoch, persists across launches
  int* abort_flag;              // host-mapped; non-zero => stop spinning (peer failure / shutdown / timeout)
  unsigned long long timeout_ns; // a barrier that waits longer sets *abort_flag = 2 and bails out (0 = wait forever)
};

// One tensor of a fused response. `offset` is the tensor's byte offset inside
// the fused buffer in WIRE dtype (128 B aligned); `count` in elements.
struct TensorDesc {
  const void* in;
  void* out;
  int64_t offset;
  int64_t count;
};

enum Variant : int { kOneShot = 0, kTwoShot = 1, kNvls = 2, kPipelined = 3 };

// Software-pipelined variant (kPipelined): per-rank words inside the flag region, past the per-CTA barrier flags and the
// Adasum scratch.  `packed` / `reduced` are written by the peers, the rest is local bookkeeping.
constexpr int kPipeAreaOffset = 96 * 1024;   // bytes from the start of the flag region
constexpr int kPipeMaxSlots = 32;
constexpr int kPipePacked = 0;                       // [kMaxPeers] chunk counter published by every rank after its pack
constexpr int kPipeReduced = 16;                     // [kMaxPeers] ... after its reduce + broadcast
constexpr int kPipePackCnt = 64;                     // [kPipeMaxSlots] CTAs that finished packing the chunk in this slot
constexpr int kPipeRedCnt = 64 + kPipeMaxSlots;      // [kPipeMaxSlots]
constexpr int kPipeUnpackCnt = 64 + 2 * kPipeMaxSlots;
constexpr int kPipeUnpackDone = 64 + 3 * kPipeMaxSlots;  // 1 word: chunks completely unpacked on this rank
constexpr int kPipeAreaWords = 64 + 3 * kPipeMaxSlots + 16;
// reduce op codes follow hvd::ReduceOp: 1 SUM (AVERAGE arrives as SUM + postscale), 3 MIN, 4 MAX, 5 PRODUCT
// dtype codes follow hvd::DataType.

struct AllreduceArgs {
  const TensorDesc* descs;      // device table (nullptr => use inline_descs)
  TensorDesc inline_descs[kInlineDescs];
  int ndesc;
  int64_t total_bytes;          // wire bytes of the fused buffer region, multiple of 128
  int64_t reduce_lo, reduce_hi; // byte range this rank must produce output for (allreduce: [0,total))
  double prescale, postscale;
  int op;
  int dtype, wire_dtype;
  int variant;
  int ctas;
  // reducescatter: outputs use their own table (offsets relative to the fused buffer)
  const TensorDesc* out_descs;  // nullptr => same as descs
  int nout;
  int oneshot_nvls;             // kOneShot only: reduce with multimem.ld_reduce on the multicast mapping instead of N peer loads
  // kPipelined only: the symmetric buffer is a ring of `pipe_slots` chunks of `pipe_chunk_bytes`; `pipe_base` is the
  // team-wide running chunk counter at the start of this launch (identical on all ranks), `pipe_use_nvls` selects the
  // in-switch reduction
  int64_t pipe_chunk_bytes;
  int64_t pipe_rblock_bytes;    // bytes one reduce CTA handles per trip (multiple of 4096)
  int pipe_slots;
  uint32_t pipe_base;
  int pipe_use_nvls;
};

// Fused allreduce / reducescatter.  Returns cudaErrorInvalidValue for
// unsupported dtype combinations.
cudaError_t LaunchAllreduce(const CommParams& cp, const AllreduceArgs& args, cudaStream_t stream);

// Zero-copy allreduce on a registered (peer-mapped) tensor: ptr[r] = the tensor's address in rank r's region as
// mapped into THIS process, mc = its multicast alias (or nullptr). In place on every rank. SUM-like ops only for
// multicast; bytes must be a multiple of 16.
struct InplaceArgs {
  void* ptr[kMaxPeers];
  void* mc;
  int64_t bytes;
  double scale;       // prescale * postscale (valid for SUM / AVERAGE)
  int op;
  int dtype;
  int use_multicast;
  int ctas;
  int chunk_bytes;    // filled in by the launcher
  int nvls_unroll;    // filled in by the launcher (4 | 8 multimem.ld_reduce in flight per thread)
};
cudaError_t LaunchInplaceAllreduce(const CommParams& cp, const InplaceArgs& args, cudaStream_t stream);

// Generic "pack -> barrier -> pull" used by allgather / broadcast / alltoall.
// send: local src -> local symmetric buffer offset; recv: peer buffer offset -> local dst.
struct CopyDesc {
  const void* src;   // send: local source; recv: unused
  void* dst;         // recv: local destination; send: unused
  int64_t offset;    // byte offset in the (sender's) symmetric buffer
  int64_t bytes;
  int peer;          // recv: which rank's buffer to pull from
  int pad;           // send: kSendMulticast = store through the multicast mapping (lands in EVERY rank's buffer at `offset`)
};
constexpr int kSendMulticast = 1;  // requires a 16 B aligned offset and a 16 B-padded window
struct ExchangeArgs {
  const CopyDesc* sends; int nsend;
  const CopyDesc* recvs; int nrecv;
  int ctas;
};
cudaError_t LaunchExchange(const CommParams& cp, const ExchangeArgs& args, cudaStream_t stream);

// Stand-alone batched pack / unpack / scale (NCCL-baseline path and odd cases).
// direction 0: tensors -> buffer (prescale), 1: buffer -> tensors (postscale)
cudaError_t LaunchPackUnpack(void* buffer, const TensorDesc* descs, int ndesc, int64_t total_bytes, int dtype,
                             int wire_dtype, double scale, int direction, int ctas, cudaStream_t stream);
cudaError_t LaunchScale(const void* in, void* out, int64_t count, int dtype, double scale, cudaStream_t stream);

// Adasum (vector-halving distance-doubling with per-tensor adaptive coefficients) entirely on the GPUs of a
// peer-mapped team; see adasum_kernels.cu.  Tensors are fp32 / fp16 / bf16; the fused vector lives in fp32.
struct AdasumArgs {
  const TensorDesc* descs; int ndesc;   // device table; offsets are byte offsets of the FP32 fused vector (128 B aligned)
  int64_t total_bytes;                  // fp32 bytes, multiple of 128
  int dtype;
  int ctas;
  int64_t scratch_stride_bytes;         // distance between the two per-level partial-dot tables in the flag region
  // Persistent single-launch variant (both non-null): local device scratch of AdasumPersistentScratchBytes(ctas, ndesc)
  // and a 16-byte grid-synchronisation block; `ctas` CTAs must be co-resident on the device.
  void* persist_scratch = nullptr;
  void* persist_sync = nullptr;
};
size_t AdasumPersistentScratchBytes(int ctas, int ndesc);
size_t AdasumPersistentSyncBytes();
constexpr int64_t kAdasumScratchStride = 30000;
constexpr int kAdasumMaxTensors = 1250;
cudaError_t LaunchAdasum(const CommParams& cp, const AdasumArgs& args, double prescale, double postscale, cudaStream_t stream);
// Single-GPU simulation support: launch only step `only` of the sequence (see adasum_kernels.cu); -1 = all.
cudaError_t LaunchAdasumStep(const CommParams& cp, const AdasumArgs& args, double prescale, double postscale, cudaStream_t stream,
                             int only);
int AdasumNumLaunches(int nranks);

// Fused multi-tensor optimizer updates (see optim_kernels.cu)
struct SgdTensor { void* param; const void* grad; void* momentum; int64_t count; };
cudaError_t LaunchFusedSgd(const SgdTensor* table, int n, int64_t max_count, float lr, float momentum, float dampening,
                           float weight_decay, int nesterov, float grad_scale, int first_step, int param_dtype,
                           int grad_dtype, cudaStream_t stream);
struct AdamTensor { void* param; const void* grad; void* exp_avg; void* exp_avg_sq; int64_t count; };
cudaError_t LaunchFusedAdamW(const AdamTensor* table, int n, int64_t max_count, float lr, float beta1, float beta2,
                             float eps, float weight_decay, float bias_c1, float bias_c2, float grad_scale,
                             int adamw, int param_dtype, int grad_dtype, cudaStream_t stream);

}  // namespace kern
}  // namespace hvd
