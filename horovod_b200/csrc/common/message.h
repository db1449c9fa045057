// Negotiation messages (Request / Response and their lists) and a compact
// little-endian binary wire codec.
//
// Parity: horovod/common/message.{h,cc} + wire/message.fbs.  The reference
// serialises with FlatBuffers (unavailable offline); this codec is a
// hand-rolled length-prefixed format (ByteWriter/ByteReader).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "common.h"

namespace hvd {

class ByteWriter {
 public:
  void u8(uint8_t v) { buf_.push_back(v); }
  void i32(int32_t v) { raw(&v, 4); }
  void i64(int64_t v) { raw(&v, 8); }
  void f64(double v) { raw(&v, 8); }
  void str(const std::string& s) { i32((int32_t)s.size()); raw(s.data(), s.size()); }
  void vec_i64(const std::vector<int64_t>& v) { i32((int32_t)v.size()); raw(v.data(), v.size() * 8); }
  void vec_i32(const std::vector<int32_t>& v) { i32((int32_t)v.size()); raw(v.data(), v.size() * 4); }
  void raw(const void* p, size_t n) { auto* c = (const uint8_t*)p; buf_.insert(buf_.end(), c, c + n); }
  std::vector<uint8_t>& data() { return buf_; }
 private:
  std::vector<uint8_t> buf_;
};

class ByteReader {
 public:
  ByteReader(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
  uint8_t u8() { need(1); return *p_++; }
  int32_t i32() { int32_t v; rd(&v, 4); return v; }
  int64_t i64() { int64_t v; rd(&v, 8); return v; }
  double f64() { double v; rd(&v, 8); return v; }
  std::string str() { int32_t n = i32(); need(n); std::string s((const char*)p_, n); p_ += n; return s; }
  std::vector<int64_t> vec_i64() { int32_t n = i32(); std::vector<int64_t> v(n); rd(v.data(), (size_t)n * 8); return v; }
  std::vector<int32_t> vec_i32() { int32_t n = i32(); std::vector<int32_t> v(n); rd(v.data(), (size_t)n * 4); return v; }
  bool done() const { return p_ >= end_; }
 private:
  void need(size_t n);
  void rd(void* o, size_t n) { need(n); if (n) memcpy(o, p_, n); p_ += n; }
  const uint8_t* p_; const uint8_t* end_;
};

struct Request {
  int32_t request_rank = 0;
  RequestType type = RequestType::ALLREDUCE;
  DataType dtype = DataType::FLOAT32;
  std::string name;
  int32_t root_rank = 0;
  int32_t device = CPU_DEVICE_ID;
  std::vector<int64_t> shape;  // for PROCESS_SET_ADD: the member ranks
  double prescale = 1.0, postscale = 1.0;
  ReduceOp reduce_op = ReduceOp::SUM;
  int32_t group_id = -1;
  int32_t group_size = 0;
  // >= 0: (region << 44 | offset), the tensor lives in registered symmetric memory; <= -2: per-rank key of a plain device
  // allocation that can be IPC-registered on the fly (ops/ipc_registry.h); -1: neither
  int64_t symm_key = -1;
  void Serialize(ByteWriter& w) const;
  static Request Parse(ByteReader& r);
};

struct RequestList {
  std::vector<Request> requests;
  bool shutdown = false;
  std::vector<uint8_t> Serialize() const;
  static RequestList Parse(const uint8_t* p, size_t n);
};

enum class ResponseType : uint8_t {
  ALLREDUCE = 0, ALLGATHER = 1, BROADCAST = 2, JOIN = 3, ADASUM = 4, ALLTOALL = 5,
  BARRIER = 6, REDUCESCATTER = 7, PROCESS_SET_ADD = 8, PROCESS_SET_REMOVE = 9, ERROR = 10, SYMM_ALLOC = 11,
};
const char* ResponseTypeName(ResponseType t);

struct Response {
  ResponseType type = ResponseType::ALLREDUCE;
  std::vector<std::string> tensor_names;
  std::string error_message;
  std::vector<int32_t> devices;       // device of the tensor on every rank of the set
  // ALLREDUCE/ADASUM/REDUCESCATTER/BROADCAST: element count per tensor.
  // ALLGATHER: for each tensor, first-dim size on every rank (names.size()*set_size).
  // PROCESS_SET_ADD: member ranks; PROCESS_SET_REMOVE: [id].
  std::vector<int64_t> tensor_sizes;
  DataType dtype = DataType::FLOAT32;
  double prescale = 1.0, postscale = 1.0;
  ReduceOp reduce_op = ReduceOp::SUM;
  int32_t last_joined_rank = -1;
  int32_t root_rank = 0;
  int32_t group_id = -1;              // not serialised beyond fusion decisions
  int64_t symm_key = -1;              // >= 0: every rank holds this tensor at the same place of the same registered region;
                                      // -2: every rank holds it in an IPC-registrable plain allocation (in place)
  bool from_cache = false;            // local, not serialised: replayed from the response cache (no rank's tensor moved)
  int64_t payload_bytes = 0;          // fusion accounting: bytes this response moves through the fusion / symmetric buffer
  void Serialize(ByteWriter& w) const;
  static Response Parse(ByteReader& r);
};

struct ResponseList {
  std::vector<Response> responses;
  bool shutdown = false;
  std::vector<uint8_t> Serialize() const;
  static ResponseList Parse(const uint8_t* p, size_t n);
};

}  // namespace hvd
