#include "message.h"
#include <stdexcept>

namespace hvd {

void ByteReader::need(size_t n) {
  if ((size_t)(end_ - p_) < n) throw std::runtime_error("hvd wire: truncated message");
}

const char* ResponseTypeName(ResponseType t) {
  if (t == ResponseType::ERROR) return "ERROR";
  return RequestTypeName((RequestType)t);
}

void Request::Serialize(ByteWriter& w) const {
  w.i32(request_rank); w.u8((uint8_t)type); w.u8((uint8_t)dtype); w.str(name);
  w.i32(root_rank); w.i32(device); w.vec_i64(shape); w.f64(prescale); w.f64(postscale);
  w.u8((uint8_t)reduce_op); w.i32(group_id); w.i32(group_size); w.i64(symm_key);
}
Request Request::Parse(ByteReader& r) {
  Request q;
  q.request_rank = r.i32(); q.type = (RequestType)r.u8(); q.dtype = (DataType)r.u8(); q.name = r.str();
  q.root_rank = r.i32(); q.device = r.i32(); q.shape = r.vec_i64(); q.prescale = r.f64(); q.postscale = r.f64();
  q.reduce_op = (ReduceOp)r.u8(); q.group_id = r.i32(); q.group_size = r.i32(); q.symm_key = r.i64();
  return q;
}
std::vector<uint8_t> RequestList::Serialize() const {
  ByteWriter w; w.u8(shutdown ? 1 : 0); w.i32((int32_t)requests.size());
  for (auto& q : requests) q.Serialize(w);
  return std::move(w.data());
}
RequestList RequestList::Parse(const uint8_t* p, size_t n) {
  ByteReader r(p, n); RequestList l; l.shutdown = r.u8() != 0; int32_t c = r.i32();
  l.requests.reserve(c);
  for (int i = 0; i < c; ++i) l.requests.push_back(Request::Parse(r));
  return l;
}

void Response::Serialize(ByteWriter& w) const {
  w.u8((uint8_t)type); w.i32((int32_t)tensor_names.size());
  for (auto& s : tensor_names) w.str(s);
  w.str(error_message); w.vec_i32(devices); w.vec_i64(tensor_sizes); w.u8((uint8_t)dtype);
  w.f64(prescale); w.f64(postscale); w.u8((uint8_t)reduce_op); w.i32(last_joined_rank); w.i32(root_rank);
  w.i32(group_id); w.i64(symm_key); w.i64(payload_bytes);
}
Response Response::Parse(ByteReader& r) {
  Response s; s.type = (ResponseType)r.u8(); int32_t n = r.i32();
  s.tensor_names.reserve(n);
  for (int i = 0; i < n; ++i) s.tensor_names.push_back(r.str());
  s.error_message = r.str(); s.devices = r.vec_i32(); s.tensor_sizes = r.vec_i64(); s.dtype = (DataType)r.u8();
  s.prescale = r.f64(); s.postscale = r.f64(); s.reduce_op = (ReduceOp)r.u8(); s.last_joined_rank = r.i32();
  s.root_rank = r.i32(); s.group_id = r.i32(); s.symm_key = r.i64(); s.payload_bytes = r.i64();
  return s;
}
std::vector<uint8_t> ResponseList::Serialize() const {
  ByteWriter w; w.u8(shutdown ? 1 : 0); w.i32((int32_t)responses.size());
  for (auto& s : responses) s.Serialize(w);
  return std::move(w.data());
}
ResponseList ResponseList::Parse(const uint8_t* p, size_t n) {
  ByteReader r(p, n); ResponseList l; l.shutdown = r.u8() != 0; int32_t c = r.i32();
  l.responses.reserve(c);
  for (int i = 0; i < c; ++i) l.responses.push_back(Response::Parse(r));
  return l;
}

}  // namespace hvd
