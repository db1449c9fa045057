#include "response_cache.h"
#include "logging.h"

namespace hvd {

void ResponseCache::set_capacity(uint32_t c) {
  clear();
  capacity_ = c;
}

void ResponseCache::clear() {
  slots_.clear(); free_.clear(); lru_.clear(); by_name_.clear();
}

ResponseCache::State ResponseCache::Cached(const Request& r) const {
  auto it = by_name_.find(r.name);
  if (it == by_name_.end()) return State::MISS;
  const Slot& s = slots_[it->second];
  bool same = s.params_valid && s.type == r.type && s.dtype == r.dtype && s.shape == r.shape &&
              s.device == r.device && s.root_rank == r.root_rank && s.prescale == r.prescale &&
              s.postscale == r.postscale && s.op == r.reduce_op && s.symm_key == r.symm_key;
  return same ? State::HIT : State::INVALID;
}

uint32_t ResponseCache::PeekBit(const std::string& name) const {
  auto it = by_name_.find(name);
  return it == by_name_.end() ? UINT32_MAX : it->second;
}

uint32_t ResponseCache::Put(const Response& single, const Request* local) {
  if (capacity_ == 0) return UINT32_MAX;
  const std::string& name = single.tensor_names[0];
  uint32_t evicted = UINT32_MAX;
  uint32_t bit;
  auto it = by_name_.find(name);
  if (it != by_name_.end()) {
    bit = it->second;
    lru_.erase(slots_[bit].lru_it);
  } else {
    if (by_name_.size() >= capacity_) {
      evicted = lru_.front();
      LOG(DEBUG) << "response cache full: evicting " << slots_[evicted].response.tensor_names[0];
      Erase(evicted);
    }
    if (!free_.empty()) { bit = free_.back(); free_.pop_back(); }
    else { bit = (uint32_t)slots_.size(); slots_.emplace_back(); }
    by_name_[name] = bit;
  }
  Slot& s = slots_[bit];
  s.used = true;
  s.response = single;
  s.params_valid = local != nullptr;
  if (local) {
    s.type = local->type; s.dtype = local->dtype; s.shape = local->shape; s.device = local->device;
    s.root_rank = local->root_rank; s.prescale = local->prescale; s.postscale = local->postscale;
    s.op = local->reduce_op; s.symm_key = local->symm_key;
  }
  lru_.push_back(bit);
  s.lru_it = std::prev(lru_.end());
  return evicted;
}

const Response& ResponseCache::GetResponse(uint32_t bit) {
  Slot& s = slots_[bit];
  lru_.erase(s.lru_it);
  lru_.push_back(bit);
  s.lru_it = std::prev(lru_.end());
  return s.response;
}

void ResponseCache::Erase(uint32_t bit) {
  if (!HasBit(bit)) return;
  Slot& s = slots_[bit];
  by_name_.erase(s.response.tensor_names[0]);
  lru_.erase(s.lru_it);
  s = Slot();
  free_.push_back(bit);
}

}  // namespace hvd
