#include "group_table.h"
namespace hvd {
int32_t GroupTable::RegisterGroup(std::vector<std::string> names) {
  std::lock_guard<std::mutex> l(mu_);
  int32_t id;
  if (!free_ids_.empty()) { id = free_ids_.front(); free_ids_.pop(); } else { id = next_id_++; }
  for (auto& n : names) name_to_id_[n] = id;
  id_to_names_[id] = std::move(names);
  return id;
}
void GroupTable::DeregisterGroup(int32_t id) {
  std::lock_guard<std::mutex> l(mu_);
  auto it = id_to_names_.find(id);
  if (it == id_to_names_.end()) return;
  for (auto& n : it->second) name_to_id_.erase(n);
  id_to_names_.erase(it);
  free_ids_.push(id);
}
std::vector<std::string> GroupTable::GetGroupTensorNames(int32_t id) const {
  std::lock_guard<std::mutex> l(mu_);
  auto it = id_to_names_.find(id);
  return it == id_to_names_.end() ? std::vector<std::string>{} : it->second;
}
int32_t GroupTable::GetGroupIDFromTensorName(const std::string& name) const {
  std::lock_guard<std::mutex> l(mu_);
  auto it = name_to_id_.find(name);
  return it == name_to_id_.end() ? -1 : it->second;
}
bool GroupTable::empty() const { std::lock_guard<std::mutex> l(mu_); return id_to_names_.empty(); }
}  // namespace hvd
