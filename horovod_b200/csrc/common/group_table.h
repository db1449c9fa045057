// group_id <-> member tensor names for grouped collectives (ids are handed out
// in call order, identical on every rank because grouped calls are collective).
// Parity: horovod/common/group_table.{h,cc}.
#pragma once
#include <mutex>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>

namespace hvd {
class GroupTable {
 public:
  int32_t RegisterGroup(std::vector<std::string> names);
  void DeregisterGroup(int32_t id);
  std::vector<std::string> GetGroupTensorNames(int32_t id) const;
  int32_t GetGroupIDFromTensorName(const std::string& name) const;
  bool empty() const;

 private:
  mutable std::mutex mu_;
  std::unordered_map<int32_t, std::vector<std::string>> id_to_names_;
  std::unordered_map<std::string, int32_t> name_to_id_;
  std::queue<int32_t> free_ids_;
  int32_t next_id_ = 0;
};
}  // namespace hvd
