#include "ipc_registry.h"
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstring>
#include <dlfcn.h>
#include "../common/logging.h"

namespace hvd {

namespace {
#define HVD_STR2(x) #x
#define HVD_STR(x) HVD_STR2(x)
// libcuda is not linked (the CPU-only test box must be able to load this library): resolve the two driver calls lazily
struct IpcDriver {
  decltype(&cuPointerGetAttribute) p_cuPointerGetAttribute = nullptr;
  decltype(&cuMemGetAddressRange) p_cuMemGetAddressRange = nullptr;
  IpcDriver() {
    void* lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libcuda.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return;
    p_cuPointerGetAttribute = (decltype(&cuPointerGetAttribute))dlsym(lib, HVD_STR(cuPointerGetAttribute));
    p_cuMemGetAddressRange = (decltype(&cuMemGetAddressRange))dlsym(lib, HVD_STR(cuMemGetAddressRange));
  }
  bool ok() const { return p_cuPointerGetAttribute && p_cuMemGetAddressRange; }
};
IpcDriver& Drv() { static IpcDriver d; return d; }
}  // namespace

int64_t IpcKeyFor(const void* ptr) {
  unsigned long long id = 0;
  if (!Drv().ok() || Drv().p_cuPointerGetAttribute(&id, CU_POINTER_ATTRIBUTE_BUFFER_ID, (CUdeviceptr)(uintptr_t)ptr) != CUDA_SUCCESS) return -1;
  // buffer ids are unique per allocation for the life of the process: (id, address) changes whenever the tensor moves
  uint64_t h = (uint64_t)id * 0x9E3779B97F4A7C15ull ^ ((uint64_t)(uintptr_t)ptr >> 4) * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 29;
  return -(int64_t)(2 + (h & ((1ull << 62) - 1)));
}

void IpcRegistry::Release(Stored& s) {
  for (auto& k : s.holds) {
    auto it = opened_.find(k);
    if (it == opened_.end()) continue;
    if (--it->second.refs <= 0) {
      cudaIpcCloseMemHandle(it->second.base);
      opened_.erase(it);
    }
  }
  s.holds.clear();
  cudaGetLastError();
}

void IpcRegistry::Clear() {
  for (auto& kv : entries_) Release(kv.second);
  entries_.clear();
  for (auto& kv : opened_) cudaIpcCloseMemHandle(kv.second.base);
  opened_.clear();
  cudaGetLastError();
}

const IpcRegistry::Entry* IpcRegistry::Find(const std::string& name) const {
  auto it = entries_.find(name);
  return it == entries_.end() ? nullptr : &it->second.e;
}

const IpcRegistry::Entry& IpcRegistry::Exchange(Transport* t, const std::string& name, const void* ptr, size_t bytes) {
  const int n = t->size(), me = t->rank();
  // wire record per rank: [ok, offset, handle (8 x int64)]
  constexpr int kRec = 10;
  int64_t mine[kRec] = {};
  CUdeviceptr base = 0;
  size_t range = 0;
  cudaIpcMemHandle_t h;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  bool ok = Drv().ok() && Drv().p_cuMemGetAddressRange(&base, &range, (CUdeviceptr)(uintptr_t)ptr) == CUDA_SUCCESS &&
            (uintptr_t)ptr + bytes <= (uintptr_t)base + range &&
            cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base) == cudaSuccess;
  if (!ok) cudaGetLastError();
  mine[0] = ok ? 1 : 0;
  mine[1] = ok ? (int64_t)((uintptr_t)ptr - (uintptr_t)base) : 0;
  if (ok) memcpy(&mine[2], &h, 64);
  std::vector<int64_t> all((size_t)n * kRec);
  t->AllgatherInts(mine, kRec, all.data());

  Stored fresh;
  fresh.e.my_ptr = ptr;
  bool usable = true;
  for (int p = 0; p < n; ++p) if (!all[(size_t)p * kRec]) usable = false;
  if (usable) {
    for (int p = 0; p < n && usable; ++p) {
      if (p == me) { fresh.e.ptr[p] = const_cast<void*>(ptr); continue; }
      Handle hk;
      memcpy(hk.data(), &all[(size_t)p * kRec + 2], 64);
      auto key = std::make_pair(p, hk);
      auto it = opened_.find(key);
      if (it == opened_.end()) {
        cudaIpcMemHandle_t ph;
        memcpy(&ph, hk.data(), 64);
        void* mapped = nullptr;
        if (cudaIpcOpenMemHandle(&mapped, ph, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          LOG(DEBUG) << "cudaIpcOpenMemHandle of rank " << p << "'s allocation failed: " << cudaGetErrorString(cudaGetLastError());
          usable = false;
          break;
        }
        it = opened_.emplace(key, Opened{mapped, 0}).first;
      }
      ++it->second.refs;
      fresh.holds.push_back(key);
      fresh.e.ptr[p] = (char*)it->second.base + all[(size_t)p * kRec + 1];
    }
  }
  // every rank must take the same path: one more agreement round on the outcome of the imports
  uint64_t okw = usable ? 1 : 0;
  t->AllreduceBits(&okw, 1, nullptr, 0);
  fresh.e.usable = okw != 0;
  auto old = entries_.find(name);
  if (old != entries_.end()) Release(old->second);   // after the new opens: an unchanged allocation keeps its mapping
  if (!fresh.e.usable) Release(fresh);
  Stored& slot = entries_[name];
  slot = std::move(fresh);
  return slot.e;
}

}  // namespace hvd
