// Host-side launch helpers shared by the .cu files.
#pragma once
#include <cuda_runtime.h>

namespace mnc {

// cudaFuncSetAttribute(cudaFuncAttributeMaxDynamicSharedMemorySize) applies to the CURRENT device
// only, and entry points such as mnc_nms_host(device_id) may be called for several devices from
// one process: remember, per device, the largest opt-in already granted for a kernel.
constexpr int kMaxDevices = 64;
struct SmemGrant {
  int granted[kMaxDevices] = {};
};

template <typename Kernel>
inline bool ensure_dynamic_smem(Kernel kernel, int bytes, SmemGrant& g) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) dev = 0;
  if (bytes <= g.granted[dev]) return true;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) != cudaSuccess)
    return false;
  g.granted[dev] = bytes;
  return true;
}

}  // namespace mnc
