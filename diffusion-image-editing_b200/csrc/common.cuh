// Shared host/device helpers for libb200edit.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/b200edit.h"

namespace b2e {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return B2E_CUDA_ERROR;
  }
  return B2E_OK;
}

#define B2E_REQUIRE(cond, status, ...) \
  do {                                 \
    if (!(cond)) {                     \
      b2e::set_error(__VA_ARGS__);     \
      return (status);                 \
    }                                  \
  } while (0)

#define B2E_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e_ = (call);                                                \
    if (e_ != cudaSuccess) {                                                \
      b2e::set_error("%s: %s", #call, cudaGetErrorString(e_));              \
      return B2E_CUDA_ERROR;                                                \
    }                                                                       \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr int kNumSMs = 148;

// ---- programmatic dependent launch (PDL).  Every kernel of the guided step calls pdl_wait() before it touches
// memory written by an earlier kernel and is launched with launch_pdl(): its CTAs may then be scheduled while the
// previous kernel's last CTAs are still running, so launch latency and the prologue (barrier init, TMEM
// allocation, descriptor prefetch) overlap the predecessor's tail.  Persistent / single-wave kernels call
// pdl_trigger() right after their prologue; multi-wave kernels rely on the implicit trigger at CTA exit.
// B2E_PDL=0 launches without the attribute (the two instructions are then no-ops).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---- streaming 128-bit accesses (data touched once: keep it out of L1, evict-first in L2)
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, const float4& v) { __stcs(p, v); }
// read-only data that IS reused across CTAs (masks, broadcast noise): default caching
__device__ __forceinline__ float4 ld_reuse(const float4* p) { return __ldg(p); }

// torch.sign semantics: (0 < x) - (x < 0); NaN -> 0
__device__ __forceinline__ float sign_torch(float x) { return (float)((0.f < x) - (x < 0.f)); }
// torch.clamp semantics: NaN propagates
__device__ __forceinline__ float clamp_torch(float x, float lo, float hi) {
  return x < lo ? lo : (x > hi ? hi : x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace b2e
