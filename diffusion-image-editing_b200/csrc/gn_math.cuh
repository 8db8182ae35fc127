// GroupNorm(+SiLU) apply arithmetic shared by the stand-alone gn_apply_kernel (csrc/unet_kernels.cu) and the fused
// A-operand transform of the convolution kernel (csrc/conv_igemm.cu, XF variants): the two paths are bit-identical.
#pragma once
#include "act_type.cuh"

namespace b2e {

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const f16x2* b = reinterpret_cast<const f16x2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 t = f16x2_to_float2(b[j]);
    f[2 * j] = t.x; f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  f16x2* b = reinterpret_cast<f16x2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) b[j] = floats_to_f16x2(f[2 * j], f[2 * j + 1]);
  return v;
}
// y * sigmoid(y) = 0.5 y + 0.5 y tanh(0.5 y): one MUFU op per element
__device__ __forceinline__ float silu_tanh(float y) {
  const float hy = 0.5f * y;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(hy));
  return fmaf(hy, th, hy);
}
// 8 channels: f16(silu(x * scale + shift))
__device__ __forceinline__ uint4 gn_apply8(const uint4& raw, const float* scale, const float* shift, bool silu) {
  float f[8];
  unpack8(raw, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float y = fmaf(f[j], scale[j], shift[j]);
    if (silu) y = silu_tanh(y);
    f[j] = y;
  }
  return pack8(f);
}

}  // namespace b2e
