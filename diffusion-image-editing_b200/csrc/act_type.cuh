// The 16-bit storage / tensor-core operand type of the engine.
//
// Default: IEEE fp16 (`tcgen05.mma kind::f16` with f16 A / B, fp32 accumulation in TMEM).  Same UTCHMMA rate as bf16 and
// 3 more mantissa bits (2^-12 instead of 2^-9 relative rounding per stored activation / weight): GroupNorm keeps the
// activations of the noise predictors and decoders O(1-100), far inside the fp16 range, and the gradients of the
// backward passes are kept in range by one power-of-two scale found on the device (csrc/unet_kernels.cu, grad_scale).
// -DB2E_ACT_BF16 rebuilds the whole library on bf16 operands (A/B measurements only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace b2e {

#ifdef B2E_ACT_BF16
typedef __nv_bfloat16 f16;
typedef __nv_bfloat162 f16x2;
constexpr int kActIsBf16 = 1;
__host__ __device__ __forceinline__ float f16_to_float(f16 v) { return __bfloat162float(v); }
__host__ __device__ __forceinline__ f16 float_to_f16(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ f16x2 floats_to_f16x2(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ float2 f16x2_to_float2(f16x2 v) { return __bfloat1622float2(v); }
#define B2E_TMA_DTYPE CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#else
typedef __half f16;
typedef __half2 f16x2;
constexpr int kActIsBf16 = 0;
__host__ __device__ __forceinline__ float f16_to_float(f16 v) { return __half2float(v); }
__host__ __device__ __forceinline__ f16 float_to_f16(float v) { return __float2half_rn(v); }
__device__ __forceinline__ f16x2 floats_to_f16x2(float a, float b) { return __floats2half2_rn(a, b); }
__device__ __forceinline__ float2 f16x2_to_float2(f16x2 v) { return __half22float2(v); }
#define B2E_TMA_DTYPE CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#endif

}  // namespace b2e
