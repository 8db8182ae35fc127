// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a): host-side plan + launcher.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace b2e {

typedef __nv_bfloat16 bf16;

constexpr int kConvBlockM = 128;  // output pixels per CTA tile (UMMA_M)
constexpr int kConvBlockK = 64;   // bf16 channels per pipeline stage (one 128B swizzle row)

struct ConvEpilogue {
  const float* bias = nullptr;     // [Cout]
  const float* temb = nullptr;     // [N][temb_stride], already offset to this layer's columns
  int temb_stride = 0;
  const bf16* residual = nullptr;  // NHWC, same shape as the output
  bf16* out_bf16 = nullptr;        // NHWC [N,Ho,Wo,Cout]
  float* out_f32_nchw = nullptr;   // NCHW [N,Cout,Ho,Wo] (network output)
  float* gn_partial = nullptr;     // reserved: fused GroupNorm statistics
};

// Everything the kernel needs that is fixed per layer; built once at model-build time.
struct ConvPlan {
  CUtensorMap map_a0, map_a1, map_b;
  int N, Ho, Wo, Cout, cout_pad;
  int Wt, Ht, Nt, w_blks, h_blks, n_blks;
  int taps, c0_chunks, c1_chunks;
  int tap_dc[9], tap_dw[9], tap_da[9], tap_dh[9];
  int block_n;  // 16, 64 or 128
  double flops;
};

struct ConvSrc {
  const bf16* ptr;  // NHWC [N,H,W,C]
  int C;
};

// Builds the TMA descriptors and tile geometry for y = conv(x0 ++ x1) with a ksize x ksize
// kernel.  stride 1: padding ksize/2.  stride 2: ksize 3, padding (0,1,0,1) (Downsample2D).
// w_packed: bf16 [cout_pad][ksize*ksize][C0+C1].
int conv_plan_build(ConvPlan* plan, ConvSrc s0, ConvSrc s1, int N, int H, int W, int ksize, int stride,
                    const bf16* w_packed, int Cout);
int conv_cout_pad(int Cout);
int conv_launch(const ConvPlan& plan, const ConvEpilogue& ep, cudaStream_t st);

// fp32 [Cout][Cin][k][k] -> bf16 [cout_pad][k][k][cin_pad] (zero padded), Cin placed at
// channel offset cin_off of a cin_total-wide K row.
int conv_pack_weight(const float* w, bf16* out, int Cout, int cout_pad, int Cin, int cin_total,
                     int ksize, cudaStream_t st);

}  // namespace b2e
