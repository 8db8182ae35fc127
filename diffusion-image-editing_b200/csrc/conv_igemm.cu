// Implicit-GEMM convolution / linear layer on the 5th-gen tensor cores (tcgen05, sm_100a).
//
//   out[n,h,w,co] = sum_{tap, c} A[n, h*s+dh(tap), w*s+dw(tap), c] * Wt[co, tap, c]  (+ epilogue)
//
// GEMM view: M = output pixels (128 per CTA: an Nt x Ht x Wt brick of the NHWC output),
// N = output channels (BN per CTA), K = taps x input channels, walked 64 channels at a time.
//  * A is never materialised (no im2col buffer): for every (tap, 64-channel chunk) one TMA
//    tiled load fetches the shifted Nt x Ht x Wt x 64 brick of the NHWC activation straight
//    into a 128B-swizzled K-major smem tile; the conv zero padding is TMA out-of-bounds fill.
//    Stride-2 convolutions view the input as (2C, W/2, 2, H/2, N) so a tap is again a box.
//    The K loop can draw chunks from two tensors (skip concatenation without a copy).
//  * B (weights, bf16 [Cout][tap][Cin]) is a plain 2-D TMA tile.
//  * one elected thread issues tcgen05.mma (UMMA 128 x BN x 16, bf16 in, fp32 accumulate in
//    TMEM); tcgen05.commit releases smem stages / signals the epilogue through mbarriers.
//  * warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 =
//    epilogue (tcgen05.ld TMEM -> registers, + bias + time-embedding + residual, bf16 NHWC or
//    fp32 NCHW store).  3-stage smem ring, two CTAs per SM so one CTA's epilogue overlaps the
//    other's main loop.
#include "conv_igemm.cuh"

#include <cudaTypedefs.h>

namespace b2e {

constexpr int kConvThreads = 192;
constexpr int kABytes = kConvBlockM * kConvBlockK * 2;  // 16 KB

template <int BN>
struct ConvCfg {
  static constexpr int kBBytes = BN * kConvBlockK * 2;
  static constexpr int kBBytesPad = (kBBytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytesPad;
  static constexpr int kStages = BN == 128 ? 3 : 4;
  static constexpr int kTmemCols = BN < 32 ? 32 : BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct ConvKParams {
  int N, Ho, Wo, Cout;
  int Wt, Ht, Nt, w_blks, h_blks;
  int taps, c0_chunks, c1_chunks;
  int tap_dc[9], tap_dw[9], tap_da[9], tap_dh[9];
  int n_tiles;
  const float* bias;
  const float* temb;
  int temb_stride;
  const bf16* residual;
  bf16* out_bf16;
  float* out_f32_nchw;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
      "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, 128B-swizzled smem tile: rows of 128 B, 8-row swizzle atoms 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                   // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: 8 rows * 128 B
  d |= (uint64_t)1 << 46;                   // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ------------------------------------------------------------------ the kernel
template <int BN>
__global__ void __launch_bounds__(kConvThreads, BN == 128 ? 2 : 2)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_b, const ConvKParams p) {
  using Cfg = ConvCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x % p.n_tiles;
  int m_tile = blockIdx.x / p.n_tiles;
  const int w_blk = m_tile % p.w_blks; m_tile /= p.w_blks;
  const int h_blk = m_tile % p.h_blks;
  const int n_blk = m_tile / p.h_blks;
  const int w0 = w_blk * p.Wt, h0 = h_blk * p.Ht, n0 = n_blk * p.Nt;
  const int chunks = p.c0_chunks + p.c1_chunks;
  const int num_kb = p.taps * chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a0);
    if (p.c1_chunks) prefetch_tmap(&map_a1);
    prefetch_tmap(&map_b);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        const int cw = w0 + p.tap_dw[tap], ch = h0 + p.tap_dh[tap], ca = p.tap_da[tap], cc = p.tap_dc[tap];
        for (int ck = 0; ck < chunks; ++ck) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + kABytes;
          mbar_expect_tx(full_bar + stage, kABytes + Cfg::kBBytes);
          if (ck < p.c0_chunks)
            tma_load_5d(sa, &map_a0, full_bar + stage, cc + ck * kConvBlockK, cw, ca, ch, n0);
          else
            tma_load_5d(sa, &map_a1, full_bar + stage, cc + (ck - p.c0_chunks) * kConvBlockK, cw, ca, ch, n0);
          tma_load_2d(sb, &map_b, full_bar + stage, (tap * chunks + ck) * kConvBlockK, n_tile * BN);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kConvBlockM, BN);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(full_bar + stage, phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint64_t adesc = make_smem_desc(sa);
        const uint64_t bdesc = make_smem_desc(sa + kABytes);
#pragma unroll
        for (int k = 0; k < kConvBlockK / 16; ++k) {
          // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        umma_commit(empty_bar + stage);  // frees this smem stage once the MMAs above retire
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  } else {
    // ===== epilogue: 4 warps, warp w owns TMEM lanes 32*(w%4) .. +31
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile = output pixel
    const int w_l = r % p.Wt, h_l = (r / p.Wt) % p.Ht, n_l = r / (p.Wt * p.Ht);
    const int n = n0 + n_l, h = h0 + h_l, w = w0 + w_l;
    const bool valid = n < p.N;
    const int64_t pix = ((int64_t)n * p.Ho + h) * p.Wo + w;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c = 0; c < BN / 16; ++c) {
      float v[16];
      tmem_ld16(taddr + c * 16, v);
      const int col0 = n_tile * BN + c * 16;
      if (!valid || col0 >= p.Cout) continue;
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.Cout) v[j] += __ldg(p.bias + col0 + j);
      }
      if (p.temb) {
        const float* t = p.temb + (int64_t)n * p.temb_stride + col0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.Cout) v[j] += __ldg(t + j);
      }
      if (p.out_bf16) {
        // Cout % 16 == 0 on this path (checked on the host)
        if (p.residual) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.Cout + col0);
          uint4 r0 = __ldg(rp), r1 = __ldg(rp + 1);
          const __nv_bfloat162* rb0 = reinterpret_cast<const __nv_bfloat162*>(&r0);
          const __nv_bfloat162* rb1 = reinterpret_cast<const __nv_bfloat162*>(&r1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 a = __bfloat1622float2(rb0[j]), b = __bfloat1622float2(rb1[j]);
            v[2 * j] += a.x; v[2 * j + 1] += a.y;
            v[8 + 2 * j] += b.x; v[8 + 2 * j + 1] += b.y;
          }
        }
        uint4 o0, o1;
        __nv_bfloat162* ob0 = reinterpret_cast<__nv_bfloat162*>(&o0);
        __nv_bfloat162* ob1 = reinterpret_cast<__nv_bfloat162*>(&o1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ob0[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
          ob1[j] = __floats2bfloat162_rn(v[8 + 2 * j], v[8 + 2 * j + 1]);
        }
        uint4* op = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.Cout + col0);
        op[0] = o0;
        op[1] = o1;
      }
      if (p.out_f32_nchw) {
        const int64_t hw = (int64_t)p.Ho * p.Wo;
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (col0 + j < p.Cout)
            p.out_f32_nchw[((int64_t)n * p.Cout + col0 + j) * hw + (int64_t)h * p.Wo + w] = v[j];
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------ weight packing
__global__ void pack_weight_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout,
                                   int Cin, int cin_total, int cin_off, int kk) {
  // out[(co*kk + t)*cin_total + cin_off + ci] = w[(co*Cin + ci)*kk + t]
  const int64_t total = (int64_t)Cout * kk * Cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int t = (int)((i / Cin) % kk);
    const int co = (int)(i / ((int64_t)Cin * kk));
    out[((int64_t)co * kk + t) * cin_total + cin_off + ci] = __float2bfloat16_rn(w[((int64_t)co * Cin + ci) * kk + t]);
  }
}

int conv_pack_weight(const float* w, bf16* out, int Cout, int cout_pad, int Cin, int cin_total,
                     int ksize, cudaStream_t st) {
  (void)cout_pad;
  const int kk = ksize * ksize;
  const int64_t total = (int64_t)Cout * kk * Cin;
  int grid = (int)((total + 255) / 256);
  if (grid > kNumSMs * 16) grid = kNumSMs * 16;
  pack_weight_kernel<<<grid, 256, 0, st>>>(w, out, Cout, Cin, cin_total, 0, kk);
  return check_launch("pack_weight");
}

// ------------------------------------------------------------------ host: TMA descriptors
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  }
  return fn;
}

static int encode_map(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims,
                      const uint64_t* strides_bytes, const uint32_t* box) {
  auto enc = get_encode();
  B2E_REQUIRE(enc, B2E_CUDA_ERROR, "cuTensorMapEncodeTiled entry point not available");
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr),
                   (const cuuint64_t*)dims, (const cuuint64_t*)strides_bytes, (const cuuint32_t*)box,
                   (const cuuint32_t*)estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2E_REQUIRE(r == CUDA_SUCCESS, B2E_CUDA_ERROR,
              "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu %llu] box [%u %u %u %u %u]",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], box[1], rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
  return B2E_OK;
}

static int pow2_divisor(int v, int cap) {
  int p = 1;
  while (p * 2 <= cap && v % (p * 2) == 0) p *= 2;
  return p;
}

int conv_cout_pad(int Cout) {
  if (Cout <= 16) return 16;
  if (Cout % 128 == 0) return Cout;
  return (Cout + 63) / 64 * 64;
}

static int encode_act_map(CUtensorMap* m, const bf16* ptr, int N, int H, int W, int C, int stride,
                          int Wt, int Ht, int Nt) {
  const uint64_t e = 2;
  uint64_t dims[5], str[4];
  uint32_t box[5] = {(uint32_t)kConvBlockK, (uint32_t)Wt, 1, (uint32_t)Ht, (uint32_t)Nt};
  if (stride == 1) {
    dims[0] = C; dims[1] = W; dims[2] = 1; dims[3] = H; dims[4] = N;
    str[0] = (uint64_t)C * e; str[1] = (uint64_t)W * C * e; str[2] = (uint64_t)W * C * e;
    str[3] = (uint64_t)H * W * C * e;
  } else {
    dims[0] = 2 * (uint64_t)C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = N;
    str[0] = 2 * (uint64_t)C * e; str[1] = (uint64_t)W * C * e; str[2] = 2 * (uint64_t)W * C * e;
    str[3] = (uint64_t)H * W * C * e;
  }
  return encode_map(m, ptr, 5, dims, str, box);
}

int conv_plan_build(ConvPlan* pl, ConvSrc s0, ConvSrc s1, int N, int H, int W, int ksize, int stride,
                    const bf16* w_packed, int Cout) {
  B2E_REQUIRE(s0.ptr && s0.C > 0 && s0.C % kConvBlockK == 0 && (s1.C % kConvBlockK == 0), B2E_UNSUPPORTED_SHAPE,
              "conv: input channels must be multiples of %d (got %d + %d)", kConvBlockK, s0.C, s1.C);
  B2E_REQUIRE((ksize == 1 || ksize == 3) && (stride == 1 || (stride == 2 && ksize == 3 && !s1.ptr)),
              B2E_UNSUPPORTED_SHAPE, "conv: unsupported ksize/stride %d/%d", ksize, stride);
  B2E_REQUIRE(stride == 1 || (H % 2 == 0 && W % 2 == 0), B2E_UNSUPPORTED_SHAPE, "conv: stride 2 needs even H, W");
  B2E_REQUIRE(aligned16(s0.ptr) && (!s1.ptr || aligned16(s1.ptr)) && aligned16(w_packed), B2E_INVALID_ARG,
              "conv: unaligned tensor");
  ConvPlan& p = *pl;
  p.N = N; p.Ho = H / stride; p.Wo = W / stride; p.Cout = Cout; p.cout_pad = conv_cout_pad(Cout);
  p.block_n = p.cout_pad <= 16 ? 16 : (p.cout_pad % 128 == 0 ? 128 : 64);
  p.Wt = pow2_divisor(p.Wo, kConvBlockM);
  p.Ht = pow2_divisor(p.Ho, kConvBlockM / p.Wt);
  p.Nt = kConvBlockM / (p.Wt * p.Ht);
  p.w_blks = p.Wo / p.Wt; p.h_blks = p.Ho / p.Ht; p.n_blks = (N + p.Nt - 1) / p.Nt;
  p.taps = ksize * ksize;
  p.c0_chunks = s0.C / kConvBlockK;
  p.c1_chunks = s1.ptr ? s1.C / kConvBlockK : 0;
  for (int t = 0; t < p.taps; ++t) {
    const int kh = t / ksize, kw = t % ksize;
    if (stride == 1) {
      p.tap_dc[t] = 0; p.tap_dw[t] = kw - ksize / 2; p.tap_da[t] = 0; p.tap_dh[t] = kh - ksize / 2;
    } else {
      p.tap_dc[t] = (kw & 1) * s0.C; p.tap_dw[t] = kw >> 1; p.tap_da[t] = kh & 1; p.tap_dh[t] = kh >> 1;
    }
  }
  int rc = encode_act_map(&p.map_a0, s0.ptr, N, H, W, s0.C, stride, p.Wt, p.Ht, p.Nt);
  if (rc) return rc;
  if (s1.ptr) {
    rc = encode_act_map(&p.map_a1, s1.ptr, N, H, W, s1.C, stride, p.Wt, p.Ht, p.Nt);
    if (rc) return rc;
  } else {
    p.map_a1 = p.map_a0;
  }
  const uint64_t ktot = (uint64_t)p.taps * (s0.C + (s1.ptr ? s1.C : 0));
  uint64_t bd[2] = {ktot, (uint64_t)p.cout_pad};
  uint64_t bs[1] = {ktot * 2};
  uint32_t bb[2] = {(uint32_t)kConvBlockK, (uint32_t)p.block_n};
  rc = encode_map(&p.map_b, w_packed, 2, bd, bs, bb);
  if (rc) return rc;
  p.flops = 2.0 * N * p.Ho * p.Wo * (double)Cout * (double)ktot;
  return B2E_OK;
}

template <int BN>
static int launch_t(const ConvPlan& pl, const ConvKParams& kp, int grid, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    B2E_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ConvCfg<BN>::kSmemBytes));
    attr_set = true;
  }
  conv_igemm_kernel<BN><<<grid, kConvThreads, ConvCfg<BN>::kSmemBytes, st>>>(pl.map_a0, pl.map_a1, pl.map_b, kp);
  return check_launch("conv_igemm");
}

int conv_launch(const ConvPlan& pl, const ConvEpilogue& ep, cudaStream_t st) {
  B2E_REQUIRE(ep.out_bf16 || ep.out_f32_nchw, B2E_INVALID_ARG, "conv: no output");
  B2E_REQUIRE(!ep.out_bf16 || pl.Cout % 16 == 0, B2E_UNSUPPORTED_SHAPE,
              "conv: bf16 NHWC output needs Cout %% 16 == 0 (got %d)", pl.Cout);
  ConvKParams kp;
  kp.N = pl.N; kp.Ho = pl.Ho; kp.Wo = pl.Wo; kp.Cout = pl.Cout;
  kp.Wt = pl.Wt; kp.Ht = pl.Ht; kp.Nt = pl.Nt; kp.w_blks = pl.w_blks; kp.h_blks = pl.h_blks;
  kp.taps = pl.taps; kp.c0_chunks = pl.c0_chunks; kp.c1_chunks = pl.c1_chunks;
  for (int t = 0; t < 9; ++t) {
    kp.tap_dc[t] = pl.tap_dc[t]; kp.tap_dw[t] = pl.tap_dw[t]; kp.tap_da[t] = pl.tap_da[t]; kp.tap_dh[t] = pl.tap_dh[t];
  }
  kp.n_tiles = pl.cout_pad / pl.block_n;
  kp.bias = ep.bias; kp.temb = ep.temb; kp.temb_stride = ep.temb_stride; kp.residual = ep.residual;
  kp.out_bf16 = ep.out_bf16; kp.out_f32_nchw = ep.out_f32_nchw;
  const int grid = pl.w_blks * pl.h_blks * pl.n_blks * kp.n_tiles;
  switch (pl.block_n) {
    case 16: return launch_t<16>(pl, kp, grid, st);
    case 64: return launch_t<64>(pl, kp, grid, st);
    default: return launch_t<128>(pl, kp, grid, st);
  }
}

}  // namespace b2e

using namespace b2e;

extern "C" int b2e_conv2d_nhwc_bf16(const void* x, const float* w, const float* bias, void* out, int64_t N,
                                    int64_t H, int64_t W, int64_t Cin, int64_t Cout, int ksize, int stride,
                                    void* stream) {
  B2E_REQUIRE(x && w && out, B2E_INVALID_ARG, "conv2d: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int cout_pad = conv_cout_pad((int)Cout);
  const size_t wbytes = (size_t)cout_pad * ksize * ksize * Cin * sizeof(bf16);
  bf16* wp = nullptr;
  B2E_CUDA(cudaMalloc(&wp, wbytes));
  int rc = B2E_OK;
  do {
    if (cudaMemsetAsync(wp, 0, wbytes, st) != cudaSuccess) { set_error("conv2d: memset failed"); rc = B2E_CUDA_ERROR; break; }
    rc = conv_pack_weight(w, wp, (int)Cout, cout_pad, (int)Cin, (int)Cin, ksize, st);
    if (rc) break;
    ConvPlan plan;
    rc = conv_plan_build(&plan, ConvSrc{(const bf16*)x, (int)Cin}, ConvSrc{nullptr, 0}, (int)N, (int)H, (int)W,
                         ksize, stride, wp, (int)Cout);
    if (rc) break;
    ConvEpilogue ep;
    ep.bias = bias;
    ep.out_bf16 = (bf16*)out;
    rc = conv_launch(plan, ep, st);
  } while (0);
  cudaStreamSynchronize(st);
  cudaFree(wp);
  return rc;
}
