// Fused element-wise / reduction kernels of the guided denoising step (HBM-bound).
//
// Every kernel streams fp32 NCHW tensors with 128-bit coalesced accesses, four
// independent 128-bit loads per input in flight per thread, and evaluates the
// reference's expressions in the reference's op order with explicitly rounded
// (non-contracted) fp32 operations, so results are bit-identical to the PyTorch path:
//   x0  = (x - sb*e) / sa                         src/diffusion_utils.py:27-31
//   xp  = sp*x0 + cdir*e (+ sigma*z)              DDIMScheduler.step / src/ddpm_inversion.py:203-240
//   x0g = (xp - sb*e) / sa ; g = -((k*sign(x0g - tau))/sa) ; xp += (mask*)g * a_t^2
//                                                  src/attr_functions.py:22-37,104-163
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace b2e {

std::atomic<long long> g_launches{0};
bool pdl_enabled() {
  static const bool on = !(getenv("B2E_PDL") && atoi(getenv("B2E_PDL")) == 0);
  return on;
}
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------------------------
struct StepK {
  float sa, sb, sp, cdir, sigma, a2, clip_range;
  int clip, has_noise, noise_batched, guide, mask_grad, mask_batched, no_step;
  int has_target[4];
  float target[4];
  float gk[4];  // k_c / sa  (== (k_c * s)/sa for s in {-1,0,1})
};

static StepK make_stepk(const b2e_guided_step_params* p) {
  StepK k;
  k.sa = p->c.sqrt_a_t; k.sb = p->c.sqrt_b_t; k.sp = p->c.sqrt_a_prev; k.cdir = p->c.dir_coef;
  k.sigma = p->c.sigma; k.a2 = p->c.a_t_sq; k.clip_range = p->clip_range;
  k.clip = p->clip; k.has_noise = p->has_noise; k.noise_batched = p->noise_batched;
  k.guide = p->guide; k.mask_grad = p->mask_grad; k.mask_batched = p->mask_batched;
  k.no_step = p->no_step;
  for (int i = 0; i < 4; ++i) {
    k.has_target[i] = p->has_target[i];
    k.target[i] = p->target[i];
    k.gk[i] = p->coef[i] / p->c.sqrt_a_t;  // host fp32 division (IEEE)
  }
  return k;
}

__device__ __forceinline__ void step_elem(float x, float e, float z, float m, int has_t, float tau,
                                          float gk, const StepK& k, float& xp_out, float& x0_out) {
  float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.sb, e)), k.sa);
  float xp = x;  // no_step: x already is the post-step sample (AttrFunc.apply called on its own)
  if (!k.no_step) {
    if (k.clip) x0 = clamp_torch(x0, -k.clip_range, k.clip_range);
    xp = __fadd_rn(__fmul_rn(k.sp, x0), __fmul_rn(k.cdir, e));
    if (k.has_noise) xp = __fadd_rn(xp, __fmul_rn(k.sigma, z));
  }
  if (k.guide && has_t) {
    float x0g = __fdiv_rn(__fsub_rn(xp, __fmul_rn(k.sb, e)), k.sa);
    float g = -__fmul_rn(gk, sign_torch(__fsub_rn(x0g, tau)));
    if (k.mask_grad) g = __fmul_rn(m, g);
    xp = __fadd_rn(xp, __fmul_rn(g, k.a2));
  }
  xp_out = xp;
  x0_out = x0;
}

constexpr int kStepThreads = 256;
constexpr int kStepUnroll = 4;  // float4 per thread per input

// per-channel parameters without dynamic indexing (keeps StepK in constant bank / registers)
__device__ __forceinline__ void chan_params(const StepK& k, int c, int& ht, float& tau, float& gk) {
  ht = c == 0 ? k.has_target[0] : c == 1 ? k.has_target[1] : c == 2 ? k.has_target[2] : k.has_target[3];
  tau = c == 0 ? k.target[0] : c == 1 ? k.target[1] : c == 2 ? k.target[2] : k.target[3];
  gk = c == 0 ? k.gk[0] : c == 1 ? k.gk[1] : c == 2 ? k.gk[2] : k.gk[3];
}

// total4 = B*C*HW/4 ; HW % 4 == 0 so the four lanes of a float4 share a channel.
// Index math is 32-bit (the host guarantees total4 < 2^31): one division per thread, then the
// (plane, offset) pair is advanced incrementally across the unrolled accesses.
__global__ void __launch_bounds__(kStepThreads, 3)
guided_step_vec4(const float4* __restrict__ x, const float4* __restrict__ e,
                 const float4* __restrict__ z, const float4* __restrict__ mask,
                 float4* __restrict__ xp, float4* __restrict__ x0o, uint32_t total4, uint32_t hw4,
                 uint32_t C, const __grid_constant__ StepK k) {
  pdl_wait();   // eps comes from the UNet's last convolution
  const uint32_t base = blockIdx.x * (uint32_t)(kStepThreads * kStepUnroll) + threadIdx.x;
  float4 vx[kStepUnroll], ve[kStepUnroll], vz[kStepUnroll], vm[kStepUnroll];
  uint32_t ch[kStepUnroll];
  const bool use_mask = k.guide && k.mask_grad;
  uint32_t plane = base / hw4, off = base - plane * hw4;
#pragma unroll
  for (int u = 0; u < kStepUnroll; ++u) {
    const uint32_t i = base + (uint32_t)u * kStepThreads;
    const uint32_t c = plane % C;
    ch[u] = c;
    if (i < total4) {
      vx[u] = ld_stream(x + i);
      ve[u] = ld_stream(e + i);
      const uint32_t bi = c * hw4 + off;  // index inside one (C,H,W) image
      if (k.has_noise) vz[u] = k.noise_batched ? ld_stream(z + i) : ld_reuse(z + bi);
      if (use_mask) vm[u] = k.mask_batched ? ld_stream(mask + i) : ld_reuse(mask + bi);
    }
    off += kStepThreads;
    while (off >= hw4) { off -= hw4; ++plane; }
  }
#pragma unroll
  for (int u = 0; u < kStepUnroll; ++u) {
    const uint32_t i = base + (uint32_t)u * kStepThreads;
    if (i < total4) {
      int ht; float tau, gk;
      chan_params(k, (int)ch[u], ht, tau, gk);
      float4 zz = k.has_noise ? vz[u] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 mm = use_mask ? vm[u] : make_float4(1.f, 1.f, 1.f, 1.f);
      float4 o, o0;
      step_elem(vx[u].x, ve[u].x, zz.x, mm.x, ht, tau, gk, k, o.x, o0.x);
      step_elem(vx[u].y, ve[u].y, zz.y, mm.y, ht, tau, gk, k, o.y, o0.y);
      step_elem(vx[u].z, ve[u].z, zz.z, mm.z, ht, tau, gk, k, o.z, o0.z);
      step_elem(vx[u].w, ve[u].w, zz.w, mm.w, ht, tau, gk, k, o.w, o0.w);
      st_stream(xp + i, o);
      if (x0o) st_stream(x0o + i, o0);
    }
  }
}

// scalar fallback for HW % 4 != 0 or unaligned pointers (same arithmetic)
__global__ void guided_step_scalar(const float* __restrict__ x, const float* __restrict__ e,
                                   const float* __restrict__ z, const float* __restrict__ mask,
                                   float* __restrict__ xp, float* __restrict__ x0o, int64_t total,
                                   int64_t chw, int64_t hw, int C, StepK k) {
  const bool use_mask = k.guide && k.mask_grad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / hw) % C);
    float zz = k.has_noise ? z[k.noise_batched ? i : i % chw] : 0.f;
    float mm = use_mask ? mask[k.mask_batched ? i : i % chw] : 1.f;
    float o, o0;
    int ht; float tau, gk;
    chan_params(k, c, ht, tau, gk);
    step_elem(x[i], e[i], zz, mm, ht, tau, gk, k, o, o0);
    xp[i] = o;
    if (x0o) x0o[i] = o0;
  }
}

// ------------------------------------------------------------------ L2-regularised variant
struct L2K {
  StepK s;
  float lam_scale;  // loss_scale * lambda (fp32 product)
  float k[4];       // k_c = loss_scale*w_c/N
};

// block-wide sum in double; result left in smem[0]
__device__ __forceinline__ void block_sum_to_double(double v, double* smem) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) smem[w] = v;
  __syncthreads();
  if (w == 0) {
    double t = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) smem[0] = t;
  }
  __syncthreads();
}

// pass 1: scheduler update (no guidance), write x_prev / x0, accumulate sum r^2 per block
__global__ void __launch_bounds__(kStepThreads)
l2reg_pass1(const float4* __restrict__ x, const float4* __restrict__ e, const float4* __restrict__ z,
            const float4* __restrict__ mask, const float4* __restrict__ xref,
            float4* __restrict__ xp, float4* __restrict__ x0o, double* __restrict__ partial,
            int64_t total4, int64_t chw4, L2K k) {
  __shared__ double red[32];
  double acc = 0.0;
  StepK s = k.s;
  s.guide = 0;
  // two float4 per thread per iteration, every load issued before the first use (the plain grid-stride loop kept one
  // iteration's 3-5 loads in flight: 0.64 of the HBM peak)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total4; i0 += 2 * stride) {
    float4 vx[2], ve[2], vz[2], vm[2], vr[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      ok[u] = i < total4;
      vz[u] = make_float4(0, 0, 0, 0);
      if (ok[u]) {
        vx[u] = ld_stream(x + i); ve[u] = ld_stream(e + i);
        if (s.has_noise) vz[u] = s.noise_batched ? ld_stream(z + i) : ld_reuse(z + (i % chw4));
        vm[u] = k.s.mask_batched ? ld_stream(mask + i) : ld_reuse(mask + (i % chw4));
        vr[u] = ld_stream(xref + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      const int64_t i = i0 + u * stride;
      float4 o, o0;
      step_elem(vx[u].x, ve[u].x, vz[u].x, 1.f, 0, 0.f, 0.f, s, o.x, o0.x);
      step_elem(vx[u].y, ve[u].y, vz[u].y, 1.f, 0, 0.f, 0.f, s, o.y, o0.y);
      step_elem(vx[u].z, ve[u].z, vz[u].z, 1.f, 0, 0.f, 0.f, s, o.z, o0.z);
      step_elem(vx[u].w, ve[u].w, vz[u].w, 1.f, 0, 0.f, 0.f, s, o.w, o0.w);
      xp[i] = o;  // re-read by pass 2: keep default caching
      if (x0o) st_stream(x0o + i, o0);
      const float xo[4] = {o.x, o.y, o.z, o.w}, ee[4] = {ve[u].x, ve[u].y, ve[u].z, ve[u].w};
      const float mm[4] = {vm[u].x, vm[u].y, vm[u].z, vm[u].w}, rr[4] = {vr[u].x, vr[u].y, vr[u].z, vr[u].w};
      float part = 0.f;     // four squares in fp32, ONE fp64 add per float4 (fp64 adds per element capped the kernel)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x0g = __fdiv_rn(__fsub_rn(xo[j], __fmul_rn(s.sb, ee[j])), s.sa);
        float r = __fsub_rn(__fsub_rn(1.f, __fmul_rn(mm[j], x0g)), rr[j]);
        part = __fadd_rn(part, __fmul_rn(r, r));
      }
      acc += (double)part;
    }
  }
  block_sum_to_double(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// pass 2: R = sqrt(sum), gradient of colour(mask*x0g) + lambda*R, update x_prev in place
__global__ void __launch_bounds__(kStepThreads)
l2reg_pass2(float4* __restrict__ xp, const float4* __restrict__ e, const float4* __restrict__ mask,
            const float4* __restrict__ xref, const double* __restrict__ partial, int n_partial,
            int64_t total4, int64_t chw4, int64_t hw4, int C, L2K k) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n_partial; i += blockDim.x) acc += partial[i];
  block_sum_to_double(acc, red);
  const float R = sqrtf((float)red[0]);
  const float t_reg = __fdiv_rn(k.lam_scale, __fmul_rn(2.f, R));  // grad / (2*sqrt(s))
  const StepK& s = k.s;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total4; i0 += 2 * stride) {
    float4 vo[2], ve[2], vm[2], vr[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      ok[u] = i < total4;
      if (ok[u]) {
        vo[u] = xp[i]; ve[u] = ld_stream(e + i);
        vm[u] = s.mask_batched ? ld_stream(mask + i) : ld_reuse(mask + (i % chw4));
        vr[u] = ld_stream(xref + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      const int64_t i = i0 + u * stride;
      const int c = (int)((i / hw4) % C);
      float xo[4] = {vo[u].x, vo[u].y, vo[u].z, vo[u].w};
      const float ee[4] = {ve[u].x, ve[u].y, ve[u].z, ve[u].w}, mm[4] = {vm[u].x, vm[u].y, vm[u].z, vm[u].w},
                  rr[4] = {vr[u].x, vr[u].y, vr[u].z, vr[u].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x0g = __fdiv_rn(__fsub_rn(xo[j], __fmul_rn(s.sb, ee[j])), s.sa);
        float img = __fmul_rn(mm[j], x0g);
        float r = __fsub_rn(__fsub_rn(1.f, img), rr[j]);
        float dimg = -__fmul_rn(t_reg, __fmul_rn(2.f, r));  // d/dimg of lambda*||r|| (r = 1-img-xref)
        if (s.has_target[c]) dimg = __fadd_rn(dimg, __fmul_rn(k.k[c], sign_torch(__fsub_rn(img, s.target[c]))));
        float g = -__fdiv_rn(__fmul_rn(dimg, mm[j]), s.sa);
        if (s.mask_grad) g = __fmul_rn(mm[j], g);
        xo[j] = __fadd_rn(xo[j], __fmul_rn(g, s.a2));
      }
      st_stream(xp + i, make_float4(xo[0], xo[1], xo[2], xo[3]));
    }
  }
}

// ------------------------------------------------------------------ generic 1/2/3-input maps
enum { OP_PRED_X0 = 0, OP_RENOISE, OP_CFG, OP_EXTRACT };

struct MapK { float a, b, c, d; };

__device__ __forceinline__ float map_elem(int op, float x, float e, const MapK& k) {
  if (op == OP_PRED_X0) return __fdiv_rn(__fsub_rn(x, __fmul_rn(k.b, e)), k.a);
  if (op == OP_RENOISE) {
    float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(k.b, e)), k.a);
    return __fadd_rn(__fmul_rn(k.c, x0), __fmul_rn(k.d, e));
  }
  /* OP_CFG: x = e_first, e = e_second */
  return __fadd_rn(x, __fmul_rn(k.a, __fsub_rn(e, x)));
}

template <int OP>
__global__ void __launch_bounds__(kStepThreads)
map2_kernel(const float* __restrict__ x, const float* __restrict__ e, float* __restrict__ out,
            int64_t n, MapK k, int vec) {
  if (vec) {
    const int64_t n4 = n >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const float4* e4 = reinterpret_cast<const float4*>(e);
    float4* o4 = reinterpret_cast<float4*>(out);
    const int64_t base = (int64_t)blockIdx.x * (kStepThreads * kStepUnroll) + threadIdx.x;
    float4 vx[kStepUnroll], ve[kStepUnroll];
#pragma unroll
    for (int u = 0; u < kStepUnroll; ++u) {
      const int64_t i = base + (int64_t)u * kStepThreads;
      if (i < n4) { vx[u] = ld_stream(x4 + i); ve[u] = ld_stream(e4 + i); }
    }
#pragma unroll
    for (int u = 0; u < kStepUnroll; ++u) {
      const int64_t i = base + (int64_t)u * kStepThreads;
      if (i < n4) {
        float4 o;
        o.x = map_elem(OP, vx[u].x, ve[u].x, k); o.y = map_elem(OP, vx[u].y, ve[u].y, k);
        o.z = map_elem(OP, vx[u].z, ve[u].z, k); o.w = map_elem(OP, vx[u].w, ve[u].w, k);
        st_stream(o4 + i, o);
      }
    }
    // tail (n % 4) handled by block 0
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
      const int64_t i = (n4 << 2) + threadIdx.x;
      out[i] = map_elem(OP, x[i], e[i], k);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
      out[i] = map_elem(OP, x[i], e[i], k);
  }
}

// z = (xm - mu)/sigma ; xm <- mu + sigma*z     (src/ddpm_inversion.py:135-169)
__global__ void __launch_bounds__(kStepThreads)
extract_noise_kernel(const float* __restrict__ x, const float* __restrict__ e, float* __restrict__ xm,
                     float* __restrict__ z, int64_t n, float sa, float sb, float sp, float cdir,
                     float sigma, int vec) {
  auto f = [&](float xv, float ev, float xmv, float& zo, float& xo) {
    float x0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(sb, ev)), sa);
    float mu = __fadd_rn(__fmul_rn(sp, x0), __fmul_rn(cdir, ev));
    zo = __fdiv_rn(__fsub_rn(xmv, mu), sigma);
    xo = __fadd_rn(mu, __fmul_rn(sigma, zo));
  };
  if (vec) {
    const int64_t n4 = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
      float4 vx[2], ve[2], vm[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < n4) {
          vx[u] = ld_stream(reinterpret_cast<const float4*>(x) + i);
          ve[u] = ld_stream(reinterpret_cast<const float4*>(e) + i);
          vm[u] = ld_stream(reinterpret_cast<const float4*>(xm) + i);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < n4) {
          float4 zo, xo;
          f(vx[u].x, ve[u].x, vm[u].x, zo.x, xo.x); f(vx[u].y, ve[u].y, vm[u].y, zo.y, xo.y);
          f(vx[u].z, ve[u].z, vm[u].z, zo.z, xo.z); f(vx[u].w, ve[u].w, vm[u].w, zo.w, xo.w);
          st_stream(reinterpret_cast<float4*>(z) + i, zo);
          st_stream(reinterpret_cast<float4*>(xm) + i, xo);
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
      const int64_t i = (n4 << 2) + threadIdx.x;
      float zo, xo;
      f(x[i], e[i], xm[i], zo, xo);
      z[i] = zo; xm[i] = xo;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
      float zo, xo;
      f(x[i], e[i], xm[i], zo, xo);
      z[i] = zo; xm[i] = xo;
    }
  }
}

// xts[t] = x0*sa[t] + noise[t]*sb[t] for t < T ; xts[T] = x0      (src/ddpm_inversion.py:31-55)
__global__ void __launch_bounds__(kStepThreads)
sample_xts_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                  const float* __restrict__ sa, const float* __restrict__ sb,
                  float* __restrict__ xts, int64_t T, int64_t chw) {
  const int64_t t = blockIdx.y;
  const float a = t < T ? sa[t] : 1.f, b = t < T ? sb[t] : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < chw;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x0[i];
    xts[t * chw + i] = t < T ? __fadd_rn(__fmul_rn(v, a), __fmul_rn(noise[t * chw + i], b)) : v;
  }
}

// out = mask*zv + (1-mask)*zo, mask broadcast over T            (src/utils.py:23-28)
// vec: chw % 4 == 0 and 16-byte aligned pointers - two float4 per thread per iteration, loads first
__global__ void __launch_bounds__(kStepThreads)
apply_mask_kernel(const float* __restrict__ mask, const float* __restrict__ zo,
                  const float* __restrict__ zv, float* __restrict__ out, int64_t n, int64_t chw, int vec) {
  auto f = [](float m, float v, float o) { return __fadd_rn(__fmul_rn(m, v), __fmul_rn(__fsub_rn(1.f, m), o)); };
  if (vec) {
    const int64_t n4 = n >> 2, chw4 = chw >> 2, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 2 * stride) {
      float4 vm[2], vv[2], vo[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < n4) {
          vm[u] = ld_reuse(reinterpret_cast<const float4*>(mask) + (i % chw4));
          vv[u] = ld_stream(reinterpret_cast<const float4*>(zv) + i);
          vo[u] = ld_stream(reinterpret_cast<const float4*>(zo) + i);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t i = i0 + u * stride;
        if (i < n4)
          st_stream(reinterpret_cast<float4*>(out) + i, make_float4(f(vm[u].x, vv[u].x, vo[u].x), f(vm[u].y, vv[u].y, vo[u].y),
                                                                    f(vm[u].z, vv[u].z, vo[u].z), f(vm[u].w, vv[u].w, vo[u].w)));
      }
    }
    return;
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = f(__ldg(mask + (i % chw)), zv[i], zo[i]);
}

// x += (mask*)g * a2
__global__ void __launch_bounds__(kStepThreads)
apply_grad_kernel(float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mask,
                  int64_t n, int64_t chw, int mask_batched, float a2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    float gv = g[i];
    if (mask) gv = __fmul_rn(mask[mask_batched ? i : i % chw], gv);
    x[i] = __fadd_rn(x[i], __fmul_rn(gv, a2));
  }
}

// (B,C,HW) fp32 -> (B,HW,C) u8 ; trunc(clamp(x/2+0.5,0,1)*255)  (src/transforms.py:8-35)
__device__ __forceinline__ uint32_t u8_of(float v) {
  v = clamp_torch(__fadd_rn(__fmul_rn(v, 0.5f), 0.5f), 0.f, 1.f);
  return (uint32_t)(uint8_t)__fmul_rn(v, 255.f);
}
// vec (C == 3, hw % 4 == 0, aligned): a thread converts four consecutive pixels - three float4 loads (one per channel
// plane), three 32-bit stores (12 interleaved bytes); two such groups per iteration with the loads first
__global__ void __launch_bounds__(kStepThreads)
to_uint8_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int64_t B, int C, int64_t hw, int vec) {
  if (vec) {
    const int64_t hw4 = hw >> 2, total4 = B * hw4, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g0 < total4; g0 += 2 * stride) {
      float4 v[2][3];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t g = g0 + u * stride;
        if (g < total4) {
          const int64_t b = g / hw4, q4 = g - b * hw4;
#pragma unroll
          for (int c = 0; c < 3; ++c) v[u][c] = ld_stream(reinterpret_cast<const float4*>(x + (b * 3 + c) * hw) + q4);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t g = g0 + u * stride;
        if (g < total4) {
          const float r[4] = {v[u][0].x, v[u][0].y, v[u][0].z, v[u][0].w}, gg[4] = {v[u][1].x, v[u][1].y, v[u][1].z, v[u][1].w},
                      bb[4] = {v[u][2].x, v[u][2].y, v[u][2].z, v[u][2].w};
          uint32_t by[12];
#pragma unroll
          for (int j = 0; j < 4; ++j) { by[3 * j] = u8_of(r[j]); by[3 * j + 1] = u8_of(gg[j]); by[3 * j + 2] = u8_of(bb[j]); }
          uint32_t* o = reinterpret_cast<uint32_t*>(out + g * 12);
#pragma unroll
          for (int w = 0; w < 3; ++w)
            __stcs(o + w, by[4 * w] | (by[4 * w + 1] << 8) | (by[4 * w + 2] << 16) | (by[4 * w + 3] << 24));
        }
      }
    }
    return;
  }
  const int64_t total = B * hw;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total;
       p += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = p / hw, q = p % hw;
    for (int c = 0; c < C; ++c) out[p * C + c] = (uint8_t)u8_of(x[(b * C + c) * hw + q]);
  }
}


// ---- analytic colour-loss gradient w.r.t. a decoded image (guidance through a latent decoder) + the latent update
struct ColorGradK { int has_target[4]; float target[4], k[4]; float lam_scale; int use_mask_pred, mask_batched; };

// pass 1 (mask_pred variant only): sum of r^2, r = 1 - mask*img - x_ref
__global__ void __launch_bounds__(kStepThreads)
color_grad_norm_kernel(const float4* __restrict__ img, const float4* __restrict__ mask, const float4* __restrict__ xref,
                       double* __restrict__ partial, int64_t total4, int64_t chw4, int mask_batched) {
  __shared__ double red[32];
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total4; i0 += 2 * stride) {
    float4 vi[2], vm[2], vr[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total4) {
        vi[u] = __ldg(img + i);      // re-read by pass 2: default caching
        vm[u] = mask_batched ? ld_stream(mask + i) : ld_reuse(mask + (i % chw4));
        vr[u] = __ldg(xref + i);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (i0 + u * stride >= total4) continue;
      const float ii[4] = {vi[u].x, vi[u].y, vi[u].z, vi[u].w}, mm[4] = {vm[u].x, vm[u].y, vm[u].z, vm[u].w},
                  rr[4] = {vr[u].x, vr[u].y, vr[u].z, vr[u].w};
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float r = __fsub_rn(__fsub_rn(1.f, __fmul_rn(mm[j], ii[j])), rr[j]);
        part = __fadd_rn(part, __fmul_rn(r, r));
      }
      acc += (double)part;
    }
  }
  block_sum_to_double(acc, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// pass 2 (or the only pass): the gradient itself
__global__ void __launch_bounds__(kStepThreads)
color_grad_kernel(const float4* __restrict__ img, const float4* __restrict__ mask, const float4* __restrict__ xref,
                  float4* __restrict__ dimg, const double* __restrict__ partial, int n_partial, int64_t total4, int64_t chw4,
                  int64_t hw4, int C, ColorGradK k) {
  __shared__ double red[32];
  float t_reg = 0.f;
  if (k.use_mask_pred) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) acc += partial[i];
    block_sum_to_double(acc, red);
    const float R = sqrtf((float)red[0]);
    t_reg = __fdiv_rn(k.lam_scale, __fmul_rn(2.f, R));     // d(lambda ||r||)/dr = lam r / R, written as the autograd chain does
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total4; i0 += 2 * stride) {
    float4 vi[2], vm[2], vr[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total4) {
        vi[u] = ld_stream(img + i);
        if (k.use_mask_pred) {
          vm[u] = k.mask_batched ? ld_stream(mask + i) : ld_reuse(mask + (i % chw4));
          vr[u] = ld_stream(xref + i);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= total4) continue;
      const int c = (int)((i / hw4) % C);
      const int ht = c == 0 ? k.has_target[0] : c == 1 ? k.has_target[1] : c == 2 ? k.has_target[2] : k.has_target[3];
      const float tau = c == 0 ? k.target[0] : c == 1 ? k.target[1] : c == 2 ? k.target[2] : k.target[3];
      const float kc = c == 0 ? k.k[0] : c == 1 ? k.k[1] : c == 2 ? k.k[2] : k.k[3];
      const float ii[4] = {vi[u].x, vi[u].y, vi[u].z, vi[u].w};
      float o[4];
      if (k.use_mask_pred) {
        const float mm[4] = {vm[u].x, vm[u].y, vm[u].z, vm[u].w}, rr[4] = {vr[u].x, vr[u].y, vr[u].z, vr[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float im = __fmul_rn(mm[j], ii[j]);
          const float r = __fsub_rn(__fsub_rn(1.f, im), rr[j]);
          float d = -__fmul_rn(t_reg, __fmul_rn(2.f, r));
          if (ht) d = __fadd_rn(d, __fmul_rn(kc, sign_torch(__fsub_rn(im, tau))));
          o[j] = __fmul_rn(d, mm[j]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = ht ? __fmul_rn(kc, sign_torch(__fsub_rn(ii[j], tau))) : 0.f;
      }
      st_stream(dimg + i, make_float4(o[0], o[1], o[2], o[3]));
    }
  }
}

// x += (mask*) (-(d * chain) / sa) * a2
__global__ void __launch_bounds__(kStepThreads)
latent_guidance_kernel(float* __restrict__ x, const float* __restrict__ d, const float* __restrict__ mask, int64_t n,
                       int64_t chw, int mask_batched, float chain, float sa, float a2) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = -__fdiv_rn(__fmul_rn(d[i], chain), sa);
    if (mask) g = __fmul_rn(mask[mask_batched ? i : i % chw], g);
    x[i] = __fadd_rn(x[i], __fmul_rn(g, a2));
  }
}

// out = a*x + b*y
__global__ void __launch_bounds__(kStepThreads)
axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
             int64_t n, float a, float b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(a, x[i]), __fmul_rn(b, y[i]));
}

// Loss values (reductions; the guidance update itself never needs them).
// mode 0: sum (x-y)^2 -> partial[block][0]
// mode 1: per channel sum |x - target_c| -> partial[block][c]   (x is (B,C,HW))
__global__ void __launch_bounds__(kStepThreads)
loss_partial_kernel(const float* __restrict__ x, const float* __restrict__ y, double* __restrict__ partial,
                    int64_t n, int64_t hw, int C, int mode, float t0, float t1, float t2, float t3) {
  __shared__ double red[32];
  double acc[4] = {0.0, 0.0, 0.0, 0.0};
  const float tg[4] = {t0, t1, t2, t3};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (mode == 0) {
      const float d = __fsub_rn(x[i], y[i]);
      acc[0] += (double)__fmul_rn(d, d);
    } else {
      const int c = (int)((i / hw) % C);
      const double v = (double)fabsf(__fsub_rn(x[i], tg[c]));
      if (c == 0) acc[0] += v; else if (c == 1) acc[1] += v; else if (c == 2) acc[2] += v; else acc[3] += v;
    }
  }
  for (int j = 0; j < 4; ++j) {
    block_sum_to_double(acc[j], red);
    if (threadIdx.x == 0) partial[(int64_t)blockIdx.x * 4 + j] = red[0];
    __syncthreads();
  }
}
// mode 0: out[0] = sqrt(sum) ; mode 1: out[c] = sum_c / count
__global__ void loss_final_kernel(const double* __restrict__ partial, int n_partial, float* __restrict__ out,
                                  int mode, double count) {
  __shared__ double red[32];
  for (int j = 0; j < 4; ++j) {
    double a = 0.0;
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) a += partial[(int64_t)i * 4 + j];
    block_sum_to_double(a, red);
    if (threadIdx.x == 0) {
      if (mode == 0) { if (j == 0) out[0] = sqrtf((float)red[0]); }
      else out[j] = (float)(red[0] / count);
    }
    __syncthreads();
  }
}

static inline int grid_for(int64_t n, int per_block, int max_blocks = kNumSMs * 16) {
  int64_t g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace b2e

using namespace b2e;

template <int OP>
static int launch_map2(const float* x, const float* e, float* out, int64_t n, MapK k, void* stream,
                       const char* what) {
  B2E_REQUIRE(x && e && out && n > 0, B2E_INVALID_ARG, "%s: bad argument", what);
  const bool vec = aligned16(x) && aligned16(e) && aligned16(out) && n >= 4;
  if (vec) {
    const int per_block = kStepThreads * kStepUnroll;
    const int64_t grid = ((n >> 2) + per_block - 1) / per_block;
    map2_kernel<OP><<<(unsigned)grid, kStepThreads, 0, (cudaStream_t)stream>>>(x, e, out, n, k, 1);
  } else {
    map2_kernel<OP><<<grid_for(n, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(x, e, out, n, k, 0);
  }
  return check_launch(what);
}

// ====================================================================== C ABI
extern "C" {

int b2e_version(void) { return B2E_VERSION; }
const char* b2e_last_error(void) { return g_err; }
int64_t b2e_launch_count(void) { return (int64_t)g_launches.load(); }

int b2e_device_check(void) {
  int dev = 0;
  B2E_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  B2E_CUDA(cudaGetDeviceProperties(&prop, dev));
  B2E_REQUIRE(prop.major == 10 && prop.minor == 0, B2E_ARCH_MISMATCH,
              "libb200edit is built for sm_100a only; device %d is sm_%d%d", dev, prop.major, prop.minor);
  return B2E_OK;
}

static int validate_step(const void* x, const void* e, const void* xp, int64_t B, int64_t C,
                         int64_t HW, const b2e_guided_step_params* p, const void* z, const void* mask) {
  B2E_REQUIRE(x && e && xp && p, B2E_INVALID_ARG, "guided_step: null pointer");
  B2E_REQUIRE(B > 0 && HW > 0 && C > 0 && C <= 4, B2E_UNSUPPORTED_SHAPE,
              "guided_step: need B>0, 1<=C<=4, HW>0 (got B=%lld C=%lld HW=%lld)", (long long)B,
              (long long)C, (long long)HW);
  B2E_REQUIRE(!p->has_noise || z, B2E_INVALID_ARG, "guided_step: eta > 0 but no noise tensor");
  B2E_REQUIRE(!(p->guide && p->mask_grad) || mask, B2E_INVALID_ARG,
              "guided_step: mask_attr_grad set but no mask");
  return B2E_OK;
}

int b2e_guided_step_f32(const float* x_t, const float* eps, const float* z, const float* mask,
                        float* x_prev, float* x0_pred, int64_t B, int64_t C, int64_t HW,
                        const b2e_guided_step_params* p, void* stream) {
  int rc = validate_step(x_t, eps, x_prev, B, C, HW, p, z, mask);
  if (rc) return rc;
  StepK k = make_stepk(p);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = B * C * HW;
  const bool vec = (HW % 4 == 0) && total / 4 < (int64_t)0x7fffffff - kStepThreads * kStepUnroll &&
                   aligned16(x_t) && aligned16(eps) && aligned16(x_prev) && (!x0_pred || aligned16(x0_pred)) &&
                   (!z || aligned16(z)) && (!mask || aligned16(mask));
  if (vec) {
    const int64_t total4 = total / 4;
    const int per_block = kStepThreads * kStepUnroll;
    const int64_t grid = (total4 + per_block - 1) / per_block;
    launch_pdl(guided_step_vec4, dim3((unsigned)grid), dim3(kStepThreads), 0, st,
               (const float4*)x_t, (const float4*)eps, (const float4*)z, (const float4*)mask,
               (float4*)x_prev, (float4*)x0_pred, (uint32_t)total4, (uint32_t)(HW / 4), (uint32_t)C, k);
  } else {
    guided_step_scalar<<<grid_for(total, kStepThreads), kStepThreads, 0, st>>>(
        x_t, eps, z, mask, x_prev, x0_pred, total, C * HW, HW, (int)C, k);
  }
  return check_launch("guided_step");
}

size_t b2e_l2reg_workspace_bytes(int64_t, int64_t, int64_t) { return sizeof(double) * kNumSMs * 8; }

int b2e_guided_step_l2reg_f32(const float* x_t, const float* eps, const float* z, const float* mask,
                              const float* x_ref, float* x_prev, float* x0_pred, int64_t B, int64_t C,
                              int64_t HW, const b2e_l2reg_params* p, void* workspace,
                              size_t workspace_bytes, void* stream) {
  B2E_REQUIRE(p, B2E_INVALID_ARG, "l2reg: null params");
  int rc = validate_step(x_t, eps, x_prev, B, C, HW, &p->base, z, mask);
  if (rc) return rc;
  B2E_REQUIRE(mask && x_ref, B2E_INVALID_ARG, "l2reg: mask and x_0 are required");
  B2E_REQUIRE(workspace && workspace_bytes >= b2e_l2reg_workspace_bytes(B, C, HW),
              B2E_WORKSPACE_TOO_SMALL, "l2reg: workspace too small");
  B2E_REQUIRE(HW % 4 == 0 && aligned16(x_t) && aligned16(eps) && aligned16(x_prev) &&
                  aligned16(mask) && aligned16(x_ref) && (!z || aligned16(z)) &&
                  (!x0_pred || aligned16(x0_pred)),
              B2E_UNSUPPORTED_SHAPE, "l2reg: needs HW %% 4 == 0 and 16-byte aligned tensors");
  L2K k;
  k.s = make_stepk(&p->base);
  k.lam_scale = p->loss_scale * p->lambda_;
  for (int i = 0; i < 4; ++i) k.k[i] = p->base.coef[i];
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total4 = B * C * HW / 4;
  B2E_REQUIRE(total4 < (int64_t)0x7fffffff, B2E_UNSUPPORTED_SHAPE, "l2reg: tensor too large for 32-bit index arithmetic");
  int grid = grid_for(total4, kStepThreads, kNumSMs * 8);
  l2reg_pass1<<<grid, kStepThreads, 0, st>>>((const float4*)x_t, (const float4*)eps, (const float4*)z,
                                             (const float4*)mask, (const float4*)x_ref,
                                             (float4*)x_prev, (float4*)x0_pred, (double*)workspace,
                                             total4, C * HW / 4, k);
  rc = check_launch("l2reg_pass1");
  if (rc) return rc;
  if (p->base.guide) {
    l2reg_pass2<<<grid, kStepThreads, 0, st>>>((float4*)x_prev, (const float4*)eps,
                                               (const float4*)mask, (const float4*)x_ref,
                                               (const double*)workspace, grid, total4, C * HW / 4,
                                               HW / 4, (int)C, k);
    rc = check_launch("l2reg_pass2");
  }
  return rc;
}

int b2e_apply_guidance_grad_f32(float* x, const float* neg_grad, const float* mask, int64_t B,
                                int64_t CHW, int mask_batched, float a_t_sq, void* stream) {
  B2E_REQUIRE(x && neg_grad && B > 0 && CHW > 0, B2E_INVALID_ARG, "apply_guidance_grad: bad argument");
  apply_grad_kernel<<<grid_for(B * CHW, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
      x, neg_grad, mask, B * CHW, CHW, mask_batched, a_t_sq);
  return check_launch("apply_guidance_grad");
}

size_t b2e_color_loss_grad_workspace_bytes(void) { return sizeof(double) * kNumSMs * 8; }

int b2e_color_loss_grad_f32(const float* img, const float* mask, const float* x_ref, float* d_img, int64_t B, int64_t C,
                            int64_t HW, const b2e_color_grad_params* p, void* workspace, size_t workspace_bytes,
                            void* stream) {
  B2E_REQUIRE(img && d_img && p && B > 0 && C > 0 && C <= 4 && HW > 0, B2E_INVALID_ARG, "color_loss_grad: bad argument");
  B2E_REQUIRE(HW % 4 == 0 && aligned16(img) && aligned16(d_img) && (!mask || aligned16(mask)) && (!x_ref || aligned16(x_ref)),
              B2E_UNSUPPORTED_SHAPE, "color_loss_grad: needs HW %% 4 == 0 and 16-byte aligned tensors");
  B2E_REQUIRE(!p->use_mask_pred || (mask && x_ref && workspace && workspace_bytes >= b2e_color_loss_grad_workspace_bytes()),
              B2E_INVALID_ARG, "color_loss_grad: the masked + L2-regularised variant needs mask, x_ref and a workspace");
  ColorGradK k;
  for (int i = 0; i < 4; ++i) { k.has_target[i] = p->has_target[i]; k.target[i] = p->target[i]; k.k[i] = p->k[i]; }
  k.lam_scale = p->lam_scale; k.use_mask_pred = p->use_mask_pred; k.mask_batched = p->mask_batched;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total4 = B * C * HW / 4, chw4 = C * HW / 4;
  B2E_REQUIRE(total4 < (int64_t)0x7fffffff, B2E_UNSUPPORTED_SHAPE, "color_loss_grad: tensor too large for 32-bit index arithmetic");
  const int grid = grid_for(total4 / 2 + 1, kStepThreads, kNumSMs * 8);
  if (p->use_mask_pred) {
    color_grad_norm_kernel<<<grid, kStepThreads, 0, st>>>((const float4*)img, (const float4*)mask, (const float4*)x_ref,
                                                          (double*)workspace, total4, chw4, p->mask_batched);
    int rc = check_launch("color_grad_norm");
    if (rc) return rc;
  }
  color_grad_kernel<<<grid, kStepThreads, 0, st>>>((const float4*)img, (const float4*)mask, (const float4*)x_ref, (float4*)d_img,
                                                   (const double*)workspace, grid, total4, chw4, HW / 4, (int)C, k);
  return check_launch("color_grad");
}

int b2e_apply_latent_guidance_f32(float* x, const float* d_latent, const float* mask, int64_t B, int64_t CHW,
                                  int mask_batched, float chain, float sqrt_a_t, float a_t_sq, void* stream) {
  B2E_REQUIRE(x && d_latent && B > 0 && CHW > 0, B2E_INVALID_ARG, "apply_latent_guidance: bad argument");
  latent_guidance_kernel<<<grid_for(B * CHW, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
      x, d_latent, mask, B * CHW, CHW, mask_batched, chain, sqrt_a_t, a_t_sq);
  return check_launch("apply_latent_guidance");
}

int b2e_pred_x0_f32(const float* x_t, const float* eps, float* x0, int64_t n, float sqrt_a_t,
                    float sqrt_b_t, void* stream) {
  return launch_map2<OP_PRED_X0>(x_t, eps, x0, n, MapK{sqrt_a_t, sqrt_b_t, 0.f, 0.f}, stream, "pred_x0");
}

int b2e_renoise_f32(const float* x, const float* eps, float* out, int64_t n, float c_a, float c_b,
                    float c_out_x0, float c_out_e, void* stream) {
  return launch_map2<OP_RENOISE>(x, eps, out, n, MapK{c_a, c_b, c_out_x0, c_out_e}, stream, "renoise");
}

int b2e_cfg_combine_f32(const float* e_first, const float* e_second, float* out, int64_t n,
                        float scale, void* stream) {
  return launch_map2<OP_CFG>(e_first, e_second, out, n, MapK{scale, 0.f, 0.f, 0.f}, stream, "cfg_combine");
}

int b2e_apply_mask_f32(const float* mask, const float* zo, const float* zv, float* out, int64_t T,
                       int64_t CHW, void* stream) {
  B2E_REQUIRE(mask && zo && zv && out && T > 0 && CHW > 0, B2E_INVALID_ARG, "apply_mask: bad argument");
  const int vec = CHW % 4 == 0 && aligned16(mask) && aligned16(zo) && aligned16(zv) && aligned16(out);
  apply_mask_kernel<<<grid_for(vec ? T * CHW / 8 : T * CHW, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
      mask, zo, zv, out, T * CHW, CHW, vec);
  return check_launch("apply_mask");
}

int b2e_to_uint8_f32(const float* x, uint8_t* out, int64_t B, int64_t C, int64_t HW, void* stream) {
  B2E_REQUIRE(x && out && B > 0 && C > 0 && HW > 0, B2E_INVALID_ARG, "to_uint8: bad argument");
  const int vec = C == 3 && HW % 4 == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(out) & 3) == 0;
  to_uint8_kernel<<<grid_for(vec ? B * HW / 8 : B * HW, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
      x, out, B, (int)C, HW, vec);
  return check_launch("to_uint8");
}

int b2e_sample_xts_f32(const float* x0, const float* noise, const float* sa, const float* sb,
                       float* xts, int64_t T, int64_t CHW, void* stream) {
  B2E_REQUIRE(x0 && noise && sa && sb && xts && T > 0 && CHW > 0, B2E_INVALID_ARG,
              "sample_xts: bad argument");
  dim3 grid(grid_for(CHW, kStepThreads, 1024), (unsigned)(T + 1));
  sample_xts_kernel<<<grid, kStepThreads, 0, (cudaStream_t)stream>>>(x0, noise, sa, sb, xts, T, CHW);
  return check_launch("sample_xts");
}

int b2e_extract_noise_f32(const float* x_t, const float* eps, float* x_tm1, float* z, int64_t n,
                          const b2e_step_coeffs* c, void* stream) {
  B2E_REQUIRE(x_t && eps && x_tm1 && z && c && n > 0, B2E_INVALID_ARG, "extract_noise: bad argument");
  const int vec = aligned16(x_t) && aligned16(eps) && aligned16(x_tm1) && aligned16(z) && n >= 4;
  extract_noise_kernel<<<grid_for(vec ? n / 4 : n, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(
      x_t, eps, x_tm1, z, n, c->sqrt_a_t, c->sqrt_b_t, c->sqrt_a_prev, c->dir_coef, c->sigma, vec);
  return check_launch("extract_noise");
}

int b2e_axpby_f32(const float* x, const float* y, float* out, int64_t n, float a, float b, void* stream) {
  B2E_REQUIRE(x && y && out && n > 0, B2E_INVALID_ARG, "axpby: bad argument");
  axpby_kernel<<<grid_for(n, kStepThreads), kStepThreads, 0, (cudaStream_t)stream>>>(x, y, out, n, a, b);
  return check_launch("axpby");
}

size_t b2e_loss_workspace_bytes(void) { return sizeof(double) * 4 * kNumSMs * 4; }

static int run_loss(const float* x, const float* y, int64_t n, int64_t hw, int C, int mode, const float* tg,
                    double count, float* out, void* ws, size_t ws_bytes, void* stream, const char* what) {
  B2E_REQUIRE(x && out && ws && n > 0, B2E_INVALID_ARG, "%s: bad argument", what);
  B2E_REQUIRE(ws_bytes >= b2e_loss_workspace_bytes(), B2E_WORKSPACE_TOO_SMALL, "%s: workspace too small", what);
  const int grid = grid_for(n, kStepThreads, kNumSMs * 4);
  loss_partial_kernel<<<grid, kStepThreads, 0, (cudaStream_t)stream>>>(x, y, (double*)ws, n, hw, C, mode, tg[0],
                                                                         tg[1], tg[2], tg[3]);
  int rc = check_launch(what);
  if (rc) return rc;
  loss_final_kernel<<<1, kStepThreads, 0, (cudaStream_t)stream>>>((const double*)ws, grid, out, mode, count);
  return check_launch(what);
}

int b2e_l2_distance_f32(const float* x, const float* y, int64_t n, float* out, void* workspace,
                        size_t workspace_bytes, void* stream) {
  B2E_REQUIRE(y, B2E_INVALID_ARG, "l2_distance: bad argument");
  const float tg[4] = {0, 0, 0, 0};
  return run_loss(x, y, n, 1, 1, 0, tg, 1.0, out, workspace, workspace_bytes, stream, "l2_distance");
}

int b2e_channel_l1_f32(const float* img, int64_t B, int64_t C, int64_t HW, const float* targets4, float* out4,
                       void* workspace, size_t workspace_bytes, void* stream) {
  B2E_REQUIRE(targets4 && C > 0 && C <= 4 && B > 0 && HW > 0, B2E_INVALID_ARG, "channel_l1: bad argument");
  return run_loss(img, nullptr, B * C * HW, HW, (int)C, 1, targets4, (double)(B * HW), out4, workspace,
                  workspace_bytes, stream, "channel_l1");
}

// Host scalar math, reference op order (this TU is compiled with -ffp-contract=off).
int b2e_step_coeffs_compute(const float* ac, int num_train, float final_alpha_cumprod, int t,
                            int t_prev, float eta, int mode, b2e_step_coeffs* out) {
  B2E_REQUIRE(ac && out && t >= 0 && t < num_train && t_prev < num_train, B2E_INVALID_ARG,
              "step_coeffs: bad timestep %d / %d", t, t_prev);
  volatile float a_t = ac[t];
  volatile float a_p = t_prev >= 0 ? ac[t_prev] : final_alpha_cumprod;
  volatile float b_t = 1.f - a_t;
  volatile float b_p = 1.f - a_p;
  volatile float q = a_t / a_p;
  volatile float om = 1.f - q;
  volatile float ratio = b_p / b_t;
  volatile float var = ratio * om;
  volatile float sd = sqrtf(var);
  volatile float sigma = eta * sd;
  volatile float inner;
  if (mode == B2E_MODE_DDIM) {
    volatile float s2 = sigma * sigma;
    inner = b_p - s2;
  } else {
    volatile float ev = eta * var;
    inner = b_p - ev;
  }
  out->sqrt_a_t = sqrtf(a_t);
  out->sqrt_b_t = sqrtf(b_t);
  out->sqrt_a_prev = sqrtf(a_p);
  out->dir_coef = sqrtf(inner);
  out->sigma = sigma;
  out->a_t_sq = a_t * a_t;
  out->variance = var;
  out->a_t = a_t;
  out->a_prev = a_p;
  return B2E_OK;
}

}  // extern "C"
