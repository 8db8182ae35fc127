// Loss heads of the network-based guidance functions: the part of NetAttrFunc / ClassifierAttrFunc that is not a
// dense network.  Each head returns the loss value AND its analytic
// gradient with respect to the network's output, so the caller seeds the network's backward pass with it
// (grad_output) instead of running softmax / sum / index autograd nodes.
//
//   NetAttrFunc.loss        src/attr_functions.py:213-219
//       p = softmax_c(logits[0]) ; area_c = sum_hw p[c] / (256*256) ; L = sum_{c in S} area_c
//       dL/dlogits[c,hw] = p[c,hw] * (1[c in S] - sum_{k in S} p[k,hw]) / (256*256)
//   ClassifierAttrFunc.loss src/attr_functions.py:237-257
//       a = logits.view(-1,40,2)[0] ; L = a[i][j] (+ (a[r][p] + score[p])^2)
//       dL/dlogits = onehot(2i+j) (+ 2 (a[r][p] + score[p]) onehot(2r+p)), other batch rows 0
#include "common.cuh"

namespace b2e {

constexpr int kHeadThreads = 256;
constexpr int kSegMaxClasses = 32;

// one thread per pixel; logits are [C][HW] (class-major), so every per-class access is coalesced
__global__ void __launch_bounds__(kHeadThreads)
seg_area_head_kernel(const float* __restrict__ logits, float* __restrict__ dlogits, double* __restrict__ partial,
                     int C, int64_t HW, uint32_t class_mask, float inv_div) {
  pdl_wait();
  __shared__ double s_red[kHeadThreads / 32];
  double acc = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    float v[kSegMaxClasses];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < kSegMaxClasses; ++c)
      if (c < C) { v[c] = __ldcs(logits + (int64_t)c * HW + p); m = fmaxf(m, v[c]); }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kSegMaxClasses; ++c)
      if (c < C) { v[c] = expf(__fsub_rn(v[c], m)); s = __fadd_rn(s, v[c]); }
    float q = 0.f;   // probability mass of the selected classes at this pixel
#pragma unroll
    for (int c = 0; c < kSegMaxClasses; ++c)
      if (c < C) { v[c] = __fdiv_rn(v[c], s); if ((class_mask >> c) & 1u) q = __fadd_rn(q, v[c]); }
    acc += (double)q;
    if (dlogits) {
#pragma unroll
      for (int c = 0; c < kSegMaxClasses; ++c)
        if (c < C) {
          const float ind = ((class_mask >> c) & 1u) ? 1.f : 0.f;
          __stcs(dlogits + (int64_t)c * HW + p, __fmul_rn(__fmul_rn(v[c], __fsub_rn(ind, q)), inv_div));
        }
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kHeadThreads / 32; ++w) t += s_red[w];
    partial[blockIdx.x] = t;
  }
}

// fixed-order final sum of the per-block partials (deterministic)
__global__ void seg_area_final_kernel(const double* __restrict__ partial, int n, float inv_div, float* __restrict__ loss) {
  pdl_wait();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += partial[i];
    *loss = (float)(t * (double)inv_div);
  }
}

__global__ void classifier_head_kernel(const float* __restrict__ logits, int64_t n, int idx, int ridx, float rscore,
                                       float* __restrict__ loss, float* __restrict__ dlogits) {
  pdl_wait();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float reg_t = 0.f;
  if (ridx >= 0) reg_t = __fadd_rn(logits[ridx], rscore);
  if (i < n && dlogits) {
    float g = 0.f;
    if (i == idx) g = 1.f;
    if (ridx >= 0 && i == ridx) g = __fadd_rn(g, __fmul_rn(2.f, reg_t));
    dlogits[i] = g;
  }
  if (i == 0 && loss) {
    float v = logits[idx];
    if (ridx >= 0) v = __fadd_rn(v, __fmul_rn(reg_t, reg_t));
    *loss = v;
  }
}

}  // namespace b2e

using namespace b2e;

extern "C" {

size_t b2e_seg_area_head_workspace_bytes(void) { return sizeof(double) * kNumSMs * 8; }

int b2e_seg_area_head_f32(const float* logits, int64_t C, int64_t HW, const int32_t* classes, int n_classes,
                          float area_divisor, float* loss, float* dlogits, void* workspace, size_t workspace_bytes,
                          void* stream) {
  B2E_REQUIRE(logits && loss && classes, B2E_INVALID_ARG, "seg_area_head: null pointer");
  B2E_REQUIRE(C >= 1 && C <= kSegMaxClasses, B2E_UNSUPPORTED_SHAPE, "seg_area_head: %lld classes (max %d)", (long long)C,
              kSegMaxClasses);
  B2E_REQUIRE(HW >= 1 && n_classes >= 0 && area_divisor > 0.f, B2E_INVALID_ARG, "seg_area_head: bad sizes");
  B2E_REQUIRE(workspace && workspace_bytes >= b2e_seg_area_head_workspace_bytes(), B2E_WORKSPACE_TOO_SMALL,
              "seg_area_head: workspace too small");
  uint32_t mask = 0;
  for (int i = 0; i < n_classes; ++i) {
    // a class listed twice counts twice in the reference (fancy indexing); refuse instead of silently differing
    B2E_REQUIRE(classes[i] >= 0 && classes[i] < C, B2E_INVALID_ARG, "seg_area_head: class id %d out of range", classes[i]);
    B2E_REQUIRE(!((mask >> classes[i]) & 1u), B2E_UNSUPPORTED_SHAPE, "seg_area_head: class id %d listed twice", classes[i]);
    mask |= 1u << classes[i];
  }
  cudaStream_t st = (cudaStream_t)stream;
  int grid = (int)((HW + kHeadThreads - 1) / kHeadThreads);
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  const float inv_div = 1.f / area_divisor;   // 65536 in the reference: exact
  launch_pdl(seg_area_head_kernel, dim3(grid), dim3(kHeadThreads), 0, st, logits, dlogits, (double*)workspace, (int)C, HW,
             mask, inv_div);
  int rc = check_launch("seg_area_head");
  if (rc) return rc;
  launch_pdl(seg_area_final_kernel, dim3(1), dim3(32), 0, st, (const double*)workspace, grid, inv_div, loss);
  return check_launch("seg_area_final");
}

int b2e_classifier_head_f32(const float* logits, int64_t n, int idx_for_class, int idx_of_interest, int reg_idx,
                            int reg_pred, float reg_score, float* loss, float* dlogits, void* stream) {
  B2E_REQUIRE(logits && (loss || dlogits), B2E_INVALID_ARG, "classifier_head: null pointer");
  B2E_REQUIRE(n >= 80 && n % 80 == 0, B2E_UNSUPPORTED_SHAPE, "classifier_head: expected B x 80 logits, got %lld", (long long)n);
  B2E_REQUIRE(idx_for_class >= 0 && idx_for_class < 40 && (idx_of_interest == 0 || idx_of_interest == 1), B2E_INVALID_ARG,
              "classifier_head: index out of range");
  B2E_REQUIRE(reg_idx < 40 && (reg_idx < 0 || reg_pred == 0 || reg_pred == 1), B2E_INVALID_ARG,
              "classifier_head: regulariser index out of range");
  const int idx = 2 * idx_for_class + idx_of_interest, ridx = reg_idx >= 0 ? 2 * reg_idx + reg_pred : -1;
  launch_pdl(classifier_head_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, logits, n, idx,
             ridx, reg_score, loss, dlogits);
  return check_launch("classifier_head");
}

}  // extern "C"
