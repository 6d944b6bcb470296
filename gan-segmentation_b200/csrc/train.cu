// Building blocks of decoder training (reference seg_solver.py:351-465).  Round 1
// ships the loss and the optimizer step; the decoder backward pass (dgrad / wgrad / BatchNorm) is not built yet.
//   softmax_ce : SoftmaxCELoss(axis=1) with sample_weight = (mask > -1) (seg_solver.py:404-407): per-sample loss =
//                mean over ALL H*W pixels of -w * log_softmax(logits)[label] (ignored pixels stay in the denominator),
//                and its gradient w * (softmax - onehot) / (H*W).
//   adam_step  : MXNet Adam on one flat fp32 bucket (all decoder parameters): bias correction folded into the
//                learning rate, rescale_grad = 1/batch (trainer.step(batch), seg_solver.py:421), eps 1e-8.
//                Runs right after the single all-reduce of the flat gradient bucket.
#include "../../include/gsx.h"
#include "gsx_internal.h"

#include <atomic>
#include <cmath>

namespace gsx {
extern std::atomic<uint64_t> g_launches;

static constexpr int kCeThreads = 256;

// grid (blocks, N); each block reduces its pixels to one partial loss (fixed order), summed by a second tiny kernel
__global__ void __launch_bounds__(kCeThreads) softmax_ce_kernel(const float* __restrict__ logits, const int* __restrict__ labels,
                                                                float* __restrict__ dlogits, float* __restrict__ partial,
                                                                int K, int HW) {
  const int n = blockIdx.y;
  const float* lg = logits + (size_t)n * K * HW;
  float* dl = dlogits ? dlogits + (size_t)n * K * HW : nullptr;
  const int* lab = labels + (size_t)n * HW;
  const float inv_hw = 1.f / (float)HW;
  float acc = 0.f;
  for (int p = blockIdx.x * kCeThreads + threadIdx.x; p < HW; p += gridDim.x * kCeThreads) {
    const int l = lab[p];
    const float w = l > -1 ? 1.f : 0.f;                 // seg_solver.py:404 (l_w == 1, :405 is a no-op)
    const int lc = l < 0 ? 0 : (l >= K ? K - 1 : l);    // pick() clips the index
    float mx = lg[p];
    for (int k = 1; k < K; ++k) mx = fmaxf(mx, lg[(size_t)k * HW + p]);
    float se = 0.f;
    for (int k = 0; k < K; ++k) se += expf(lg[(size_t)k * HW + p] - mx);
    const float lse = mx + logf(se);
    acc += w * (lse - lg[(size_t)lc * HW + p]);
    if (dl) {
      for (int k = 0; k < K; ++k) {
        const float sm = expf(lg[(size_t)k * HW + p] - lse);
        dl[(size_t)k * HW + p] = w * (sm - (k == lc ? 1.f : 0.f)) * inv_hw;
      }
    }
  }
  __shared__ float red[kCeThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kCeThreads / 32; ++w) s += red[w];
    partial[(size_t)n * gridDim.x + blockIdx.x] = s * inv_hw;
  }
}

__global__ void ce_finish_kernel(const float* __restrict__ partial, float* __restrict__ loss, int blocks) {
  const int n = blockIdx.x;
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < blocks; ++b) s += partial[(size_t)n * blocks + b];
    loss[n] = s;
  }
}

__global__ void adam_step_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, size_t n, float lr_t, float beta1, float beta2, float eps,
                                 float wd, float rescale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * rescale + wd * w[i];
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    w[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace gsx

using namespace gsx;

extern "C" int gsx_softmax_ce(const float* logits_dev, const int* labels_dev, int n, int num_classes, int h, int w,
                              float* loss_dev, float* dlogits_dev, float* scratch_dev, size_t scratch_floats,
                              gsx_stream stream) {
  if (!logits_dev || !labels_dev || !loss_dev || !scratch_dev || n <= 0) { set_error("bad argument"); return -1; }
  const int HW = h * w;
  const int blocks = std::min(256, (HW + kCeThreads - 1) / kCeThreads);
  if (scratch_floats < (size_t)n * blocks) { set_error("scratch too small (needs n*256 floats)"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  softmax_ce_kernel<<<dim3(blocks, n), kCeThreads, 0, st>>>(logits_dev, labels_dev, dlogits_dev, scratch_dev, num_classes, HW);
  ce_finish_kernel<<<n, 32, 0, st>>>(scratch_dev, loss_dev, blocks);
  g_launches += 2;
  return cuda_ok(cudaGetLastError(), "softmax_ce") ? 0 : -2;
}

extern "C" int gsx_adam_step(float* w_dev, const float* g_dev, float* m_dev, float* v_dev, size_t count, int t, float lr,
                             float beta1, float beta2, float eps, float wd, float rescale_grad, gsx_stream stream) {
  if (!w_dev || !g_dev || !m_dev || !v_dev || t < 1) { set_error("bad argument"); return -1; }
  // MXNet Adam: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
  const double c1 = 1.0 - std::pow((double)beta1, t), c2 = 1.0 - std::pow((double)beta2, t);
  const float lr_t = (float)(lr * std::sqrt(c2) / c1);
  const int blocks = (int)std::min<size_t>(1184, (count + 255) / 256);
  adam_step_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(w_dev, g_dev, m_dev, v_dev, count, lr_t, beta1, beta2,
                                                                          eps, wd, rescale_grad);
  g_launches++;
  return cuda_ok(cudaGetLastError(), "adam_step") ? 0 : -2;
}
