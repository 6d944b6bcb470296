// Mapping MLP + truncation + AdaIN affine ("styles") in fp32.
//   mapping: PixelNorm (networks_stylegan.py:558-565) then 8 x [DenseW 512->512 + LeakyReLU(0.2)] (:128-139)
//   truncation lerp w_l = avg*(1-psi_l) + w*psi_l (:158-163, :180-189) folded into the operand load of
//   the affine layers (:244,:252), which are evaluated for all style layers in one launch.
// 4.2 + 10.4 MFLOP per sample: latency/bandwidth-bound on the fp32 weights, so CUDA cores, one warp per
// output unit, weights in registers, the (transformed) activations of up to 32 samples in shared memory.
#include "gsx_internal.h"
#include "ptx.cuh"

namespace gsx {

static constexpr int kDnThreads = 256;       // 8 warps -> 8 output units per block
static constexpr int kDnSamples = 32;        // samples staged per pass
static constexpr int kDnMaxK = 512;

__global__ void __launch_bounds__(kDnThreads) dense_kernel(const DenseArgs a) {
  pdl_launch_dependents();
  extern __shared__ float xs[];              // [kDnSamples][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * 8 + warp;
  const int K = a.K;
  float wr[kDnMaxK / 32];
#pragma unroll
  for (int j = 0; j < kDnMaxK / 32; ++j) wr[j] = (u < a.U && lane + 32 * j < K) ? a.W[(size_t)u * K + lane + 32 * j] : 0.f;
  const float bu = (u < a.U && a.b) ? a.b[u] : 0.f;
  pdl_wait();                                  // x (and a per-call psi) come from earlier work in the stream
  float psi = 1.f;
  if (a.psi) psi = a.psi[a.unit_layer[blockIdx.x * 8]];

  for (int nbase = 0; nbase < a.N; nbase += kDnSamples) {
    const int ns = min(kDnSamples, a.N - nbase);
    __syncthreads();
    for (int i = threadIdx.x; i < ns * K; i += kDnThreads) {
      const int n = i / K, k = i - n * K;
      float v = a.x[(size_t)(nbase + n) * K + k];
      if (a.psi) v = a.latent_avg[k] * (1.f - psi) + v * psi;
      xs[i] = v;
    }
    __syncthreads();
    if (a.pixelnorm) {
      // one warp per sample: x * rsqrt(mean(x^2) + 1e-8)
      for (int n = warp; n < ns; n += 8) {
        float ss = 0.f;
        for (int k = lane; k < K; k += 32) ss += xs[n * K + k] * xs[n * K + k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float r = rsqrtf(ss / (float)K + 1e-8f);
        for (int k = lane; k < K; k += 32) xs[n * K + k] *= r;
      }
      __syncthreads();
    }
    if (u < a.U) {
      for (int n = 0; n < ns; ++n) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < kDnMaxK / 32; ++j) acc = fmaf(wr[j], (lane + 32 * j < K) ? xs[n * K + lane + 32 * j] : 0.f, acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) {
          float y = acc + bu;
          if (a.lrelu) y = y > 0.f ? y : 0.2f * y;
          a.y[(size_t)(nbase + n) * a.U + u] = y;
        }
      }
    }
  }
}

void launch_dense(const DenseArgs& a, cudaStream_t st) {
  static bool configured[64] = {false};        // the attribute is per device
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 0 ? 0 : (dev > 63 ? 63 : dev);
  const size_t smem = (size_t)kDnSamples * a.K * sizeof(float);
  if (!configured[dev]) {
    cudaFuncSetAttribute(dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    configured[dev] = true;
  }
  launch_pdl(dense_kernel, dim3((a.U + 7) / 8), dim3(kDnThreads), smem, st, a);
}

}  // namespace gsx
