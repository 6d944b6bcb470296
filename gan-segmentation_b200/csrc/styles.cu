// Mapping MLP + truncation + AdaIN affine ("styles") in fp32.
//   mapping: PixelNorm (networks_stylegan.py:558-565) then 8 x [DenseW 512->512 + LeakyReLU(0.2)] (:128-139)
//   truncation lerp w_l = avg*(1-psi_l) + w*psi_l (:158-163, :180-189) folded into the operand load of
//   the affine layers (:244,:252), which are evaluated for all style layers in one launch.
// 4.2 + 10.4 MFLOP per sample: latency/bandwidth-bound on the fp32 weights, so CUDA cores, one warp per
// output unit, weights in registers, the (transformed) activations of up to 32 samples in shared memory.
#include "gsx_internal.h"
#include "ptx.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

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

// ------------------------------------------------------------------------------------------------------------------
// The whole mapping path in ONE cooperative launch: PixelNorm, the 8 DenseW + LeakyReLU layers, and the AdaIN affine of
// every style layer (truncation folded in).  The nine dependent launches of dense_kernel cost 0.28 ms per step for
// 0.5 GFLOP (r01: 64-block grids, ~23 us of latency each); here 128 blocks x 4 warps = one warp per mapping unit walk the
// layers with a grid barrier in between (the next layer's weight row is already in registers when the barrier opens),
// then share out the style units.   lerp is applied to the dot products: W.(avg(1-psi) + w psi) = (1-psi) W.avg + psi W.w.
// ------------------------------------------------------------------------------------------------------------------
static constexpr int kMapBlocks = 128, kMapWarps = 4;

__global__ void __launch_bounds__(kMapWarps * 32) mapping_kernel(const MapArgs a) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(128) float xs[];          // [kDnSamples][512]
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * kMapWarps + warp;          // mapping unit of this warp (512 warps in all)
  constexpr int K = kDnMaxK, KJ = K / 32;
  float wr[KJ];
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  uint32_t par = 0;
  // activations of up to 32 samples -> shared memory as ONE bulk copy (a 128-thread load loop took ~20 us per layer:
  // 128 dependent-latency rounds for 64 KB)
  auto stage = [&](const float* src, int ns) {
    __syncthreads();                                     // everybody is done with the previous contents
    if (threadIdx.x == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes (other CTAs' outputs, this CTA's smem reads) before the bulk copy
      mbar_expect_tx(&bar, (uint32_t)(ns * K * sizeof(float)));
      bulk_load(xs, src, (uint32_t)(ns * K * sizeof(float)), &bar);
    }
    mbar_wait(&bar, par);
    par ^= 1u;
  };
  for (int layer = 0; layer < 8; ++layer) {
#pragma unroll
    for (int j = 0; j < KJ; ++j) wr[j] = a.W[layer][(size_t)u * K + lane + 32 * j];
    const float bu = a.b[layer][u];
    if (layer > 0) grid.sync();                          // the previous layer's outputs are complete
    const float* x = layer == 0 ? a.z : (layer & 1 ? a.ya : a.yb);
    float* y = layer & 1 ? a.yb : a.ya;
    for (int nbase = 0; nbase < a.N; nbase += kDnSamples) {
      const int ns = min(kDnSamples, a.N - nbase);
      stage(x + (size_t)nbase * K, ns);
      if (layer == 0) {                                   // PixelNorm: x * rsqrt(mean(x^2) + 1e-8), one warp per sample
        for (int n = warp; n < ns; n += kMapWarps) {
          float ss = 0.f;
          for (int k = lane; k < K; k += 32) ss += xs[n * K + k] * xs[n * K + k];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
          const float r = rsqrtf(ss / (float)K + 1e-8f);
          for (int k = lane; k < K; k += 32) xs[n * K + k] *= r;
        }
        __syncthreads();
      }
      // 32 independent dot products (one per sample) per warp, then ONE transposing butterfly: lane n ends up with
      // sample n's sum (a dependent FMA chain + 5 shuffles per sample was ~200 cycles x 32 samples per layer)
      float vals[32];
#pragma unroll
      for (int n = 0; n < 32; ++n) {
        float acc = 0.f;
        if (n < ns) {
#pragma unroll
          for (int j = 0; j < KJ; ++j) acc = fmaf(wr[j], xs[n * K + lane + 32 * j], acc);
        }
        vals[n] = acc;
      }
      const float tot = warp_transpose_reduce32(vals, lane);
      if (lane < ns) {
        const float v = tot + bu;
        y[(size_t)(nbase + lane) * K + u] = v > 0.f ? v : 0.2f * v;
      }
    }
  }
  // styles: unit s of style layer l = affine_l(avg*(1-psi_l) + w*psi_l)    (networks_stylegan.py:180-189, :252)
  const float* wfin = a.yb;                                // layer 7 wrote yb
  const int total_warps = gridDim.x * kMapWarps;
  float av[KJ], wn[KJ];
#pragma unroll
  for (int j = 0; j < KJ; ++j) {
    av[j] = a.latent_avg[lane + 32 * j];
    wn[j] = u < a.S ? a.Waff[(size_t)u * K + lane + 32 * j] : 0.f;      // first style unit's row: in flight across the barrier
  }
  grid.sync();
  for (int nbase = 0; nbase < a.N; nbase += kDnSamples) {
    const int ns = min(kDnSamples, a.N - nbase);
    stage(wfin + (size_t)nbase * K, ns);
    for (int s = u; s < a.S; s += total_warps) {
      float dav = 0.f;
      const int s_next = s + total_warps < a.S ? s + total_warps : (nbase + kDnSamples < a.N ? u : s);
#pragma unroll
      for (int j = 0; j < KJ; ++j) {
        wr[j] = wn[j];
        wn[j] = a.Waff[(size_t)s_next * K + lane + 32 * j];            // next unit's row while this one is evaluated
        dav = fmaf(wr[j], av[j], dav);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dav += __shfl_xor_sync(0xffffffffu, dav, o);
      const float psi = a.psi[a.unit_layer[s]];
      const float base = dav * (1.f - psi) + a.baff[s];
      float vals[32];
#pragma unroll
      for (int n = 0; n < 32; ++n) {
        float acc = 0.f;
        if (n < ns) {
#pragma unroll
          for (int j = 0; j < KJ; ++j) acc = fmaf(wr[j], xs[n * K + lane + 32 * j], acc);
        }
        vals[n] = acc;
      }
      const float tot = warp_transpose_reduce32(vals, lane);
      if (lane < ns) a.styles[(size_t)(nbase + lane) * a.S + s] = fmaf(psi, tot, base);
    }
  }
}

bool launch_mapping(const MapArgs& a, cudaStream_t st) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev = dev < 0 ? 0 : (dev > 63 ? 63 : dev);
  const size_t smem = (size_t)kDnSamples * kDnMaxK * sizeof(float);
  if (!configured[dev]) {
    cudaFuncSetAttribute(mapping_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    configured[dev] = true;
  }
  void* args[] = {const_cast<MapArgs*>(&a)};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(mapping_kernel), dim3(kMapBlocks), dim3(kMapWarps * 32), args, smem, st) ==
         cudaSuccess;
}

}  // namespace gsx
