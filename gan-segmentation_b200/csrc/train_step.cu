// One decoder-training iteration behind ONE C-ABI call (reference seg_solver.py:386-421, generator frozen):
// train-mode forward (BatchNorm with batch statistics, LeakyReLU, Dropout(0.5) after every cvt block), SoftmaxCE with
// sample weight (mask > -1), hand-derived backward, gradients into one flat fp32 bucket (then: a single all-reduce and
// gsx_adam_step).  Everything stays resident in the caller's workspace as blocked 16-bit tensors; no host synchronisation,
// no library kernels:
//   forward convs / data gradients  shiftconv_kernel (tcgen05), the data gradient as the same conv with transposed, flipped
//                                    weights; 16-bit operand streams re-packed from the fp32 master weights every step by
//                                    pack_kernel (a gather through tables built once, plan.cpp: pack_conv_sources)
//   weight gradients                 wgrad_kernel (tcgen05 split-K GEMM over pixels, wgrad.cu) straight into the bucket
//   BatchNorm / LeakyReLU / Dropout  the kernels below, on the blocked layout; per-channel reductions are two-level with a
//                                    fixed order (bit-reproducible); dropout masks are Philox bits recomputed in backward
//   loss                             softmax_ce_blocked_kernel: loss + (H*W-scaled) gradient directly as a blocked tensor
// The loss gradient is scaled by H*W (see gsx_softmax_ce); every gradient in the bucket carries that factor.
#include "../../include/gsx.h"
#include "gsx_internal.h"
#include "ptx.cuh"

#include <atomic>
#include <cmath>
#include <cstring>
#include <map>

namespace gsx {
extern std::atomic<uint64_t> g_launches;
const char* last_error_cstr();

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ void t_unpack8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#if GSX_FP16
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
#else
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
#endif
    f[2 * k] = v.x; f[2 * k + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 t_pack8(const float (&f)[8]) {
  __align__(16) act_t o[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = to_act(f[k]);
  return *reinterpret_cast<const uint4*>(o);
}
// 8 dropout keep-bits for the 8 channels of (sample n, channel block cb, pixel p) of dropout site `site`
__device__ __forceinline__ uint32_t drop_bits(unsigned long long seed, int site, int n, int cb, int p) {
  uint32_t c[4] = {(uint32_t)p, (uint32_t)cb, (uint32_t)site, (uint32_t)n};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return c[0] & 0xFFu;
}

// block-wide fixed-order reduction of 16 per-thread values -> dst[(cb*8 + ch)*2 + which]
__device__ __forceinline__ void t_block_reduce16(float (&v)[16], float* dst, int nthreads) {
  __shared__ float red[32][17];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) red[warp][i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float s = 0.f;
    for (int w = 0; w < (nthreads >> 5); ++w) s += red[w][threadIdx.x];
    dst[(threadIdx.x & 7) * 2 + (threadIdx.x >> 3)] = s;
  }
}

// ------------------------------------------------------------------------------------------------ kernels
__global__ void pack_kernel(const float* __restrict__ w, const int4* __restrict__ idx, act_t* __restrict__ out, size_t count) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
    const int4 s = __ldg(idx + i);
    float v = 0.f;
    if (s.x >= 0) v += w[s.x];
    if (s.y >= 0) v += w[s.y];
    if (s.z >= 0) v += w[s.z];
    if (s.w >= 0) v += w[s.w];
    out[i] = to_act(v);
  }
}

// BatchNorm statistics from the per-(sample, tile) partial sums of launch_stats: one thread per channel.
// bnp[c] = {a = gamma*rstd, b = beta - mean*a, mean, rstd}; moving statistics <- 0.9*moving + 0.1*batch (biased variance).
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int NT, int C, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                                   float4* __restrict__ bnp) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s1 = 0.0, s2 = 0.0;
  for (int t = 0; t < NT; ++t) { s1 += partial[((size_t)t * C + c) * 2]; s2 += partial[((size_t)t * C + c) * 2 + 1]; }
  const double mean = s1 / count, var = fmax(s2 / count - mean * mean, 0.0);
  const float rstd = (float)(1.0 / sqrt(var + 1e-5));
  const float a = gamma[c] * rstd;
  bnp[c] = make_float4(a, beta[c] - (float)mean * a, (float)mean, rstd);
  rmean[c] = 0.9f * rmean[c] + 0.1f * (float)mean;
  rvar[c] = 0.9f * rvar[c] + 0.1f * (float)var;
}

// "last block finishes" helper for the two-level per-channel reductions: every block publishes its partial, takes a ticket;
// the block that draws the last ticket sums all partials in their fixed order (so the result does not depend on which block
// that is) and resets the counter.  Saves one tiny dependent launch per reduction (~75 per training step).
__device__ __forceinline__ bool t_last_block(unsigned int* counter) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y;
    s_last = atomicAdd(counter, 1u) == total - 1;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

// Fixed-order sums of NT partial pairs per channel by the whole (256-thread) block: thread (part, c) adds the pairs part,
// part + parts, ...; the parts are then combined in order.  f(c, s0, s1) runs in the thread that owns channel c.  (One thread
// per channel walking all NT pairs left 16 threads with 512 dependent-latency rounds: 45 of the 52 us of a 1024^2 launch.)
template <class F>
__device__ __forceinline__ void t_sum_partials(const float* partial, int NT, int C, F f) {
  __shared__ double sh[2][256];
  for (int c0 = 0; c0 < C; c0 += 256) {
    const int cw = min(256, C - c0), parts = 256 / cw;
    const int c = c0 + (int)threadIdx.x % cw, part = (int)threadIdx.x / cw;
    double s0 = 0.0, s1 = 0.0;
    if (part < parts)
      for (int t = part; t < NT; t += parts) {
        const float2 v = __ldcg(reinterpret_cast<const float2*>(partial + ((size_t)t * C + c) * 2));
        s0 += v.x; s1 += v.y;
      }
    sh[0][threadIdx.x] = s0; sh[1][threadIdx.x] = s1;
    __syncthreads();
    if (part == 0) {
      for (int q = 1; q < parts; ++q) { s0 += sh[0][q * cw + (c - c0)]; s1 += sh[1][q * cw + (c - c0)]; }
      f(c, s0, s1);
    }
    __syncthreads();
  }
}

struct BnStatArgs {
  const act_t* z; float* partial; unsigned int* counter;
  int C, N, HW;
  const float* gamma; const float* beta; float* rmean; float* rvar; float4* bnp;     // gamma == null: channel sums only -> out
  float* out; int c_real;
};
// per-(sample, tile) channel sums / sums of squares of a blocked tensor; the last block turns them into the BatchNorm
// coefficients bnp[c] = {a = gamma*rstd, b = beta - mean*a, mean, rstd} and updates the moving statistics (0.9 / 0.1, biased
// variance) -- or, without gamma, writes the channel sums (bias gradients)
__global__ void __launch_bounds__(256) bn_stats_kernel(const BnStatArgs a) {
  const int plane = blockIdx.y, cb = plane / a.N, n = plane - cb * a.N;
  const act_t* src = a.z + (size_t)plane * a.HW * 8;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < a.HW; p += gridDim.x * 256) {
    float f[8];
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(src + (size_t)p * 8)), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] += f[i]; acc[8 + i] = fmaf(f[i], f[i], acc[8 + i]); }
  }
  const int T = gridDim.x;
  t_block_reduce16(acc, a.partial + (((size_t)n * T + blockIdx.x) * a.C + cb * 8) * 2, 256);
  if (!t_last_block(a.counter)) return;
  const int NT = a.N * T;
  const double count = (double)a.N * a.HW;
  t_sum_partials(a.partial, NT, a.C, [&](int c, double s1, double s2) {
    if (!a.gamma) { if (c < a.c_real) a.out[c] = (float)s1; return; }
    const double mean = s1 / count, var = fmax(s2 / count - mean * mean, 0.0);
    const float rstd = (float)(1.0 / sqrt(var + 1e-5));
    const float ca = a.gamma[c] * rstd;
    a.bnp[c] = make_float4(ca, a.beta[c] - (float)mean * ca, (float)mean, rstd);
    a.rmean[c] = 0.9f * a.rmean[c] + 0.1f * (float)mean;
    a.rvar[c] = 0.9f * a.rvar[c] + 0.1f * (float)var;
  });
  if (threadIdx.x == 0) *a.counter = 0;
}

struct BnFwdArgs {
  const act_t* z; act_t* y; const float4* bnp;
  int C, N, H, W;
  int drop_site; unsigned long long seed;      // drop_site < 0: no dropout
  const unsigned long long* seed_dev;          // non-null: the seed is read from device memory (CUDA-graph replays)
  const act_t* addsrc;                         // blocked [C/8][N][H/2][W/2][8], nearest-upsampled and added, or null
};
// y = Dropout(LeakyReLU(BN(z))) (+ up2(addsrc)): one pass over the blocked tensor
__global__ void __launch_bounds__(256) bn_fwd_kernel(const BnFwdArgs a) {
  const int plane = blockIdx.y, cb = plane / a.N, n = plane - cb * a.N;
  const int HW = a.H * a.W;
  float ca[8], cc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float4 v = a.bnp[cb * 8 + i]; ca[i] = v.x; cc[i] = v.y; }
  const act_t* z = a.z + (size_t)plane * HW * 8;
  act_t* y = a.y + (size_t)plane * HW * 8;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
    float f[8];
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(z + (size_t)p * 8)), f);
    uint32_t keep = 0xFFu;
    if (a.drop_site >= 0) keep = drop_bits(a.seed_dev ? *a.seed_dev : a.seed, a.drop_site, n, cb, p);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float pre = fmaf(f[i], ca[i], cc[i]);
      float v = fmaxf(pre, 0.2f * pre);
      if (a.drop_site >= 0) v = ((keep >> i) & 1u) ? 2.f * v : 0.f;
      f[i] = v;
    }
    if (a.addsrc) {
      const int py = p / a.W, px = p - py * a.W;
      float r[8];
      t_unpack8(__ldg(reinterpret_cast<const uint4*>(a.addsrc + ((size_t)plane * (HW >> 2) + (size_t)(py >> 1) * (a.W >> 1) + (px >> 1)) * 8)), r);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] += r[i];
    }
    *reinterpret_cast<uint4*>(y + (size_t)p * 8) = t_pack8(f);
  }
}

struct BnBwdArgs {
  const act_t* z; const act_t* dy; act_t* dz; const float4* bnp; const float2* dparam;   // dparam[c] = {dbeta, dgamma}
  float* partial;                                                                        // [N*T][C][2]
  int C, N, HW, T;
  int drop_site; unsigned long long seed;
  float m;
  const unsigned long long* seed_dev;
  unsigned int* counter; float* dgamma_out; float* dbeta_out; float2* dparam_w;      // last-block finalize of the reduction
};
// g = dy * dropout' * lrelu'(pre);  partial sums of g (-> dbeta) and g * xhat (-> dgamma) per (sample, block)
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnBwdArgs a) {
  const int plane = blockIdx.y, cb = plane / a.N, n = plane - cb * a.N;
  float ca[8], cc[8], mu[8], rs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float4 v = a.bnp[cb * 8 + i]; ca[i] = v.x; cc[i] = v.y; mu[i] = v.z; rs[i] = v.w; }
  const act_t* z = a.z + (size_t)plane * a.HW * 8;
  const act_t* dy = a.dy + (size_t)plane * a.HW * 8;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < a.HW; p += gridDim.x * 256) {
    float f[8], d[8];
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(z + (size_t)p * 8)), f);
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (size_t)p * 8)), d);
    uint32_t keep = 0xFFu;
    if (a.drop_site >= 0) keep = drop_bits(a.seed_dev ? *a.seed_dev : a.seed, a.drop_site, n, cb, p);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float pre = fmaf(f[i], ca[i], cc[i]);
      float g = d[i] * (pre > 0.f ? 1.f : 0.2f);
      if (a.drop_site >= 0) g = ((keep >> i) & 1u) ? 2.f * g : 0.f;
      acc[i] += g;
      acc[8 + i] = fmaf(g, (f[i] - mu[i]) * rs[i], acc[8 + i]);
    }
  }
  t_block_reduce16(acc, a.partial + (((size_t)n * a.T + blockIdx.x) * a.C + cb * 8) * 2, 256);
  if (!t_last_block(a.counter)) return;
  // dbeta / dgamma = fixed-order sums of the partials -> the gradient bucket and dparam (read by bn_bwd_apply_kernel)
  const int NT = a.N * a.T;
  t_sum_partials(a.partial, NT, a.C, [&](int c, double s0, double s1) {
    a.dbeta_out[c] = (float)s0; a.dgamma_out[c] = (float)s1;
    a.dparam_w[c] = make_float2((float)s0, (float)s1);
  });
  if (threadIdx.x == 0) *a.counter = 0;
}
// (stand-alone form of the finalize, kept for reference / tests)
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ partial, int NT, int C, float* __restrict__ dgamma_out,
                                       float* __restrict__ dbeta_out, float2* __restrict__ dparam) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int t = 0; t < NT; ++t) { s0 += partial[((size_t)t * C + c) * 2]; s1 += partial[((size_t)t * C + c) * 2 + 1]; }
  dbeta_out[c] = (float)s0; dgamma_out[c] = (float)s1;
  dparam[c] = make_float2((float)s0, (float)s1);
}
// dz = gamma*rstd/m * (m*g - dbeta - xhat*dgamma)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdArgs a) {
  const int plane = blockIdx.y, cb = plane / a.N, n = plane - cb * a.N;
  float ca[8], cc[8], mu[8], rs[8], db[8], dg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 v = a.bnp[cb * 8 + i]; ca[i] = v.x; cc[i] = v.y; mu[i] = v.z; rs[i] = v.w;
    const float2 d = a.dparam[cb * 8 + i]; db[i] = d.x; dg[i] = d.y;
  }
  const act_t* z = a.z + (size_t)plane * a.HW * 8;
  const act_t* dy = a.dy + (size_t)plane * a.HW * 8;
  act_t* dz = a.dz + (size_t)plane * a.HW * 8;
  const float inv_m = 1.f / a.m;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < a.HW; p += gridDim.x * 256) {
    float f[8], d[8];
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(z + (size_t)p * 8)), f);
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(dy + (size_t)p * 8)), d);
    uint32_t keep = 0xFFu;
    if (a.drop_site >= 0) keep = drop_bits(a.seed_dev ? *a.seed_dev : a.seed, a.drop_site, n, cb, p);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float pre = fmaf(f[i], ca[i], cc[i]);
      float g = d[i] * (pre > 0.f ? 1.f : 0.2f);
      if (a.drop_site >= 0) g = ((keep >> i) & 1u) ? 2.f * g : 0.f;
      const float xh = (f[i] - mu[i]) * rs[i];
      f[i] = ca[i] * inv_m * (a.m * g - db[i] - xh * dg[i]);
    }
    *reinterpret_cast<uint4*>(dz + (size_t)p * 8) = t_pack8(f);
  }
}

// SoftmaxCE (sample weight = label > -1, mean over ALL pixels) on fp32 NCHW logits; the gradient, times grad_scale, leaves
// as a blocked 16-channel tensor (classes in channels 0..K-1, the rest zero) for the data / weight gradient kernels.
__global__ void __launch_bounds__(256) softmax_ce_blocked_kernel(const float* __restrict__ logits, const int* __restrict__ labels,
                                                                 act_t* __restrict__ dl, float* __restrict__ partial, int N, int K,
                                                                 int HW, float grad_scale) {
  const int n = blockIdx.y;
  const float* lg = logits + (size_t)n * K * HW;
  const int* lab = labels + (size_t)n * HW;
  const float inv_hw = 1.f / (float)HW, gs = grad_scale * inv_hw;
  float acc = 0.f;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
    const int l = lab[p];
    const float w = l > -1 ? 1.f : 0.f;
    const int lc = l < 0 ? 0 : (l >= K ? K - 1 : l);
    float v[16];
    float mx = -3.4e38f;
#pragma unroll
    for (int k = 0; k < 16; ++k) { v[k] = k < K ? lg[(size_t)k * HW + p] : -3.4e38f; mx = fmaxf(mx, v[k]); }
    float se = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) se += k < K ? expf(v[k] - mx) : 0.f;
    const float lse = mx + logf(se);
    float picked = 0.f;
    float g[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k == lc) picked = v[k];
      g[k] = k < K ? w * (expf(v[k] - lse) - (k == lc ? 1.f : 0.f)) * gs : 0.f;
    }
    acc += w * (lse - picked);
    float g0[8], g1[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { g0[k] = g[k]; g1[k] = g[8 + k]; }
    *reinterpret_cast<uint4*>(dl + (((size_t)0 * N + n) * HW + p) * 8) = t_pack8(g0);
    *reinterpret_cast<uint4*>(dl + (((size_t)1 * N + n) * HW + p) * 8) = t_pack8(g1);
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    partial[(size_t)n * gridDim.x + blockIdx.x] = s * inv_hw;
  }
}
__global__ void ce_sum_kernel(const float* __restrict__ partial, float* __restrict__ loss, int blocks) {
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int b = 0; b < blocks; ++b) s += partial[(size_t)blockIdx.x * blocks + b];
    loss[blockIdx.x] = s;
  }
}

// out[cb][n][y][x] = sum of the 2x2 block of in (+ addend): adjoint of the nearest-x2 upsampling
__global__ void __launch_bounds__(256) sumpool2_blocked_kernel(const act_t* __restrict__ in, const act_t* __restrict__ addend,
                                                               act_t* __restrict__ out, int planes, int H, int W) {
  const int plane = blockIdx.y;
  const int HW = H * W;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW; p += gridDim.x * 256) {
    const int y = p / W, x = p - y * W;
    const act_t* src = in + ((size_t)plane * 4 * HW + (size_t)(2 * y) * (2 * W) + 2 * x) * 8;
    float a[8], b[8], s[8];
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(src)), a);
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(src + 8)), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = a[i] + b[i];
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(src + (size_t)2 * W * 8)), a);
    t_unpack8(__ldg(reinterpret_cast<const uint4*>(src + (size_t)2 * W * 8 + 8)), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] += a[i] + b[i];
    if (addend) {
      t_unpack8(__ldg(reinterpret_cast<const uint4*>(addend + ((size_t)plane * HW + p) * 8)), a);
#pragma unroll
      for (int i = 0; i < 8; ++i) s[i] += a[i];
    }
    *reinterpret_cast<uint4*>(out + ((size_t)plane * HW + p) * 8) = t_pack8(s);
  }
  (void)planes;
}
// nearest-x2 upsampling of a blocked tensor (the input of the up-conv's weight gradient)
__global__ void __launch_bounds__(256) upsample2_blocked_kernel(const act_t* __restrict__ in, act_t* __restrict__ out, int H, int W) {
  const int plane = blockIdx.y;
  const int HW4 = 4 * H * W;
  for (int p = blockIdx.x * 256 + threadIdx.x; p < HW4; p += gridDim.x * 256) {
    const int y = p / (2 * W), x = p - y * (2 * W);
    *reinterpret_cast<uint4*>(out + ((size_t)plane * HW4 + p) * 8) =
        __ldg(reinterpret_cast<const uint4*>(in + ((size_t)plane * H * W + (size_t)(y >> 1) * W + (x >> 1)) * 8));
  }
}
// bias gradient: sum over samples and tiles of the per-(sample, tile) channel sums of launch_stats
__global__ void chan_sum_finalize_kernel(const float* __restrict__ partial, int NT, int C, int c_real, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= c_real) return;
  double s = 0.0;
  for (int t = 0; t < NT; ++t) s += partial[((size_t)t * C + c) * 2];
  out[c] = (float)s;
}
__global__ void zero_kernel(float* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = 0.f;
}
__global__ void dropout_mask_kernel(float* out, int N, int C, int HW, int site, unsigned long long seed) {
  const size_t total = (size_t)N * C * HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % HW), c = (int)((i / HW) % C), n = (int)(i / ((size_t)HW * C));
    out[i] = (float)((drop_bits(seed, site, n, c >> 3, p) >> (c & 7)) & 1u);
  }
}

static int ew_grid(int HW) { return std::max(1, std::min((HW + 255) / 256, 2048)); }
// tiles of a two-level per-channel reduction: ~4 blocks per SM over all planes, at least 1024 pixels per block
static int t_tiles(int HW, int planes) { return std::max(1, std::min(std::min((HW + 1023) / 1024, 512), (1184 + planes - 1) / planes)); }

}  // namespace gsx

using namespace gsx;

// =================================================================================================================
namespace {

struct TConv {
  ConvLayer fwd, dgrad;
  bool has_dgrad = false;
  int mode = CONV3, k = 3, cin0 = 0, cin1 = 0, cout = 0, cout_pad = 0, H = 0, W = 0;    // H, W: input resolution
  size_t w_off = 0, b_off = 0;                 // offsets (floats) in the parameter / gradient bucket
  size_t pk_fwd = 0, pk_dgrad = 0;             // offsets (elements) in the packed-operand buffer
};
struct TBn {
  int C = 0;
  size_t gamma_off = 0, beta_off = 0, rmean_off = 0, rvar_off = 0;
  float4* bnp = nullptr;
  float2* dparam = nullptr;
};
struct TLevel {
  int H, W, cin, f, fnext, c0, c1;
  bool has_sc = false, last = false;
  TConv cvt, conv_a, conv_b, sc, fin;
  TBn bn_cvt, bn_a, bn_b;
};

}  // namespace

struct gsx_train {
  gsx_dec_cfg cfg;
  int n = 1, nf = 0, K = 2;
  bool use_dropout = true;
  std::vector<TLevel> levels;
  std::map<std::string, std::pair<size_t, size_t>> layout;      // reference name -> (offset, count)
  std::vector<std::string> order;
  size_t n_learn = 0, n_total = 0;
  act_t* wpack_all = nullptr;
  int4* pack_idx = nullptr;
  size_t pack_count = 0;
  float* bn_mem = nullptr;
  unsigned int* counter = nullptr;                   // ticket counters of the last-block-finishes reductions, one per branch (zero between launches)
  // Branches of one step (gsx_train_step): the cvt block of every level (forward, and its BatchNorm backward) on its own
  // stream, the weight gradients on two more; joined to the caller's stream by events, so that a CUDA-graph capture of the
  // step records the dependency graph instead of one chain.
  std::vector<cudaStream_t> lvl_streams;
  cudaStream_t wg_streams[2] = {nullptr, nullptr};
  cudaStream_t sc_stream = nullptr;                  // the 1x1 shortcut: forward conv, and its gradient path (sum-pool + data gradient)
  std::vector<cudaEvent_t> events;
  const unsigned long long* seed_dev = nullptr;     // gsx_train_set_seed_buffer
  int sms = 148;
};

namespace {

struct TWs {
  std::vector<act_t*> feat, z_cvt, y_cvt, z_a, y_a, z_b, prev, sc;
  std::vector<act_t*> dzc;                                      // per level: the cvt branch runs beside the main chain
  act_t *upx, *g_out[2], *g1, *g2, *g3, *g_up, *g_sc, *g_in1, *dlog;
  std::vector<float*> stats;                                    // per branch (index = branch id, see gsx_train_step)
  float *wg_scratch[2], *logits, *ce_partial;
  size_t total;
};

size_t t_align(size_t v) { return (v + 1023) / 1024 * 1024; }

TWs train_layout(const gsx_train* h, void* base, bool own_feats) {
  TWs w;
  uint8_t* b = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = b + off; off = t_align(off + bytes); return p; };
  const int nf = h->nf, N = h->n;
  size_t max_hi = 0, max_lo = 0, max_up = 0, max_wg = 0;
  for (int i = 0; i < nf; ++i) {
    const TLevel& l = h->levels[i];
    const size_t px = (size_t)N * l.H * l.W;
    w.feat.push_back(own_feats ? reinterpret_cast<act_t*>(take(px * l.cin * 2)) : nullptr);
    w.z_cvt.push_back(reinterpret_cast<act_t*>(take(px * l.f * 2)));
    w.y_cvt.push_back(reinterpret_cast<act_t*>(take(px * l.f * 2)));
    w.dzc.push_back(reinterpret_cast<act_t*>(take(px * l.f * 2)));
    if (!l.last) {
      w.z_a.push_back(reinterpret_cast<act_t*>(take(px * 4 * l.fnext * 2)));
      w.y_a.push_back(reinterpret_cast<act_t*>(take(px * 4 * l.fnext * 2)));
      w.z_b.push_back(reinterpret_cast<act_t*>(take(px * 4 * l.fnext * 2)));
      w.prev.push_back(reinterpret_cast<act_t*>(take(px * 4 * l.fnext * 2)));
      w.sc.push_back(l.has_sc ? reinterpret_cast<act_t*>(take(px * l.fnext * 2)) : nullptr);
      max_hi = std::max(max_hi, px * 4 * (size_t)std::max(l.fnext, 16));
      max_up = std::max(max_up, px * 4 * (size_t)(l.c0 + l.c1));
      max_lo = std::max(max_lo, px * (size_t)std::max(l.c0 + l.c1, l.fnext));
    } else {
      max_lo = std::max(max_lo, px * (size_t)std::max(l.c0 + l.c1, 16));
    }
    auto wg = [&](const TConv& c) { if (c.cout) max_wg = std::max(max_wg, wgrad_scratch_floats(c.k, std::max(c.cin0, c.cin1 ? c.cin0 + c.cin1 : c.cin0), c.cout_pad, h->sms)); };
    wg(l.cvt); wg(l.conv_a); wg(l.conv_b); wg(l.sc); wg(l.fin);
  }
  w.upx = reinterpret_cast<act_t*>(take(max_up * 2));
  w.g_up = reinterpret_cast<act_t*>(take(max_up * 2));
  for (int i = 0; i < 2; ++i) w.g_out[i] = reinterpret_cast<act_t*>(take(std::max(max_lo, max_hi) * 2));
  w.g1 = reinterpret_cast<act_t*>(take(max_hi * 2));
  w.g2 = reinterpret_cast<act_t*>(take(max_hi * 2));
  w.g3 = reinterpret_cast<act_t*>(take(max_hi * 2));
  w.g_sc = reinterpret_cast<act_t*>(take(max_lo * 2));
  w.g_in1 = reinterpret_cast<act_t*>(take(max_lo * 2));
  const TLevel& top = h->levels[nf - 1];
  w.dlog = reinterpret_cast<act_t*>(take((size_t)N * top.H * top.W * 16 * 2));
  w.logits = reinterpret_cast<float*>(take((size_t)N * h->K * top.H * top.W * 4));
  // reduction partials: one block writes 16 floats, a launch has at most 1184 + (C/8) N <= 1184 + 64 N blocks (t_tiles)
  for (int b = 0; b < nf + 3; ++b) w.stats.push_back(reinterpret_cast<float*>(take((size_t)(1184 + 64 * N) * 16 * 4 * 2)));
  w.ce_partial = reinterpret_cast<float*>(take((size_t)N * 256 * 4));
  for (int i = 0; i < 2; ++i) w.wg_scratch[i] = reinterpret_cast<float*>(take(max_wg * 4));
  w.total = off;
  return w;
}

void add_param(gsx_train* h, const std::string& name, size_t count, size_t* off_out, bool learnable) {
  (void)learnable;
  h->layout[name] = {h->n_total, count};
  h->order.push_back(name);
  *off_out = h->n_total;
  h->n_total += count;
}

bool plan_tconv(TConv& c, int mode, int k, int H, int W, int cin0, int cin1, int cout, int argmax, bool need_dgrad) {
  c.mode = mode; c.k = k; c.H = H; c.W = W; c.cin0 = cin0; c.cin1 = cin1; c.cout = cout;
  c.cout_pad = (cout + 15) / 16 * 16;
  set_error("");
  plan_conv(c.fwd, mode, H, W, cin0, cin1, cout, argmax, nullptr);
  if (*gsx_last_error()) return false;
  c.has_dgrad = need_dgrad;
  if (need_dgrad) {
    const int Ho = mode == UPCONV3 ? 2 * H : H, Wo = mode == UPCONV3 ? 2 * W : W;
    plan_conv(c.dgrad, k == 3 ? CONV3 : CONV1, Ho, Wo, c.cout_pad, 0, cin0 + cin1, 0, nullptr);
    if (*gsx_last_error()) return false;
  }
  return true;
}

}  // namespace

extern "C" int gsx_train_create(const gsx_dec_cfg* cfg, int n, int use_dropout, gsx_train** out) {
  if (!cfg || !out || n <= 0 || cfg->num_levels < 1 || cfg->num_levels > 16) { set_error("bad argument"); return -1; }
  if (!cfg->use_bn) { set_error("gsx_train: use_bn = 0 is not built (the reference config trains with BatchNorm)"); return -1; }
  int dev = 0;
  if (!cuda_ok(cudaGetDevice(&dev), "cudaGetDevice (no CPU fallback)")) return -2;
  gsx_train* h = new gsx_train();
  h->cfg = *cfg; h->n = n; h->nf = cfg->num_levels; h->K = cfg->features[cfg->num_levels]; h->use_dropout = use_dropout != 0;
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, dev);
  const int nf = h->nf;
  // ---- layers and the layout of the flat parameter bucket (learnable first, then the BatchNorm moving statistics)
  for (int i = 0; i < nf; ++i) {
    TLevel l{};
    l.H = cfg->base_y << i; l.W = cfg->base_x << i;
    l.cin = cfg->in_channels[i]; l.f = cfg->features[i]; l.fnext = cfg->features[i + 1];
    l.c0 = l.f; l.c1 = i > 0 ? l.f : 0;
    l.last = (i == nf - 1);
    const std::string cv = "cvt_block_" + std::to_string(i);
    if (!plan_tconv(l.cvt, CONV3, 3, l.H, l.W, l.cin, 0, l.f, 0, false)) { delete h; return -1; }
    add_param(h, cv + ".0.weight", (size_t)l.f * l.cin * 9, &l.cvt.w_off, true);
    add_param(h, cv + ".0.bias", l.f, &l.cvt.b_off, true);
    l.bn_cvt.C = l.f;
    add_param(h, cv + ".1.gamma", l.f, &l.bn_cvt.gamma_off, true);
    add_param(h, cv + ".1.beta", l.f, &l.bn_cvt.beta_off, true);
    const int cm = l.c0 + l.c1;
    if (!l.last) {
      const std::string mb = "main_block_" + std::to_string(i) + ".1";
      if (!plan_tconv(l.conv_a, UPCONV3, 3, l.H, l.W, l.c0, l.c1, l.fnext, 0, true)) { delete h; return -1; }
      add_param(h, mb + ".base_layers.0.weight", (size_t)l.fnext * cm * 9, &l.conv_a.w_off, true);
      add_param(h, mb + ".base_layers.0.bias", l.fnext, &l.conv_a.b_off, true);
      l.bn_a.C = l.fnext;
      add_param(h, mb + ".base_layers.1.gamma", l.fnext, &l.bn_a.gamma_off, true);
      add_param(h, mb + ".base_layers.1.beta", l.fnext, &l.bn_a.beta_off, true);
      if (!plan_tconv(l.conv_b, CONV3, 3, 2 * l.H, 2 * l.W, l.fnext, 0, l.fnext, 0, true)) { delete h; return -1; }
      add_param(h, mb + ".base_layers.3.weight", (size_t)l.fnext * l.fnext * 9, &l.conv_b.w_off, true);
      add_param(h, mb + ".base_layers.3.bias", l.fnext, &l.conv_b.b_off, true);
      l.bn_b.C = l.fnext;
      add_param(h, mb + ".base_layers.4.gamma", l.fnext, &l.bn_b.gamma_off, true);
      add_param(h, mb + ".base_layers.4.beta", l.fnext, &l.bn_b.beta_off, true);
      l.has_sc = (l.fnext != cm);
      if (l.has_sc) {
        if (!plan_tconv(l.sc, CONV1, 1, l.H, l.W, l.c0, l.c1, l.fnext, 0, true)) { delete h; return -1; }
        add_param(h, mb + ".shortcut.0.weight", (size_t)l.fnext * cm, &l.sc.w_off, true);
        add_param(h, mb + ".shortcut.0.bias", l.fnext, &l.sc.b_off, true);
      } else if (l.c1 > 0) { set_error("identity shortcut over a concatenated input is not supported"); delete h; return -1; }
    } else {
      const std::string mb = "main_block_" + std::to_string(i) + ".0";
      if (!plan_tconv(l.fin, CONV3, 3, l.H, l.W, l.c0, l.c1, l.fnext, l.fnext, i > 0)) { delete h; return -1; }
      add_param(h, mb + ".weight", (size_t)l.fnext * cm * 9, &l.fin.w_off, true);
      add_param(h, mb + ".bias", l.fnext, &l.fin.b_off, true);
    }
    h->levels.push_back(l);
  }
  h->n_learn = h->n_total;
  for (int i = 0; i < nf; ++i) {
    TLevel& l = h->levels[i];
    const std::string cv = "cvt_block_" + std::to_string(i);
    add_param(h, cv + ".1.running_mean", l.f, &l.bn_cvt.rmean_off, false);
    add_param(h, cv + ".1.running_var", l.f, &l.bn_cvt.rvar_off, false);
    if (!l.last) {
      const std::string mb = "main_block_" + std::to_string(i) + ".1";
      add_param(h, mb + ".base_layers.1.running_mean", l.fnext, &l.bn_a.rmean_off, false);
      add_param(h, mb + ".base_layers.1.running_var", l.fnext, &l.bn_a.rvar_off, false);
      add_param(h, mb + ".base_layers.4.running_mean", l.fnext, &l.bn_b.rmean_off, false);
      add_param(h, mb + ".base_layers.4.running_var", l.fnext, &l.bn_b.rvar_off, false);
    }
  }
  // ---- gather tables of every packed operand stream (forward and data-gradient weights)
  std::vector<int4> idx_all;
  auto add_pack = [&](TConv& c) -> bool {
    if (!c.cout) return true;
    std::vector<int> src;
    if (!pack_conv_sources(c.fwd, src)) { set_error("train: unsupported conv mode"); return false; }
    c.pk_fwd = idx_all.size();
    c.fwd.wpack_elems = src.size() / 4;
    for (size_t e = 0; e < src.size() / 4; ++e) {
      int4 v = make_int4(src[4 * e], src[4 * e + 1], src[4 * e + 2], src[4 * e + 3]);
      int* pv = &v.x;
      for (int j = 0; j < 4; ++j) if (pv[j] >= 0) pv[j] += (int)c.w_off;
      idx_all.push_back(v);
    }
    if (c.has_dgrad) {
      // data gradient = the same kind of conv with Wd[ci][co][ky][kx] = W[co][ci][k-1-ky][k-1-kx]; the gradient's channels
      // are padded to a multiple of 16 (zero weights there)
      if (!pack_conv_sources(c.dgrad, src)) { set_error("train: unsupported dgrad mode"); return false; }
      c.pk_dgrad = idx_all.size();
      c.dgrad.wpack_elems = src.size() / 4;
      const int cin = c.cin0 + c.cin1, k = c.k, kk = k * k;
      for (size_t e = 0; e < src.size() / 4; ++e) {
        int s = src[4 * e];                           // plain CONV3 / CONV1: one term
        int4 v = make_int4(-1, -1, -1, -1);
        if (s >= 0) {
          const int t = s % kk, cof = (s / kk) % c.cout_pad, cif = s / (kk * c.cout_pad);
          const int ky = t / k, kx = t % k;
          if (cof < c.cout) v.x = (int)c.w_off + ((cof * cin + cif) * k + (k - 1 - ky)) * k + (k - 1 - kx);
        }
        idx_all.push_back(v);
      }
    }
    return true;
  };
  for (auto& l : h->levels) {
    if (!add_pack(l.cvt) || !add_pack(l.conv_a) || !add_pack(l.conv_b) || !add_pack(l.sc) || !add_pack(l.fin)) { delete h; return -1; }
  }
  h->pack_count = idx_all.size();
  bool ok = cuda_ok(cudaMalloc(&h->pack_idx, idx_all.size() * sizeof(int4)), "cudaMalloc") &&
            cuda_ok(cudaMemcpy(h->pack_idx, idx_all.data(), idx_all.size() * sizeof(int4), cudaMemcpyHostToDevice), "H2D") &&
            cuda_ok(cudaMalloc(&h->wpack_all, idx_all.size() * sizeof(act_t)), "cudaMalloc");
  // per-BatchNorm coefficient / gradient scratch + tap tables
  size_t bn_floats = 0;
  for (auto& l : h->levels) bn_floats += 6 * (size_t)(l.bn_cvt.C + l.bn_a.C + l.bn_b.C);
  ok = ok && cuda_ok(cudaMalloc(&h->bn_mem, std::max<size_t>(bn_floats, 1) * sizeof(float)), "cudaMalloc");
  const int n_branches = h->nf + 3;                  // caller's stream, two weight-gradient streams, one per level
  ok = ok && cuda_ok(cudaMalloc(&h->counter, n_branches * sizeof(unsigned int)), "cudaMalloc") &&
       cuda_ok(cudaMemset(h->counter, 0, n_branches * sizeof(unsigned int)), "memset");
  h->lvl_streams.resize(h->nf, nullptr);
  for (int i = 0; ok && i < h->nf; ++i) ok = cuda_ok(cudaStreamCreateWithFlags(&h->lvl_streams[i], cudaStreamNonBlocking), "stream");
  for (int i = 0; ok && i < 2; ++i) ok = cuda_ok(cudaStreamCreateWithFlags(&h->wg_streams[i], cudaStreamNonBlocking), "stream");
  ok = ok && cuda_ok(cudaStreamCreateWithFlags(&h->sc_stream, cudaStreamNonBlocking), "stream");
  h->events.resize((size_t)h->nf * 20 + 16, nullptr);       // created up front: nothing is allocated while a step is being captured
  for (size_t i = 0; ok && i < h->events.size(); ++i) ok = cuda_ok(cudaEventCreateWithFlags(&h->events[i], cudaEventDisableTiming), "event");
  if (!ok) { delete h; return -2; }
  float* bm = h->bn_mem;
  auto bn_take = [&](TBn& b) { if (!b.C) return; b.bnp = reinterpret_cast<float4*>(bm); bm += 4 * b.C; b.dparam = reinterpret_cast<float2*>(bm); bm += 2 * b.C; };
  auto finish = [&](TConv& c) -> bool {
    if (!c.cout) return true;
    c.fwd.wpack_dev = h->wpack_all + c.pk_fwd;
    std::vector<int4> taps(4 * kMaxSlots);
    build_tap_table(c.fwd.g, taps.data());
    if (!cuda_ok(cudaMalloc(&c.fwd.taps_dev, taps.size() * sizeof(int4)), "cudaMalloc")) return false;
    cudaMemcpy(c.fwd.taps_dev, taps.data(), taps.size() * sizeof(int4), cudaMemcpyHostToDevice);
    if (c.has_dgrad) {
      c.dgrad.wpack_dev = h->wpack_all + c.pk_dgrad;
      build_tap_table(c.dgrad.g, taps.data());
      if (!cuda_ok(cudaMalloc(&c.dgrad.taps_dev, taps.size() * sizeof(int4)), "cudaMalloc")) return false;
      cudaMemcpy(c.dgrad.taps_dev, taps.data(), taps.size() * sizeof(int4), cudaMemcpyHostToDevice);
    }
    return true;
  };
  for (auto& l : h->levels) {
    bn_take(l.bn_cvt); bn_take(l.bn_a); bn_take(l.bn_b);
    if (!finish(l.cvt) || !finish(l.conv_a) || !finish(l.conv_b) || !finish(l.sc) || !finish(l.fin)) { delete h; return -2; }
  }
  set_error("");
  *out = h;
  return 0;
}

extern "C" void gsx_train_destroy(gsx_train* h) {
  if (!h) return;
  for (auto& l : h->levels)
    for (TConv* c : {&l.cvt, &l.conv_a, &l.conv_b, &l.sc, &l.fin}) { cudaFree(c->fwd.taps_dev); cudaFree(c->dgrad.taps_dev); }
  cudaFree(h->wpack_all); cudaFree(h->pack_idx); cudaFree(h->bn_mem); cudaFree(h->counter);
  for (cudaStream_t s : h->lvl_streams) if (s) cudaStreamDestroy(s);
  for (cudaStream_t s : h->wg_streams) if (s) cudaStreamDestroy(s);
  if (h->sc_stream) cudaStreamDestroy(h->sc_stream);
  for (cudaEvent_t e : h->events) if (e) cudaEventDestroy(e);
  delete h;
}

extern "C" int gsx_train_param_count(const gsx_train* h, size_t* learnable, size_t* total) {
  if (!h) { set_error("null handle"); return -1; }
  if (learnable) *learnable = h->n_learn;
  if (total) *total = h->n_total;
  return (int)h->order.size();
}
extern "C" int gsx_train_param_info(const gsx_train* h, int index, const char** name, size_t* offset, size_t* count) {
  if (!h || index < 0 || index >= (int)h->order.size()) { set_error("bad index"); return -1; }
  const auto& e = h->layout.at(h->order[index]);
  *name = h->order[index].c_str(); *offset = e.first; *count = e.second;
  return 0;
}
extern "C" int gsx_train_set_seed_buffer(gsx_train* h, const uint64_t* seed_dev) {
  if (!h) { set_error("null handle"); return -1; }
  h->seed_dev = reinterpret_cast<const unsigned long long*>(seed_dev);
  return 0;
}
extern "C" int gsx_train_workspace_bytes(const gsx_train* h, size_t* bytes) {
  if (!h || !bytes) { set_error("bad argument"); return -1; }
  *bytes = train_layout(h, nullptr, true).total;
  return 0;
}
extern "C" int gsx_train_dropout_mask(const gsx_train* h, int level, uint64_t seed, float* out_dev, gsx_stream stream) {
  if (!h || level < 0 || level >= h->nf || !out_dev) { set_error("bad argument"); return -1; }
  const TLevel& l = h->levels[level];
  dropout_mask_kernel<<<256, 256, 0, static_cast<cudaStream_t>(stream)>>>(out_dev, h->n, l.f, l.H * l.W, level, seed);
  return cuda_ok(cudaGetLastError(), "dropout_mask") ? 0 : -2;
}

namespace {

bool t_conv_fwd(const TConv& c, int N, const act_t* x0, const act_t* x1, act_t* out, const float* bias, cudaStream_t st, const char* label) {
  ConvEpi e{};
  const bool up = c.mode == UPCONV3;
  e.out = out; e.Ho = up ? 2 * c.H : c.H; e.Wo = up ? 2 * c.W : c.W; e.up = up ? 1 : 0; e.flags = 0; e.Cout = c.cout; e.bias = bias;
  return run_conv_layer(c.fwd, N, x0, x1, e, st, label);
}
bool t_conv_dgrad(const TConv& c, int N, const act_t* dy, act_t* dx, cudaStream_t st, const char* label) {
  ConvEpi e{};
  const bool up = c.mode == UPCONV3;
  e.out = dx; e.Ho = up ? 2 * c.H : c.H; e.Wo = up ? 2 * c.W : c.W; e.flags = 0; e.Cout = c.cin0 + c.cin1;
  return run_conv_layer(c.dgrad, N, dy, nullptr, e, st, label);
}
// A branch of the step: its stream and its own reduction scratch + ticket counter (branches run concurrently).
struct TBranch { cudaStream_t st; float* stats; unsigned int* counter; };

void t_bn_stats(const gsx_train* h, const TBn& b, const act_t* z, int HW, const float* p, float* r, const TBranch& br) {
  // enough blocks to fill the GPU at batch 1: (C/8)*N planes x T tiles; the last block finalizes
  const int T = t_tiles(HW, (b.C / 8) * h->n);
  cudaStream_t st = br.st; float* stats = br.stats;
  BnStatArgs a{z, stats, br.counter, b.C, h->n, HW, p + b.gamma_off, p + b.beta_off, r + b.rmean_off, r + b.rvar_off, b.bnp, nullptr, 0};
  bn_stats_kernel<<<dim3(T, (b.C / 8) * h->n), 256, 0, st>>>(a);
  g_launches++;
}
void t_bn_fwd(const gsx_train* h, const TBn& b, const act_t* z, act_t* y, int H, int W, int site, uint64_t seed, const act_t* addsrc, cudaStream_t st) {
  BnFwdArgs a{z, y, b.bnp, b.C, h->n, H, W, site, seed, h->seed_dev, addsrc};
  bn_fwd_kernel<<<dim3(ew_grid(H * W), (b.C / 8) * h->n), 256, 0, st>>>(a);
  g_launches++;
}
// dz (may alias dy) from dy, and dgamma / dbeta into the gradient bucket
void t_bn_bwd(const gsx_train* h, const TBn& b, const act_t* z, const act_t* dy, act_t* dz, int HW, int site, uint64_t seed, float* g,
              const TBranch& br) {
  const int T = t_tiles(HW, (b.C / 8) * h->n);
  cudaStream_t st = br.st; float* stats = br.stats;
  BnBwdArgs a{z, dy, dz, b.bnp, b.dparam, stats, b.C, h->n, HW, T, site, seed, (float)h->n * (float)HW, h->seed_dev,
              br.counter, g + b.gamma_off, g + b.beta_off, b.dparam};
  bn_bwd_reduce_kernel<<<dim3(T, (b.C / 8) * h->n), 256, 0, st>>>(a);
  bn_bwd_apply_kernel<<<dim3(ew_grid(HW), (b.C / 8) * h->n), 256, 0, st>>>(a);
  g_launches += 2;
}
bool t_wgrad(const gsx_train* h, const TConv& c, const act_t* x0, const act_t* x1, const act_t* dy, float* g, float* scratch, cudaStream_t st,
             int Hx, int Wx) {
  const int cin = c.cin0 + c.cin1;
  if (!launch_wgrad(c.k, h->n, Hx, Wx, c.cin0, c.cout_pad, c.cout, x0, dy, g + c.w_off, 0, cin, 1.f, scratch, st)) return false;
  if (c.cin1 && !launch_wgrad(c.k, h->n, Hx, Wx, c.cin1, c.cout_pad, c.cout, x1, dy, g + c.w_off, c.cin0, cin, 1.f, scratch, st)) return false;
  return true;
}
void t_bias_grad(const gsx_train* h, const act_t* dy, int Cpad, int Creal, int HW, float* out, const TBranch& br) {
  const int T = t_tiles(HW, (Cpad / 8) * h->n);
  cudaStream_t st = br.st; float* stats = br.stats;
  BnStatArgs a{dy, stats, br.counter, Cpad, h->n, HW, nullptr, nullptr, nullptr, nullptr, nullptr, out, Creal};
  bn_stats_kernel<<<dim3(T, (Cpad / 8) * h->n), 256, 0, st>>>(a);
  g_launches++;
}

}  // namespace

extern "C" int gsx_train_step(gsx_train* h, const float* params_dev, float* grads_dev, const float* const* feats_f32_dev,
                              const gsx_synth* synth, const void* synth_ws, const int* labels_dev, uint64_t dropout_seed,
                              float* loss_dev, uint8_t* pred_mask_dev, float* grad_scale_out, void* ws, size_t ws_bytes,
                              gsx_stream stream) {
  (void)synth; (void)synth_ws;
  if (!h || !params_dev || !grads_dev || !labels_dev || !loss_dev || !ws) { set_error("bad argument"); return -1; }
  if (!feats_f32_dev) { set_error("gsx_train_step: features must be given as fp32 NCHW device arrays"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nf = h->nf, N = h->n;
  TWs w = train_layout(h, ws, true);
  if (w.total > ws_bytes) { set_error("training workspace too small"); return -1; }
  pdl_set_for_work(-1);                                    // plain stream order between the kernels of the step
  const float* P = params_dev;
  float* R = const_cast<float*>(params_dev);               // the moving statistics live behind the learnable prefix
  float* G = grads_dev;
  const bool drop = h->use_dropout;
  set_error("");

  // Branches.  At batch 1 most of the ~300 kernels of a step fill a fraction of the GPU (a 512-channel cvt conv at 4^2..32^2
  // is ONE CTA for 25-80 us), and more than half of them are leaves of the dependency graph: the cvt block of a level needs
  // only that level's features, nothing but the optimizer waits for a weight gradient, and the cvt block's backward ends in
  // one.  So: the caller's stream runs the main chain (res-blocks forward, loss, data gradients), every level's cvt block
  // runs on its own stream, the weight gradients on two more; events join them.  Captured into a CUDA graph this is the
  // step's dependency DAG.  Each branch has its own reduction scratch, ticket counter and (weight gradients) partial buffer;
  // a scratch tensor that a side branch reads is not overwritten by the main chain before that read is done (ev_* below).
  const TBranch main_br{st, w.stats[0], h->counter};
  const TBranch wg_br[2] = {{h->wg_streams[0], w.stats[1], h->counter + 1}, {h->wg_streams[1], w.stats[2], h->counter + 2}};
  std::vector<TBranch> lvl_br;
  for (int i = 0; i < nf; ++i) lvl_br.push_back(TBranch{h->lvl_streams[i], w.stats[3 + i], h->counter + 3 + i});
  size_t next_event = 0;
  bool ev_ok = true;
  auto record = [&](cudaStream_t on) -> cudaEvent_t {
    if (next_event >= h->events.size()) { ev_ok = false; return nullptr; }
    cudaEvent_t e = h->events[next_event++];
    ev_ok = ev_ok && cudaEventRecord(e, on) == cudaSuccess;
    return e;
  };
  auto wait = [&](cudaStream_t on, cudaEvent_t e) { if (e) ev_ok = ev_ok && cudaStreamWaitEvent(on, e, 0) == cudaSuccess; };
  auto link = [&](cudaStream_t from, cudaStream_t to) { wait(to, record(from)); };

  // 16-bit operand streams of every conv (forward + data gradient) from the fp32 master weights
  pack_kernel<<<std::max(1, std::min((int)((h->pack_count + 255) / 256), 1184)), 256, 0, st>>>(P, h->pack_idx, h->wpack_all, h->pack_count);
  g_launches++;
  // gradient bucket: every entry is written exactly once below, except the biases in front of a BatchNorm (exactly zero)
  if (!cuda_ok(cudaMemsetAsync(G, 0, h->n_learn * sizeof(float), st), "zero grads")) return -2;
  {
    cudaEvent_t start = record(st);                        // operands packed, bucket zeroed, the caller's earlier work done
    for (int i = 0; i < nf; ++i) wait(lvl_br[i].st, start);
    for (int i = 0; i < 2; ++i) wait(wg_br[i].st, start);
    wait(h->sc_stream, start);
  }

  // ---------------------------------------------------------------- forward (train mode)
  for (int i = 0; i < nf; ++i) {                           // cvt blocks: one branch per level
    const TLevel& l = h->levels[i];
    const TBranch& br = lvl_br[i];
    const int HW = l.H * l.W;
    launch_nchw_to_blocked(feats_f32_dev[i], w.feat[i], l.cin, N, HW, br.st); g_launches++;
    if (!t_conv_fwd(l.cvt, N, w.feat[i], nullptr, w.z_cvt[i], P + l.cvt.b_off, br.st, "t.cvt")) return -2;
    t_bn_stats(h, l.bn_cvt, w.z_cvt[i], HW, P, R, br);
    t_bn_fwd(h, l.bn_cvt, w.z_cvt[i], w.y_cvt[i], l.H, l.W, drop ? i : -1, dropout_seed, nullptr, br.st);
  }
  for (int i = 0; i < nf; ++i) {                           // main chain
    const TLevel& l = h->levels[i];
    const int HW = l.H * l.W;
    link(lvl_br[i].st, st);                                // y_cvt[i]
    const act_t* x0 = i > 0 ? w.prev[i - 1] : w.y_cvt[i];
    const act_t* x1 = i > 0 ? w.y_cvt[i] : nullptr;
    if (!l.last) {
      const act_t* sc = x0;
      cudaEvent_t ev_scf = nullptr;
      if (l.has_sc) {                                      // the 1x1 shortcut runs beside conv_a -> BN -> conv_b -> BN
        link(st, h->sc_stream);                            // x0, x1
        if (!t_conv_fwd(l.sc, N, x0, x1, w.sc[i], P + l.sc.b_off, h->sc_stream, "t.shortcut")) return -2;
        ev_scf = record(h->sc_stream);
        sc = w.sc[i];
      }
      if (!t_conv_fwd(l.conv_a, N, x0, x1, w.z_a[i], P + l.conv_a.b_off, st, "t.conv_a")) return -2;
      t_bn_stats(h, l.bn_a, w.z_a[i], 4 * HW, P, R, main_br);
      t_bn_fwd(h, l.bn_a, w.z_a[i], w.y_a[i], 2 * l.H, 2 * l.W, -1, 0, nullptr, st);
      if (!t_conv_fwd(l.conv_b, N, w.y_a[i], nullptr, w.z_b[i], P + l.conv_b.b_off, st, "t.conv_b")) return -2;
      t_bn_stats(h, l.bn_b, w.z_b[i], 4 * HW, P, R, main_br);
      wait(st, ev_scf);
      // prev_{i+1} = up2(shortcut) + lrelu(BN(z_b))      (networks_seg.py:43-46, the 1x1 conv commutes with the upsampling)
      t_bn_fwd(h, l.bn_b, w.z_b[i], w.prev[i], 2 * l.H, 2 * l.W, -1, 0, sc, st);
    } else {
      ConvEpi e{};
      e.Ho = l.H; e.Wo = l.W; e.flags = EPI_ARGMAX; e.Cout = l.fnext; e.bias = P + l.fin.b_off;
      e.mask = pred_mask_dev ? pred_mask_dev : reinterpret_cast<unsigned char*>(w.dlog);     // (scratch when not wanted)
      e.logits = w.logits; e.num_classes = h->K;
      if (!run_conv_layer(l.fin.fwd, N, x0, x1, e, st, "t.final")) return -2;
    }
  }
  // ---------------------------------------------------------------- loss
  const TLevel& top = h->levels[nf - 1];
  const int HWt = top.H * top.W;
  const float gscale = (float)HWt;
  {
    const int blocks = std::min(256, (HWt + 255) / 256);
    softmax_ce_blocked_kernel<<<dim3(blocks, N), 256, 0, st>>>(w.logits, labels_dev, w.dlog, w.ce_partial, N, h->K, HWt, gscale);
    ce_sum_kernel<<<N, 32, 0, st>>>(w.ce_partial, loss_dev, blocks);
    g_launches += 2;
  }
  if (grad_scale_out) *grad_scale_out = gscale;

  // ---------------------------------------------------------------- backward
  // main chain: BatchNorm backward -> data gradient, level by level; weight gradients branch off to wg_br[0] (the res-block's
  // two convs) and wg_br[1] (shortcut, final conv, cvt), the cvt block's BatchNorm backward to the level's own stream.
  const act_t* d_prev_out = nullptr;        // gradient w.r.t. prev_{i+1} while level i is processed
  cudaEvent_t ev_g1 = nullptr, ev_g3 = nullptr, ev_gsc = nullptr;      // last side-branch read of the scratch tensors g1 / g3 / g_sc
  cudaEvent_t ev_sum_done = nullptr;                                   // the level's final sum-pool (reads g_sc / g_in1) is enqueued-complete
  std::vector<cudaEvent_t> ev_cvt_bwd(nf + 2, nullptr);               // level i's cvt branch has read its slice of g_out[i & 1]
  for (int i = nf - 1; i >= 0; --i) {
    const TLevel& l = h->levels[i];
    const int HW = l.H * l.W, cm = l.c0 + l.c1;
    const act_t* x0 = i > 0 ? w.prev[i - 1] : w.y_cvt[i];
    const act_t* x1 = i > 0 ? w.y_cvt[i] : nullptr;
    act_t* dxin = w.g_out[i & 1];           // [cm/8][N][H][W][8]: first c0 channels -> prev_i, rest -> cvt_i
    if (l.last) {
      if (i == 0) { set_error("single-level decoders are not supported by the training step"); return -1; }
      link(st, wg_br[1].st);                                                                            // dlog
      if (!t_wgrad(h, l.fin, x0, x1, w.dlog, G, w.wg_scratch[1], wg_br[1].st, l.H, l.W)) return -2;
      t_bias_grad(h, w.dlog, 16, l.fnext, HW, G + l.fin.b_off, wg_br[1]);
      if (!t_conv_dgrad(l.fin, N, w.dlog, dxin, st, "t.final.dgrad")) return -2;
    } else {
      // shortcut branch on its own stream: sum-pool of the incoming gradient, (shortcut conv: weight gradient on wg_br[1],) data gradient
      const act_t* addend = w.g_sc;
      cudaEvent_t ev_sc_done = nullptr;
      {
        cudaStream_t ss = h->sc_stream;
        link(st, ss);                                      // d_prev_out
        wait(ss, ev_gsc);                                  // the previous level's weight gradient has read g_sc
        if (!l.has_sc) wait(ss, ev_sum_done);              // ... and, without a shortcut conv, its final sum-pool reads g_sc directly
        sumpool2_blocked_kernel<<<dim3(ew_grid(HW), (l.fnext / 8) * N), 256, 0, ss>>>(d_prev_out, nullptr, w.g_sc, 0, l.H, l.W);
        g_launches++;
        if (l.has_sc) {
          link(ss, wg_br[1].st);
          if (!t_wgrad(h, l.sc, x0, x1, w.g_sc, G, w.wg_scratch[1], wg_br[1].st, l.H, l.W)) return -2;
          t_bias_grad(h, w.g_sc, l.sc.cout_pad, l.fnext, HW, G + l.sc.b_off, wg_br[1]);
          ev_gsc = record(wg_br[1].st);
          wait(ss, ev_sum_done);                           // the previous level's final sum-pool has read g_in1
          if (!t_conv_dgrad(l.sc, N, w.g_sc, w.g_in1, ss, "t.shortcut.dgrad")) return -2;
          addend = w.g_in1;
        }
        ev_sc_done = record(ss);
      }
      // second conv of the res-block
      wait(st, ev_g1);
      t_bn_bwd(h, l.bn_b, w.z_b[i], d_prev_out, w.g1, 4 * HW, -1, 0, G, main_br);                      // dz_b
      link(st, wg_br[0].st);
      if (!launch_wgrad(3, N, 2 * l.H, 2 * l.W, l.fnext, l.conv_b.cout_pad, l.fnext, w.y_a[i], w.g1, G + l.conv_b.w_off, 0, l.fnext, 1.f,
                        w.wg_scratch[0], wg_br[0].st)) return -2;
      ev_g1 = record(wg_br[0].st);
      // (bias in front of a BatchNorm: sum(dz) == 0 exactly -- the bucket was zeroed at the start of the step)
      if (!t_conv_dgrad(l.conv_b, N, w.g1, w.g2, st, "t.conv_b.dgrad")) return -2;                     // dy_a
      // first conv (nearest-x2 + 3x3)
      wait(st, ev_g3);
      t_bn_bwd(h, l.bn_a, w.z_a[i], w.g2, w.g3, 4 * HW, -1, 0, G, main_br);                            // dz_a
      link(st, wg_br[0].st);
      {
        cudaStream_t ws0 = wg_br[0].st;
        upsample2_blocked_kernel<<<dim3(ew_grid(4 * HW), (l.c0 / 8) * N), 256, 0, ws0>>>(x0, w.upx, l.H, l.W);
        if (l.c1) upsample2_blocked_kernel<<<dim3(ew_grid(4 * HW), (l.c1 / 8) * N), 256, 0, ws0>>>(x1, w.upx + (size_t)N * 4 * HW * l.c0, l.H, l.W);
        g_launches += 2;
        if (!launch_wgrad(3, N, 2 * l.H, 2 * l.W, cm, l.conv_a.cout_pad, l.fnext, w.upx, w.g3, G + l.conv_a.w_off, 0, cm, 1.f, w.wg_scratch[0], ws0))
          return -2;
        ev_g3 = record(ws0);
      }
      if (!t_conv_dgrad(l.conv_a, N, w.g3, w.g_up, st, "t.conv_a.dgrad")) return -2;                   // gradient at 2H x 2W
      wait(st, ev_sc_done);                 // the shortcut branch's gradient (addend)
      wait(st, ev_cvt_bwd[i + 2]);          // g_out[i & 1] was level i+2's dxin
      sumpool2_blocked_kernel<<<dim3(ew_grid(HW), (cm / 8) * N), 256, 0, st>>>(w.g_up, addend, dxin, 0, l.H, l.W);
      g_launches++;
      ev_sum_done = record(st);
    }
    // cvt block: the last l.f channels of dxin (all of them at level 0)
    const act_t* d_c = dxin + (size_t)N * HW * l.c0 * (i > 0 ? 1 : 0);
    link(st, lvl_br[i].st);
    t_bn_bwd(h, l.bn_cvt, w.z_cvt[i], d_c, w.dzc[i], HW, drop ? i : -1, dropout_seed, G, lvl_br[i]);
    ev_cvt_bwd[i] = record(lvl_br[i].st);
    wait(wg_br[1].st, ev_cvt_bwd[i]);
    if (!launch_wgrad(3, N, l.H, l.W, l.cin, l.cvt.cout_pad, l.f, w.feat[i], w.dzc[i], G + l.cvt.w_off, 0, l.cin, 1.f, w.wg_scratch[1], wg_br[1].st))
      return -2;
    d_prev_out = dxin;                      // its first c0 channel blocks = gradient w.r.t. prev_i (level i-1's output)
  }
  // join: the step is complete on the caller's stream
  for (int i = 0; i < nf; ++i) link(lvl_br[i].st, st);
  for (int i = 0; i < 2; ++i) link(wg_br[i].st, st);
  link(h->sc_stream, st);
  if (!ev_ok) { set_error("train step: event record / wait failed"); return -2; }
  return cuda_ok(cudaGetLastError(), "train step") ? 0 : -2;
}
