// Building blocks of decoder training (reference seg_solver.py:351-465): loss, weight
// gradient, BatchNorm / upsample forward+backward kernels and the optimizer step; gan-segmentation_b200/decoder_training.py
// composes them (with the conv kernel of shiftconv.cu for the forward pass and the data gradients).
//   softmax_ce : SoftmaxCELoss(axis=1) with sample_weight = (mask > -1) (seg_solver.py:404-407): per-sample loss =
//                mean over ALL H*W pixels of -w * log_softmax(logits)[label] (ignored pixels stay in the denominator),
//                and its gradient w * (softmax - onehot) / (H*W), times `grad_scale`: the backward pass runs its data
//                gradients through 16-bit tensors, 1/(H*W) = 9.5e-7 at 1024^2 is below the fp16 normal range, so the
//                caller passes grad_scale = H*W and folds 1/grad_scale into Adam's rescale_grad.
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
                                                                int K, int HW, float grad_scale) {
  const int n = blockIdx.y;
  const float* lg = logits + (size_t)n * K * HW;
  float* dl = dlogits ? dlogits + (size_t)n * K * HW : nullptr;
  const int* lab = labels + (size_t)n * HW;
  const float inv_hw = 1.f / (float)HW;
  const float gs = grad_scale * inv_hw;       // grad_scale = H*W keeps the gradient O(1) for the 16-bit backward pass
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
        dl[(size_t)k * HW + p] = w * (sm - (k == lc ? 1.f : 0.f)) * gs;
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

namespace gsx {

// ----------------------------------------------------------------------------------------------------------------
// Weight gradient of a 'same' 3x3 / 1x1 convolution on blocked 16-bit tensors (first version: CUDA cores, fp32
// accumulation, deterministic).   dW[co][ci][ky][kx] = sum_{n,y,x} dY[n,co,y,x] * X[n,ci,y+ky-P,x+kx-P],  db[co] = sum dY.
// grid (pixel tiles, N, (Cout/8)*(Cin/8)); a block owns one 8x8 (co,ci) block over a 16x32 pixel tile: thread =
// (co,ci) pair x 4 row phases, partial sums per (tile, sample) go to HBM and a second kernel adds them in fixed order.
// ----------------------------------------------------------------------------------------------------------------
static constexpr int kWgTY = 16, kWgTX = 32, kWgThreads = 256;

__device__ __forceinline__ void wg_unpack8(uint4 v, float* f) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#if GSX_FP16
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
#else
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
#endif
    f[2 * k] = t.x; f[2 * k + 1] = t.y;
  }
}

template <int K>
__global__ void __launch_bounds__(kWgThreads) conv_wgrad_kernel(const act_t* __restrict__ x, const act_t* __restrict__ dy,
                                                               float* __restrict__ partial, int N, int H, int W,
                                                               int CinB, int CoutB, int tiles_x) {
  constexpr int P = K / 2, SY = kWgTY + 2 * P, SX = kWgTX + 2 * P;
  __shared__ float sx[SY][SX][8];
  __shared__ float sdy[kWgTY][kWgTX][8];
  __shared__ float red[4][64][10];
  const int tile = blockIdx.x, n = blockIdx.y;
  const int cob = blockIdx.z / CinB, cib = blockIdx.z - cob * CinB;
  const int ty0 = (tile / tiles_x) * kWgTY, tx0 = (tile % tiles_x) * kWgTX;
  const size_t plane = (size_t)H * W;
  const act_t* xb = x + ((size_t)cib * N + n) * plane * 8;
  const act_t* db = dy + ((size_t)cob * N + n) * plane * 8;
  for (int i = threadIdx.x; i < SY * SX; i += kWgThreads) {
    const int r = i / SX, c = i - r * SX;
    const int yy = ty0 + r - P, xx = tx0 + c - P;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (yy >= 0 && yy < H && xx >= 0 && xx < W) wg_unpack8(__ldg(reinterpret_cast<const uint4*>(xb + ((size_t)yy * W + xx) * 8)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) sx[r][c][k] = f[k];
  }
  for (int i = threadIdx.x; i < kWgTY * kWgTX; i += kWgThreads) {
    const int r = i / kWgTX, c = i - r * kWgTX;
    const int yy = ty0 + r, xx = tx0 + c;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (yy < H && xx < W) wg_unpack8(__ldg(reinterpret_cast<const uint4*>(db + ((size_t)yy * W + xx) * 8)), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) sdy[r][c][k] = f[k];
  }
  __syncthreads();
  const int pair = threadIdx.x & 63, sub = threadIdx.x >> 6;
  const int co = pair >> 3, ci = pair & 7;
  float acc[K * K], accb = 0.f;
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = 0.f;
  for (int r = sub; r < kWgTY; r += 4) {
    for (int c = 0; c < kWgTX; ++c) {
      const float d = sdy[r][c][co];
      accb += d;
#pragma unroll
      for (int ky = 0; ky < K; ++ky)
#pragma unroll
        for (int kx = 0; kx < K; ++kx) acc[ky * K + kx] = fmaf(d, sx[r + ky][c + kx][ci], acc[ky * K + kx]);
    }
  }
#pragma unroll
  for (int t = 0; t < K * K; ++t) red[sub][pair][t] = acc[t];
  red[sub][pair][9] = accb;
  __syncthreads();
  // fixed-order sum of the 4 row phases; entry 9 = bias partial (meaningful for ci == 0 of the first input block)
  for (int i = threadIdx.x; i < 64 * 10; i += kWgThreads) {
    const int pr = i / 10, t = i - pr * 10;
    if (t < K * K || t == 9) {
      const float v = red[0][pr][t] + red[1][pr][t] + red[2][pr][t] + red[3][pr][t];
      const size_t E = (size_t)CoutB * CinB * 640;
      partial[((size_t)tile * N + n) * E + ((size_t)blockIdx.z * 64 + pr) * 10 + t] = v;
    }
  }
}

template <int K>
__global__ void conv_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, float* __restrict__ dbias,
                                         int P, int Cin, int Cout) {
  const int CinB = Cin / 8, CoutB = Cout / 8;
  const size_t E = (size_t)CoutB * CinB * 640;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if ((size_t)e >= E) return;
  const int t = e % 10, pr = (e / 10) % 64, blk = e / 640;
  const int cob = blk / CinB, cib = blk - cob * CinB;
  const int co = cob * 8 + (pr >> 3), ci = cib * 8 + (pr & 7);
  const bool is_w = t < K * K, is_b = (t == 9 && cib == 0 && (pr & 7) == 0);
  if (!is_w && !is_b) return;
  float s = 0.f;
  for (int p = 0; p < P; ++p) s += partial[(size_t)p * E + e];          // fixed order: bit-reproducible
  if (is_w) dw[((size_t)co * Cin + ci) * K * K + t] = s;
  else if (dbias) dbias[co] = s;
}


// ----------------------------------------------------------------------------------------------------------------
// Train-mode BatchNorm + LeakyReLU(0.2) (+ Dropout(0.5) mask) forward / backward and the nearest-x2 upsample pair, on
// fp32 NCHW tensors (first version of the decoder training path: the single-operator hooks work on fp32 NCHW).
// Per-channel reductions: per-block double partials, summed in a fixed order -> bit-reproducible.
// ----------------------------------------------------------------------------------------------------------------
static constexpr int kRedThreads = 256, kRedBlocks = 32;

struct BnArgs {
  const float* z; const float* dy; const float* drop; const float* gamma; const float* beta;
  const float* mean; const float* rstd;
  int N, C, HW;
};

template <bool BWD>
__global__ void __launch_bounds__(kRedThreads) chan_reduce_kernel(BnArgs a, double* __restrict__ partial) {
  const int c = blockIdx.y;
  const size_t total = (size_t)a.N * a.HW;
  double s0 = 0.0, s1 = 0.0;
  float mu = 0.f, rs = 0.f, ga = 0.f, be = 0.f;
  if (BWD) { mu = a.mean[c]; rs = a.rstd[c]; ga = a.gamma[c]; be = a.beta[c]; }
  for (size_t i = (size_t)blockIdx.x * kRedThreads + threadIdx.x; i < total; i += (size_t)gridDim.x * kRedThreads) {
    const size_t n = i / a.HW, p = i - n * a.HW;
    const size_t off = (n * a.C + c) * a.HW + p;
    const float z = a.z[off];
    if (!BWD) { s0 += z; s1 += (double)z * z; }
    else {
      const float xh = (z - mu) * rs;
      float g = a.dy[off] * ((ga * xh + be) > 0.f ? 1.f : 0.2f);
      if (a.drop) g *= 2.f * a.drop[off];
      s0 += g; s1 += (double)g * xh;
    }
  }
  __shared__ double r0[kRedThreads], r1[kRedThreads];
  r0[threadIdx.x] = s0; r1[threadIdx.x] = s1;
  __syncthreads();
  for (int o = kRedThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { r0[threadIdx.x] += r0[threadIdx.x + o]; r1[threadIdx.x] += r1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { partial[((size_t)c * gridDim.x + blockIdx.x) * 2] = r0[0]; partial[((size_t)c * gridDim.x + blockIdx.x) * 2 + 1] = r1[0]; }
}

// STATS: out0 = mean, out1 = biased variance, out2 = rstd (eps 1e-5).  BWD: out0 = dbeta, out1 = dgamma.
template <bool BWD>
__global__ void chan_finalize_kernel(const double* __restrict__ partial, int blocks, int C, double count, float* out0, float* out1,
                                     float* out2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int b = 0; b < blocks; ++b) { s0 += partial[((size_t)c * blocks + b) * 2]; s1 += partial[((size_t)c * blocks + b) * 2 + 1]; }
  if (BWD) { out0[c] = (float)s0; out1[c] = (float)s1; }
  else {
    const double m = s0 / count, v = fmax(s1 / count - m * m, 0.0);
    out0[c] = (float)m; out1[c] = (float)v; out2[c] = (float)(1.0 / sqrt(v + 1e-5));
  }
}

__global__ void bn_lrelu_fwd_kernel(BnArgs a, float* __restrict__ y) {
  const size_t total = (size_t)a.N * a.C * a.HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / a.HW) % a.C);
    const float pre = a.gamma[c] * ((a.z[i] - a.mean[c]) * a.rstd[c]) + a.beta[c];
    float v = pre > 0.f ? pre : 0.2f * pre;
    if (a.drop) v *= 2.f * a.drop[i];
    y[i] = v;
  }
}

__global__ void bn_lrelu_bwd_kernel(BnArgs a, const float* __restrict__ dbeta, const float* __restrict__ dgamma, float* __restrict__ dz) {
  const size_t total = (size_t)a.N * a.C * a.HW;
  const float m = (float)a.N * (float)a.HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / a.HW) % a.C);
    const float xh = (a.z[i] - a.mean[c]) * a.rstd[c];
    float g = a.dy[i] * ((a.gamma[c] * xh + a.beta[c]) > 0.f ? 1.f : 0.2f);
    if (a.drop) g *= 2.f * a.drop[i];
    dz[i] = a.gamma[c] * a.rstd[c] / m * (m * g - dbeta[c] - xh * dgamma[c]);
  }
}

__global__ void upsample2_kernel(const float* __restrict__ x, float* __restrict__ y, size_t planes, int H, int W) {
  const size_t total = planes * (size_t)(2 * H) * (2 * W);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % (2 * W)), Y = (int)((i / (2 * W)) % (2 * H));
    const size_t pl = i / ((size_t)4 * H * W);
    y[i] = x[(pl * H + (Y >> 1)) * W + (X >> 1)];
  }
}
__global__ void sumpool2_kernel(const float* __restrict__ dy, float* __restrict__ dx, size_t planes, int H, int W) {
  const size_t total = planes * (size_t)H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int X = (int)(i % W), Y = (int)((i / W) % H);
    const size_t pl = i / ((size_t)H * W);
    const float* r0 = dy + (pl * 2 * H + 2 * Y) * 2 * W + 2 * X;
    dx[i] = (r0[0] + r0[1]) + (r0[2 * W] + r0[2 * W + 1]);
  }
}

static int ew_blocks(size_t total) { return (int)std::min<size_t>(148 * 8, (total + 255) / 256); }

}  // namespace gsx

using namespace gsx;

extern "C" int gsx_op_conv_wgrad(int k, int n, int h, int w, int cin, int cout, const float* x_dev, const float* dy_dev,
                                 float* dw_dev, float* db_dev, gsx_stream stream) {
  if ((k != 1 && k != 3) || cin % 8 || cout % 8 || n <= 0 || !x_dev || !dy_dev || !dw_dev) { set_error("bad argument"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int tiles_x = (w + kWgTX - 1) / kWgTX, tiles_y = (h + kWgTY - 1) / kWgTY, tiles = tiles_x * tiles_y;
  const int CinB = cin / 8, CoutB = cout / 8;
  const size_t E = (size_t)CoutB * CinB * 640, P = (size_t)tiles * n;
  act_t *xb = nullptr, *dyb = nullptr;
  float* partial = nullptr;
  xb = static_cast<act_t*>(pool_get((size_t)n * cin * h * w * sizeof(act_t)));
  dyb = static_cast<act_t*>(pool_get((size_t)n * cout * h * w * sizeof(act_t)));
  partial = static_cast<float*>(pool_get(P * E * sizeof(float)));
  bool ok = xb && dyb && partial;
  if (!ok) set_error("conv_wgrad: out of device memory");
  if (ok) {
    launch_nchw_to_blocked(x_dev, xb, cin, n, h * w, st);
    launch_nchw_to_blocked(dy_dev, dyb, cout, n, h * w, st);
    dim3 grid(tiles, n, CoutB * CinB);
    const int rb = (int)((E + 255) / 256);
    if (k == 3) {
      conv_wgrad_kernel<3><<<grid, kWgThreads, 0, st>>>(xb, dyb, partial, n, h, w, CinB, CoutB, tiles_x);
      conv_wgrad_reduce_kernel<3><<<rb, 256, 0, st>>>(partial, dw_dev, db_dev, (int)P, cin, cout);
    } else {
      conv_wgrad_kernel<1><<<grid, kWgThreads, 0, st>>>(xb, dyb, partial, n, h, w, CinB, CoutB, tiles_x);
      conv_wgrad_reduce_kernel<1><<<rb, 256, 0, st>>>(partial, dw_dev, db_dev, (int)P, cin, cout);
    }
    g_launches += 4;
    ok = cuda_ok(cudaStreamSynchronize(st), "conv_wgrad");
  }
  pool_put(xb); pool_put(dyb); pool_put(partial);
  return ok ? 0 : -2;
}

extern "C" int gsx_op_upsample2(const float* x_dev, float* y_dev, int n, int c, int h, int w, gsx_stream stream) {
  if (!x_dev || !y_dev) { set_error("bad argument"); return -1; }
  const size_t total = (size_t)n * c * 4 * h * w;
  upsample2_kernel<<<ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(x_dev, y_dev, (size_t)n * c, h, w);
  g_launches++;
  return cuda_ok(cudaGetLastError(), "upsample2") ? 0 : -2;
}
extern "C" int gsx_op_sumpool2(const float* dy_dev, float* dx_dev, int n, int c, int h, int w, gsx_stream stream) {
  if (!dy_dev || !dx_dev) { set_error("bad argument"); return -1; }
  const size_t total = (size_t)n * c * h * w;
  sumpool2_kernel<<<ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(dy_dev, dx_dev, (size_t)n * c, h, w);
  g_launches++;
  return cuda_ok(cudaGetLastError(), "sumpool2") ? 0 : -2;
}
// stats_dev: [3][C] = mean, biased variance, rstd (written by fwd, read by bwd).  drop_dev: {0,1} mask or NULL.
extern "C" int gsx_op_bn_lrelu_fwd(const float* z_dev, const float* gamma_dev, const float* beta_dev, const float* drop_dev,
                                   float* y_dev, float* stats_dev, int n, int c, int hw, gsx_stream stream) {
  if (!z_dev || !gamma_dev || !beta_dev || !y_dev || !stats_dev) { set_error("bad argument"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(pool_get((size_t)c * kRedBlocks * 2 * sizeof(double)));
  if (!partial) { set_error("bn_lrelu_fwd: out of device memory"); return -2; }
  BnArgs a{z_dev, nullptr, drop_dev, gamma_dev, beta_dev, stats_dev, stats_dev + 2 * c, n, c, hw};
  chan_reduce_kernel<false><<<dim3(kRedBlocks, c), kRedThreads, 0, st>>>(a, partial);
  chan_finalize_kernel<false><<<(c + 127) / 128, 128, 0, st>>>(partial, kRedBlocks, c, (double)n * hw, stats_dev, stats_dev + c, stats_dev + 2 * c);
  bn_lrelu_fwd_kernel<<<ew_blocks((size_t)n * c * hw), 256, 0, st>>>(a, y_dev);
  g_launches += 3;
  const bool ok = cuda_ok(cudaStreamSynchronize(st), "bn_lrelu_fwd");
  pool_put(partial);
  return ok ? 0 : -2;
}
// dparam_dev: [2][C] = dbeta, dgamma.
extern "C" int gsx_op_bn_lrelu_bwd(const float* dy_dev, const float* z_dev, const float* stats_dev, const float* gamma_dev,
                                   const float* beta_dev, const float* drop_dev, float* dz_dev, float* dparam_dev, int n, int c,
                                   int hw, gsx_stream stream) {
  if (!dy_dev || !z_dev || !stats_dev || !gamma_dev || !beta_dev || !dz_dev || !dparam_dev) { set_error("bad argument"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* partial = static_cast<double*>(pool_get((size_t)c * kRedBlocks * 2 * sizeof(double)));
  if (!partial) { set_error("bn_lrelu_bwd: out of device memory"); return -2; }
  BnArgs a{z_dev, dy_dev, drop_dev, gamma_dev, beta_dev, stats_dev, stats_dev + 2 * c, n, c, hw};
  chan_reduce_kernel<true><<<dim3(kRedBlocks, c), kRedThreads, 0, st>>>(a, partial);
  chan_finalize_kernel<true><<<(c + 127) / 128, 128, 0, st>>>(partial, kRedBlocks, c, (double)n * hw, dparam_dev, dparam_dev + c, nullptr);
  bn_lrelu_bwd_kernel<<<ew_blocks((size_t)n * c * hw), 256, 0, st>>>(a, dparam_dev, dparam_dev + c, dz_dev);
  g_launches += 3;
  const bool ok = cuda_ok(cudaStreamSynchronize(st), "bn_lrelu_bwd");
  pool_put(partial);
  return ok ? 0 : -2;
}

extern "C" int gsx_softmax_ce(const float* logits_dev, const int* labels_dev, int n, int num_classes, int h, int w,
                              float* loss_dev, float* dlogits_dev, float grad_scale, float* scratch_dev,
                              size_t scratch_floats, gsx_stream stream) {
  if (!logits_dev || !labels_dev || !loss_dev || !scratch_dev || n <= 0) { set_error("bad argument"); return -1; }
  const int HW = h * w;
  const int blocks = std::min(256, (HW + kCeThreads - 1) / kCeThreads);
  if (scratch_floats < (size_t)n * blocks) { set_error("scratch too small (needs n*256 floats)"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  softmax_ce_kernel<<<dim3(blocks, n), kCeThreads, 0, st>>>(logits_dev, labels_dev, dlogits_dev, scratch_dev, num_classes, HW, grad_scale);
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
