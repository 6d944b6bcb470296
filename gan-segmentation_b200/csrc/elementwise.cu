// HBM-bound passes of the generate path on the blocked activation layout [C/8][N][H][W][8] act_t.
// All global accesses are 128-bit, consecutive lanes on consecutive pixels.
//   pass1 : Blur (networks_stylegan.py:200-236) + AddNoise (:302-304) + Bias (:544) + LeakyReLU(0.2)
//           (:38-40) + InstanceNorm sum/sumsq (warp-shuffle reduction) in one read + one write.
//   apply : InstanceNorm finalize + AdaIN modulation (:254-262) [+ ToRGB (:118-126) + the uint8
//           image transform (image_generator.py:76-84) on the last layer].
//   layout converters, Philox noise / latent generators, instance-norm statistics.
#include "gsx_internal.h"
#include "ptx.cuh"

namespace gsx {

__device__ int d_pdl_late_ew = 1;          // see shiftconv.cu: d_pdl_late_conv
void set_pdl_late_ew(int v) { cudaMemcpyToSymbol(d_pdl_late_ew, &v, sizeof(int)); }

__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#if GSX_FP16
    const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
#else
    const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[k]));
#endif
    f[2 * k] = v.x;
    f[2 * k + 1] = v.y;
  }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  uint32_t r;
#if GSX_FP16
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));     // saturates instead of inf
#else
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
#endif
  return r;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack2(f[0], f[1]); o.y = pack2(f[2], f[3]); o.z = pack2(f[4], f[5]); o.w = pack2(f[6], f[7]);
  return o;
}

// Block-wide reduction of 16 per-thread values (8 channel sums, 8 sums of squares) in a fixed order,
// then one plain store per value: dst[ch*2 + which] = block total.  No atomics -> bit-reproducible.
template <int NV>
__device__ __forceinline__ void block_reduce_store(float (&v)[NV], float* dst0, float* dst1, int nthreads) {
  __shared__ float red[32][NV + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[warp][i] = v[i];
  }
  __syncthreads();
  const int nw = nthreads >> 5;
  if (threadIdx.x < NV) {
    float s = 0.f;
    for (int w = 0; w < nw; ++w) s += red[w][threadIdx.x];
    // values 0..7 = channel sums, 8..15 = channel sums of squares
    const int ch = threadIdx.x & 7;
    ((threadIdx.x < 8 ? dst0 : dst1) + ch * 2)[0] = s;
  }
}

// ------------------------------------------------------------------------------------------ pass1
// One warp owns a strip of 30 output columns (32 loaded columns: one halo column each side, so the
// horizontal blur taps come from the neighbouring lanes by shuffle) and walks down kP1Rows rows with a
// 3-row register window for the vertical taps: 1 coalesced 512-B load per output row instead of 9.
static constexpr int kP1Warps = 4;
static constexpr int kP1Cols = 30;
static constexpr int kP1Rows = 32;
static constexpr int kP1Ring = 8;      // ring slots per lane
static constexpr int kP1Ahead = 6;     // rows in flight per lane

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kP1Warps * 32) pass1_kernel(const Pass1Args a, int strips, int rowblocks) {
  const int pdl_late = d_pdl_late_ew;
  if (!pdl_late) pdl_launch_dependents();
  pdl_wait();
  const int plane = blockIdx.y;                 // cb * N + n
  const int cb = plane / a.N, n = plane - cb * a.N;
  const int HW = a.H * a.W;
  const act_t* in = a.in + ((size_t)(a.in_broadcast ? cb : plane) * HW) * 8;
  act_t* out = a.out + ((size_t)plane * HW) * 8;
  const float* noise = a.noise ? a.noise + (size_t)n * HW : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ __align__(16) uint4 s_rows[kP1Warps][kP1Ring][32];
  __shared__ float s_noise[kP1Warps][kP1Ring][32];

  float ns[8], bs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    ns[i] = a.nscale ? a.nscale[cb * 8 + i] : 0.f;
    bs[i] = a.bias ? a.bias[cb * 8 + i] : 0.f;
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;

  const int item = blockIdx.x * kP1Warps + warp;            // (rowblock, strip)
  if (item < strips * rowblocks) {
    const int strip = item % strips, rb = item / strips;
    const int x = strip * kP1Cols - 1 + lane;               // column this lane loads
    const bool xin = x >= 0 && x < a.W;
    const bool xout = lane >= 1 && lane <= kP1Cols && x < a.W;
    const int y0 = rb * kP1Rows, y1 = min(a.H, y0 + kP1Rows);
    // Rows stream through a per-lane shared-memory ring filled by cp.async kP1Ahead rows ahead of their use
    // (zero-filled outside the image = the blur's zero padding): deep prefetch without spending registers,
    // which is what a latency-bound streaming pass needs (bytes in flight, not occupancy).
    uint4* ring = &s_rows[warp][0][lane];
    float* nring = &s_noise[warp][0][lane];
    const int lead = a.blur ? 1 : 0;                         // the vertical tap below needs row y+1
    auto issue = [&](int j) {                                // j-th consumed row: image row y0 + lead + j - ...
      const int y = y0 - lead + j;                           // rows y0-1 (blur) .. y1-1+lead
      const bool ok = xin && y >= 0 && y < a.H && y <= y1 - 1 + lead;
      cp_async16(ring + (j % kP1Ring) * 32, ok ? (const void*)(in + ((size_t)y * a.W + x) * 8) : (const void*)in, ok ? 16 : 0);
      const int yn = y0 + j;                                 // noise row for output row y0 + j
      const bool nok = noise && xout && yn < y1;
      cp_async4(nring + (j % kP1Ring) * 32, nok ? (const void*)(noise + (size_t)yn * a.W + x) : (const void*)in, nok ? 4 : 0);
      cp_async_commit();
    };
    const int n_rows = (y1 - y0) + 2 * lead;                 // rows to stream
#pragma unroll
    for (int j = 0; j < kP1Ahead; ++j) issue(j);
    float r0[8], r1[8], r2[8];
    int j = 0;
    if (a.blur) {                                            // prime the 3-row window with rows y0-1, y0
      cp_async_wait<kP1Ahead - 1>(); unpack8(ring[(0 % kP1Ring) * 32], r0); issue(kP1Ahead);
      cp_async_wait<kP1Ahead - 1>(); unpack8(ring[(1 % kP1Ring) * 32], r1); issue(kP1Ahead + 1);
      j = 2;
    }
    for (int y = y0; y < y1; ++y, ++j) {
      cp_async_wait<kP1Ahead - 1>();
      const uint4 cur = ring[(j % kP1Ring) * 32];
      const float nz = nring[((y - y0) % kP1Ring) * 32];
      issue(j + kP1Ahead);
      float v[8];
      if (a.blur) {
        unpack8(cur, r2);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float vb = r0[i] + 2.f * r1[i] + r2[i];                     // vertical [1,2,1]
          const float l = __shfl_up_sync(0xffffffffu, vb, 1), r = __shfl_down_sync(0xffffffffu, vb, 1);
          v[i] = (l + 2.f * vb + r) * (1.f / 16.f);                         // horizontal [1,2,1], /16
          r0[i] = r1[i]; r1[i] = r2[i];
        }
      } else {
        unpack8(cur, v);
      }
      if (xout) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float t = v[i] + ns[i] * nz + bs[i];
          t = fmaxf(t, 0.2f * t);
          v[i] = t;
          acc[i] += t;
          acc[8 + i] = fmaf(t, t, acc[8 + i]);
        }
        *reinterpret_cast<uint4*>(out + ((size_t)y * a.W + x) * 8) = pack8(v);
      }
    }
    cp_async_wait<0>();
    (void)n_rows;
  }
  if (pdl_late) pdl_launch_dependents();
  if (a.stats) {
    float* st = a.stats + (((size_t)n * gridDim.x + blockIdx.x) * a.C + cb * 8) * 2;
    block_reduce_store<16>(acc, st, st + 1, kP1Warps * 32);
  }
}

static void pass1_shape(int H, int W, int& strips, int& rowblocks, int& blocks) {
  strips = (W + kP1Cols - 1) / kP1Cols;
  rowblocks = (H + kP1Rows - 1) / kP1Rows;
  blocks = (strips * rowblocks + kP1Warps - 1) / kP1Warps;
}

int pass1_tiles(int H, int W) {
  int s, r, b;
  pass1_shape(H, W, s, r, b);
  return b;
}

void launch_pass1(const Pass1Args& a, cudaStream_t st) {
  int strips, rowblocks, blocks;
  pass1_shape(a.H, a.W, strips, rowblocks, blocks);
  dim3 grid(blocks, (a.C / 8) * a.N);
  launch_pdl(pass1_kernel, grid, dim3(kP1Warps * 32), 0, st, a, strips, rowblocks);
}

// ------------------------------------------------------------------------------------------ stats
__global__ void __launch_bounds__(256) stats_kernel(const act_t* in, float* stats, int C, int N, int HW) {
  pdl_launch_dependents();
  pdl_wait();
  const int plane = blockIdx.y;
  const int cb = plane / N, n = plane - cb * N;
  const act_t* src = in + (size_t)plane * HW * 8;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int pix = blockIdx.x * 256 + threadIdx.x; pix < HW; pix += gridDim.x * 256) {
    float f[8];
    unpack8(ld_dep_u4(src + (size_t)pix * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] += f[i]; acc[8 + i] += f[i] * f[i]; }
  }
  float* st = stats + (((size_t)n * gridDim.x + blockIdx.x) * C + cb * 8) * 2;
  block_reduce_store<16>(acc, st, st + 1, 256);
}

int stats_tiles(int HW) { return min(16, (HW + 255) / 256); }

void launch_stats(const act_t* in, float* stats_partial, int C, int N, int HW, cudaStream_t st, int tiles) {
  dim3 grid(tiles > 0 ? tiles : stats_tiles(HW), (C / 8) * N);
  launch_pdl(stats_kernel, grid, dim3(256), 0, st, in, stats_partial, C, N, HW);
}

// ------------------------------------------------------------------------------------------ finalize
// One warp per (n, c): lane l sums tiles l, l+32, ... (fixed order), then a fixed butterfly adds the 32 partials,
// so the result is bit-reproducible and the tile loop is 32-way parallel (544 tiles at 1024^2).
__global__ void __launch_bounds__(256) finalize_kernel(const float* partial, int T, int N, int C, float inv_hw,
                                                       const float* styles, int style_stride, int style_off,
                                                       float* coef) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // (n, c)
  const int lane = threadIdx.x & 31;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C;
  float s1 = 0.f, s2 = 0.f;
  const float2* p = reinterpret_cast<const float2*>(partial) + ((size_t)n * T * C + c);
  for (int t = lane; t < T; t += 32) {
    const float2 v = ld_dep_f2(p + (size_t)t * C);
    s1 += v.x;
    s2 += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  if (lane != 0) return;
  if (!styles) { coef[(size_t)i * 2] = s1; coef[(size_t)i * 2 + 1] = s2; return; }
  const float mean = s1 * inv_hw;
  const float var = fmaxf(s2 * inv_hw - mean * mean, 0.f);  // biased variance (InstanceNorm)
  const float rstd = rsqrtf(var + 1e-5f);                   // gluon InstanceNorm eps
  const float* sty = styles + (size_t)n * style_stride + style_off;
  const float a = rstd * (ld_dep_f32(sty + c) + 1.f);                    // ys + 1   (networks_stylegan.py:262)
  coef[(size_t)i * 2] = a;
  coef[(size_t)i * 2 + 1] = ld_dep_f32(sty + C + c) - mean * a;          // yb
}

void launch_finalize(const float* partial, int T, int N, int C, int HW, const float* styles, int style_stride,
                     int style_off, float* coef, cudaStream_t st) {
  const int warps = N * C;
  launch_pdl(finalize_kernel, dim3((warps * 32 + 255) / 256), dim3(256), 0, st, partial, T, N, C, 1.f / (float)HW, styles,
             style_stride, style_off, coef);
}

// ------------------------------------------------------------------------------------------ apply
__device__ __forceinline__ void adain_coeffs(const ApplyArgs& a, int n, int c, float /*inv_hw*/, float& ca, float& cb_) {
  const float2 v = *reinterpret_cast<const float2*>(a.coef + ((size_t)n * a.C + c) * 2);
  ca = v.x;
  cb_ = v.y;
}

static constexpr int kApThreads = 256;
static constexpr int kApPixPerThread = 8;

__global__ void __launch_bounds__(kApThreads) apply_kernel(const ApplyArgs a) {
  const int pdl_late = d_pdl_late_ew;
  if (!pdl_late) pdl_launch_dependents();
  pdl_wait();
  const int plane = blockIdx.y;
  const int cb = plane / a.N, n = plane - cb * a.N;
  const int HW = a.H * a.W;
  const float inv_hw = 1.f / (float)HW;
  float ca[8], cc[8];
  if (a.coef) {
#pragma unroll
    for (int i = 0; i < 8; ++i) adain_coeffs(a, n, cb * 8 + i, inv_hw, ca[i], cc[i]);
  } else {
    // finalize inlined: warp w owns channel cb*8 + w; lane l sums tiles l, l+32, ..., then the same butterfly as
    // finalize_kernel -> bit-identical coefficients, recomputed by every block of the plane (T <= a few hundred float2 loads)
    __shared__ float2 s_coef[8];
    const int wch = threadIdx.x >> 5, lane = threadIdx.x & 31, c = cb * 8 + wch;
    float s1 = 0.f, s2 = 0.f;
    const float* p = a.partial + ((size_t)n * a.T * a.C + c) * 2;
    for (int t = lane; t < a.T; t += 32) {
      const float2 v = ld_dep_f2(p + (size_t)t * a.C * 2);
      s1 += v.x;
      s2 += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) {
      const float mean = s1 * inv_hw;
      const float var = fmaxf(s2 * inv_hw - mean * mean, 0.f);
      const float rstd = rsqrtf(var + 1e-5f);
      const float* sty = a.styles + (size_t)n * a.style_stride + a.style_off;
      const float aa = rstd * (ld_dep_f32(sty + c) + 1.f);
      s_coef[wch] = make_float2(aa, ld_dep_f32(sty + a.C + c) - mean * aa);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) { ca[i] = s_coef[i].x; cc[i] = s_coef[i].y; }
  }
  const act_t* in = a.in + (size_t)plane * HW * 8;
  act_t* out = a.out + (size_t)plane * HW * 8;
  const int base = blockIdx.x * (kApThreads * kApPixPerThread);
  uint4 r[kApPixPerThread];
#pragma unroll
  for (int it = 0; it < kApPixPerThread; ++it) {
    const int pix = base + it * kApThreads + threadIdx.x;
    if (pix < HW) r[it] = ld_dep_u4(in + (size_t)pix * 8);
  }
#pragma unroll
  for (int it = 0; it < kApPixPerThread; ++it) {
    const int pix = base + it * kApThreads + threadIdx.x;
    if (pix < HW) {
      float f[8];
      unpack8(r[it], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], ca[i], cc[i]);
      if (a.out) *reinterpret_cast<uint4*>(out + (size_t)pix * 8) = pack8(f);
      if (a.out_nchw_f32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a.out_nchw_f32[((size_t)n * a.C + cb * 8 + i) * HW + pix] = f[i];
      }
    }
  }
  if (pdl_late) pdl_launch_dependents();
}

// Last layer: all channels of a pixel are needed for ToRGB, so one thread owns a pixel and walks the channel blocks.
__global__ void __launch_bounds__(256) apply_rgb_kernel(const ApplyArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];                 // ca[C], cc[C], wrgb[nc*C]
  float* s_ca = sm;
  float* s_cc = sm + a.C;
  float* s_w = sm + 2 * a.C;
  const int n = blockIdx.y;
  const int HW = a.H * a.W;
  const float inv_hw = 1.f / (float)HW;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) adain_coeffs(a, n, c, inv_hw, s_ca[c], s_cc[c]);
  for (int i = threadIdx.x; i < a.nc * a.C; i += blockDim.x) s_w[i] = a.wrgb[i];
  __syncthreads();
  const int CB = a.C / 8;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    float rgb[4] = {0.f, 0.f, 0.f, 0.f};
    for (int cb = 0; cb < CB; ++cb) {
      const size_t off = (((size_t)cb * a.N + n) * HW + pix) * 8;
      float f[8];
      unpack8(ld_dep_u4(a.in + off), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        f[i] = fmaf(f[i], s_ca[cb * 8 + i], s_cc[cb * 8 + i]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (k < a.nc) rgb[k] = fmaf(s_w[k * a.C + cb * 8 + i], f[i], rgb[k]);
      }
      if (a.out) *reinterpret_cast<uint4*>(a.out + off) = pack8(f);     // (null: AdaIN folded into the consumers, only the image leaves)
      if (a.out_nchw_f32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a.out_nchw_f32[((size_t)n * a.C + cb * 8 + i) * HW + pix] = f[i];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < a.nc) {
        const float v = rgb[k] + a.brgb[k];
        if (a.img_f32) a.img_f32[((size_t)n * a.nc + k) * HW + pix] = v;
        if (a.img_u8) {
          // image_generator.py:76-84: (x - (-1)) / 2 -> clip [0,1] -> *255 -> truncate to uint8
          float u = (v + 1.f) / 2.f;
          u = fminf(fmaxf(u, 0.f), 1.f);
          a.img_u8[((size_t)n * HW + pix) * a.nc + k] = (unsigned char)(255.f * u);
        }
      }
    }
  }
}

// ToRGB alone (AdaIN 2 of the last block folded into its consumers: the normalised feature is never written).  The
// modulation is linear, so it moves into the 1x1 weights: rgb_k = sum_c (Wrgb[k][c] a_c) t_c + (sum_c Wrgb[k][c] b_c + brgb_k).
// One thread = PIX pixels, all CB x PIX 16-byte loads issued before the first use; 3 bytes out per pixel.
template <int CB, int PIX>
__global__ void __launch_bounds__(256) rgb_kernel(const ApplyArgs a) {
  const int pdl_late = d_pdl_late_ew;
  if (!pdl_late) pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_w[4][CB * 8];
  __shared__ float s_b[4];
  const int n = blockIdx.y;
  const int HW = a.H * a.W;
  if (threadIdx.x < 4 * CB * 8) {
    const int k = threadIdx.x / (CB * 8), c = threadIdx.x - k * (CB * 8);
    s_w[k][c] = k < a.nc ? a.wrgb[k * a.C + c] * a.coef[((size_t)n * a.C + c) * 2] : 0.f;
  }
  if (threadIdx.x < 4) {
    float acc = threadIdx.x < a.nc ? a.brgb[threadIdx.x] : 0.f;
    if (threadIdx.x < a.nc)
      for (int c = 0; c < a.C; ++c) acc = fmaf(a.wrgb[threadIdx.x * a.C + c], a.coef[((size_t)n * a.C + c) * 2 + 1], acc);
    s_b[threadIdx.x] = acc;
  }
  __syncthreads();
  const int base = blockIdx.x * (256 * PIX) + threadIdx.x;
  uint4 r[PIX][CB];
#pragma unroll
  for (int i = 0; i < PIX; ++i) {
    const int pix = base + i * 256;
#pragma unroll
    for (int cb = 0; cb < CB; ++cb)
      if (pix < HW) r[i][cb] = ld_dep_u4(a.in + (((size_t)cb * a.N + n) * HW + pix) * 8);
  }
#pragma unroll
  for (int i = 0; i < PIX; ++i) {
    const int pix = base + i * 256;
    if (pix >= HW) continue;
    float rgb[4] = {s_b[0], s_b[1], s_b[2], s_b[3]};
#pragma unroll
    for (int cb = 0; cb < CB; ++cb) {
      float f[8];
      unpack8(r[i][cb], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) rgb[k] = fmaf(s_w[k][cb * 8 + j], f[j], rgb[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < a.nc) {
        if (a.img_f32) a.img_f32[((size_t)n * a.nc + k) * HW + pix] = rgb[k];
        if (a.img_u8) {
          // image_generator.py:76-84: (x - (-1)) / 2 -> clip [0,1] -> *255 -> truncate to uint8
          float u = (rgb[k] + 1.f) / 2.f;
          u = fminf(fmaxf(u, 0.f), 1.f);
          a.img_u8[((size_t)n * HW + pix) * a.nc + k] = (unsigned char)(255.f * u);
        }
      }
    }
  }
}

void launch_apply(const ApplyArgs& a, cudaStream_t st) {
  const int HW = a.H * a.W;
  if (a.wrgb && !a.out && !a.out_nchw_f32 && (a.C == 16 || a.C == 32 || a.C == 64)) {
    const int pixn = a.C == 16 ? 8 : (a.C == 32 ? 4 : 2);
    dim3 grid((HW + 256 * pixn - 1) / (256 * pixn), a.N);
    if (a.C == 16) launch_pdl(rgb_kernel<2, 8>, grid, dim3(256), 0, st, a);
    else if (a.C == 32) launch_pdl(rgb_kernel<4, 4>, grid, dim3(256), 0, st, a);
    else launch_pdl(rgb_kernel<8, 2>, grid, dim3(256), 0, st, a);
    return;
  }
  if (a.wrgb) {
    dim3 grid(min((HW + 255) / 256, 4096), a.N);
    const size_t smem = (size_t)(2 * a.C + a.nc * a.C) * sizeof(float);
    launch_pdl(apply_rgb_kernel, grid, dim3(256), smem, st, a);
  } else {
    dim3 grid((HW + kApThreads * kApPixPerThread - 1) / (kApThreads * kApPixPerThread), (a.C / 8) * a.N);
    launch_pdl(apply_kernel, grid, dim3(kApThreads), 0, st, a);
  }
}

// ------------------------------------------------------------------------------------------ layout
__global__ void blocked_to_nchw_kernel(const act_t* in, float* out, int C, int N, int HW) {
  pdl_launch_dependents();
  pdl_wait();
  const int plane = blockIdx.y;
  const int cb = plane / N, n = plane - cb * N;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    float f[8];
    unpack8(ld_dep_u4(in + ((size_t)plane * HW + pix) * 8), f);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[((size_t)n * C + cb * 8 + i) * HW + pix] = f[i];
  }
}
__global__ void nchw_to_blocked_kernel(const float* in, act_t* out, int C, int N, int HW) {
  pdl_launch_dependents();
  pdl_wait();
  const int plane = blockIdx.y;
  const int cb = plane / N, n = plane - cb * N;
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < HW; pix += gridDim.x * blockDim.x) {
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = ld_dep_f32(in + ((size_t)n * C + cb * 8 + i) * HW + pix);
    *reinterpret_cast<uint4*>(out + ((size_t)plane * HW + pix) * 8) = pack8(f);
  }
}
void launch_blocked_to_nchw(const act_t* in, float* out, int C, int N, int HW, cudaStream_t st) {
  dim3 grid(min((HW + 255) / 256, 1024), (C / 8) * N);
  blocked_to_nchw_kernel<<<grid, 256, 0, st>>>(in, out, C, N, HW);
}
void launch_nchw_to_blocked(const float* in, act_t* out, int C, int N, int HW, cudaStream_t st) {
  dim3 grid(min((HW + 255) / 256, 1024), (C / 8) * N);
  nchw_to_blocked_kernel<<<grid, 256, 0, st>>>(in, out, C, N, HW);
}

// ------------------------------------------------------------------------------------------ RNG
// Philox4x32-10 keyed by the seed; counter = (element/4, stream id, global sample index lo, hi).
// The result depends only on (seed, global sample index, stream, element), never on the batch
// split or the GPU count.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  // hardware-approximate log / sin / cos: this generator DEFINES the noise stream (there is no reference
  // stream to match -- MXNet's sampler is not reproducible outside MXNet), it only has to be N(0,1) and cheap
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;      // (0,1]
  const float u2 = (float)b * 2.3283064365386963e-10f;               // [0,1)
  const float r = sqrtf(-2.f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c; z1 = r * s;
}
__global__ void fill_normal_kernel(float* out, size_t per_sample, int N, uint64_t seed, uint64_t first_sample,
                                   uint32_t stream_id, const unsigned long long* first_dev) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t quads = (per_sample + 3) / 4;
  const int n = blockIdx.y;
  if (first_dev) first_sample = *first_dev;         // CUDA-graph replays: the running sample index lives in HBM
  const uint64_t gs = first_sample + (uint64_t)n;
  for (size_t qd = (size_t)blockIdx.x * blockDim.x + threadIdx.x; qd < quads; qd += (size_t)gridDim.x * blockDim.x) {
    uint32_t c[4] = {(uint32_t)qd, stream_id, (uint32_t)gs, (uint32_t)(gs >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float z[4];
    box_muller(c[0], c[1], z[0], z[1]);
    box_muller(c[2], c[3], z[2], z[3]);
    float* dst = out + (size_t)n * per_sample + qd * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (qd * 4 + i < per_sample) dst[i] = z[i];
  }
}
void launch_fill_noise(float* out, size_t plane_elems, int N, uint64_t seed, uint64_t first_sample, int layer,
                       cudaStream_t st, const unsigned long long* first_dev) {
  const size_t quads = (plane_elems + 3) / 4;
  dim3 grid((unsigned)min((size_t)1024, (quads + 255) / 256), N);
  fill_normal_kernel<<<grid, 256, 0, st>>>(out, plane_elems, N, seed, first_sample, (uint32_t)layer, first_dev);
}
// all noise planes of a forward pass in one launch: the quads (4 values) of all layers of a sample form one index
// space, so that every block has the same amount of work (a grid dimension per layer left ~70 % of the blocks of
// the low-resolution layers empty and the launch cost 0.26 ms for 0.36 GB)
struct NoiseIndex { size_t qstart[25]; int nlayers; };
__global__ void __launch_bounds__(256) fill_noise_all_kernel(NoisePlanes pl, NoiseIndex ix, uint64_t seed, uint64_t first_sample,
                                                             const unsigned long long* first_dev) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y;
  if (first_dev) first_sample = *first_dev;
  const uint64_t gs = first_sample + (uint64_t)n;
  const size_t total = ix.qstart[ix.nlayers];
  for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
    int layer = ix.nlayers - 1;                          // most quads belong to the last layers: scan from the top
    while (q < ix.qstart[layer]) --layer;
    const size_t qd = q - ix.qstart[layer];
    const size_t per_sample = pl.elems[layer];
    uint32_t c[4] = {(uint32_t)qd, (uint32_t)layer, (uint32_t)gs, (uint32_t)(gs >> 32)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    float z[4];
    box_muller(c[0], c[1], z[0], z[1]);
    box_muller(c[2], c[3], z[2], z[3]);
    float* dst = pl.ptr[layer] + (size_t)n * per_sample + qd * 4;
    if (qd * 4 + 3 < per_sample) {
      *reinterpret_cast<float4*>(dst) = make_float4(z[0], z[1], z[2], z[3]);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (qd * 4 + i < per_sample) dst[i] = z[i];
    }
  }
}
void launch_fill_noise_all(const NoisePlanes& pl, int nlayers, int N, uint64_t seed, uint64_t first_sample, cudaStream_t st,
                           const unsigned long long* first_dev) {
  NoiseIndex ix{};
  ix.nlayers = nlayers;
  size_t acc = 0;
  for (int l = 0; l < nlayers; ++l) { ix.qstart[l] = acc; acc += (pl.elems[l] + 3) / 4; }
  ix.qstart[nlayers] = acc;
  // ~4 quads per thread; at least one block per sample
  const size_t blocks = (acc + 1023) / 1024;
  dim3 grid((unsigned)(blocks < 1 ? 1 : (blocks > 4096 ? 4096 : blocks)), N);
  launch_pdl(fill_noise_all_kernel, grid, dim3(256), 0, st, pl, ix, seed, first_sample, first_dev);
}

void launch_fill_latents(float* z, int N, int Z, uint64_t seed, uint64_t first_sample, cudaStream_t st,
                         const unsigned long long* first_dev) {
  dim3 grid(1, N);
  fill_normal_kernel<<<grid, 128, 0, st>>>(z, (size_t)Z, N, seed, first_sample, 0xFFFFu, first_dev);
}
__global__ void advance_counter_kernel(unsigned long long* counter, unsigned long long by) { pdl_launch_dependents(); pdl_wait(); *counter += by; }
void launch_advance_counter(unsigned long long* counter, unsigned long long by, cudaStream_t st) {
  advance_counter_kernel<<<1, 1, 0, st>>>(counter, by);
}


// ----------------------------------------------------------------------------------------------------------------
// Border correction of the folded transposed-conv + blur (DECONV4B, plan.cpp).  The reference blurs the CROPPED
// deconv output D (networks_stylegan.py:16-17: Conv2DTranspose(4, stride 2, pad 1) then Blur with zero padding), the
// folded 3x3-per-phase composite also sees the ring of D just outside the crop (rows -1 and Ho, columns -1 and Wo).
// E = the blur taps that land on that ring; the conv epilogue subtracts it on the 1-pixel output border:
//   e_rows[n][s][X][co] = 1/4 * sum_dx bl[dx] D[s ? Ho : -1][X+dx]                      (X+dx in -1..Wo)
//   e_cols[n][s][Y][co] = 1/4 * sum_dy bl[dy] D[Y+dy][s ? Wo : -1]   for 0 <= Y+dy < Ho (corner terms live in e_rows)
// A ring value only sees one input row/column: <= 2 taps x Cin MACs.  grid (segments, 4 sides, N).
// ----------------------------------------------------------------------------------------------------------------
static constexpr int kBorderSeg = 30;                 // border pixels per block (ring positions: +2)
static constexpr int kBorderThreads = 128;

__global__ void __launch_bounds__(kBorderThreads) deconv_border_kernel(const act_t* x, const float* __restrict__ wt,
                                                                       float* __restrict__ e_rows, float* __restrict__ e_cols,
                                                                       int N, int Cin, int Cout, int H, int W,
                                                                       const float* coef) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float ring[];                      // [kBorderSeg + 2][Cout]
  const int side = blockIdx.y, n = blockIdx.z;         // 0 top, 1 bottom, 2 left, 3 right
  const int Ho = 2 * H, Wo = 2 * W;
  const bool is_row = side < 2;
  const int len = is_row ? Wo : Ho;
  const int p0 = blockIdx.x * kBorderSeg;
  if (p0 >= len) return;
  const size_t plane = (size_t)H * W;
  for (int idx = threadIdx.x; idx < (kBorderSeg + 2) * Cout; idx += kBorderThreads) {
    const int r = idx / Cout, co = idx - r * Cout;
    const int P = p0 - 1 + r;                          // position along the side, ring coordinates
    int Yr, Xr;
    bool live;
    if (is_row) { Yr = side == 0 ? -1 : Ho; Xr = P; live = P >= -1 && P <= Wo; }
    else        { Xr = side == 2 ? -1 : Wo; Yr = P; live = P >= 0 && P < Ho; }
    float acc = 0.f;
    if (live) {
      // D[Y][X], Y = 2i'+p':  sum_a in[i'+p'-1+a] * w[a == 0 ? 3-p' : 1-p']
      const int ip = (Yr + 2) / 2 - 1, pp = (Yr + 2) & 1, jp = (Xr + 2) / 2 - 1, qp = (Xr + 2) & 1;
      for (int ay = 0; ay < 2; ++ay) {
        const int iy = ip + pp - 1 + ay;
        if (iy < 0 || iy >= H) continue;
        const int ky = ay == 0 ? 3 - pp : 1 - pp;
        for (int ax = 0; ax < 2; ++ax) {
          const int ix = jp + qp - 1 + ax;
          if (ix < 0 || ix >= W) continue;
          const int kx = ax == 0 ? 3 - qp : 1 - qp;
          const float* wk = wt + ((size_t)(ky * 4 + kx) * Cin) * Cout + co;
          for (int cb = 0; cb < Cin / 8; ++cb) {
            const uint4 v = ld_dep_u4(x + (((size_t)cb * N + n) * plane + (size_t)iy * W + ix) * 8);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#if GSX_FP16
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w4[k]));
#else
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w4[k]));
#endif
              float x0 = f.x, x1 = f.y;
              if (coef) {          // the stored tensor is the un-normalised t: x = a*t + b (AdaIN folded into the consumers)
                const float4 ab = ld_dep_f4(coef + ((size_t)n * Cin + cb * 8 + 2 * k) * 2);
                x0 = fmaf(x0, ab.x, ab.y); x1 = fmaf(x1, ab.z, ab.w);
              }
              acc = fmaf(x0, __ldg(wk + (size_t)(cb * 8 + 2 * k) * Cout), acc);
              acc = fmaf(x1, __ldg(wk + (size_t)(cb * 8 + 2 * k + 1) * Cout), acc);
            }
          }
        }
      }
    }
    ring[idx] = acc;
  }
  __syncthreads();
  float* dst = (is_row ? e_rows : e_cols) + ((size_t)n * 2 + (side & 1)) * len * Cout;
  for (int idx = threadIdx.x; idx < kBorderSeg * Cout; idx += kBorderThreads) {
    const int q = idx / Cout, co = idx - q * Cout;
    const int P = p0 + q;
    if (P >= len) break;
    const float v = 0.25f * (0.25f * ring[q * Cout + co] + 0.5f * ring[(q + 1) * Cout + co] + 0.25f * ring[(q + 2) * Cout + co]);
    dst[(size_t)P * Cout + co] = v;
  }
}

void launch_deconv_border(const act_t* x, const float* wt, float* e_rows, float* e_cols, int N, int Cin, int Cout, int H,
                          int W, cudaStream_t st, const float* coef) {
  const int len = 2 * (H > W ? H : W);
  dim3 grid((len + kBorderSeg - 1) / kBorderSeg, 4, N);
  launch_pdl(deconv_border_kernel, grid, dim3(kBorderThreads), (kBorderSeg + 2) * Cout * sizeof(float), st, x, wt, e_rows, e_cols, N, Cin,
             Cout, H, W, coef);
}

}  // namespace gsx
