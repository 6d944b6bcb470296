// Weight gradient of a stride-1 'same' k x k convolution (k = 3 or 1) on tcgen05 tensor cores -- the contraction of
// decoder training (reference seg_solver.py:411-412, err.backward() through nn.Conv2D, networks_seg.py:14-38,68,91):
//
//     dW[co][ci][ky][kx] = sum over (n, y, x) of  dY[n, co, y, x] * X[n, ci, y + ky - p, x + kx - p]
//
// is a GEMM whose K dimension is the PIXELS.  In the blocked activation layout [C/8][N][H][W][8] a pixel is one 16-byte
// vector of 8 channels, consecutive pixels are 16 bytes apart: read with pixels as K that is exactly the canonical
// MN-major, no-swizzle UMMA operand layout (core matrix = 8 pixels x 8 channels = 128 contiguous bytes; K groups 128 B
// apart, channel blocks one plane apart).  So the same TMA boxes as the forward pass feed the MMAs without a transpose:
//   A = X   (M = input channels, up to 128 per CTA; M = 64 MMAs when the layer has <= 64 of them: half the operand bytes --
//            their accumulator row i sits in TMEM lane 32*(i/16) + i%16, probed by tools/wgrad_m64_probe.py), read through a
//            descriptor whose start address is shifted by the
//            tap offset ((ky-1)*BW + (kx-1)) * 16 B -- the forward kernel's trick, now along K;
//   B = dY  (N = output channels, 16..64), rows of the same pixel tile;
//   D[tap]  = [ci][co] fp32 in TMEM, one 128 x Cout block per tap (9 * Cout <= 512 columns), accumulated over ALL pixel
//            tiles of the CTA: a split-K GEMM over pixels with one partial per CTA, summed in a fixed order by
//            wgrad_reduce_kernel (deterministic two-level reduction, no float atomics).
// The halo columns of the dY tile (neighbour tiles' pixels) are zeroed in shared memory before the MMAs read them, so
// that every pixel contributes exactly once; out-of-image rows / columns are zero-filled by TMA.
#include "../../include/gsx.h"
#include "gsx_internal.h"
#include "ptx.cuh"

#include <atomic>

namespace gsx {
extern std::atomic<uint64_t> g_launches;

struct WgradGeom {
  int H, W, N, Cin, Cout, K;          // K = 3 or 1
  int BW, TH;                         // box width (multiple of 16, >= tile width + 2), tile rows
  int TW;                             // tile width = BW - 2 (or W when the whole row fits)
  int tiles_x, tiles_y, n_tiles;      // pixel tiles (per sample), total = tiles_x * tiles_y * N
  int cbx, cbo;                       // channel blocks of X per CTA (<= 16), of dY
  int m_tiles;                        // Cin / 128 (rounded up): blockIdx.y
  int x_bytes, y_bytes, stage_bytes;  // per stage
  int x_cb_stride, y_cb_stride;       // bytes between channel blocks in smem
  int q_steps;                        // K steps of 16 pixels per tile = TH * BW / 16
  int rows;                           // valid accumulator rows per CTA = min(Cin, 128): only these reach the partials
  int m64;                            // 0: M = 128 MMAs; 1 / 2: M = 64 (Cin <= 64: half the A operand bytes per MMA), accumulator row i in
                                      //   TMEM lane i (1) or lane 32*(i/16) + i%16 (2)
  int kxm;                            // 1: layers with <= 16 input channels -- the kx taps ride along M: A descriptors with SBO = 16 B, so the
                                      //   8 row groups of an M = 64 MMA are the SAME channel block shifted by 0..7 pixels (0..2 used);
                                      //   one MMA per (ky, channel block) instead of one per tap: 6 instead of 9 at 16 channels
};

struct WgradParams {
  CUtensorMap tm_x, tm_y;
  WgradGeom g;
  float* partial;                     // [gridDim.y][gridDim.x][taps][128][Cout]
};

static constexpr int kWgStages = 2;
static constexpr int kWgPad = 128;     // zeroed bytes in front of every X stage (tap (0,0) starts one pixel before it)

struct __align__(16) WgHeader {
  uint64_t full[kWgStages], empty[kWgStages], done;
  uint32_t tmem_base, pad;
};

// MN-major, no-swizzle shared-memory descriptor: LBO = bytes between K groups of 8 (128), SBO = bytes between the
// 8-channel blocks along M / N.
__device__ __forceinline__ uint64_t wg_desc(uint32_t addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__global__ void __launch_bounds__(192, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const WgradGeom& g = p.g;
  WgHeader* hdr = reinterpret_cast<WgHeader*>(smem);
  uint8_t* stage0 = smem + 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int taps = g.K * g.K;
  const int tmem_cols = 512;
  // zero the whole stage area once: pads in front of / behind the tiles stay zero for the kernel's lifetime (K steps and
  // the M over-read touch them), everything else is overwritten by TMA
  for (int i = threadIdx.x * 16; i < kWgStages * g.stage_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(stage0 + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&hdr->full[s], 1); mbar_init(&hdr->empty[s], 1); }
    mbar_init(&hdr->done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tm_x);
    tma_prefetch_desc(&p.tm_y);
  }
  if (warp == 1) { tmem_alloc(&hdr->tmem_base, tmem_cols); tmem_relinquish(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;
  const int mt = blockIdx.y;                               // 128-channel block of Cin
  // contiguous range of pixel tiles for this CTA
  const int t0 = (int)((long long)blockIdx.x * g.n_tiles / gridDim.x), t1 = (int)((long long)(blockIdx.x + 1) * g.n_tiles / gridDim.x);
  const int per_sample = g.tiles_x * g.tiles_y;

  if (warp == 0) {
    if (elect_one()) {
      int it = 0;
      for (int t = t0; t < t1; ++t, ++it) {
        const int s = it % kWgStages, round = it / kWgStages;
        if (round > 0) mbar_wait_relaxed(&hdr->empty[s], (uint32_t)((round - 1) & 1));
        const int n = t / per_sample, r = t - n * per_sample;
        const int ty = r / g.tiles_x, tx = r - ty * g.tiles_x;
        const int x0 = tx * g.TW, y0 = ty * g.TH;
        uint8_t* st = stage0 + (size_t)s * g.stage_bytes;
        mbar_expect_tx(&hdr->full[s], (uint32_t)(g.x_bytes + g.y_bytes));
        const int P = g.K / 2;
        tma_load_4d(st + kWgPad, &p.tm_x, &hdr->full[s], (x0 - 1) * 2, y0 - P, n, mt * 16);
        tma_load_4d(st + kWgPad + g.cbx * g.x_cb_stride + kWgPad, &p.tm_y, &hdr->full[s], (x0 - 1) * 2, y0, n, 0);
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer
    uint32_t idesc = umma_idesc_16bit(g.m64 ? 64 : 128, (uint32_t)g.Cout, GSX_FP16 ? 0u : 1u) | (1u << 15) | (1u << 16);   // A, B MN-major
    int it = 0;
    bool first = true;
    const int P = g.K / 2;
    for (int t = t0; t < t1; ++t, ++it) {
      const int s = it % kWgStages;
      mbar_wait(&hdr->full[s], (uint32_t)((it / kWgStages) & 1));
      uint8_t* st = stage0 + (size_t)s * g.stage_bytes;
      uint8_t* xs = st + kWgPad;
      uint8_t* ys = xs + g.cbx * g.x_cb_stride + kWgPad;
      // halo columns of the dY tile belong to the neighbouring tiles: zero them (generic-proxy stores, then the fence
      // that makes them visible to the tensor core's async-proxy reads)
      for (int i = lane; i < g.cbo * g.TH * 2; i += 32) {
        const int cb = i / (g.TH * 2), rr = (i >> 1) % g.TH, side = i & 1;
        const int xl = side ? g.TW + 1 : 0;
        if (xl < g.BW) *reinterpret_cast<uint4*>(ys + (size_t)cb * g.y_cb_stride + ((size_t)rr * g.BW + xl) * 16) = make_uint4(0, 0, 0, 0);
      }
      // columns right of the tile inside the box (box wider than tile + 2) also belong to the neighbour
      for (int i = lane; i < g.cbo * g.TH * (g.BW - g.TW - 2); i += 32) {
        const int wgap = g.BW - g.TW - 2;
        const int cb = i / (g.TH * wgap), rem = i - cb * (g.TH * wgap), rr = rem / wgap, xl = g.TW + 2 + rem % wgap;
        *reinterpret_cast<uint4*>(ys + (size_t)cb * g.y_cb_stride + ((size_t)rr * g.BW + xl) * 16) = make_uint4(0, 0, 0, 0);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      tc_fence_after();
      // The issuing thread must be close to straight-line code (shiftconv.cu): descriptor low words per tap are built
      // once per tile, the K loop only adds 16 positions (16 x 16 B = 16 descriptor units) per step.
      const uint32_t xa = smem_u32(xs), ya = smem_u32(ys);
      const uint32_t lbo = (128u >> 4) << 16;
      uint32_t a_lo[9];
#pragma unroll
      for (int tp = 0; tp < 9; ++tp) {
        const int ky = tp / g.K, kx = tp - ky * g.K;
        const int shift = tp < taps ? ky * g.BW + (kx - P) : 0;     // X position of dY position q: q + ky*BW + kx - P
        a_lo[tp] = (((xa + (uint32_t)(shift * 16)) >> 4) & 0x3FFF) | lbo;
      }
      const uint32_t b_lo0 = ((ya >> 4) & 0x3FFF) | lbo;
      const uint32_t a_hi = (uint32_t)(wg_desc(0, (uint32_t)g.x_cb_stride) >> 32), b_hi = (uint32_t)(wg_desc(0, (uint32_t)g.y_cb_stride) >> 32);
      const uint32_t ncol = (uint32_t)g.Cout;
      if (elect_one()) {
        uint32_t acc = first ? 0u : 1u;
        if (g.kxm) {
          // accumulator (ky, cb) at columns (ky * cbx + cb) * Cout; its row kx * 8 + c = tap (ky, kx), channel cb * 8 + c
          const uint32_t a_hi_kx = (uint32_t)(wg_desc(0, 16u) >> 32);
          uint32_t a_lo_kx[6];
#pragma unroll
          for (int a = 0; a < 6; ++a) {
            const int ky = a / g.cbx, cb = a - ky * g.cbx;
            a_lo_kx[a] = a < 3 * g.cbx ? ((((xa + (uint32_t)(cb * g.x_cb_stride) + (uint32_t)((ky * g.BW - P) * 16)) >> 4) & 0x3FFF) | lbo) : 0u;
          }
          const int n_acc = 3 * g.cbx;
          for (int q = 0; q < g.q_steps; ++q) {
            const uint32_t ko = (uint32_t)q * 16u;
#pragma unroll
            for (int a = 0; a < 6; ++a)
              if (a < n_acc) umma_f16kind_lohi(tmem_base + (uint32_t)a * ncol, a_lo_kx[a] + ko, a_hi_kx, b_lo0 + ko, b_hi, idesc, acc);
            acc = 1u;
          }
        } else if (taps == 9) {
          for (int q = 0; q < g.q_steps; ++q) {
            const uint32_t ko = (uint32_t)q * 16u;
#pragma unroll
            for (int tp = 0; tp < 9; ++tp) umma_f16kind_lohi(tmem_base + (uint32_t)tp * ncol, a_lo[tp] + ko, a_hi, b_lo0 + ko, b_hi, idesc, acc);
            acc = 1u;
          }
        } else {
          for (int q = 0; q < g.q_steps; ++q) {
            const uint32_t ko = (uint32_t)q * 16u;
            umma_f16kind_lohi(tmem_base, a_lo[0] + ko, a_hi, b_lo0 + ko, b_hi, idesc, acc);
            acc = 1u;
          }
        }
        umma_commit(&hdr->empty[s]);
      }
      __syncwarp();
      first = false;
    }
    if (elect_one()) umma_commit(&hdr->done);
    __syncwarp();
  } else {
    // ---- epilogue: 4 warps, after the last MMA: TMEM -> partial[cta][tap][ci][co]
    const int quarter = warp & 3;
    if (t1 > t0) {
      mbar_wait_relaxed(&hdr->done, 0);
      tc_fence_after();
    }
    float* out = p.partial + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * taps) * g.rows * g.Cout;
    // accumulator row held by this thread's TMEM lane
    int row = quarter * 32 + lane;
    if (g.m64 == 2) row = lane < 16 ? quarter * 16 + lane : 128;
    const bool quarter_live = g.m64 == 2 ? quarter * 16 < g.rows : quarter * 32 < g.rows;
    const bool live = row < g.rows && quarter_live;
    if (g.kxm) {
      // rows 0..23 of every (ky, cb) accumulator: kx = row / 8, channel cb * 8 + row % 8 (M = 64: rows 0..15 in lanes 0..15 of
      // quarter 0, rows 16..23 in lanes 0..7 of quarter 1)
      const bool live_kx = row < 24 && quarter < 2;
      for (int a = 0; a < 3 * g.cbx && quarter < 2; ++a)
        for (int c = 0; c < g.Cout; c += 16) {
          uint32_t v[16];
          if (t1 > t0) {
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(a * g.Cout + c), v);
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0u;
          }
          const int ky = a / g.cbx, cb = a - ky * g.cbx, kx = row >> 3, ci = cb * 8 + (row & 7);
          float4* o = reinterpret_cast<float4*>(out + ((size_t)(ky * 3 + kx) * g.rows + ci) * g.Cout + c);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (live_kx) o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
        }
    } else
    for (int tp = 0; tp < taps && quarter_live; ++tp)
      for (int c = 0; c < g.Cout; c += 16) {
        uint32_t v[16];
        if (t1 > t0) {
          tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(tp * g.Cout + c), v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0u;
        }
        float4* o = reinterpret_cast<float4*>(out + ((size_t)tp * g.rows + row) * g.Cout + c);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (live) o[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]), __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
      }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// dW[co][ci][ky][kx] = scale * sum over CTAs of partial[mt][cta][tap][ci % 128][co].  One block per (tap, ci): thread
// (part, co) adds the CTAs part, part + parts, ... (reads coalesced over co), the parts are combined in a fixed order ->
// bit-reproducible.  (One thread per output walking all CTAs was 19 us for the 148-CTA layers: 2 blocks, 148 dependent-
// latency rounds of scattered 4-byte reads.)
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int ctas, int m_tiles,
                                                           int taps, int Cin, int Cout, int cout_real, int cin_off, int cin_total,
                                                           float scale, int rows) {
  __shared__ float sh[256];
  (void)m_tiles; (void)Cin;
  const int tp = blockIdx.x % taps, ci = blockIdx.x / taps;
  const int mt = ci / 128, r = ci - mt * 128;
  const int parts = 256 / Cout;
  const int co = threadIdx.x % Cout, part = threadIdx.x / Cout;
  float s = 0.f;
  if (part < parts)
    for (int c = part; c < ctas; c += parts) s += partial[(((size_t)(mt * ctas + c) * taps + tp) * rows + r) * Cout + co];
  sh[threadIdx.x] = s;
  __syncthreads();
  if (part == 0 && co < cout_real) {
    for (int q = 1; q < parts; ++q) s += sh[q * Cout + co];
    dw[((size_t)co * cin_total + cin_off + ci) * taps + tp] = s * scale;
  }
}

static int g_wg_sms = 0;
int g_wgrad_kxm = 1;        // gsx_set_option("wgrad_kxm", 0/1): kx taps along M for layers with <= 16 input channels
int g_wgrad_m64 = 2;        // gsx_set_option("wgrad_m64", 0 / 2): M = 64 MMAs for layers with <= 64 input channels (default on)

bool plan_wgrad(WgradGeom& g, int K, int N, int H, int W, int Cin, int Cout) {
  g = WgradGeom{};
  g.K = K; g.N = N; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
  if ((K != 1 && K != 3) || Cin % 8 || Cout % 16 || Cout > 56 || Cout * K * K > 512) return false;
  g.m64 = (Cin <= 64) ? g_wgrad_m64 : 0;
  g.kxm = (K == 3 && Cin <= 16 && g.m64 == 2 && g_wgrad_kxm) ? 1 : 0;
  const int m_blocks = g.kxm ? 1 : (g.m64 ? 8 : 16);       // channel blocks of A one MMA reads, whatever Cin is (kxm: 7 pixels past its own)
  int BW = 16;
  while (BW < W + 2 && BW < 128) BW <<= 1;
  g.BW = BW;
  g.TW = (W + 2 <= BW) ? W : BW - 2;
  g.tiles_x = (W + g.TW - 1) / g.TW;
  g.cbx = std::min(Cin / 8, 16);
  g.cbo = Cout / 8;
  g.m_tiles = (Cin + 127) / 128;
  const int halo = K - 1;
  int TH = std::min(H, 32);
  for (;; --TH) {
    const long xb = (long)g.cbx * (TH + halo) * BW * 16, yb = (long)g.cbo * TH * BW * 16;
    // the MMA reads 16 (M = 128) or 8 (M = 64) channel blocks of A whatever Cin is: the over-read must stay inside the stage
    const long xs_stride = (long)(TH + halo) * BW * 16;
    const long stage = kWgPad + std::max(xb, m_blocks * xs_stride) + kWgPad + yb + 512;
    if (kWgStages * stage + 1024 <= 200 * 1024 || TH == 1) {
      g.TH = TH;
      g.x_cb_stride = (int)xs_stride; g.y_cb_stride = TH * BW * 16;
      g.x_bytes = (int)xb; g.y_bytes = (int)yb;
      g.stage_bytes = (int)((stage + 1023) / 1024 * 1024);
      break;
    }
  }
  if (kWgStages * (long)g.stage_bytes + 1024 > 227 * 1024) return false;
  g.tiles_y = (H + g.TH - 1) / g.TH;
  g.n_tiles = g.tiles_x * g.tiles_y * N;
  g.q_steps = g.TH * BW / 16;
  g.rows = std::min(Cin, 128);
  return true;
}

// X, dY: blocked 16-bit tensors.  dw: [cout_real][cin_total][K][K] fp32, the slice [cin_off, cin_off + Cin) is written.
// scratch: wgrad_scratch_floats(...) floats.
size_t wgrad_scratch_floats(int K, int Cin, int Cout, int sms) {
  const int m_tiles = (Cin + 127) / 128;
  const int ctas = std::max(1, sms / m_tiles);
  return (size_t)m_tiles * ctas * K * K * 128 * Cout;
}

bool launch_wgrad(int K, int N, int H, int W, int Cin, int Cout, int cout_real, const act_t* x, const act_t* dy, float* dw,
                  int cin_off, int cin_total, float scale, float* scratch, cudaStream_t st) {
  WgradParams p;
  if (!plan_wgrad(p.g, K, N, H, W, Cin, Cout)) { set_error("wgrad: unsupported shape"); return false; }
  const WgradGeom& g = p.g;
  if (!g_wg_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&g_wg_sms, cudaDevAttrMultiProcessorCount, dev); }
  set_error("");
  make_act_tensormap(&p.tm_x, x, Cin, N, H, W, g.BW, g.TH + K - 1, 1, g.cbx);
  make_act_tensormap(&p.tm_y, dy, Cout, N, H, W, g.BW, g.TH, 1, g.cbo);
  if (*gsx_last_error()) return false;
  // at least ~4 pixel tiles per CTA: every CTA costs a TMEM drain and a partial of taps x rows x Cout floats
  const int ctas = std::max(1, std::min((g.n_tiles + 3) / 4, g_wg_sms / g.m_tiles));
  p.partial = scratch;
  static bool configured[64] = {false};
  int dev = 0; cudaGetDevice(&dev); dev = dev < 0 ? 0 : (dev > 63 ? 63 : dev);
  if (!configured[dev]) { cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); configured[dev] = true; }
  const size_t smem = 1024 + (size_t)kWgStages * g.stage_bytes;
  wgrad_kernel<<<dim3(ctas, g.m_tiles), 192, smem, st>>>(p);
  wgrad_reduce_kernel<<<K * K * Cin, 256, 0, st>>>(scratch, dw, ctas, g.m_tiles, K * K, Cin, Cout, cout_real, cin_off, cin_total, scale, g.rows);
  g_launches += 2;
  return cuda_ok(cudaGetLastError(), "wgrad launch");
}

}  // namespace gsx

using namespace gsx;

// Test / tuning hook: fp32 NCHW in (converted to the blocked 16-bit layout inside), dw [cout][cin][k][k] out.
extern "C" int gsx_op_conv_wgrad_tc(int k, int n, int h, int w, int cin, int cout, const float* x_dev, const float* dy_dev,
                                    float* dw_dev, gsx_stream stream) {
  if (!x_dev || !dy_dev || !dw_dev || n <= 0) { set_error("bad argument"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cout_p = (cout + 15) / 16 * 16;
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  act_t* xb = static_cast<act_t*>(pool_get((size_t)n * cin * h * w * sizeof(act_t)));
  act_t* dyb = static_cast<act_t*>(pool_get((size_t)n * cout_p * h * w * sizeof(act_t)));
  float* scratch = static_cast<float*>(pool_get(wgrad_scratch_floats(k, cin, cout_p, sms) * sizeof(float)));
  float* dyp = nullptr;
  bool ok = xb && dyb && scratch;
  if (ok && cout_p != cout) {
    dyp = static_cast<float*>(pool_get((size_t)n * cout_p * h * w * sizeof(float)));
    ok = dyp != nullptr;
    if (ok) {
      cudaMemsetAsync(dyp, 0, (size_t)n * cout_p * h * w * sizeof(float), st);
      cudaMemcpy2DAsync(dyp, (size_t)cout_p * h * w * sizeof(float), dy_dev, (size_t)cout * h * w * sizeof(float),
                        (size_t)cout * h * w * sizeof(float), n, cudaMemcpyDeviceToDevice, st);
    }
  }
  if (!ok) set_error("conv_wgrad_tc: out of device memory");
  if (ok) {
    launch_nchw_to_blocked(x_dev, xb, cin, n, h * w, st);
    launch_nchw_to_blocked(dyp ? dyp : dy_dev, dyb, cout_p, n, h * w, st);
    ok = launch_wgrad(k, n, h, w, cin, cout_p, cout, xb, dyb, dw_dev, 0, cin, 1.0f, scratch, st);
    ok = cuda_ok(cudaStreamSynchronize(st), "conv_wgrad_tc") && ok;
  }
  pool_put(xb); pool_put(dyb); pool_put(scratch); pool_put(dyp);
  return ok ? 0 : -2;
}
