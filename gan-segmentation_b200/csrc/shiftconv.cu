// Shift-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Every convolution on the generate path -- the generator's 3x3 convs (reference
// networks_stylegan.py:446-457 used at :24,:46), the nearest-x2 + 3x3 conv (:27 + :24), the 4x4
// stride-2 transposed conv (:460-476 used at :16) and all decoder convs (networks_seg.py:14-38,68,91)
// -- is computed as  out[p, :] = sum_t  in[p + shift_t, :] * W_t  over a zero-haloed input tile:
//
//   * TMA loads one {x, y, n, channel-block} box of the blocked activation layout per k-chunk;
//     out-of-image halo elements are zero-filled by the TMA unit (the conv's zero padding).
//     In shared memory the box *is* the K-major no-swizzle UMMA operand layout: row = pixel
//     (flattened over the box, pitch BW), 16 B = 8 channels.
//   * the A operand of filter tap (dy,dx) is the same box read through a descriptor whose start
//     address is shifted by (dy*BW+dx)*16 B -- no im2col copy, each input byte is staged once.
//   * nearest-x2+conv and the transposed conv are four output phases of 2x2 taps on the low-res
//     input (16 tap/phase pairs); the phases either share one CTA (thin layers; 4 accumulator
//     groups) or are spread over blockIdx.z (wide layers).
//   * accumulators live in TMEM (128 lanes x N_tile fp32 columns per 128-pixel MMA tile); one
//     thread issues tcgen05.mma, tcgen05.commit releases smem stages / publishes the accumulators.
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b), fuse noise*scale + bias + leaky-ReLU
//     (+ residual add, + InstanceNorm sum/sumsq, or + argmax) and store 16-bit activations with 16 B per thread.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue.
#include "gsx_internal.h"
#include "ptx.cuh"

namespace gsx {

static constexpr int kThreads = 192;
static constexpr int kHeaderBytes = kConvHeaderBytes;   // barriers + tmem slot + stats scratch
static constexpr int kMaxStages = 8;

struct __align__(16) SmemHeader {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t accum_full;
  uint32_t tmem_base;
  uint32_t pad;
  float stats[4][2 * 256];     // [epilogue warp][channel in N_tile][sum, sumsq] -- one slot per warp, no atomics
};
static_assert(sizeof(SmemHeader) <= kHeaderBytes, "header too large");

__device__ __forceinline__ float lrelu02(float v) { return v > 0.f ? v : 0.2f * v; }

__device__ __forceinline__ uint32_t pack_x2(float a, float b) {
#if GSX_FP16
  a = fminf(fmaxf(a, -65504.f), 65504.f);
  b = fminf(fmaxf(b, -65504.f), 65504.f);
  __half2 h = __floats2half2_rn(a, b);
#else
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
#endif
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_x2(uint32_t w) {
#if GSX_FP16
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
#else
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
#endif
}

__global__ void __launch_bounds__(kThreads, 1) shiftconv_kernel(const __grid_constant__ ConvParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
  uint8_t* a_base = smem + kHeaderBytes;
  const ConvGeom& g = p.g;
  uint8_t* b_base = a_base + (size_t)g.stages * g.a_stage_stride;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile coordinates
  int t = blockIdx.x;
  const int tx = t % g.tiles_x; t /= g.tiles_x;
  const int ty = t % g.tiles_y; t /= g.tiles_y;
  const int tn = t;
  const int x0 = tx * g.TW, y0 = ty * g.TH, n0 = tn * g.NB;
  const int ntile = blockIdx.y;
  const int phase_z = g.phase_grid ? (int)blockIdx.z : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&hdr->full[s], 1);
      mbar_init(&hdr->empty[s], 1);
    }
    mbar_init(&hdr->accum_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&p.tm[0]);
    if (g.kch0 < g.n_k) tma_prefetch_desc(&p.tm[1]);
  }
  for (int i = threadIdx.x; i < 4 * 2 * 256; i += kThreads) (&hdr->stats[0][0])[i] = 0.f;
  if (warp == 1) {
    tmem_alloc(&hdr->tmem_base, (uint32_t)g.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = hdr->tmem_base;

  if (warp == 0) {
    // ================================ TMA producer =================================
    if (lane == 0) {
      const uint32_t stage_bytes = (uint32_t)(g.a_stage_bytes + g.b_stage_bytes);
      const act_t* wsrc = p.wpack + ((size_t)(phase_z * g.n_ntiles + ntile) * g.n_k) * (size_t)(g.b_stage_bytes / 2);
      for (int kc = 0; kc < g.n_k; ++kc) {
        const int s = kc % g.stages;
        const int it = kc / g.stages;
        if (it > 0) mbar_wait(&hdr->empty[s], (uint32_t)((it - 1) & 1));
        mbar_expect_tx(&hdr->full[s], stage_bytes);
        const int src = kc < g.kch0 ? 0 : 1;
        const int cb0 = (src ? kc - g.kch0 : kc) * g.CBK;
        // dim0 is in 8-byte units (2 per pixel) so that the inner box extent reaches 128 pixels
        tma_load_4d(a_base + (size_t)s * g.a_stage_stride, &p.tm[src], &hdr->full[s], (x0 - 1) * 2, y0 - 1, n0, cb0);
        bulk_load(b_base + (size_t)s * g.b_stage_bytes, wsrc + (size_t)kc * (g.b_stage_bytes / 2),
                  (uint32_t)g.b_stage_bytes, &hdr->full[s]);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===================================
    const uint32_t idesc = umma_idesc_16bit(128, (uint32_t)g.N_tile, GSX_FP16 ? 0u : 1u);
    const uint64_t a_hi = umma_desc_hi((uint32_t)g.cb_stride_bytes, 128);
    const uint64_t b_hi = umma_desc_hi((uint32_t)g.N_tile * 16, 128);
    const int k16_per_chunk = g.CBK >> 1;
    const uint32_t b_tile_bytes = (uint32_t)g.N_tile * 32;
    for (int kc = 0; kc < g.n_k; ++kc) {
      const int s = kc % g.stages;
      const int it = kc / g.stages;
      mbar_wait(&hdr->full[s], (uint32_t)(it & 1));
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(a_base + (size_t)s * g.a_stage_stride);
        const uint32_t b_addr = smem_u32(b_base + (size_t)s * g.b_stage_bytes);
        for (int slot = 0; slot < g.n_slots; ++slot) {
          const int grp = g.slot_group[slot];
          const uint32_t shift_bytes = (uint32_t)g.slot_shift[phase_z][slot] * 16u;
          for (int j = 0; j < k16_per_chunk; ++j) {
            const uint64_t bdesc = umma_desc(b_hi, b_addr + (uint32_t)(slot * k16_per_chunk + j) * b_tile_bytes);
            const uint32_t acc = (kc > 0 || j > 0 || !g.slot_first[slot]) ? 1u : 0u;
            const uint32_t a_k = a_addr + (uint32_t)(2 * j) * (uint32_t)g.cb_stride_bytes + shift_bytes;
            for (int mt = 0; mt < g.n_mtiles; ++mt) {
              const uint64_t adesc = umma_desc(a_hi, a_k + (uint32_t)mt * 2048u);
              umma_f16kind(tmem_base + (uint32_t)((grp * g.n_mtiles + mt) * g.N_tile), adesc, bdesc, idesc, acc);
            }
          }
        }
        umma_commit(&hdr->empty[s]);
        if (kc == g.n_k - 1) umma_commit(&hdr->accum_full);
      }
      __syncwarp();
    }
  } else {
    // ================================ epilogue =====================================
    const ConvEpi& e = p.e;
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int box_pix = g.BH * g.BW;
    const bool do_stats = (e.flags & EPI_STATS) != 0;
    const bool do_act = (e.flags & EPI_LRELU) != 0;
    const size_t plane_out = (size_t)e.Ho * e.Wo;

    mbar_wait(&hdr->accum_full, 0);
    tc_fence_after();

    if (e.flags & EPI_ARGMAX) {
      for (int mt = 0; mt < g.n_mtiles; ++mt) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * g.N_tile), v);
        tmem_ld_wait();
        const int q = mt * 128 + row;
        const int nb = q / box_pix;
        const int rem = q - nb * box_pix;
        const int yl = rem / g.BW, xl = rem - yl * g.BW;
        const int n = n0 + nb, y = y0 + yl, x = x0 + xl;
        if (nb < g.NB && yl < g.TH && xl < g.TW && n < g.N && y < g.H && x < g.W) {
          float best = __uint_as_float(v[0]) + (e.bias ? e.bias[0] : 0.f);
          int arg = 0;
          const size_t pix = (size_t)y * e.Wo + x;
          if (e.logits) e.logits[((size_t)n * e.num_classes) * plane_out + pix] = best;
#pragma unroll
          for (int c = 1; c < 16; ++c) {
            if (c < e.num_classes) {
              const float lv = __uint_as_float(v[c]) + (e.bias ? e.bias[c] : 0.f);
              if (e.logits) e.logits[((size_t)n * e.num_classes + c) * plane_out + pix] = lv;
              if (lv > best) { best = lv; arg = c; }       // first maximum wins (seg_solver.py:326)
            }
          }
          e.mask[(size_t)n * plane_out + pix] = (unsigned char)arg;
        }
      }
    } else {
      const int n_chunks = g.N_tile >> 4;
      for (int cc = 0; cc < n_chunks; ++cc) {
        const int c0 = ntile * g.N_tile + cc * 16;          // first output channel of this chunk
        float bias_r[16], ns_r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          bias_r[i] = e.bias ? __ldg(e.bias + c0 + i) : 0.f;
          ns_r[i] = e.nscale ? __ldg(e.nscale + c0 + i) : 0.f;
        }
        float s1[16], s2[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { s1[i] = 0.f; s2[i] = 0.f; }

        for (int grp = 0; grp < g.n_groups; ++grp) {
          const int ph = g.phase_grid ? phase_z : grp;
          const int py = ph >> 1, px = ph & 1;
          for (int mt = 0; mt < g.n_mtiles; ++mt) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) +
                          (uint32_t)((grp * g.n_mtiles + mt) * g.N_tile + cc * 16), v);
            tmem_ld_wait();
            const int q = mt * 128 + row;
            const int nb = q / box_pix;
            const int rem = q - nb * box_pix;
            const int yl = rem / g.BW, xl = rem - yl * g.BW;
            const int n = n0 + nb;
            int y = y0 + yl, x = x0 + xl;
            const bool valid = nb < g.NB && yl < g.TH && xl < g.TW && n < g.N && y < g.H && x < g.W;
            if (valid) {
              if (e.up) { y = 2 * y + py; x = 2 * x + px; }
              const size_t pix = (size_t)y * e.Wo + x;
              float f[16];
              const float nz = e.noise ? __ldg(e.noise + (size_t)n * plane_out + pix) : 0.f;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float a = __uint_as_float(v[i]) + ns_r[i] * nz + bias_r[i];
                f[i] = do_act ? lrelu02(a) : a;
              }
              if (e.addsrc) {
                const size_t plane_lo = (size_t)(e.Ho >> 1) * (e.Wo >> 1);
                const size_t pl = (size_t)(y >> 1) * (e.Wo >> 1) + (x >> 1);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint4 r = *reinterpret_cast<const uint4*>(
                      e.addsrc + (((size_t)((c0 >> 3) + h) * g.N + n) * plane_lo + pl) * 8);
                  const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const float2 b2 = unpack_x2(w4[k]);
                    f[h * 8 + 2 * k] += b2.x;
                    f[h * 8 + 2 * k + 1] += b2.y;
                  }
                }
              }
              if (do_stats) {
#pragma unroll
                for (int i = 0; i < 16; ++i) { s1[i] += f[i]; s2[i] += f[i] * f[i]; }
              }
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (c0 + h * 8 < e.Cout) {
                  uint4 o;
                  o.x = pack_x2(f[h * 8 + 0], f[h * 8 + 1]);
                  o.y = pack_x2(f[h * 8 + 2], f[h * 8 + 3]);
                  o.z = pack_x2(f[h * 8 + 4], f[h * 8 + 5]);
                  o.w = pack_x2(f[h * 8 + 6], f[h * 8 + 7]);
                  *reinterpret_cast<uint4*>(e.out + (((size_t)((c0 >> 3) + h) * g.N + n) * plane_out + pix) * 8) = o;
                }
              }
            }
          }
        }
        if (do_stats) {
          // 32 values (16 sums, 16 sums of squares) x 32 lanes -> lane L ends up holding value L fully reduced
          float vals[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) { vals[i] = s1[i]; vals[16 + i] = s2[i]; }
#pragma unroll
          for (int step = 0; step < 5; ++step) {
            const int half = 16 >> step;                      // values kept per lane after this step
            const int bit = 16 >> step;                       // lane bit deciding which half is kept
            const bool upper = (lane & bit) != 0;
#pragma unroll
            for (int i = 0; i < half; ++i) {
              const float keep = upper ? vals[half + i] : vals[i];
              const float send = upper ? vals[i] : vals[half + i];
              vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
            }
          }
          // lane's bits (16,8,4,2,1) selected halves successively -> value index = lane
          const int vi = lane;                                 // 0..15 sums, 16..31 sumsq
          const int ch = cc * 16 + (vi & 15);
          hdr->stats[warp - 2][ch * 2 + (vi >> 4)] += vals[0];     // this warp's own slot: plain add
        }
      }
      if (do_stats) {
        named_bar_sync(1, 128);                                // the 4 epilogue warps
        const int tid = threadIdx.x - 64;
        // fused stats need a single sample per CTA (NB == 1); enforced by the callers.  Fixed summation
        // order + one partial per (sample, tile): bit-reproducible, finalize_kernel adds the tiles up.
        const int tile_in_sample = ty * g.tiles_x + tx;
        for (int i = tid; i < g.N_tile * 2; i += 128) {
          const int ch = ntile * g.N_tile + (i >> 1);
          const float s = (hdr->stats[0][i] + hdr->stats[1][i]) + (hdr->stats[2][i] + hdr->stats[3][i]);
          if (ch < e.Cout) e.stats[(((size_t)n0 * e.stats_T + tile_in_sample) * e.Cout + ch) * 2 + (i & 1)] = s;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

void launch_shiftconv(const ConvParams& p, cudaStream_t st) {
  static int configured_smem = 0;
  if (p.g.smem_bytes > configured_smem) {
    cudaFuncSetAttribute(shiftconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    configured_smem = 227 * 1024;
  }
  dim3 grid((unsigned)(p.g.tiles_x * p.g.tiles_y * p.g.tiles_n), (unsigned)p.g.n_ntiles, p.g.phase_grid ? 4u : 1u);
  shiftconv_kernel<<<grid, kThreads, p.g.smem_bytes, st>>>(p);
}

}  // namespace gsx
