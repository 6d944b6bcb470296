// Shift-GEMM convolution on tcgen05 tensor cores (sm_100a), persistent and warp-specialized.
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
//     groups) or are separate work items (wide layers).
//   * accumulators live in TMEM (128 lanes x N_tile fp32 columns per 128-pixel MMA tile), double
//     buffered when they fit twice, so the MMAs of tile i+1 overlap the epilogue of tile i.
//   * one CTA per SM loops over its work items (static round-robin); for thin layers the packed
//     weights are loaded once per CTA and stay resident in shared memory.
//   * epilogue warps read TMEM (tcgen05.ld 32x32b), fuse noise*scale + bias + leaky-ReLU
//     (+ residual add, + InstanceNorm sum/sumsq, or + argmax) and store 16-bit activations, 16 B per thread.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2.. = 4*G epilogue warps
// (G warps per TMEM lane quarter).
#include "gsx_internal.h"
#include "ptx.cuh"

#include <cstdlib>

// tuning builds only (make TUNING=1 -> -DGSX_TUNING=1): the role-isolation switches of g.dbg (env GSX_DBG, tools/roles.sh)
// exist in the kernel; the release build compiles them out.  -DGSX_EPI_LITE=1 compiles every epilogue body out.
#ifndef GSX_TUNING
#define GSX_TUNING 0
#endif
#ifndef GSX_EPI_LITE
#define GSX_EPI_LITE 0
#endif

namespace gsx {

// 1: the big kernels release their programmatic dependents when a CTA has finished its work instead of at its start
// (gsx_set_option("pdl", 3)): the successor's launch latency still overlaps this kernel's tail, but its CTAs do not sit on
// shared memory while this kernel runs
__device__ int d_pdl_late_conv = 1;
void set_pdl_late_conv(int v) { cudaMemcpyToSymbol(d_pdl_late_conv, &v, sizeof(int)); }

static constexpr int kMaxStages = 8;

struct __align__(16) SmemHeader {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t bres_full;
  uint64_t aux_full[2];
  uint64_t aux_empty[2];
  uint32_t tmem_base;
  uint32_t stats_cnt[2];      // epilogue warps that have finished the statistics of the tile in slot buffer 0 / 1
  uint2 sched[64 + 16];       // per (phase, slot, k16 step): {A offset, B offset} in 16-byte units (+16: group over-read)
  uint2 sched2[64 + 16];      //   ... {first accumulator column, instruction descriptor}: variable-N plans issue narrower MMAs
};
static_assert(sizeof(SmemHeader) <= kConvHeaderBytes, "header too large");

__device__ __forceinline__ float lrelu02(float v) { return fmaxf(v, 0.2f * v); }

__device__ __forceinline__ uint32_t pack_x2(float a, float b) {
  uint32_t r;
#if GSX_FP16
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));     // saturates instead of inf
#else
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
#endif
  return r;
}
__device__ __forceinline__ float2 unpack_x2(uint32_t w) {
#if GSX_FP16
  return __half22float2(*reinterpret_cast<const __half2*>(&w));
#else
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
#endif
}

// 32 values (16 sums, 16 sums of squares) x 32 lanes -> lane L returns value L fully reduced over the warp
__device__ __forceinline__ float warp_reduce_32x32(const float (&s1)[16], const float (&s2)[16], int lane) {
  float vals[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) { vals[i] = s1[i]; vals[16 + i] = s2[i]; }
  return warp_transpose_reduce32(vals, lane);
}

// argmax over the classes of the output phases held in one 16-column chunk (CT columns per phase): first maximum
// wins (seg_solver.py:326); mask bytes of the two px phases leave as one 2-byte store.
template <int CT>
__device__ __forceinline__ void argmax_phases(const uint32_t (&v)[16], const ConvEpi& e, int cc, int n, int y, int x) {
  constexpr int PPC = 16 / CT;                      // phases per chunk
  const size_t plane_out = (size_t)e.Ho * e.Wo;
  unsigned char arg[PPC];
#pragma unroll
  for (int p = 0; p < PPC; ++p) {
    const int ph = cc * PPC + p;
    const size_t pix = (size_t)(2 * y + (ph >> 1)) * e.Wo + 2 * x + (ph & 1);
    float best = __uint_as_float(v[p * CT]) + (e.bias ? __ldg(e.bias) : 0.f);
    int a = 0;
    if (e.logits) e.logits[((size_t)n * e.num_classes) * plane_out + pix] = best;
#pragma unroll
    for (int c = 1; c < CT; ++c) {
      if (c < e.num_classes) {
        const float lv = __uint_as_float(v[p * CT + c]) + (e.bias ? __ldg(e.bias + c) : 0.f);
        if (e.logits) e.logits[((size_t)n * e.num_classes + c) * plane_out + pix] = lv;
        if (lv > best) { best = lv; a = c; }
      }
    }
    arg[p] = (unsigned char)a;
  }
  if (PPC >= 2) {
#pragma unroll
    for (int p = 0; p < PPC; p += 2) {
      const int ph = cc * PPC + p;
      const size_t pix = (size_t)(2 * y + (ph >> 1)) * e.Wo + 2 * x;
      *reinterpret_cast<uchar2*>(e.mask + (size_t)n * plane_out + pix) = make_uchar2(arg[p], arg[p + (PPC >= 2 ? 1 : 0)]);
    }
  } else {
    const int ph = cc;
    e.mask[(size_t)n * plane_out + (size_t)(2 * y + (ph >> 1)) * e.Wo + 2 * x + (ph & 1)] = arg[0];
  }
}

// CNT MMAs (one schedule group) for every 128-row MMA tile of the work item, as straight-line code.
template <int CNT, bool VARN>
__device__ __forceinline__ void issue_group(uint32_t d, int n_mtiles, uint32_t n_tile, uint32_t mt_desc, const uint32_t (&ab)[16],
                                            const uint32_t (&bk)[16], const uint32_t (&dc)[16], const uint32_t (&id)[16], uint32_t a_hi,
                                            uint32_t b_hi, uint32_t idesc, uint32_t first_acc) {
  uint32_t moff = 0;
  for (int mt = 0; mt < n_mtiles; ++mt, moff += mt_desc, d += n_tile) {
#pragma unroll
    for (int k = 0; k < CNT; ++k)
      umma_f16kind_lohi(VARN ? d + dc[k] : d, ab[k] + moff, a_hi, bk[k], b_hi, VARN ? id[k] : idesc, k == 0 ? first_acc : 1u);
  }
}

struct TileCoord { int x0, y0, n0, ntile, phase, tile_in_sample; };
struct Loc { int nb, yl, xl, n, y, x; bool valid; };     // a GEMM row of a tile: tile-local and global coordinates

// a / d for 0 <= a < 2^24 via the float reciprocal (+ one correction step): ~8 instructions instead of the ~35 of an
// integer division -- decode_tile runs once per work item in all three roles.
__device__ __forceinline__ int fast_div(int a, int d, int& rem) {
  int q = __float2int_rz(__int2float_rz(a) * __frcp_rn(__int2float_rz(d)));
  int r = a - q * d;
  if (r < 0) { --q; r += d; }
  if (r >= d) { ++q; r -= d; }
  rem = r;
  return q;
}

__device__ __forceinline__ TileCoord decode_tile(const ConvGeom& g, int t) {
  TileCoord c;
  const int sp = g.tiles_x * g.tiles_y * g.tiles_n;
  int s;
  const int rest = fast_div(t, sp, s);
  c.phase = fast_div(rest, g.n_ntiles, c.ntile);
  int tx, ty;
  s = fast_div(s, g.tiles_x, tx);
  s = fast_div(s, g.tiles_y, ty);
  c.x0 = tx * g.TW; c.y0 = ty * g.TH; c.n0 = s * g.NB;
  c.tile_in_sample = ty * g.tiles_x + tx;
  return c;
}

// GEN = the generator's epilogues (noise, InstanceNorm statistics, deconv+blur border correction); the decoder / raw
// variant compiles those paths out (127 instead of 168 registers).  G = epilogue warps per TMEM lane quarter.
// MODE = which epilogue the instantiation contains (one per kernel: the three bodies together cost instruction-cache
// misses -- 11 % of the warp stalls of the r01 kernel were "no instruction").
enum { kEpiGeneric = 0, kEpiUpCols = 1, kEpiArgmax = 2 };
template <int G, bool GEN, int MODE, bool VARN>
__global__ void __launch_bounds__(64 + 128 * G, 1) shiftconv_kernel(const __grid_constant__ ConvParams p) {
  constexpr int kEpiWarps = 4 * G;
  constexpr int kEpiThreads = 128 * G;
  extern __shared__ __align__(1024) uint8_t smem[];
  SmemHeader* hdr = reinterpret_cast<SmemHeader*>(smem);
  const ConvGeom& g = p.g;
  float* stats_slots = reinterpret_cast<float*>(smem + kConvHeaderBytes);     // [2][kEpiWarps][2*N_tile]
  uint8_t* a_base = smem + g.a_off;
  uint8_t* b_base = a_base + (size_t)g.stages * g.a_stage_stride;

  const int pdl_late = d_pdl_late_conv;
  if (!pdl_late) pdl_launch_dependents();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = g.tiles_x * g.tiles_y * g.tiles_n * g.n_ntiles * (g.phase_grid ? 4 : 1);
  const int nbuf = g.acc_bufs;
  const int cols_per_buf = g.n_mtiles * g.N_tile;
  // Work items of this CTA: round-robin over the grid, or -- per-sample layers -- one contiguous range, so that a CTA
  // crosses a sample boundary (= reloads its resident weight set and bias) at most a few times per launch.
  const int t_first = g.per_sample ? (int)((long long)blockIdx.x * total_tiles / gridDim.x) : (int)blockIdx.x;
  const int t_last = g.per_sample ? (int)((long long)(blockIdx.x + 1) * total_tiles / gridDim.x) : total_tiles;
  const int t_step = g.per_sample ? 1 : (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < g.stages; ++s) {
      mbar_init(&hdr->full[s], 1);
      mbar_init(&hdr->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hdr->tmem_full[b], 1);
      mbar_init(&hdr->tmem_empty[b], kEpiWarps);
    }
    mbar_init(&hdr->bres_full, 1);
    hdr->stats_cnt[0] = 0; hdr->stats_cnt[1] = 0;
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hdr->aux_full[b], 1);
      mbar_init(&hdr->aux_empty[b], kEpiWarps);
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.tm[0]);
    if (g.kch0 < g.n_k) tma_prefetch_desc(&p.tm[1]);
    if (g.s2d) {
      for (int pl = 1; pl < 4; ++pl) tma_prefetch_desc(&p.tm_pl[0][pl]);
    }
    if (g.aux_kind) tma_prefetch_desc(&p.tm_aux);
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + 80) {
    // MMA schedule: entry (phase, slot, j) = {tap shift + j k16-steps of A, (slot * k16_per_chunk + j) B tiles}
    const int i = threadIdx.x - 64, k16pc = g.CBK >> 1, len = g.n_slots * k16pc;
    const int nph = g.phase_grid ? 4 : 1;
    uint2 v = make_uint2(0u, 0u), v2 = make_uint2(0u, umma_idesc_16bit(128, (uint32_t)g.N_tile, GSX_FP16 ? 0u : 1u));
    if (i < nph * len) {
      const int ph = i / len, idx = i - ph * len, slot = idx / k16pc, j = idx - slot * k16pc;
      const int4 tp = __ldg(p.taps + ph * kMaxSlots + slot);       // {A shift bytes, first column, columns (0 = all), -}
      v.x = ((uint32_t)tp.x >> 4) + (uint32_t)j * ((2u * (uint32_t)g.cb_stride_bytes) >> 4);
      v.y = (uint32_t)idx * (((uint32_t)g.N_tile * 32) >> 4) + (uint32_t)tp.y;     // B rows are 16 bytes apart: + first column
      v2.x = (uint32_t)tp.y;
      if (tp.z > 0) v2.y = umma_idesc_16bit(128, (uint32_t)tp.z, GSX_FP16 ? 0u : 1u);
    }
    hdr->sched[i] = v;
    hdr->sched2[i] = v2;
  }
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
    if (elect_one()) {
      const size_t b_stage_elems = (size_t)(g.b_stage_bytes / 2);
      if (g.b_resident && !g.per_sample) {                     // all k-chunks of the (single) weight set
        mbar_expect_tx(&hdr->bres_full, (uint32_t)(g.n_k * g.b_stage_bytes));
        bulk_load(b_base, p.wpack, (uint32_t)(g.n_k * g.b_stage_bytes), &hdr->bres_full);
      }
      const uint32_t stage_bytes = (uint32_t)(g.a_stage_bytes + (g.b_resident ? 0 : g.b_stage_bytes));
      pdl_wait();                                   // activations / aux tiles come from earlier kernels of the stream
      int it = 0, tlp = 0;
      int w_sample = -1;                            // sample whose weight set is resident (per-sample layers)
      for (int t = t_first; t < t_last; t += t_step, ++tlp) {
        const TileCoord tc = decode_tile(g, t);
        if (g.aux_kind) {
          // per-tile epilogue operand (noise plane tile / residual tile), double buffered on its own barriers
          const int ab = tlp & 1;
          if (tlp >= 2) mbar_wait_relaxed(&hdr->aux_empty[ab], (uint32_t)(((tlp >> 1) - 1) & 1));
          mbar_expect_tx(&hdr->aux_full[ab], (uint32_t)g.aux_bytes_tx);
          uint8_t* dst = smem + g.aux_off + (size_t)ab * g.aux_bytes;
          if (g.aux_kind == 1) tma_load_3d(dst, &p.tm_aux, &hdr->aux_full[ab], tc.x0 << g.aux_up, tc.y0 << g.aux_up, tc.n0);
          else tma_load_4d(dst, &p.tm_aux, &hdr->aux_full[ab], (tc.x0 >> g.aux_shift) * 2, tc.y0 >> g.aux_shift, tc.n0,
                           tc.ntile * (g.cout_tile >> 3));
        }
        if (g.b_resident && g.per_sample && tc.n0 != w_sample) {
          // the resident weight set belongs to a sample: at a sample boundary wait until the MMAs that read the old
          // set have finished (the "empty" commit of the last chunk issued covers all earlier ones), then reload
          if (it > 0) mbar_wait_relaxed(&hdr->empty[(it - 1) % g.stages], (uint32_t)(((it - 1) / g.stages) & 1));
          w_sample = tc.n0;
          mbar_expect_tx(&hdr->bres_full, (uint32_t)(g.n_k * g.b_stage_bytes));
          bulk_load(b_base, p.wpack + (size_t)tc.n0 * p.wpack_n_stride, (uint32_t)(g.n_k * g.b_stage_bytes), &hdr->bres_full);
        }
        const act_t* wsrc = p.wpack + (size_t)tc.n0 * p.wpack_n_stride + ((size_t)(tc.phase * g.n_ntiles + tc.ntile) * g.n_k) * b_stage_elems;
        for (int kc = 0; kc < g.n_k; ++kc, ++it) {
          const int s = it % g.stages;
          const int round = it / g.stages;
          if (round > 0) mbar_wait_relaxed(&hdr->empty[s], (uint32_t)((round - 1) & 1));
          if (GSX_TUNING && (g.dbg & 4)) { mbar_arrive(&hdr->full[s]); continue; }
          mbar_expect_tx(&hdr->full[s], stage_bytes);
          const int src = kc < g.kch0 ? 0 : 1;
          const int cb0 = (src ? kc - g.kch0 : kc) * g.CBK;
          if (g.s2d) {
            // the 4 input phase planes of the block tile, each a dense [cb][nb][BH][BW] operand plane
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) {
              uint8_t* dst = a_base + (size_t)s * g.a_stage_stride + (size_t)pl * g.plane_stride;
              if (g.in_planar)        // stored phase-planar: dense rows
                tma_load_4d(dst, &p.tm_pl[src][pl], &hdr->full[s], (tc.x0 - 1) * 2, tc.y0 - 1, tc.n0, cb0);
              else                    // gathered from the pixel-interleaved layout: 16-byte inner extent (slow)
                tma_load_5d(dst, &p.tm_pl[src][pl], &hdr->full[s], 0, tc.x0 - 1, tc.y0 - 1, tc.n0, cb0);
            }
          } else {
            // dim0 is in 8-byte units (2 per pixel) so that the inner box extent reaches 128 pixels
            tma_load_4d(a_base + (size_t)s * g.a_stage_stride, &p.tm[src], &hdr->full[s], (tc.x0 - 1) * 2, tc.y0 - 1,
                        tc.n0, cb0);
          }
          if (!g.b_resident)
            bulk_load(b_base + (size_t)s * g.b_stage_bytes, wsrc + (size_t)kc * b_stage_elems, (uint32_t)g.b_stage_bytes,
                      &hdr->full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ===================================
    // The issuing thread, not the tensor pipe, bounds thin layers unless its loop is nearly straight-line code: the
    // pipe takes an M=128,K=16 MMA every 39 / 48 / 64 cycles at N = 16 / 64 / 128 whatever operands change between
    // instructions (tools/umma_bench4.cu), a loop with dependent descriptor arithmetic sustained one per 65-180.
    // So the per-chunk schedule {A offset, B offset} of up to 16 (slot, k16) pairs sits in registers (built once in
    // shared memory, hdr->sched) and each MMA tile replays it fully unrolled.
    uint32_t idesc = umma_idesc_16bit(128, (uint32_t)g.N_tile, GSX_FP16 ? 0u : 1u);
    uint32_t a_hi = (uint32_t)(umma_desc_hi((uint32_t)g.cb_stride_bytes, 128) >> 32);
    uint32_t b_hi = (uint32_t)(umma_desc_hi((uint32_t)g.N_tile * 16, 128) >> 32);
    const uint32_t a_lbo = (((uint32_t)g.cb_stride_bytes >> 4) & 0x3FFF) << 16;      // low-word part of the A descriptor
    const uint32_t b_lbo = ((((uint32_t)g.N_tile * 16) >> 4) & 0x3FFF) << 16;
    int n_k = g.n_k, stages = g.stages, n_mtiles = g.n_mtiles;
    const int sched_len = (GSX_TUNING && (g.dbg & 2)) ? 0 : g.n_slots * (g.CBK >> 1);      // MMAs per (k-chunk, MMA tile)
    const int b_res = g.b_resident;
    uint32_t a_stride = (uint32_t)g.a_stage_stride, b_stride = (uint32_t)g.b_stage_bytes;
    uint32_t n_tile = (uint32_t)g.N_tile;
    uint32_t mt_desc = (uint32_t)g.mt_stride;                          // MMA-tile pitch in 16-B units (= positions)
    keep_in_reg(mt_desc); keep_in_reg(idesc); keep_in_reg(a_hi); keep_in_reg(b_hi); keep_in_reg(n_k); keep_in_reg(stages);
    keep_in_reg(n_mtiles); keep_in_reg(a_stride); keep_in_reg(b_stride); keep_in_reg(n_tile);
    const uint32_t a_smem = smem_u32(a_base), b_smem = smem_u32(b_base);
    uint32_t bres_par = 0;
    if (g.b_resident && !g.per_sample) mbar_wait(&hdr->bres_full, 0);
    // Everything between the last MMA of a chunk and the first of the next is time the tensor pipe may run dry
    // (measured: ~1 us per work item with integer divisions and the tile decode in this path), so the stage /
    // buffer bookkeeping is incremental and only the phase (schedule selector, phase_grid layers) is decoded.
    int tl = 0;
    int cur_key = -1;
    int s = 0;                       // stage of the next chunk
    uint32_t full_par = 0;           // parity its "full" barrier completes with
    int buf = 0;                     // accumulator buffer of the next work item
    uint32_t empty_par = 1;          // parity of the "tmem_empty" wait = (round - 1) & 1, round = tl / nbuf
    const int sp_nt = g.tiles_x * g.tiles_y * g.tiles_n * g.n_ntiles;
    const bool phased = g.phase_grid != 0;
    uint32_t aa[16], bb[16], dc[16], id[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { aa[i] = 0; bb[i] = 0; dc[i] = 0; id[i] = idesc; }
    int w_next = t_first;                           // first tile of the next sample (per-sample resident weights)
    const int tiles_per_sample = g.tiles_x * g.tiles_y;
    for (int t = t_first; t < t_last; t += t_step, ++tl) {
      int phase = 0;
      if (phased) { int rem; phase = fast_div(t, sp_nt, rem); }
      if (b_res && g.per_sample && t >= w_next) {     // first tile of a sample: its weight set has to be resident
        mbar_wait(&hdr->bres_full, bres_par);
        bres_par ^= 1u;
        w_next = (t / tiles_per_sample + 1) * tiles_per_sample;
      }
      if (tl >= nbuf) mbar_wait(&hdr->tmem_empty[buf], empty_par);
      tc_fence_after();
      const uint32_t acc_base = tmem_base + (uint32_t)(buf * cols_per_buf);
      for (int kc = 0; kc < n_k; ++kc) {
        if (!(GSX_TUNING && (g.dbg & 8))) mbar_wait(&hdr->full[s], full_par);
        // descriptor low words: start address (16-B units) | LBO << 16
        const uint32_t a_lo0 = (((a_smem + (uint32_t)s * a_stride) >> 4) & 0x3FFF) | a_lbo;
        const uint32_t b_lo0 = (((b_smem + (uint32_t)(b_res ? kc : s) * b_stride) >> 4) & 0x3FFF) | b_lbo;
        for (int g0 = 0; g0 < sched_len; g0 += 16) {
          const int key = phase * 64 + g0;                              // schedule group held in aa / bb
          if (key != cur_key) {
            cur_key = key;
            const uint2* sc = hdr->sched + phase * sched_len + g0;      // (reads past the end are never issued)
            const uint2* sc2 = hdr->sched2 + phase * sched_len + g0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint2 v = sc[i];
              aa[i] = v.x; bb[i] = v.y;
              if (VARN) { const uint2 v2 = sc2[i]; dc[i] = v2.x; id[i] = v2.y; }
            }
          }
          const int cnt = sched_len - g0;
          // absolute descriptor low words of this chunk's schedule entries (the MMA tile offset is added per MMA)
          uint32_t ab[16], bk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) { ab[i] = a_lo0 + aa[i]; bk[i] = b_lo0 + bb[i]; }
          const uint32_t mt_step = mt_desc;
          if (elect_one()) {
            const uint32_t first_acc = (kc > 0 || g0 > 0) ? 1u : 0u;   // the very first MMA of a tile overwrites
            // straight-line bodies for the schedule lengths the planner produces; predicated fallback otherwise
            if (cnt >= 16)      issue_group<16, VARN>(acc_base, n_mtiles, n_tile, mt_step, ab, bk, dc, id, a_hi, b_hi, idesc, first_acc);
            else if (cnt == 9)  issue_group<9, VARN>(acc_base, n_mtiles, n_tile, mt_step, ab, bk, dc, id, a_hi, b_hi, idesc, first_acc);
            else if (cnt == 4)  issue_group<4, VARN>(acc_base, n_mtiles, n_tile, mt_step, ab, bk, dc, id, a_hi, b_hi, idesc, first_acc);
            else if (cnt == 8)  issue_group<8, VARN>(acc_base, n_mtiles, n_tile, mt_step, ab, bk, dc, id, a_hi, b_hi, idesc, first_acc);
            else if (cnt == 2)  issue_group<2, VARN>(acc_base, n_mtiles, n_tile, mt_step, ab, bk, dc, id, a_hi, b_hi, idesc, first_acc);
            else if (cnt == 1)  issue_group<1, VARN>(acc_base, n_mtiles, n_tile, mt_step, ab, bk, dc, id, a_hi, b_hi, idesc, first_acc);
            else {
              uint32_t moff = 0, d = acc_base;
              for (int mt = 0; mt < n_mtiles; ++mt, moff += mt_step, d += n_tile) {
#pragma unroll
                for (int k = 0; k < 16; ++k)
                  if (k < cnt) umma_f16kind_lohi(VARN ? d + dc[k] : d, ab[k] + moff, a_hi, bk[k], b_hi, VARN ? id[k] : idesc, k == 0 ? first_acc : 1u);
              }
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(&hdr->empty[s]);
          if (kc == n_k - 1) umma_commit(&hdr->tmem_full[buf]);
        }
        __syncwarp();
        if (++s == stages) { s = 0; full_par ^= 1u; }
      }
      if (++buf == nbuf) { buf = 0; empty_par ^= 1u; }
    }
  } else {
    // ================================ epilogue =====================================
    const ConvEpi& e = p.e;
    const int ew = warp - 2;                      // 0 .. kEpiWarps-1
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int egrp = ew >> 2;                     // which of the G warps sharing this quarter
    const int row = quarter * 32 + lane;
    const bool do_stats = GEN && (e.flags & EPI_STATS) != 0;
    const bool do_act = (e.flags & EPI_LRELU) != 0;
    const size_t plane_out = (size_t)e.Ho * e.Wo;
    const int n_chunks = g.N_tile >> 4;
    const int cpp = g.cout_tile >> 4;             // 16-channel chunks per phase block (== n_chunks unless up_cols)
    const int n_units = g.n_mtiles;
    const int slot_floats = 2 * g.cout_tile;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    pdl_wait();                                     // before the first global read / write of this role

    // per-channel epilogue operands (bias, noise scale) of all output channels of the layer: staged once per CTA in
    // shared memory and read as packed fp32 pairs (broadcast LDS.128) -- 32 registers less than holding the current
    // 16-channel block in registers, and no reload per tile
    float* chan_s = reinterpret_cast<float*>(smem + g.chan_off);      // [2][chan_n]: bias, noise scale
    for (int i = threadIdx.x - 64; i < g.chan_n; i += kEpiThreads) {
      chan_s[i] = (e.bias && i < e.Cout) ? __ldg(e.bias + i) : 0.f;
      chan_s[g.chan_n + i] = (GEN && e.nscale && i < e.Cout) ? __ldg(e.nscale + i) : 0.f;
    }
    named_bar_sync(1, kEpiThreads);
    const ulonglong2* const ns_s2 = reinterpret_cast<const ulonglong2*>(chan_s + g.chan_n);
    // per-sample layers (the producer's AdaIN folded into this conv): bias per accumulator column of the tile's sample,
    // in a private copy per warp -> refreshed by the warp itself whenever its tile belongs to another sample
    const bool per_sample = g.per_sample != 0;
    float* const bias_w = reinterpret_cast<float*>(smem + g.bias_w_off) + (size_t)ew * g.bias_cols;
    int bias_sample = -1;

    int tl = 0;
    for (int t = t_first; t < t_last; t += t_step, ++tl) {
      const TileCoord tc = decode_tile(g, t);
      const int buf = tl % nbuf;
      float* my_slot = stats_slots + ((size_t)(tl & 1) * kEpiWarps + ew) * slot_floats;
      mbar_wait_relaxed<64>(&hdr->tmem_full[buf], (uint32_t)((tl / nbuf) & 1));
      tc_fence_after();
      if (do_stats) {
        // (after the wait: the accumulators of this tile exist only once every warp has handed back the tile two
        //  before -- i.e. after the last reader of this slot buffer has finished, see the combine below)
        for (int i = lane; i < slot_floats; i += 32) my_slot[i] = 0.f;
        __syncwarp();
      }
      if (per_sample && tc.n0 != bias_sample) {
        bias_sample = tc.n0;
        __syncwarp();
        for (int i = lane; i < g.bias_cols; i += 32) bias_w[i] = ld_dep_f32(e.bias_n + (size_t)tc.n0 * g.bias_cols + i);
        __syncwarp();
      }
      // image-border class of this tile's rows (per-sample layers: the folded AdaIN shift must not flow in through taps
      // that fall outside the image; interior tiles skip the whole test)
      const bool tile_edge = per_sample && e.bdelta != nullptr &&
                             (tc.y0 == 0 || tc.y0 + g.TH >= g.H || tc.x0 == 0 || tc.x0 + g.TW >= g.W);
      const int ab = tl & 1;
      const uint8_t* aux = smem + g.aux_off + (size_t)ab * g.aux_bytes;
      if (g.aux_kind) mbar_wait_relaxed<32>(&hdr->aux_full[ab], (uint32_t)((tl >> 1) & 1));
      const uint32_t acc_base = tmem_base + (uint32_t)(buf * cols_per_buf) + lane_base;

      // position of this thread's row inside MMA tile `mt`:  q = mt*128 + row  ->  (nb, yl, xl)
      auto locate = [&](int mt) -> Loc {
        Loc l;
        const int q = mt * g.mt_stride + row;
        l.nb = (int)__umulhi((uint32_t)q, g.magic_box);
        const int rem = q - l.nb * (g.BH * g.BW);
        l.yl = (int)__umulhi((uint32_t)rem, g.magic_bw);
        l.xl = rem - l.yl * g.BW;
        l.n = tc.n0 + l.nb; l.y = tc.y0 + l.yl; l.x = tc.x0 + l.xl;
        l.valid = row < g.mt_stride && l.nb < g.NB && l.yl < g.TH && l.xl < g.TW && l.n < g.N && l.y < g.H && l.x < g.W;
        return l;
      };

      if (GSX_EPI_LITE || (GSX_TUNING && (g.dbg & 1))) {
      } else if (MODE == kEpiArgmax && g.up_cols) {
        // s2d final conv: the 4 output phases of a block are column groups of cout_tile (4, 8 or 16) classes
        for (int u = egrp; u < n_units; u += G) {
          const Loc lc = locate(u);
          const bool valid = lc.valid;
          const int n = lc.n, y = lc.y, x = lc.x;
          for (int cc = 0; cc < n_chunks; ++cc) {
            uint32_t v[16];
            tmem_ld16(acc_base + (uint32_t)(u * g.N_tile + cc * 16), v);
            tmem_ld_wait();
            if (valid) {
              if (g.cout_tile == 4) argmax_phases<4>(v, e, cc, n, y, x);
              else if (g.cout_tile == 8) argmax_phases<8>(v, e, cc, n, y, x);
              else argmax_phases<16>(v, e, cc, n, y, x);
            }
          }
        }
      } else if (MODE == kEpiArgmax) {
        for (int mt = egrp; mt < g.n_mtiles; mt += G) {
          uint32_t v[16];
          tmem_ld16(acc_base + (uint32_t)(mt * g.N_tile), v);
          tmem_ld_wait();
          const Loc lc = locate(mt);
          const int n = lc.n, y = lc.y, x = lc.x;
          if (lc.valid) {
            float best = __uint_as_float(v[0]) + (e.bias ? __ldg(e.bias) : 0.f);
            int arg = 0;
            const size_t pix = (size_t)y * e.Wo + x;
            if (e.logits) e.logits[((size_t)n * e.num_classes) * plane_out + pix] = best;
#pragma unroll
            for (int c = 1; c < 16; ++c) {
              if (c < e.num_classes) {
                const float lv = __uint_as_float(v[c]) + (e.bias ? __ldg(e.bias + c) : 0.f);
                if (e.logits) e.logits[((size_t)n * e.num_classes + c) * plane_out + pix] = lv;
                if (lv > best) { best = lv; arg = c; }       // first maximum wins (seg_solver.py:326)
              }
            }
            e.mask[(size_t)n * plane_out + pix] = (unsigned char)arg;
          }
        }
      } else if (MODE == kEpiUpCols) {
        // up-conv with the 4 phases as column blocks: the px=0 / px=1 outputs of a low-res pixel are adjacent in
        // the output row, so both phases are processed together and leave as ONE 32-byte store per channel block
        // (full sectors instead of two half-sector writes).  GEN adds the generator's first-half epilogue for the
        // folded deconv+blur: border correction, noise (tile staged at output resolution), statistics.
        // Instruction-count notes (ncu r01: this loop issued 24 thread-instructions per output value): the row is
        // located once per MMA tile (not per phase), the per-channel operands are (re)loaded only when the channel
        // block changes, the border test is a per-tile flag, and the arithmetic runs on packed fp32 pairs.
        const f32x2 slope2 = pk2(do_act ? 0.2f : 1.0f, do_act ? 0.2f : 1.0f);
        const bool has_res = !GEN && (g.aux_kind == 2 || e.addsrc != nullptr);
        const bool tile_border = GEN && e.e_rows != nullptr &&
                                 (tc.y0 == 0 || tc.y0 + g.TH >= g.H || tc.x0 == 0 || tc.x0 + g.TW >= g.W);
        for (int c16 = 0; c16 < cpp; ++c16) {
          const int c0 = tc.ntile * g.cout_tile + c16 * 16;
          if (c0 >= e.Cout) break;
          f32x2 s1[8], s2[8];
          if (GEN) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s1[i] = 0ull; s2[i] = 0ull; }
          }
          for (int u = egrp; u < n_units; u += G) {
            const Loc lc = locate(u);
            const uint32_t tbase = acc_base + (uint32_t)(u * g.N_tile + c16 * 16);
            // output pixel (2y+py, 2x) of this row; the noise tile is staged at output resolution
            const size_t pix0 = (size_t)(2 * lc.y) * e.Wo + 2 * lc.x;
            act_t* const obase = e.out + ((size_t)(c0 >> 3) * g.N + lc.n) * plane_out * 8;
            const float* const nzp = reinterpret_cast<const float*>(aux) + ((size_t)(lc.nb * 2 * g.TH + 2 * lc.yl) * (2 * g.TW) + 2 * lc.xl);
#pragma unroll
            for (int py = 0; py < 2; ++py) {
              uint32_t va[16], vb[16];
              const int blkA = phase_block(VARN ? 1 : 0, py, 0), blkB = phase_block(VARN ? 1 : 0, py, 1);    // accumulator blocks of the 2 phases
              tmem_ld16(tbase + (uint32_t)(blkA * cpp * 16), va);
              tmem_ld16(tbase + (uint32_t)(blkB * cpp * 16), vb);
              const int Y = 2 * lc.y + py;
              const size_t pix = pix0 + (size_t)py * e.Wo;
              f32x2 nz0 = 0ull, nz1 = 0ull;
              if (GEN && lc.valid) {
                float2 t2 = make_float2(0.f, 0.f);
                if (g.aux_kind == 1) t2 = *reinterpret_cast<const float2*>(nzp + (size_t)py * (2 * g.TW));
                else if (e.noise) t2 = ld_dep_f2(e.noise + (size_t)lc.n * plane_out + pix);
                nz0 = pk2(t2.x, t2.x); nz1 = pk2(t2.y, t2.y);
              }
              tmem_ld_wait();
              if (tile_edge) {
                const int cls = (lc.y == 0 ? 0 : (lc.y == g.H - 1 ? 2 : 1)) * 3 + (lc.x == 0 ? 0 : (lc.x == g.W - 1 ? 2 : 1));
                if (lc.valid && cls != 4) {
                  const float* dp = e.bdelta + ((size_t)lc.n * 9 + cls) * g.bias_cols + c16 * 16;
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    va[i] = __float_as_uint(__uint_as_float(va[i]) + ld_dep_f32(dp + blkA * cpp * 16 + i));
                    vb[i] = __float_as_uint(__uint_as_float(vb[i]) + ld_dep_f32(dp + blkB * cpp * 16 + i));
                  }
                }
              }
              if (tile_border) {
                // 1-pixel output border of the folded deconv+blur: subtract what the blur would have read from
                // outside the cropped deconv output (only tiles on the image border get here)
                const bool brow = lc.valid && (Y == 0 || Y == e.Ho - 1);
                const bool bl = lc.valid && lc.x == 0, br = lc.valid && lc.x == g.W - 1;
                if (brow) {
                  const float* er = e.e_rows + (((size_t)lc.n * 2 + (Y ? 1 : 0)) * e.Wo + 2 * lc.x) * e.Cout + c0;
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    va[i] = __float_as_uint(__uint_as_float(va[i]) - ld_dep_f32(er + i));
                    vb[i] = __float_as_uint(__uint_as_float(vb[i]) - ld_dep_f32(er + e.Cout + i));
                  }
                }
                if (bl) {
                  const float* ec = e.e_cols + (((size_t)lc.n * 2) * e.Ho + Y) * e.Cout + c0;
#pragma unroll
                  for (int i = 0; i < 16; ++i) va[i] = __float_as_uint(__uint_as_float(va[i]) - ld_dep_f32(ec + i));
                }
                if (br) {
                  const float* ec = e.e_cols + (((size_t)lc.n * 2 + 1) * e.Ho + Y) * e.Cout + c0;
#pragma unroll
                  for (int i = 0; i < 16; ++i) vb[i] = __float_as_uint(__uint_as_float(vb[i]) - ld_dep_f32(ec + i));
                }
              }
              if (lc.valid) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  // residual (s2d conv_b): one block-resolution vector per 8 channels, shared by the 4 output phases
                  uint4 rv = make_uint4(0, 0, 0, 0);
                  if (has_res) {
                    if (g.aux_kind == 2)
                      rv = reinterpret_cast<const uint4*>(aux)[((size_t)((c16 * 2 + h) * g.NB + lc.nb) * g.aux_bh + lc.yl) * g.aux_bw + lc.xl];
                    else
                      rv = ld_dep_u4(
                          e.addsrc + ((((size_t)(c0 >> 3) + h) * g.N + lc.n) * (plane_out >> 2) + (size_t)lc.y * (e.Wo >> 1) + lc.x) * 8);
                  }
                  const uint32_t r4[4] = {rv.x, rv.y, rv.z, rv.w};
                  // bias: per channel (shared table), or per column of the two phases (per-sample layers)
                  const ulonglong2* bpa = reinterpret_cast<const ulonglong2*>(per_sample ? bias_w + (blkA * cpp + c16) * 16 : chan_s + c0) + 2 * h;
                  const ulonglong2* bpb = reinterpret_cast<const ulonglong2*>(per_sample ? bias_w + (blkB * cpp + c16) * 16 : chan_s + c0) + 2 * h;
                  const ulonglong2 bq0 = bpa[0], bq1 = bpa[1], bq2 = bpb[0], bq3 = bpb[1];
                  const f32x2 bias4[4] = {bq0.x, bq0.y, bq1.x, bq1.y};
                  const f32x2 bias4b[4] = {bq2.x, bq2.y, bq3.x, bq3.y};
                  f32x2 ns4[4] = {0ull, 0ull, 0ull, 0ull};
                  if (GEN) {
                    const ulonglong2 nq0 = ns_s2[(c0 >> 2) + 2 * h], nq1 = ns_s2[(c0 >> 2) + 2 * h + 1];
                    ns4[0] = nq0.x; ns4[1] = nq0.y; ns4[2] = nq1.x; ns4[3] = nq1.y;
                  }
                  uint32_t o[8];
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const int j = h * 4 + k, i0 = 2 * j;
                    f32x2 a = add2(pk2u(va[i0], va[i0 + 1]), bias4[k]);
                    f32x2 b = add2(pk2u(vb[i0], vb[i0 + 1]), bias4b[k]);
                    if (GEN) { a = fma2(ns4[k], nz0, a); b = fma2(ns4[k], nz1, b); }
                    const f32x2 am = mul2(a, slope2), bm = mul2(b, slope2);
                    float a0, a1, b0, b1, m0, m1, m2, m3;
                    upk2(a, a0, a1); upk2(b, b0, b1); upk2(am, m0, m1); upk2(bm, m2, m3);
                    a0 = fmaxf(a0, m0); a1 = fmaxf(a1, m1); b0 = fmaxf(b0, m2); b1 = fmaxf(b1, m3);
                    if (has_res) {
                      const float2 r2 = unpack_x2(r4[k]);
                      a0 += r2.x; a1 += r2.y; b0 += r2.x; b1 += r2.y;
                    }
                    if (GEN) {
                      const f32x2 ar = pk2(a0, a1), br2 = pk2(b0, b1);
                      s1[j] = add2(s1[j], add2(ar, br2));
                      s2[j] = fma2(ar, ar, fma2(br2, br2, s2[j]));
                    }
                    o[k] = pack_x2(a0, a1);
                    o[4 + k] = pack_x2(b0, b1);
                  }
                  if (e.out_planar) {       // planes (py,0) and (py,1) of the block-resolution grid
                    act_t* pb = obase + ((size_t)h * g.N * plane_out + ((size_t)(2 * py) * g.H + lc.y) * g.W + lc.x) * 8;
                    *reinterpret_cast<uint4*>(pb) = make_uint4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<uint4*>(pb + (size_t)g.H * g.W * 8) = make_uint4(o[4], o[5], o[6], o[7]);
                  } else {
                    st_global_256(obase + ((size_t)h * g.N * plane_out + pix) * 8, o);
                  }
                }
              }
            }
          }
          if (do_stats) {
            float f1[16], f2[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) { upk2(s1[j], f1[2 * j], f1[2 * j + 1]); upk2(s2[j], f2[2 * j], f2[2 * j + 1]); }
            my_slot[(c16 * 16 + (lane & 15)) * 2 + (lane >> 4)] += warp_reduce_32x32(f1, f2, lane);
          }
        }
      } else {
        const f32x2 slope2 = pk2(do_act ? 0.2f : 1.0f, do_act ? 0.2f : 1.0f);
        const bool has_res = g.aux_kind == 2 || e.addsrc != nullptr;
        for (int cc = 0; cc < n_chunks; ++cc) {
          const int cl = cc * 16;                                // first channel of the chunk inside the CTA's block
          const int c0 = tc.ntile * g.cout_tile + cl;            // first output channel of this chunk
          if (c0 >= e.Cout) break;
          f32x2 s1[8], s2[8];
          if (GEN) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s1[i] = 0ull; s2[i] = 0ull; }
          }
          // MMA tiles are dealt round-robin to the G warps of a quarter.  The per-pixel operands of the epilogue
          // (noise value, residual vectors) come from the tile the producer staged in smem by TMA; the global
          // loads below are only the fallback for layers the planner could not stage.
          for (int u = egrp; u < n_units; u += G) {
            uint32_t v[16];
            tmem_ld16(acc_base + (uint32_t)(u * g.N_tile + cc * 16), v);
            const Loc lc = locate(u);
            int y = lc.y, x = lc.x;
            if (e.up) { y = 2 * y + (tc.phase >> 1); x = 2 * x + (tc.phase & 1); }
            const size_t pix = (size_t)y * e.Wo + x;
            // phase-planar output: plane (y&1, x&1) of the block grid
            const size_t pix_st = e.out_planar ? ((size_t)((y & 1) * 2 + (x & 1)) * (e.Ho >> 1) + (y >> 1)) * (e.Wo >> 1) + (x >> 1) : pix;
            f32x2 nz = 0ull;
            uint4 add0 = make_uint4(0, 0, 0, 0), add1 = add0;
            if (lc.valid) {
              if (GEN) {
                float t1 = 0.f;
                if (g.aux_kind == 1) t1 = reinterpret_cast<const float*>(aux)[(lc.nb * g.TH + lc.yl) * g.TW + lc.xl];
                else if (e.noise) t1 = ld_dep_f32(e.noise + (size_t)lc.n * plane_out + pix);
                nz = pk2(t1, t1);
              }
              if (g.aux_kind == 2) {
                // residual tile [cb][nb][row/2][col/2] of 16-B vectors at half resolution
                const int yy = (y >> 1) - (tc.y0 >> 1), xx = (x >> 1) - (tc.x0 >> 1);
                const uint4* ap = reinterpret_cast<const uint4*>(aux) +
                                  ((size_t)((cl >> 3) * g.NB + lc.nb) * g.aux_bh + yy) * g.aux_bw + xx;
                add0 = ap[0];
                add1 = ap[(size_t)g.NB * g.aux_bh * g.aux_bw];
              } else if (e.addsrc) {
                const size_t plane_lo = (size_t)(e.Ho >> 1) * (e.Wo >> 1);
                const size_t pl = (size_t)(y >> 1) * (e.Wo >> 1) + (x >> 1);
                const act_t* ap = e.addsrc + (((size_t)(c0 >> 3) * g.N + lc.n) * plane_lo + pl) * 8;
                add0 = ld_dep_u4(ap);
                add1 = ld_dep_u4(ap + (size_t)g.N * plane_lo * 8);
              }
            }
            tmem_ld_wait();
            if (tile_edge) {
              const int cls = (lc.y == 0 ? 0 : (lc.y == g.H - 1 ? 2 : 1)) * 3 + (lc.x == 0 ? 0 : (lc.x == g.W - 1 ? 2 : 1));
              if (lc.valid && cls != 4) {
                const float* dp = e.bdelta + ((size_t)lc.n * 9 + cls) * g.bias_cols + cc * 16;
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + ld_dep_f32(dp + i));
              }
            }
            if (lc.valid) {
              const uint32_t w8[8] = {add0.x, add0.y, add0.z, add0.w, add1.x, add1.y, add1.z, add1.w};
              uint32_t o[8];
              f32x2 bias8[8], ns8[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const ulonglong2 bq = reinterpret_cast<const ulonglong2*>(per_sample ? bias_w + cc * 16 : chan_s + c0)[q];
                bias8[2 * q] = bq.x; bias8[2 * q + 1] = bq.y;
                ns8[2 * q] = 0ull; ns8[2 * q + 1] = 0ull;
                if (GEN) { const ulonglong2 nq = ns_s2[(c0 >> 2) + q]; ns8[2 * q] = nq.x; ns8[2 * q + 1] = nq.y; }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                f32x2 a = add2(pk2u(v[2 * j], v[2 * j + 1]), bias8[j]);
                if (GEN) a = fma2(ns8[j], nz, a);
                const f32x2 am = mul2(a, slope2);
                float a0, a1, m0, m1;
                upk2(a, a0, a1); upk2(am, m0, m1);
                a0 = fmaxf(a0, m0); a1 = fmaxf(a1, m1);
                if (has_res) {
                  const float2 b2 = unpack_x2(w8[j]);
                  a0 += b2.x; a1 += b2.y;
                }
                if (GEN) {
                  const f32x2 ar = pk2(a0, a1);
                  s1[j] = add2(s1[j], ar);
                  s2[j] = fma2(ar, ar, s2[j]);
                }
                o[j] = pack_x2(a0, a1);
              }
              act_t* op = e.out + (((size_t)(c0 >> 3) * g.N + lc.n) * plane_out + pix_st) * 8;
              *reinterpret_cast<uint4*>(op) = make_uint4(o[0], o[1], o[2], o[3]);
              *reinterpret_cast<uint4*>(op + (size_t)g.N * plane_out * 8) = make_uint4(o[4], o[5], o[6], o[7]);
            }
          }
          if (do_stats) {
            // value index = lane: 0..15 channel sums, 16..31 sums of squares; this warp's own slot: plain add
            float f1[16], f2[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) { upk2(s1[j], f1[2 * j], f1[2 * j + 1]); upk2(s2[j], f2[2 * j], f2[2 * j + 1]); }
            my_slot[(cl + (lane & 15)) * 2 + (lane >> 4)] += warp_reduce_32x32(f1, f2, lane);
          }
        }
      }

      if (do_stats) {
        // Fused stats need a single sample per CTA tile (NB == 1); enforced by the callers.  One partial per
        // (sample, tile), summed over the per-warp slots in a fixed order by whichever warp finishes the tile last
        // (a counter instead of a barrier: nobody waits) -> bit-reproducible, finalize_kernel adds the tiles up.
        // The slot buffer (tl & 1) is reused two tiles later; by then every warp has handed this tile's accumulator
        // back (below, AFTER the combine), which the MMAs of that later tile wait for.
        __syncwarp();
        uint32_t arrived = 0;
        if (lane == 0) { __threadfence_block(); arrived = atomicAdd(&hdr->stats_cnt[tl & 1], 1u); }
        arrived = __shfl_sync(0xffffffffu, arrived, 0);
        if (arrived == (uint32_t)kEpiWarps - 1) {
          __threadfence_block();
          const float* slots = stats_slots + (size_t)(tl & 1) * kEpiWarps * slot_floats;
          for (int i = lane; i < slot_floats; i += 32) {
            const int ch = tc.ntile * g.cout_tile + (i >> 1);
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kEpiWarps; ++w) s += slots[w * slot_floats + i];
            if (ch < e.Cout) e.stats[(((size_t)tc.n0 * e.stats_T + tc.tile_in_sample) * e.Cout + ch) * 2 + (i & 1)] = s;
          }
          __syncwarp();
          if (lane == 0) hdr->stats_cnt[tl & 1] = 0;
        }
      }

      // accumulator buffer drained -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&hdr->tmem_empty[buf]);
        if (g.aux_kind) mbar_arrive(&hdr->aux_empty[ab]);
      }
    }
  }

  if (pdl_late) pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)g.tmem_cols);
}

// cudaFuncSetAttribute is per device: one process may drive several GPUs (ImageGenerator(gpu_ids=[0,1,..])).
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev < 0 ? 0 : (dev > 63 ? 63 : dev);
}

template <int G, bool GEN, int MODE, bool VARN = false>
static void launch_g(const ConvParams& p, int grid, cudaStream_t st) {
  static bool configured[64] = {false};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaFuncSetAttribute(shiftconv_kernel<G, GEN, MODE, VARN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    configured[dev] = true;
  }
  launch_pdl_kind(1, shiftconv_kernel<G, GEN, MODE, VARN>, dim3(grid), dim3(64 + 128 * G), (size_t)p.g.smem_bytes, st, p);
}

void launch_shiftconv(const ConvParams& p, cudaStream_t st) {
  static int num_sms[64] = {0};
  const int dev = current_device();
  if (!num_sms[dev]) cudaDeviceGetAttribute(&num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
  const ConvGeom& g = p.g;
  const int total = g.tiles_x * g.tiles_y * g.tiles_n * g.n_ntiles * (g.phase_grid ? 4 : 1);
  const int grid = total < num_sms[dev] * g.ctas_per_sm ? total : num_sms[dev] * g.ctas_per_sm;
#if GSX_TUNING
  static const int dbg = getenv("GSX_DBG") ? atoi(getenv("GSX_DBG")) : 0;
  ConvParams q = p;
  if (dbg) q.g.dbg = dbg;
  const ConvParams& pp = q;
#else
  const ConvParams& pp = p;
#endif
  const bool gen = p.e.noise != nullptr || p.e.nscale != nullptr || (p.e.flags & EPI_STATS) != 0 || p.e.e_rows != nullptr;
  const bool g4 = g.epi_groups == 4;
  if (p.e.flags & EPI_ARGMAX) { if (g4) launch_g<4, false, kEpiArgmax>(pp, grid, st); else launch_g<2, false, kEpiArgmax>(pp, grid, st); }
  else if (gen && g.up_cols) {                                                  // generator epilogues: noise + statistics (+ border)
    if (g.varn) launch_g<2, true, kEpiUpCols, true>(pp, grid, st); else launch_g<2, true, kEpiUpCols>(pp, grid, st);
  }
  else if (gen) launch_g<2, true, kEpiGeneric>(pp, grid, st);
  else if (g.up_cols && g.varn) launch_g<2, false, kEpiUpCols, true>(pp, grid, st);   // variable-N MMAs: its own instantiation, so that
                                                                                      //   the uniform-N issue loop of every other layer is untouched
  else if (g.up_cols) { if (g4) launch_g<4, false, kEpiUpCols>(pp, grid, st); else launch_g<2, false, kEpiUpCols>(pp, grid, st); }
  else { if (g4) launch_g<4, false, kEpiGeneric>(pp, grid, st); else launch_g<2, false, kEpiGeneric>(pp, grid, st); }
}

}  // namespace gsx
