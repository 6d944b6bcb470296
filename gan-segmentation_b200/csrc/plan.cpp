// Host side of the shift-GEMM convolution: tile planning, weight packing into the UMMA B-operand
// stream, TMA tensor-map encoding.  Compiled by nvcc as C++ (no device code).
#include "gsx_internal.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace gsx {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error_cstr() { return g_err.c_str(); }
bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return false;
}

// Programmatic dependent launch.  Round 1: every kernel released its dependents at its START; that pays when the kernels
// are short (batch 1 at 256^2: 1.19 -> 1.07 ms per step) but costs ~2 % at the large-batch operating point (a conv CTA that
// becomes resident early pins the shared-memory carve-out while the HBM-bound pass before it drains), so large passes used
// it only for conv -> conv chains (mode 1).  Round 2 (mode 3, the default): the big kernels (convs, pass1, apply, ToRGB)
// release their dependents when a CTA has FINISHED its work, so a successor only overlaps its launch latency and prologue
// with the predecessor's tail -- and every launch carries the attribute.  Measured (device / end-to-end samples/s):
// FFHQ 3129 / 3033 -> 3115 / 3074, cars 6980 / 6802 -> 7047 / 6962, bedrooms batch 1 901 / 893 -> 923 / 915.
static thread_local bool g_pdl_small = false, g_pdl_chain = false, g_prev_conv = false;
int g_pdl_mode = 3;     // gsx_set_option("pdl", v): 0 off / 1 by size (+ conv chains) / 2 always / 3 always, big kernels trigger late /
                        //   4 = policy of 1 with the late trigger
void pdl_set_for_work(double top_level_pixels) {
  if (top_level_pixels < 0) { g_pdl_small = g_pdl_chain = false; return; }      // plain stream order (the training step)
  const int mode = g_pdl_mode;
  g_pdl_small = mode == 2 || mode == 3 || ((mode == 1 || mode == 4) && top_level_pixels <= 2.0 * 1024 * 1024);
  g_pdl_chain = mode != 0;
}
bool pdl_enabled(int kind) {
  const bool on = g_pdl_small || (g_pdl_chain && kind == 1 && g_prev_conv);
  g_prev_conv = (kind == 1);
  return on;
}

namespace {
struct PoolBlock { void* p; size_t size; bool used; int dev; };
thread_local std::vector<PoolBlock> g_pool;
int cur_dev() { int d = 0; cudaGetDevice(&d); return d; }
}
// Blocks belong to the device they were allocated on: one host thread may drive several GPUs
// (ImageGenerator(gpu_ids=[0,1,..]) / torch.cuda.device guards), and a block of another device must never be handed out.
void* pool_get(size_t bytes) {
  if (bytes == 0) bytes = 256;
  const int dev = cur_dev();
  int best = -1;
  for (int i = 0; i < (int)g_pool.size(); ++i) {
    const PoolBlock& b = g_pool[i];
    if (!b.used && b.dev == dev && b.size >= bytes && b.size <= 2 * bytes + (1 << 20) && (best < 0 || b.size < g_pool[best].size)) best = i;
  }
  if (best >= 0) { g_pool[best].used = true; return g_pool[best].p; }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    pool_release();                                   // out of memory: drop the cached blocks and retry once
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  }
  g_pool.push_back({p, bytes, true, dev});
  return p;
}
void pool_put(void* p) {
  if (!p) return;
  for (auto& b : g_pool)
    if (b.p == p) { b.used = false; return; }
  cudaFree(p);
}
void pool_release() {
  const int dev = cur_dev();
  std::vector<PoolBlock> keep;
  for (auto& b : g_pool) {
    if (b.used) { keep.push_back(b); continue; }
    if (b.dev != dev) cudaSetDevice(b.dev);           // free under the owning device
    cudaFree(b.p);
    if (b.dev != dev) cudaSetDevice(dev);
  }
  g_pool.swap(keep);
}

int g_plan_epi_groups = 2;
int g_plan_varn = 1;
static const int kSmemLimit = 227 * 1024;
static const int kHeader = kConvHeaderBytes;
static const int kSlack = 8192;     // garbage-tolerant over-read of the last (partial) MMA tile

static int ceil_div(int a, int b) { return (a + b - 1) / b; }
static int pow2_cols(int c) {
  int p = 32;
  while (p < c) p <<= 1;
  return p;
}

void plan_conv(ConvLayer& L, int mode, int H, int W, int cin0, int cin1, int cout, int argmax_classes,
               const PlanOverride* ov, int aux_kind, int in_planar, int per_sample) {
  L.mode = mode; L.H = H; L.W = W; L.cin0 = cin0; L.cin1 = cin1; L.cout = cout;
  ConvGeom& g = L.g;
  std::memset(&g, 0, sizeof(g));
  // Thin 3x3 layers (<= 32 channels in and out) are bound by the NUMBER of MMAs: an M=128,K=16 kind::f16 MMA costs
  // 42 cycles of the tensor pipe at N=16, 50 at N=64, 64 at N=128 (tools/umma_bench2.cu) -- almost independent of N.
  // Space-to-depth: rows of the GEMM are 2x2 pixel blocks, the 4 input phases are 4x the K planes and the 4 output
  // phases 4x the N columns.  A block tap (ty,tx) in {-1,0,1}^2 only needs the input phases within one pixel of the
  // block (1, 2 or 4 of them), 16 (tap, phase-plane) pairs in all: 16*Cin/16 MMAs per 512 output pixels instead of
  // 36*Cin/16 -- 2.25x fewer, each at N = 4*Cout.
  // Measured (r01, FFHQ batch 32): the phase-plane gather (TMA boxes with a 16-byte inner extent) is slow -- loads
  // alone take as long as the whole dense-layout kernel -- so by default only the final conv + argmax uses it
  // (0.85 -> 0.73 ms); GSX_S2D=1 enables it for every eligible layer.
  static const int s2d_all = tune_env("GSX_S2D") ? atoi(tune_env("GSX_S2D")) : 0;
  int s2d = (mode == CONV3 && cin0 + cin1 <= 32 && cout <= 32 && H % 2 == 0 && W % 2 == 0 && H >= 16 && W >= 16 &&
             (argmax_classes > 0 || in_planar || (s2d_all && cout <= 16 && aux_kind != 2)) && !tune_env("GSX_NO_S2D")) ? 1 : 0;
  if (ov && ov->s2d >= 0 && mode == CONV3 && H % 2 == 0 && W % 2 == 0) s2d = ov->s2d;
  if (in_planar && !s2d) { set_error("plan_conv: a phase-planar input needs the space-to-depth plan"); return; }
  if (s2d) { H /= 2; W /= 2; }               // from here on H, W, TH, TW count blocks
  g.H = H; g.W = W;
  const bool up = (mode == UPCONV3 || mode == DECONV4 || mode == DECONV4B);
  const int cb0 = cin0 / 8, cb1 = cin1 / 8, cbt = cb0 + cb1;

  int N_tile = argmax_classes > 0 ? 16 : std::min(cout, 128);
  if (s2d && argmax_classes > 0) N_tile = argmax_classes <= 4 ? 4 : (argmax_classes <= 8 ? 8 : 16);
  if (ov && ov->N_tile > 0) N_tile = ov->N_tile;
  // phases as work items for cout >= 64 (stacked along N they would need the whole 9-shift weight set per tile:
  // 590 KB streamed per 256 pixels at 128->64, measured 0.36 ms vs 0.25 ms), as column blocks of one MMA below
  int phase_grid = (up && cout >= 64) ? 1 : 0;
  if (ov && ov->phase_grid >= 0 && up) phase_grid = ov->phase_grid;
  if (mode == DECONV4B) {
    if (cout > 64) { set_error("plan_conv: DECONV4B needs cout <= 64 (phases stacked along N)"); return; }
    phase_grid = 0;
  }
  const int up_cols = ((up && !phase_grid) || s2d) ? 1 : 0;
  const int cout_tile = N_tile;
  if (up_cols) N_tile = 4 * cout_tile;        // phases stacked along the MMA N dimension (<= 256)
  // (A variant stacking the 3 horizontal taps of a 3x3 layer along N -- 3 MMAs per k16 step, the epilogue forming
  //  out[q] = acc[q][0] + acc[q+1][1] + acc[q+2][2] by shuffles -- was parity-green in round 1 but 2-2.5x slower and cost
  //  the plain epilogue 40 registers just by being compiled in; removed, git history "hstack".  The override field
  //  stays in the ABI of the tuning hook.)
  if (ov && ov->hstack > 0) { set_error("plan_conv: hstack mode is not compiled into this build"); return; }
  const int mt_stride = 128;
  const int G = 1;
  const bool thin = (cin0 + cin1) <= 64 && cout <= 64;
  // thin layers: 256 columns per accumulator buffer so that two buffers fit (MMA / epilogue overlap);
  // wide layers: all 512 columns for one buffer (more MMA rows per weight load; the epilogue is a small
  // fraction of the k-loop there)
  int budget_cols = (thin && N_tile < 128) ? 256 : 512;     // (N >= 128: 4 tiles in one buffer beat 2+2, measured)
  if (ov && ov->acc_bufs == 1) budget_cols = 512;
  if (ov && ov->acc_bufs == 2) budget_cols = 256;
  int max_mt = std::max(1, budget_cols / (G * N_tile));
  max_mt = std::min(max_mt, 32);
  if (ov && ov->max_mtiles > 0) max_mt = std::min(max_mt, ov->max_mtiles);
  // 8 epilogue warps.  16 were measured in round 1 (one and two MMA issuers, dense and space-to-depth plans) and again in
  // round 2 (kernels without the generator epilogue: 3076 vs 3152 samples/s; generator epilogue of the wide layers: spills
  // under its register cap, g7.conv2 0.252 -> 0.320 ms) -- never faster.  gsx_set_option("epi_groups", 4) keeps the
  // 16-warp instantiations of the kernels without the generator epilogue reachable for A/B runs.
  const int epi_groups = g_plan_epi_groups;
  if (ov && ov->epi_groups > 0 && ov->epi_groups != 2) { set_error("plan_conv: only epi_groups = 2 is built"); return; }
  const int stats_bytes = (2 * 4 * epi_groups * 2 * cout_tile * 4 + 1023) / 1024 * 1024;
  // per-channel epilogue operands of every output channel (bias, noise scale), staged once per CTA
  const int chan_n = std::max(16, ceil_div(argmax_classes > 0 ? argmax_classes : cout, cout_tile) * cout_tile);
  // per-sample bias: one private copy per epilogue warp (no cross-warp synchronisation when the sample changes)
  const int bias_cols = per_sample ? ceil_div(cout, cout_tile) * N_tile : 0;
  const int chan_bytes = (2 * chan_n * 4 + 4 * epi_groups * bias_cols * 4 + 1023) / 1024 * 1024;
  const int hdr_bytes = kHeader + stats_bytes + chan_bytes;

  int TW = (W <= 126) ? W : ((W % 64 == 0) ? 64 : 126);
  if (ov && ov->TW > 0) TW = ov->TW;
  const int BW = TW + 2;
  const int cap = max_mt * mt_stride;
  int NB = 1, TH;
  if (TW == W && (H + 2) * BW * 2 <= cap) {
    TH = H;
    NB = std::min(cap / ((H + 2) * BW), 64);
    // deep-K layers on tiny images (4x4 / 8x8, 256+ input channels): one MMA tile per work item, so that the
    // batch spreads over many CTAs instead of one CTA walking a 4608-deep K loop for everybody
    if (cin0 + cin1 >= 256) NB = std::max(1, std::min(NB, mt_stride / ((H + 2) * BW)));
  } else {
    int th_max = (cap - TW) / BW + 1;
    th_max = std::max(1, std::min(th_max, std::min(H, 254)));
    const int ty = ceil_div(H, th_max);
    TH = ceil_div(H, ty);
  }
  if (ov && ov->TH > 0) TH = ov->TH;
  if (ov && ov->NB > 0) NB = ov->NB;

  const int n_slots = s2d ? 16 : (mode == CONV3) ? 9 : (mode == CONV1 ? 1 : (phase_grid ? 4 : 9));
  const int planes = s2d ? 4 : 1;

  // aux staging: plain 3x3 layers (noise / residual tile) and the folded deconv (noise tile at output resolution);
  // the TMA box row must be a multiple of 16 bytes
  const int aux_up = (mode == DECONV4B || s2d) ? 1 : 0;
  if (mode != CONV3 && mode != DECONV4B) aux_kind = 0;
  if (mode == DECONV4B && aux_kind == 2) aux_kind = 0;
  if (aux_kind == 1 && ((((W << aux_up) * 4) % 16) != 0 || (((TW << aux_up) * 4) % 16) != 0)) aux_kind = 0;
  if (aux_kind == 2 && !s2d && (TW & 1)) aux_kind = 0;
  // k-chunk depth and pipeline stages under the shared-memory limit
  int CBK = 0, stages = 0, b_resident = 0;
  long aux_bytes = 0;
  for (;;) {
    const int BH = TH + 2;
    aux_bytes = aux_kind == 1 ? ((long)NB * TH * TW * 4) << (2 * aux_up)
              : aux_kind == 2 ? (s2d ? (long)NB * TH * TW * 16 * (cout_tile / 8)
                                     : (long)NB * (TH / 2 + 1) * (TW / 2) * 16 * (cout_tile / 8)) : 0;
    aux_bytes = (aux_bytes + 127) / 128 * 128;
    int cbs[4] = {8, 4, 2, 0};
    if (ov && ov->CBK > 0) { cbs[0] = ov->CBK; cbs[1] = 0; }
    bool found = false;
    const long lim = kSmemLimit;                          // one persistent CTA per SM
    for (int ci = 0; cbs[ci] && !found; ++ci) {
      const int c = cbs[ci];
      if (c > cbt || cb0 % c || (cb1 && cb1 % c)) continue;
      const int n_k = cbt / c;
      // TMA smem destinations are 128-B aligned (s2d: every phase plane is its own destination)
      const long a_st = planes * (((long)NB * BH * BW * 16 * c + 127) / 128 * 128);
      const long b_st = (long)n_slots * (c / 2) * N_tile * 32;
      if ((long)NB * BH * BW * 16 >= (1 << 18)) continue;                 // LBO field is 14 bits of 16-byte units
      // the stage ring runs across work items, so even single-chunk layers want >= 2 stages.
      // Weights stay resident in smem (loaded once per CTA) when one CTA only ever needs one weight set
      // and it leaves room for >= 2 activation stages.
      const bool one_set = !phase_grid && (argmax_classes > 0 || cout <= cout_tile);
      const int s_hi = (ov && ov->stages > 0) ? ov->stages : (n_k == 1 ? 3 : 4);
      const int s_lo = (ov && ov->stages > 0) ? ov->stages : 2;
      for (int pass = 0; pass < 2 && !found; ++pass) {
        const bool bres = (pass == 0) && one_set && (long)n_k * b_st <= 160 * 1024;
        if (pass == 0 && !bres) continue;
        for (int s = s_hi; s >= s_lo && s >= 1; --s) {
          const long tot = hdr_bytes + 2 * aux_bytes + (bres ? s * a_st + n_k * b_st : s * (a_st + b_st)) + kSlack;
          if (tot <= lim) { CBK = c; stages = s; b_resident = bres ? 1 : 0; found = true; break; }
        }
      }
    }
    if (found) break;
    if (NB > 1) { NB = std::max(1, NB / 2); continue; }
    if (TH > 1) { const int t2 = ceil_div(H, ceil_div(H, TH) + 1); TH = (t2 < TH) ? t2 : TH - 1; continue; }
    set_error("plan_conv: no feasible tiling");
    return;
  }

  g.TH = TH; g.TW = TW; g.NB = NB; g.BH = TH + 2; g.BW = BW;
  g.tiles_x = ceil_div(W, TW); g.tiles_y = ceil_div(H, TH);
  g.CBK = CBK; g.kch0 = cb0 / CBK; g.n_k = cbt / CBK;
  g.N_tile = N_tile; g.n_ntiles = ceil_div(argmax_classes > 0 ? argmax_classes : cout, cout_tile);
  g.up_cols = up_cols; g.cout_tile = cout_tile;
  g.n_groups = G; g.n_slots = n_slots; g.phase_grid = phase_grid; g.stages = stages;
  g.n_mtiles = ceil_div(((NB - 1) * g.BH + TH - 1) * BW + TW, mt_stride);
  g.mt_stride = mt_stride;
  g.cb_stride_bytes = NB * g.BH * BW * 16;
  g.a_stage_bytes = g.cb_stride_bytes * CBK * planes;
  g.s2d = s2d; g.in_planar = in_planar;
  // variable-N MMAs for the stacked-phase up-convs with >= 32 output channels (measured r02: d6.conv_a 0.367 -> 0.344 ms,
  // g9.deconv 0.400 -> 0.381; at 16 channels the narrower MMAs save less than the per-MMA issue overhead they add:
  // d7.conv_a 0.748 -> 0.773; gsx_set_option("varn", 2) forces it everywhere, 0 switches it off)
  g.varn = (up_cols && !s2d && (mode == UPCONV3 || mode == DECONV4) && (g_plan_varn == 2 || (g_plan_varn == 1 && cout_tile >= 32))) ? 1 : 0;
  g.plane_stride = (g.cb_stride_bytes * CBK + 127) / 128 * 128;
  g.a_stage_stride = planes * g.plane_stride;
  g.b_stage_bytes = n_slots * (CBK / 2) * N_tile * 32;
  const int cols = G * g.n_mtiles * N_tile;
  g.acc_bufs = (2 * cols <= 512) ? 2 : 1;
  if (ov && ov->acc_bufs == 1) g.acc_bufs = 1;
  g.tmem_cols = pow2_cols(g.acc_bufs * cols);
  if (cols > 512) { set_error("plan_conv: TMEM budget exceeded"); return; }
  g.b_resident = b_resident;
  g.epi_groups = epi_groups;
  g.ctas_per_sm = 1;
  g.chan_off = kHeader + stats_bytes; g.chan_n = chan_n;
  g.per_sample = per_sample; g.bias_cols = bias_cols; g.bias_w_off = g.chan_off + 2 * chan_n * 4;
  if (per_sample && (NB != 1 || phase_grid || g.n_ntiles != 1 || argmax_classes > 0)) { set_error("plan_conv: per-sample weights need NB = 1, one N tile, no phase grid"); return; }
  g.aux_kind = aux_kind; g.aux_off = hdr_bytes; g.aux_bytes = (int)aux_bytes;
  g.aux_bw = s2d ? TW : TW / 2; g.aux_bh = s2d ? TH : TH / 2 + 1;
  g.aux_up = aux_up; g.aux_shift = s2d ? 0 : 1;
  g.aux_bytes_tx = aux_kind == 1 ? (NB * TH * TW * 4) << (2 * aux_up)
                 : aux_kind == 2 ? NB * g.aux_bh * g.aux_bw * 16 * (cout_tile / 8) : 0;
  g.a_off = hdr_bytes + 2 * (int)aux_bytes;
  g.magic_box = (unsigned)((0x100000000ULL + (unsigned long long)(g.BH * BW) - 1) / (unsigned long long)(g.BH * BW));
  g.magic_bw = (unsigned)((0x100000000ULL + (unsigned long long)BW - 1) / (unsigned long long)BW);
  g.smem_bytes = g.a_off + stages * g.a_stage_stride + (g.b_resident ? g.n_k : stages) * g.b_stage_bytes + kSlack;

  for (int ph = 0; ph < 4; ++ph)
    for (int s = 0; s < kMaxSlots; ++s) g.slot_shift[ph][s] = 0;
  if (s2d) {
    // per axis the (block tap t, input phase p) pairs that lie within one pixel of the block: (-1,1) (0,0) (0,1) (+1,0)
    static const int kT[4] = {-1, 0, 0, 1}, kP[4] = {1, 0, 1, 0};
    const int plane_pos = g.plane_stride / 16;                     // positions (16 B) per phase plane of a stage
    for (int sy = 0; sy < 4; ++sy)
      for (int sx = 0; sx < 4; ++sx) {
        const int s = sy * 4 + sx;
        g.slot_shift[0][s] = (kP[sy] * 2 + kP[sx]) * plane_pos + (kT[sy] + 1) * BW + (kT[sx] + 1);
        g.slot_group[s] = 0; g.slot_first[s] = (s == 0);
      }
  } else if (mode == CONV3 || up_cols) {
    for (int s = 0; s < 9; ++s) {
      int ky, kx;
      slot_yx(g, s, &ky, &kx);
      g.slot_shift[0][s] = (short)(ky * BW + kx); g.slot_group[s] = 0; g.slot_first[s] = (s == 0);
      g.slot_n0[s] = 0; g.slot_n[s] = 4;
      if (g.varn) {
        // phases fed by shift (ky,kx): py in {ky-1, ky} /\ {0,1}, px likewise; as a range of the cyclic block order
        const int py0 = ky == 2 ? 1 : 0, py1 = ky == 0 ? 0 : 1, px0 = kx == 2 ? 1 : 0, px1 = kx == 0 ? 0 : 1;
        int lo = 4, hi = -1, cnt = 0;
        for (int py = py0; py <= py1; ++py)
          for (int px = px0; px <= px1; ++px) { const int b = phase_block(1, py, px); lo = std::min(lo, b); hi = std::max(hi, b); ++cnt; }
        if (hi - lo + 1 == cnt) { g.slot_n0[s] = (signed char)lo; g.slot_n[s] = (signed char)cnt; }      // contiguous (all but (1,0))
      }
    }
  } else if (mode == CONV1) {
    g.slot_shift[0][0] = (short)(BW + 1); g.slot_group[0] = 0; g.slot_first[0] = 1;
  } else {
    for (int ph = 0; ph < 4; ++ph) {
      const int py = ph >> 1, px = ph & 1;
      for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
          const short sh = (short)((py + a) * BW + (px + b));
          const int s = a * 2 + b;
          g.slot_shift[ph][s] = sh; g.slot_group[s] = 0; g.slot_first[s] = (s == 0);
        }
    }
  }
}

void build_tap_table(const ConvGeom& g, int4* out) {
  // {A shift in bytes, first accumulator column, columns (0 = all N_tile), 0}
  for (int ph = 0; ph < 4; ++ph)
    for (int s = 0; s < kMaxSlots; ++s)
      out[ph * kMaxSlots + s] = make_int4(g.slot_shift[ph][s] * 16, g.varn ? g.slot_n0[s] * g.cout_tile : 0,
                                          g.varn ? g.slot_n[s] * g.cout_tile : 0, 0);
}

void finish_geom_for_batch(ConvGeom& g, int N) {
  g.N = N;
  g.tiles_n = ceil_div(N, g.NB);
}

// Tap sets of the nearest-x2 + 3x3 conv folded onto the low-res grid:
// output row 2y+py reads upsampled rows 2y+py+ky-1 = low-res row y + (py-1+a).
static void up3_taps(int p, int a, int* k, int* nk) {
  if (p == 0) { if (a == 0) { k[0] = 0; *nk = 1; } else { k[0] = 1; k[1] = 2; *nk = 2; } }
  else        { if (a == 0) { k[0] = 0; k[1] = 1; *nk = 2; } else { k[0] = 2; *nk = 1; } }
}

void slot_offsets(const ConvLayer& L, int* dy, int* dx) {
  static const int kT[4] = {-1, 0, 0, 1};
  for (int s = 0; s < L.g.n_slots; ++s) {
    if (L.g.s2d) { dy[s] = kT[s / 4]; dx[s] = kT[s % 4]; }
    else if (L.mode == CONV1) { dy[s] = 0; dx[s] = 0; }
    else { int sy, sx; slot_yx(L.g, s, &sy, &sx); dy[s] = sy - 1; dx[s] = sx - 1; }
  }
}

void pack_conv_weights(const ConvLayer& L, const float* w, std::vector<act_t>& out, std::vector<float>* out_f32) {
  const ConvGeom& g = L.g;
  const int cin = L.cin0 + L.cin1, cout = L.cout;
  const int nz = g.phase_grid ? 4 : 1;
  const int k16pc = g.CBK / 2;
  const size_t tile = (size_t)g.N_tile * 16;
  out.assign((size_t)nz * g.n_ntiles * g.n_k * g.n_slots * k16pc * tile, to_act(0.f));
  if (out_f32) out_f32->assign(out.size(), 0.f);
  const bool up = (L.mode == UPCONV3 || L.mode == DECONV4 || L.mode == DECONV4B);

  auto wval = [&](int z, int slot, int co, int ci) -> float {   // co: row of the MMA weight tile
    if (g.s2d) {
      // row block q = output phase (qy,qx); slot = (tap, input phase) pair per axis; pixel tap k = 2t + p - q
      static const int kT[4] = {-1, 0, 0, 1}, kP[4] = {1, 0, 1, 0};
      const int q = co / g.cout_tile, c = co % g.cout_tile;
      if (c >= cout) return 0.f;
      const int ky = 2 * kT[slot / 4] + kP[slot / 4] - (q >> 1), kx = 2 * kT[slot % 4] + kP[slot % 4] - (q & 1);
      if (ky < -1 || ky > 1 || kx < -1 || kx > 1) return 0.f;
      return w[(((size_t)c * cin + ci) * 3 + ky + 1) * 3 + kx + 1];
    }
    if (L.mode == CONV3) { const int ky = slot / 3, kx = slot % 3; return w[(((size_t)co * cin + ci) * 3 + ky) * 3 + kx]; }
    if (L.mode == CONV1) return w[(size_t)co * cin + ci];
    if (L.mode == DECONV4B) {
      // composite of the transposed conv and the blur on the low-res grid.  1-D: out[2i+p] = sum_d bl[d] D[2i+p+d],
      // D[2i'+p'] = sum_a in[i'+p'-1+a] * w[a == 0 ? 3-p' : 1-p'];  slot = (ty+1)*3 + (tx+1), input offset t in -1..1
      const int ph = co / g.cout_tile, c = co % g.cout_tile;
      if (c >= cout) return 0.f;
      const int py = ph >> 1, px = ph & 1, ty = slot / 3 - 1, tx = slot % 3 - 1;
      static const double bl[3] = {0.25, 0.5, 0.25};
      double acc = 0.0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int ay = 0; ay < 2; ++ay) {
          const int Yp = py + dy + 2, ip = Yp / 2 - 1, pp = Yp & 1;      // 2i+py+dy = 2(i+ip)+pp
          if (ip + pp - 1 + ay != ty) continue;
          const int ky = ay == 0 ? 3 - pp : 1 - pp;
          for (int dx = -1; dx <= 1; ++dx)
            for (int ax = 0; ax < 2; ++ax) {
              const int Xp = px + dx + 2, jp = Xp / 2 - 1, qp = Xp & 1;
              if (jp + qp - 1 + ax != tx) continue;
              const int kx = ax == 0 ? 3 - qp : 1 - qp;
              acc += bl[dy + 1] * bl[dx + 1] * (double)w[(((size_t)ci * cout + c) * 4 + ky) * 4 + kx];
            }
        }
      return (float)acc;
    }
    int ph, a, b;
    if (g.phase_grid) { ph = z; a = slot >> 1; b = slot & 1; }
    else {
      // up_cols: slot = input shift (dy,dx) in box coordinates, row block of the weight tile = phase
      ph = block_phase(g.varn, co / g.cout_tile); co = co % g.cout_tile;
      int sy, sx;
      slot_yx(g, slot, &sy, &sx);
      a = sy - (ph >> 1); b = sx - (ph & 1);
      if (a < 0 || a > 1 || b < 0 || b > 1) return 0.f;       // this phase does not read that shift
    }
    const int py = ph >> 1, px = ph & 1;
    if (L.mode == UPCONV3) {
      int ky[2], kx[2], nky, nkx;
      up3_taps(py, a, ky, &nky); up3_taps(px, b, kx, &nkx);
      float s = 0.f;
      for (int i = 0; i < nky; ++i)
        for (int j = 0; j < nkx; ++j) s += w[(((size_t)co * cin + ci) * 3 + ky[i]) * 3 + kx[j]];
      return s;
    }
    // DECONV4: weight (Cin, Cout, 4, 4); out[2y'+py] gathers in[y'+py-1+a] with ky = 3-py (a=0) / 1-py (a=1)
    const int ky = a == 0 ? 3 - py : 1 - py, kx = b == 0 ? 3 - px : 1 - px;
    return w[(((size_t)ci * cout + co) * 4 + ky) * 4 + kx];
  };
  (void)up;
  size_t base = 0;
  for (int z = 0; z < nz; ++z)
    for (int nt = 0; nt < g.n_ntiles; ++nt)
      for (int kc = 0; kc < g.n_k; ++kc)
        for (int slot = 0; slot < g.n_slots; ++slot)
          for (int j = 0; j < k16pc; ++j, base += tile)
            for (int nr = 0; nr < g.N_tile; ++nr) {
              const int co = g.up_cols ? nr : nt * g.N_tile + nr;
              if (!g.up_cols && co >= cout) continue;
              for (int k = 0; k < 16; ++k) {
                const int ci = (kc * g.CBK + 2 * j) * 8 + k;
                const float wv = wval(z, slot, co, ci);
                const size_t idx = base + (size_t)(k >> 3) * (g.N_tile * 8) + (size_t)nr * 8 + (k & 7);
                out[idx] = to_act(wv);
                if (out_f32) (*out_f32)[idx] = wv;
              }
            }
}

// The same packing as pack_conv_weights, as a gather table: every packed element is the sum of up to 4 entries of the
// reference weight tensor (Cout,Cin,k,k) (one for plain / 1x1 / space-to-depth convs, up to four merged taps for the
// nearest-x2 + 3x3 phases); -1 = no term.  Training re-packs the 16-bit operands from the fp32 master weights on the
// device every step (train_step.cu: pack_kernel).  CONV3 / CONV1 / UPCONV3 only.
bool pack_conv_sources(const ConvLayer& L, std::vector<int>& src) {
  const ConvGeom& g = L.g;
  if (L.mode != CONV3 && L.mode != CONV1 && L.mode != UPCONV3) return false;
  const int cin = L.cin0 + L.cin1, cout = L.cout;
  const int nz = g.phase_grid ? 4 : 1;
  const int k16pc = g.CBK / 2;
  const size_t tile = (size_t)g.N_tile * 16;
  src.assign((size_t)nz * g.n_ntiles * g.n_k * g.n_slots * k16pc * tile * 4, -1);
  auto terms = [&](int z, int slot, int co, int ci, int* out) -> void {
    out[0] = out[1] = out[2] = out[3] = -1;
    if (g.s2d) {
      static const int kT[4] = {-1, 0, 0, 1}, kP[4] = {1, 0, 1, 0};
      const int q = co / g.cout_tile, c = co % g.cout_tile;
      if (c >= cout) return;
      const int ky = 2 * kT[slot / 4] + kP[slot / 4] - (q >> 1), kx = 2 * kT[slot % 4] + kP[slot % 4] - (q & 1);
      if (ky < -1 || ky > 1 || kx < -1 || kx > 1) return;
      out[0] = ((c * cin + ci) * 3 + ky + 1) * 3 + kx + 1;
      return;
    }
    if (L.mode == CONV3) { if (co < cout) out[0] = (co * cin + ci) * 9 + slot; return; }
    if (L.mode == CONV1) { if (co < cout) out[0] = co * cin + ci; return; }
    int ph, a, b;
    if (g.phase_grid) { ph = z; a = slot >> 1; b = slot & 1; }
    else {
      ph = block_phase(g.varn, co / g.cout_tile); co = co % g.cout_tile;
      int sy, sx;
      slot_yx(g, slot, &sy, &sx);
      a = sy - (ph >> 1); b = sx - (ph & 1);
      if (a < 0 || a > 1 || b < 0 || b > 1) return;
    }
    if (co >= cout) return;
    int ky[2], kx[2], nky, nkx, n = 0;
    up3_taps(ph >> 1, a, ky, &nky); up3_taps(ph & 1, b, kx, &nkx);
    for (int i = 0; i < nky; ++i)
      for (int j = 0; j < nkx; ++j) out[n++] = ((co * cin + ci) * 3 + ky[i]) * 3 + kx[j];
  };
  size_t base = 0;
  for (int z = 0; z < nz; ++z)
    for (int nt = 0; nt < g.n_ntiles; ++nt)
      for (int kc = 0; kc < g.n_k; ++kc)
        for (int slot = 0; slot < g.n_slots; ++slot)
          for (int j = 0; j < k16pc; ++j, base += tile)
            for (int nr = 0; nr < g.N_tile; ++nr) {
              const int co = g.up_cols ? nr : nt * g.N_tile + nr;
              for (int k = 0; k < 16; ++k) {
                const int ci = (kc * g.CBK + 2 * j) * 8 + k;
                terms(z, slot, co, ci, &src[(base + (size_t)(k >> 3) * (g.N_tile * 8) + (size_t)nr * 8 + (k & 7)) * 4]);
              }
            }
  return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

void make_noise_tensormap(CUtensorMap* tm, const void* base, int N, int H, int W, int boxW, int boxH, int boxN) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return; }
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
  const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxN};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (noise) failed (%d) N=%d H=%d W=%d box=%dx%dx%d", (int)r, N, H, W, boxW, boxH, boxN);
    set_error(buf);
  }
}

void make_plane_tensormap(CUtensorMap* tm, const void* base, int C, int N, int H, int W, int py, int px, int boxW,
                          int boxH, int boxN, int boxCB) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return; }
  // pixels (2yb+py, 2xb+px): dim0 = the two 8-byte halves of a pixel, dim1 = xb (pitch 32 B), dim2 = yb (two rows)
  const char* b = static_cast<const char*>(base) + ((size_t)py * W + px) * 16;
  const cuuint64_t dims[5] = {2, (cuuint64_t)(W / 2), (cuuint64_t)(H / 2), (cuuint64_t)N, (cuuint64_t)(C / 8)};
  const cuuint64_t strides[4] = {32, (cuuint64_t)W * 32, (cuuint64_t)H * W * 16, (cuuint64_t)N * H * W * 16};
  const cuuint32_t box[5] = {2, (cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxN, (cuuint32_t)boxCB};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 5, const_cast<char*>(b), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (plane) failed (%d) C=%d N=%d H=%d W=%d box=%dx%dx%dx%d", (int)r, C, N,
             H, W, boxW, boxH, boxN, boxCB);
    set_error(buf);
  }
}

void make_planar_tensormap(CUtensorMap* tm, const void* base, int C, int N, int H, int W, int plane, int boxW, int boxH,
                           int boxN, int boxCB) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return; }
  const int Hb = H / 2, Wb = W / 2;
  const char* b = static_cast<const char*>(base) + (size_t)plane * Hb * Wb * 16;
  const cuuint64_t dims[4] = {(cuuint64_t)Wb * 2, (cuuint64_t)Hb, (cuuint64_t)N, (cuuint64_t)(C / 8)};
  const cuuint64_t strides[3] = {(cuuint64_t)Wb * 16, (cuuint64_t)H * W * 16, (cuuint64_t)N * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)boxW * 2, (cuuint32_t)boxH, (cuuint32_t)boxN, (cuuint32_t)boxCB};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<char*>(b), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (planar) failed (%d) C=%d N=%d H=%d W=%d box=%dx%dx%dx%d", (int)r, C, N,
             H, W, boxW, boxH, boxN, boxCB);
    set_error(buf);
  }
}

void make_input_tensormaps(ConvParams& p, const ConvLayer& L, int N, const void* x0, const void* x1) {
  const ConvGeom& g = p.g;
  if (g.s2d && g.in_planar) {
    for (int pl = 0; pl < 4; ++pl) {
      make_planar_tensormap(&p.tm_pl[0][pl], x0, L.cin0, N, L.H, L.W, pl, g.BW, g.BH, g.NB, g.CBK);
      if (L.cin1 > 0) make_planar_tensormap(&p.tm_pl[1][pl], x1, L.cin1, N, L.H, L.W, pl, g.BW, g.BH, g.NB, g.CBK);
      else p.tm_pl[1][pl] = p.tm_pl[0][pl];
    }
    p.tm[0] = p.tm_pl[0][0];
    p.tm[1] = p.tm_pl[1][0];
    return;
  }
  if (g.s2d) {
    for (int pl = 0; pl < 4; ++pl) {
      make_plane_tensormap(&p.tm_pl[0][pl], x0, L.cin0, N, L.H, L.W, pl >> 1, pl & 1, g.BW, g.BH, g.NB, g.CBK);
      if (L.cin1 > 0) make_plane_tensormap(&p.tm_pl[1][pl], x1, L.cin1, N, L.H, L.W, pl >> 1, pl & 1, g.BW, g.BH, g.NB, g.CBK);
      else p.tm_pl[1][pl] = p.tm_pl[0][pl];
    }
    p.tm[0] = p.tm_pl[0][0];
    p.tm[1] = p.tm_pl[1][0];
    return;
  }
  make_act_tensormap(&p.tm[0], x0, L.cin0, N, L.H, L.W, g.BW, g.BH, g.NB, g.CBK);
  if (L.cin1 > 0) make_act_tensormap(&p.tm[1], x1, L.cin1, N, L.H, L.W, g.BW, g.BH, g.NB, g.CBK);
  else p.tm[1] = p.tm[0];
  for (int s = 0; s < 2; ++s)
    for (int pl = 0; pl < 4; ++pl) p.tm_pl[s][pl] = p.tm[0];
}

void make_act_tensormap(CUtensorMap* tm, const void* base, int C, int N, int H, int W, int boxW, int boxH, int boxN,
                        int boxCB) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return; }
  // dim0 counts 8-byte units (2 per pixel) so that a 256-element box row spans 128 pixels
  const cuuint64_t dims[4] = {(cuuint64_t)W * 2, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)(C / 8)};
  const cuuint64_t strides[3] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)N * H * W * 16};
  const cuuint32_t box[4] = {(cuuint32_t)boxW * 2, (cuuint32_t)boxH, (cuuint32_t)boxN, (cuuint32_t)boxCB};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d) C=%d N=%d H=%d W=%d box=%dx%dx%dx%d", (int)r, C, N, H,
             W, boxW, boxH, boxN, boxCB);
    set_error(buf);
  }
}

}  // namespace gsx
