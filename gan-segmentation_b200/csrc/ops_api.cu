// Single-operator entry points of the C ABI (tests and tuning sweeps).  fp32 NCHW device tensors in
// and out; blocked act_t temporaries are allocated per call, so these are not hot-path functions.
#include "../../include/gsx.h"
#include "gsx_internal.h"

#include <atomic>
#include <vector>

namespace gsx {
const char* last_error_cstr();
extern std::atomic<uint64_t> g_launches;

struct Tmp {
  std::vector<void*> ptrs;
  cudaStream_t st = nullptr;
  bool has_stream = false;
  // A block goes back to the pool only when nothing enqueued by this call can still touch it -- also on the early
  // error returns, where copies / conversion kernels may already be in flight (the success paths have synchronised,
  // for them this is a no-op).
  ~Tmp() {
    if (has_stream) cudaStreamSynchronize(st);
    for (void* p : ptrs) pool_put(p);
  }
  template <class T> T* get(size_t n) {
    void* p = pool_get(std::max<size_t>(n, 1) * sizeof(T));
    if (!p) return nullptr;
    ptrs.push_back(p);
    return static_cast<T*>(p);
  }
};
}  // namespace gsx

using namespace gsx;

extern "C" int gsx_op_conv(int mode, int n, int h, int w, int cin0, int cin1, int cout, const float* x0_dev,
                           const float* x1_dev, const float* w_host, const float* bias_dev, const float* nscale_dev,
                           const float* noise_dev, int flags, const float* addsrc_dev, float* out_dev, float* stats_dev,
                           uint8_t* mask_dev, float* logits_dev, int num_classes, const gsx_plan_override* ov,
                           int* plan_out, int repeat, float* ms_out, gsx_stream stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool up = (mode == UPCONV3 || mode == DECONV4 || mode == DECONV4B);
  const bool argmax = (flags & EPI_ARGMAX) != 0;
  const int Ho = up ? 2 * h : h, Wo = up ? 2 * w : w;
  Tmp tmp;
  tmp.st = st; tmp.has_stream = true;
  set_error("");
  ConvLayer L;
  PlanOverride po{};
  if (ov) { po.TH = ov->TH; po.TW = ov->TW; po.NB = ov->NB; po.CBK = ov->CBK; po.N_tile = ov->N_tile; po.stages = ov->stages; po.phase_grid = ov->phase_grid; po.epi_groups = ov->epi_groups; po.acc_bufs = ov->acc_bufs; po.max_mtiles = ov->max_mtiles; po.hstack = ov->hstack; po.s2d = ov->s2d; }
  plan_conv(L, mode, h, w, cin0, cin1, cout, argmax ? num_classes : 0, ov ? &po : nullptr,
            noise_dev ? 1 : (addsrc_dev ? 2 : 0));
  if (*last_error_cstr()) return -1;
  std::vector<act_t> packed;
  pack_conv_weights(L, w_host, packed);
  act_t* wp = tmp.get<act_t>(packed.size());
  act_t* xb0 = tmp.get<act_t>((size_t)n * cin0 * h * w);
  act_t* xb1 = cin1 ? tmp.get<act_t>((size_t)n * cin1 * h * w) : nullptr;
  act_t* ob = argmax ? nullptr : tmp.get<act_t>((size_t)n * cout * Ho * Wo);
  act_t* ab = addsrc_dev ? tmp.get<act_t>((size_t)n * cout * (Ho / 2) * (Wo / 2)) : nullptr;
  if (!wp || !xb0 || (cin1 && !xb1) || (!argmax && !ob) || (addsrc_dev && !ab)) { set_error("cudaMalloc failed"); return -2; }
  if (!cuda_ok(cudaMemcpyAsync(wp, packed.data(), packed.size() * sizeof(act_t), cudaMemcpyHostToDevice, st), "H2D weights")) return -2;
  launch_nchw_to_blocked(x0_dev, xb0, cin0, n, h * w, st);
  if (cin1) launch_nchw_to_blocked(x1_dev, xb1, cin1, n, h * w, st);
  if (addsrc_dev) launch_nchw_to_blocked(addsrc_dev, ab, cout, n, (Ho / 2) * (Wo / 2), st);
  L.wpack_dev = wp;
  std::vector<int4> taps_h(4 * kMaxSlots);
  build_tap_table(L.g, taps_h.data());
  int4* taps_d = tmp.get<int4>(taps_h.size());
  if (!taps_d) { set_error("cudaMalloc failed"); return -2; }
  cudaMemcpyAsync(taps_d, taps_h.data(), taps_h.size() * sizeof(int4), cudaMemcpyHostToDevice, st);

  ConvParams p;
  p.g = L.g;
  finish_geom_for_batch(p.g, n);
  p.wpack = wp;
  p.wpack_n_stride = 0;
  p.taps = taps_d;
  ConvEpi e{};
  e.out = ob; e.Ho = Ho; e.Wo = Wo; e.up = up ? 1 : 0; e.flags = flags; e.Cout = cout;
  e.bias = bias_dev; e.nscale = nscale_dev; e.noise = noise_dev; e.addsrc = ab;
  e.mask = mask_dev; e.logits = logits_dev; e.num_classes = num_classes;
  const bool want_stats = (flags & EPI_STATS) != 0;
  const bool fused_stats = want_stats && p.g.NB == 1;
  if (want_stats && !fused_stats) e.flags &= ~EPI_STATS;
  const int stats_T = fused_stats ? p.g.tiles_x * p.g.tiles_y : stats_tiles(Ho * Wo);
  float* partial = want_stats ? tmp.get<float>((size_t)n * stats_T * cout * 2) : nullptr;
  if (want_stats && !partial) { set_error("cudaMalloc failed"); return -2; }
  e.stats = fused_stats ? partial : nullptr;
  e.stats_T = stats_T;
  if (mode == DECONV4B) {
    // border correction of the folded deconv+blur: weights rearranged to [ky][kx][Cin][Cout]
    std::vector<float> wt((size_t)16 * cin0 * cout);
    for (int ci = 0; ci < cin0; ++ci)
      for (int co = 0; co < cout; ++co)
        for (int k = 0; k < 16; ++k) wt[((size_t)k * cin0 + ci) * cout + co] = w_host[((size_t)ci * cout + co) * 16 + k];
    float* wt_d = tmp.get<float>(wt.size());
    float* er = tmp.get<float>((size_t)n * 2 * Wo * cout);
    float* ec = tmp.get<float>((size_t)n * 2 * Ho * cout);
    if (!wt_d || !er || !ec) { set_error("cudaMalloc failed"); return -2; }
    cudaMemcpyAsync(wt_d, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);                       // wt is a local
    launch_deconv_border(xb0, wt_d, er, ec, n, cin0, cout, h, w, st); g_launches++;
    e.e_rows = er; e.e_cols = ec;
  }
  p.e = e;
  make_input_tensormaps(p, L, n, xb0, xb1);
  if (p.g.aux_kind == 1) make_noise_tensormap(&p.tm_aux, noise_dev, n, Ho, Wo, p.g.TW << p.g.aux_up, p.g.TH << p.g.aux_up, p.g.NB);
  else if (p.g.aux_kind == 2) make_act_tensormap(&p.tm_aux, ab, cout, n, Ho / 2, Wo / 2, p.g.aux_bw, p.g.aux_bh, p.g.NB, p.g.cout_tile / 8);
  else p.tm_aux = p.tm[0];
  if (*last_error_cstr()) return -1;
  if (plan_out) {
    const ConvGeom& g = p.g;
    const int vals[16] = {g.TH, g.TW, g.NB, g.CBK, g.N_tile, g.stages, g.phase_grid, g.n_mtiles, g.n_k, g.tmem_cols,
                          g.smem_bytes, g.tiles_x * g.tiles_y * g.tiles_n, g.n_ntiles, g.n_groups * 100 + g.acc_bufs * 10 + g.epi_groups, g.n_slots, g.BW};
    for (int i = 0; i < 16; ++i) plan_out[i] = vals[i];
  }
  launch_shiftconv(p, st); g_launches++;
  if (!cuda_ok(cudaGetLastError(), "shiftconv launch")) return -2;
  if (want_stats && !fused_stats) launch_stats(ob, partial, cout, n, Ho * Wo, st);
  if (want_stats) launch_finalize(partial, stats_T, n, cout, Ho * Wo, nullptr, 0, 0, stats_dev, st);
  if (!argmax && out_dev) launch_blocked_to_nchw(ob, out_dev, cout, n, Ho * Wo, st);
  if (!cuda_ok(cudaStreamSynchronize(st), "gsx_op_conv")) return -2;
  if (repeat > 0 && ms_out) {
    // timing loop for tuning sweeps: stats accumulate garbage here, outputs are idempotent
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, st);
    for (int i = 0; i < repeat; ++i) launch_shiftconv(p, st);
    cudaEventRecord(e1, st);
    if (!cuda_ok(cudaEventSynchronize(e1), "timing")) return -2;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / repeat;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    g_launches += repeat;
  }
  return 0;
}

extern "C" void gsx_op_release_cache(void) { pool_release(); }

extern "C" int gsx_plan_query(int mode, int h, int w, int cin0, int cin1, int cout, int num_classes,
                              const gsx_plan_override* ov, int* plan_out) {
  set_error("");
  ConvLayer L;
  PlanOverride po{};
  if (ov) { po.TH = ov->TH; po.TW = ov->TW; po.NB = ov->NB; po.CBK = ov->CBK; po.N_tile = ov->N_tile; po.stages = ov->stages; po.phase_grid = ov->phase_grid; po.epi_groups = ov->epi_groups; po.acc_bufs = ov->acc_bufs; po.max_mtiles = ov->max_mtiles; po.hstack = ov->hstack; po.s2d = ov->s2d; }
  plan_conv(L, mode, h, w, cin0, cin1, cout, num_classes, ov ? &po : nullptr);
  if (*last_error_cstr()) return -1;
  const ConvGeom& g = L.g;
  const int vals[16] = {g.TH, g.TW, g.NB, g.CBK, g.N_tile, g.stages, g.phase_grid, g.n_mtiles, g.n_k, g.tmem_cols,
                        g.smem_bytes, g.tiles_x * g.tiles_y, g.n_ntiles, g.n_groups * 100 + g.acc_bufs * 10 + g.epi_groups, g.n_slots, g.BW};
  for (int i = 0; i < 16; ++i) plan_out[i] = vals[i];
  return 0;
}

extern "C" int gsx_op_pass1(int n, int c, int h, int w, const float* x_dev, int blur, int in_broadcast,
                            const float* nscale_dev, const float* bias_dev, const float* noise_dev, float* out_dev,
                            float* stats_dev, gsx_stream stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Tmp tmp;
  tmp.st = st; tmp.has_stream = true;
  const int nin = in_broadcast ? 1 : n;
  act_t* xb = tmp.get<act_t>((size_t)nin * c * h * w);
  act_t* ob = tmp.get<act_t>((size_t)n * c * h * w);
  if (!xb || !ob) { set_error("cudaMalloc failed"); return -2; }
  launch_nchw_to_blocked(x_dev, xb, c, nin, h * w, st);
  const int T = pass1_tiles(h, w);
  float* partial = stats_dev ? tmp.get<float>((size_t)n * T * c * 2) : nullptr;
  Pass1Args a{};
  a.in = xb; a.out = ob; a.C = c; a.N = n; a.H = h; a.W = w; a.blur = blur; a.in_broadcast = in_broadcast;
  a.nscale = nscale_dev; a.bias = bias_dev; a.noise = noise_dev; a.stats = partial;
  launch_pass1(a, st); g_launches++;
  if (stats_dev) launch_finalize(partial, T, n, c, h * w, nullptr, 0, 0, stats_dev, st);
  launch_blocked_to_nchw(ob, out_dev, c, n, h * w, st);
  return cuda_ok(cudaStreamSynchronize(st), "gsx_op_pass1") ? 0 : -2;
}

extern "C" int gsx_op_apply(int n, int c, int h, int w, const float* x_dev, const float* stats_dev,
                            const float* styles_dev, const float* wrgb_dev, const float* brgb_dev, int nc, float* out_dev,
                            float* img_f32_dev, uint8_t* img_u8_dev, gsx_stream stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Tmp tmp;
  tmp.st = st; tmp.has_stream = true;
  act_t* xb = tmp.get<act_t>((size_t)n * c * h * w);
  act_t* ob = tmp.get<act_t>((size_t)n * c * h * w);
  if (!xb || !ob) { set_error("cudaMalloc failed"); return -2; }
  launch_nchw_to_blocked(x_dev, xb, c, n, h * w, st);
  float* coef = tmp.get<float>((size_t)n * c * 2);
  if (!coef) { set_error("cudaMalloc failed"); return -2; }
  launch_finalize(stats_dev, 1, n, c, h * w, styles_dev, 2 * c, 0, coef, st);
  ApplyArgs a{};
  a.in = xb; a.out = ob; a.C = c; a.N = n; a.H = h; a.W = w; a.coef = coef;
  a.wrgb = wrgb_dev; a.brgb = brgb_dev; a.nc = nc;
  a.img_f32 = img_f32_dev; a.img_u8 = img_u8_dev;
  launch_apply(a, st); g_launches++;
  launch_blocked_to_nchw(ob, out_dev, c, n, h * w, st);
  return cuda_ok(cudaStreamSynchronize(st), "gsx_op_apply") ? 0 : -2;
}

extern "C" int gsx_op_fill_normal(float* out_dev, size_t per_sample, int n, uint64_t seed, uint64_t first_sample,
                                  int stream_id, gsx_stream stream) {
  launch_fill_noise(out_dev, per_sample, n, seed, first_sample, stream_id, static_cast<cudaStream_t>(stream));
  g_launches++;
  return cuda_ok(cudaGetLastError(), "fill_normal") ? 0 : -2;
}
