// C ABI of libgsx (include/gsx.h): handles, parameter folding/packing, forward orchestration.
#include "../../include/gsx.h"
#include "gsx_internal.h"

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>

namespace gsx {
const char* last_error_cstr();
std::atomic<uint64_t> g_launches{0};

struct HostTensor {
  std::vector<float> v;
  std::vector<int64_t> shape;
};

static int ilog2_exact(long v) {
  int r = 0;
  while ((1L << r) < v) ++r;
  return ((1L << r) == v) ? r : -1;
}

// Legacy prefix names (networks_stylegan.py:16-54,125,136,244-247) -> structural names.
static std::string canonical_name(const std::string& in) {
  std::string name = in;
  if (name.rfind("arg:", 0) == 0 || name.rfind("aux:", 0) == 0) name = name.substr(4);
  if (name.find('.') != std::string::npos) return name;
  if (name == "constant_tensor" || name == "latent_avg" || name == "truncation_psi") return name;
  int i;
  char rest[128];
  if (sscanf(name.c_str(), "mp_dense_%d_%127s", &i, rest) == 2) return "mapping." + std::to_string(2 * i + 1) + "." + rest;
  long scale;
  if (sscanf(name.c_str(), "%ld_%127s", &scale, rest) == 2) {
    const int r = ilog2_exact(scale);
    if (r < 0) return name;
    const std::string R = std::to_string(r), s = rest;
    auto tail = [&](const char* pre) { return s.substr(strlen(pre)); };
    if (s.rfind("conv_1_", 0) == 0) return "net" + R + ".block0." + tail("conv_1_");
    if (s.rfind("deconv_1_", 0) == 0) return "net" + R + ".block0." + tail("deconv_1_");
    if (s == "blur_1_w_kernel") return "net" + R + ".blur.w_kernel";
    if (s == "noise_1_scale_factors") return "net" + R + ".block1.0.scale_factors";
    if (s == "bias_1_bias") return "net" + R + ".block1.1.bias";
    if (s == "noise_2_scale_factors") return "net" + R + ".block2.1.scale_factors";
    if (s == "bias_2_bias") return "net" + R + ".block2.2.bias";
    if (s.rfind("conv_2_", 0) == 0) return "net" + R + ".block2.0." + tail("conv_2_");
    if (s.rfind("conv_to_rgb_", 0) == 0) return "to_rgb" + R + ".0." + tail("conv_to_rgb_");
    for (int k = 1; k <= 2; ++k) {
      const std::string a = "adain_" + std::to_string(k) + "_dense_affine_", b = "adain_" + std::to_string(k) + "_norm_";
      if (s.rfind(a, 0) == 0) return "net" + R + ".adain" + std::to_string(k) + ".affine." + s.substr(a.size());
      if (s.rfind(b, 0) == 0) return "net" + R + ".adain" + std::to_string(k) + ".instance." + s.substr(b.size());
    }
  }
  return name;
}

template <class T>
static T* dev_upload(const std::vector<T>& h) {
  T* d = nullptr;
  if (!cuda_ok(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)), "cudaMalloc")) return nullptr;
  if (!cuda_ok(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy H2D")) return nullptr;
  return d;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Optional per-launch timing (bench.py's per-layer roofline table): CUDA events around every launch
// of a forward pass, on the caller's stream.  Off on the hot path.
struct ProfRec { std::string label; double bytes, flops, flops_exec; std::string kernel; cudaEvent_t e0, e1; };
static std::vector<ProfRec> g_prof;
static bool g_prof_on = false;
struct ProfScope {
  cudaStream_t st; bool on;
  // flops: dense count of the reference formulation; flops_exec: what the tensor cores are actually asked to do (phase /
  // space-to-depth decompositions, zero weight blocks of stacked phases included; tile halo rows excluded); kernel: the
  // __global__ function behind the label
  ProfScope(const std::string& label, double bytes, double flops, cudaStream_t s, double flops_exec = 0, const char* kernel = "")
      : st(s), on(g_prof_on) {
    if (!on) return;
    ProfRec r{label, bytes, flops, flops_exec, kernel, nullptr, nullptr};
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
    g_prof.push_back(r);
  }
  ~ProfScope() { if (on) cudaEventRecord(g_prof.back().e1, st); }
};

struct Arena {              // carves a caller-owned workspace
  uint8_t* base; size_t off = 0;
  explicit Arena(void* b) : base(static_cast<uint8_t*>(b)) {}
  template <class T> T* take(size_t n) {
    T* p = reinterpret_cast<T*>(base + off);
    off = align_up(off + n * sizeof(T), 1024);
    return p;
  }
};

// Per-sample operands of a conv that has its input's AdaIN folded in (modulate.cu); carved from the caller's workspace.
struct ModBufs { act_t* w = nullptr; float* bias_n = nullptr; float* bdelta = nullptr; };
static ModBufs take_mod(Arena& a, const ConvLayer& L, int N) {
  ModBufs m;
  m.w = a.take<act_t>((size_t)N * L.wpack_elems);
  m.bias_n = a.take<float>((size_t)N * L.g.bias_cols);
  m.bdelta = a.take<float>((size_t)N * 9 * L.g.bias_cols);
  return m;
}
static int g_opt_fold_apply = 1;           // gsx_set_option("fold_apply", 0/1), read when a handle is finalized
static int g_opt_dec_branch_kpx = 1 << 20;  // ... for levels with at most this many thousand pixels in the batch ("dec_branch_kpx"; measured: branching every level is best -- FFHQ batch 32: 96 K px 3147, 600 K 3169, all 3178 samples/s)
static int g_opt_dec_branches = 1;         // gsx_set_option("dec_branches", 0/1): decoder cvt blocks / shortcuts of the small levels on side streams
// Set by the fused generate calls around gsx_synth_forward: the ToRGB pass (a read of the last tensor and the image write,
// HBM-bound) goes to the generator's side stream and runs beside the decoder, which does not need the image; the fused
// call joins it before it returns.  gsx_synth_forward on its own always finishes the image on the caller's stream.
static thread_local bool t_defer_rgb = false;
static int g_opt_defer_rgb = 1;            // gsx_set_option("defer_rgb", 0/1)
static int g_opt_inline_finalize = 1;      // gsx_set_option("inline_finalize", 0/1): AdaIN coefficients computed in the apply pass's prologue (no finalize launch)
static int g_opt_fold_deconv_maxc = 16;    // gsx_set_option("fold_deconv_maxc", C): deconv+blur folded into one kernel up to C channels

// Runs one planned conv layer: builds the tensor maps for this batch / these buffers and launches.
static bool run_conv(const ConvLayer& L, int N, const act_t* x0, const act_t* x1, const ConvEpi& epi, cudaStream_t st,
                     const char* label = "conv", const ModBufs* mod = nullptr) {
  // algorithmic work of this layer: 2 B x (input + output elements), upsample/concat folded,
  // +4 B x noise plane, uint8 mask; dense FLOPs of the reference formulation
  const double in_el = (double)N * (L.cin0 + L.cin1) * L.H * L.W, out_px = (double)N * epi.Ho * epi.Wo;
  double bytes = 2.0 * in_el + ((epi.flags & EPI_ARGMAX) ? out_px : 2.0 * out_px * L.cout);
  if (epi.noise) bytes += 4.0 * out_px;
  if (epi.addsrc) bytes += 2.0 * out_px / 4 * L.cout;
  const double taps = L.mode == CONV1 ? 1 : ((L.mode == DECONV4 || L.mode == DECONV4B) ? 4 : 9);    // per OUTPUT pixel
  const double flops = 2.0 * out_px * taps * (L.cin0 + L.cin1) * L.cout;
  // executed: every GEMM row (input-resolution pixel, or 2x2 block in the space-to-depth plan; x4 phase work items) times
  // slots x Cin x all N columns
  const double rows = (double)N * L.g.H * L.g.W * (L.g.phase_grid ? 4 : 1);
  const double flops_exec = 2.0 * rows * L.g.n_slots * (L.cin0 + L.cin1) * L.g.N_tile * L.g.n_ntiles;
  ProfScope ps(label, bytes, flops, st, flops_exec, "shiftconv_kernel");
  ConvParams p;
  p.g = L.g;
  finish_geom_for_batch(p.g, N);
  p.e = epi;
  p.e.out_planar = L.out_planar;
  p.wpack = L.wpack_dev;
  p.wpack_n_stride = 0;
  if (L.g.per_sample) {
    if (!mod) { set_error("run_conv: per-sample layer without modulated operands"); return false; }
    p.wpack = mod->w; p.wpack_n_stride = L.wpack_elems;
    p.e.bias = nullptr; p.e.bias_n = mod->bias_n; p.e.bdelta = mod->bdelta;
  }
  p.taps = L.taps_dev;
  set_error("");
  make_input_tensormaps(p, L, N, x0, x1);
  if (p.g.aux_kind == 1) {
    if (!epi.noise) p.g.aux_kind = 0;
    else make_noise_tensormap(&p.tm_aux, epi.noise, N, epi.Ho, epi.Wo, p.g.TW << p.g.aux_up, p.g.TH << p.g.aux_up, p.g.NB);
  } else if (p.g.aux_kind == 2) {
    if (!epi.addsrc) p.g.aux_kind = 0;
    else make_act_tensormap(&p.tm_aux, epi.addsrc, L.cout, N, epi.Ho / 2, epi.Wo / 2, p.g.aux_bw, p.g.aux_bh, p.g.NB,
                            p.g.cout_tile / 8);
  }
  if (!p.g.aux_kind) p.tm_aux = p.tm[0];
  if (*last_error_cstr()) return false;
  launch_shiftconv(p, st);
  g_launches++;
  return cuda_ok(cudaGetLastError(), "shiftconv launch");
}

bool run_conv_layer(const ConvLayer& L, int N, const act_t* x0, const act_t* x1, const ConvEpi& epi, cudaStream_t st,
                    const char* label) {
  return run_conv(L, N, x0, x1, epi, st, label, nullptr);
}

static bool upload_conv(ConvLayer& L, const float* w) {
  std::vector<act_t> packed;
  std::vector<float> packed_f32;
  pack_conv_weights(L, w, packed, L.g.per_sample ? &packed_f32 : nullptr);
  L.wpack_elems = packed.size();
  L.wpack_dev = dev_upload(packed);
  if (L.g.per_sample) {
    L.wf32_dev = dev_upload(packed_f32);
    if (!L.wf32_dev) return false;
  }
  std::vector<int4> taps(4 * kMaxSlots);
  build_tap_table(L.g, taps.data());
  L.taps_dev = dev_upload(taps);
  return L.wpack_dev != nullptr && L.taps_dev != nullptr;
}

}  // namespace gsx

using namespace gsx;

// =============================================================================================
// generator
// =============================================================================================
struct SynthBlock {
  int r, C, Cin, H, W;
  ConvLayer conv1, conv2;                 // conv1 unused at r == 2
  float *ns1 = nullptr, *b1 = nullptr, *ns2 = nullptr, *b2 = nullptr;
  bool fold = false;                      // conv1 = deconv + blur + noise/bias/lrelu/stats in one kernel (DECONV4B)
  // AdaIN folded into the consumers (modulate.cu): mod2 = conv_2 reads the un-normalised first half t1 with per-sample
  // weights (no apply1 pass); t2 = the block's output stays un-normalised in feat[] and every consumer (next block's
  // conv_1, the decoder's cvt conv, ToRGB) folds AdaIN 2 in (no apply2 pass); mod1 = conv_1 consumes such a tensor.
  bool mod1 = false, mod2 = false, t2 = false;
  float* wt = nullptr;                    // fold: scaled deconv weights [4][4][Cin][C] fp32 for the border correction
};

struct gsx_synth {
  gsx_synth_cfg cfg;
  int L, nlayers, S_total;
  std::map<std::string, HostTensor> params;
  bool finalized = false;
  std::vector<SynthBlock> blocks;
  std::vector<int> style_off;             // per style layer, offset into the styles row
  float* d_map_w[8] = {nullptr};
  float* d_map_b[8] = {nullptr};
  float *d_aff_w = nullptr, *d_aff_b = nullptr;
  int* d_unit_layer = nullptr;
  float *d_latent_avg = nullptr, *d_psi = nullptr, *d_wrgb = nullptr, *d_brgb = nullptr;
  act_t* d_const = nullptr;
  int last_n = 0;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};   // gsx_generate_host pipelining
  unsigned long long* d_counter = nullptr;   // running global sample index in HBM (gsx_synth_device_counter; CUDA-graph replays)
  cudaStream_t side = nullptr;               // the Philox noise fill runs here, concurrently with the mapping network
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_rgb_fork = nullptr, ev_rgb_done = nullptr;     // ToRGB deferred to the side stream by the fused generate calls
  bool rgb_pending = false;

  int nf(int r) const {
    const int f = (int)(cfg.fmap_base / std::pow(2.0, (r - 1) * (double)cfg.fmap_decay));
    return std::min(f, cfg.fmap_max);
  }
  void hw(int r, int& h, int& w) const { h = cfg.base_scale_y << (r - 2); w = cfg.base_scale_x << (r - 2); }
};

struct SynthWs {
  float *z, *wa, *wb, *styles, *psi;
  std::vector<float*> noise;
  std::vector<float*> partial;            // per style layer: per-tile (sum, sumsq) partials [N][T][C][2]
  std::vector<float*> coef;               // per style layer: AdaIN coefficients [N][C][2]
  std::vector<int> stats_T;
  std::vector<act_t*> feat;
  act_t *bufA, *bufB;
  float *e_rows, *e_cols;                 // border corrections of the folded deconv+blur layers
  std::vector<ModBufs> mod1, mod2;        // per block: per-sample operands of conv_1 / conv_2 (AdaIN folded in)
  size_t total;
};

// Tiles per sample the statistics of style layer l arrive in (must match what the producers write).
static int synth_stats_tiles(const gsx_synth* h, int l);

static SynthWs synth_layout(const gsx_synth* h, int N, void* base) {
  SynthWs w;
  Arena a(base);
  const int Z = h->cfg.latent_size;
  w.z = a.take<float>((size_t)N * Z);
  w.wa = a.take<float>((size_t)N * Z);
  w.wb = a.take<float>((size_t)N * Z);
  w.styles = a.take<float>((size_t)N * h->S_total);
  w.psi = a.take<float>(h->nlayers);
  for (int l = 0; l < h->nlayers; ++l) {
    const int C = h->nf(2 + l / 2), T = synth_stats_tiles(h, l);
    w.stats_T.push_back(T);
    w.partial.push_back(a.take<float>((size_t)N * T * C * 2));
    w.coef.push_back(a.take<float>((size_t)N * C * 2));
  }
  size_t maxact = 0;
  for (int r = 2; r <= h->L; ++r) {
    int hh, ww;
    h->hw(r, hh, ww);
    const size_t plane = (size_t)N * hh * ww;
    w.noise.push_back(a.take<float>(plane));
    w.noise.push_back(a.take<float>(plane));
    w.feat.push_back(a.take<act_t>(plane * h->nf(r)));
    maxact = std::max(maxact, plane * h->nf(r));
  }
  w.bufA = a.take<act_t>(maxact);
  w.bufB = a.take<act_t>(maxact);
  size_t maxe = 0;
  for (const auto& b : h->blocks)
    if (b.fold) maxe = std::max(maxe, (size_t)N * 2 * std::max(b.H, b.W) * b.C);
  w.e_rows = a.take<float>(maxe);
  w.e_cols = a.take<float>(maxe);
  for (const auto& b : h->blocks) {
    w.mod1.push_back(b.mod1 ? take_mod(a, b.conv1, N) : ModBufs());
    w.mod2.push_back(b.mod2 ? take_mod(a, b.conv2, N) : ModBufs());
  }
  w.total = a.off;
  return w;
}

static int synth_stats_tiles(const gsx_synth* h, int l) {
  int hh, ww;
  h->hw(2 + l / 2, hh, ww);
  if ((l & 1) == 0) {
    if ((size_t)(l / 2) < h->blocks.size() && h->blocks[l / 2].fold) {
      const ConvGeom& g = h->blocks[l / 2].conv1.g;
      return g.tiles_x * g.tiles_y;
    }
    return pass1_tiles(hh, ww);
  }
  if ((size_t)(l / 2) < h->blocks.size()) {
    const ConvGeom& g = h->blocks[l / 2].conv2.g;
    if (g.NB == 1) return g.tiles_x * g.tiles_y;
  }
  return stats_tiles(hh * ww);
}

extern "C" const char* gsx_last_error(void) { return last_error_cstr(); }
extern "C" int gsx_abi_version(void) { return 2; }
extern "C" uint64_t gsx_launch_count(void) { return g_launches.load(); }
extern "C" int gsx_set_option(const char* name, int value) {
  if (name && std::strcmp(name, "fold_apply") == 0) { g_opt_fold_apply = value != 0; return 0; }
  if (name && std::strcmp(name, "fold_deconv_maxc") == 0) { g_opt_fold_deconv_maxc = value; return 0; }
  if (name && std::strcmp(name, "dec_branches") == 0) { g_opt_dec_branches = value != 0; return 0; }
  if (name && std::strcmp(name, "defer_rgb") == 0) { g_opt_defer_rgb = value != 0; return 0; }
  if (name && std::strcmp(name, "inline_finalize") == 0) { g_opt_inline_finalize = value != 0; return 0; }
  if (name && std::strcmp(name, "dec_branch_kpx") == 0 && value >= 0) { g_opt_dec_branch_kpx = value; return 0; }
  if (name && std::strcmp(name, "wgrad_m64") == 0 && value >= 0 && value <= 2) { g_wgrad_m64 = value; return 0; }
  if (name && std::strcmp(name, "wgrad_kxm") == 0) { g_wgrad_kxm = value != 0; return 0; }
  if (name && std::strcmp(name, "pdl") == 0 && value >= 0 && value <= 4) {
    g_pdl_mode = value;
    set_pdl_late_conv(value >= 3); set_pdl_late_ew(value >= 3);
    return 0;
  }
  if (name && std::strcmp(name, "varn") == 0 && value >= 0 && value <= 2) { g_plan_varn = value; return 0; }
  if (name && std::strcmp(name, "epi_groups") == 0 && (value == 2 || value == 4)) { g_plan_epi_groups = value; return 0; }
  set_error(std::string("unknown option: ") + (name ? name : "(null)"));
  return -1;
}

extern "C" int gsx_synth_create(const gsx_synth_cfg* cfg, gsx_synth** out) {
  if (!cfg || !out) { set_error("null argument"); return -1; }
  if (cfg->latent_size != 512 || cfg->max_res_log2 < 2 || cfg->max_res_log2 > 12 || cfg->channels > 4) {
    set_error("unsupported generator config (latent_size must be 512, channels <= 4)");
    return -1;
  }
  int dev;
  if (!cuda_ok(cudaGetDevice(&dev), "cudaGetDevice (libgsx has no CPU fallback)")) return -2;
  gsx_synth* h = new gsx_synth();
  h->cfg = *cfg;
  h->L = cfg->max_res_log2;
  h->nlayers = 2 * (h->L - 1);
  int off = 0;
  for (int l = 0; l < h->nlayers; ++l) {
    h->style_off.push_back(off);
    off += 2 * h->nf(2 + l / 2);
  }
  h->S_total = off;
  for (int r = 2; r <= h->L; ++r)
    if (h->nf(r) % 16) { set_error("channel counts must be multiples of 16"); delete h; return -1; }
  *out = h;
  return 0;
}

extern "C" void gsx_synth_destroy(gsx_synth* h) {
  if (!h) return;
  for (int i = 0; i < 2; ++i) { if (h->ev_done[i]) cudaEventDestroy(h->ev_done[i]); if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]); }
  cudaFree(h->d_counter);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->ev_rgb_fork) cudaEventDestroy(h->ev_rgb_fork);
  if (h->ev_rgb_done) cudaEventDestroy(h->ev_rgb_done);
  for (int i = 0; i < 8; ++i) { cudaFree(h->d_map_w[i]); cudaFree(h->d_map_b[i]); }
  cudaFree(h->d_aff_w); cudaFree(h->d_aff_b); cudaFree(h->d_unit_layer); cudaFree(h->d_latent_avg);
  cudaFree(h->d_psi); cudaFree(h->d_wrgb); cudaFree(h->d_brgb); cudaFree(h->d_const);
  for (auto& b : h->blocks) {
    cudaFree(b.conv1.wpack_dev); cudaFree(b.conv2.wpack_dev);
    cudaFree(b.conv1.taps_dev); cudaFree(b.conv2.taps_dev);
    cudaFree(b.conv1.wf32_dev); cudaFree(b.conv2.wf32_dev);
    cudaFree(b.ns1); cudaFree(b.b1); cudaFree(b.ns2); cudaFree(b.b2); cudaFree(b.wt);
  }
  delete h;
}

static int store_param(std::map<std::string, HostTensor>& params, const char* name, const float* data,
                       const int64_t* shape, int ndim) {
  if (!name || !data || ndim < 0 || ndim > 8) { set_error("bad parameter"); return -1; }
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) { t.shape.push_back(shape[i]); n *= (size_t)shape[i]; }
  t.v.assign(data, data + n);
  params[canonical_name(name)] = std::move(t);
  return 0;
}

extern "C" int gsx_synth_set_param(gsx_synth* h, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!h) { set_error("null handle"); return -1; }
  h->finalized = false;
  return store_param(h->params, name, data, shape, ndim);
}

static const HostTensor* need(const std::map<std::string, HostTensor>& p, const std::string& name, size_t elems) {
  auto it = p.find(name);
  if (it == p.end()) { set_error("missing parameter: " + name); return nullptr; }
  if (it->second.v.size() != elems) {
    set_error("parameter " + name + " has " + std::to_string(it->second.v.size()) + " elements, expected " +
              std::to_string(elems));
    return nullptr;
  }
  return &it->second;
}

extern "C" int gsx_synth_finalize(gsx_synth* h) {
  if (!h) { set_error("null handle"); return -1; }
  const int Z = h->cfg.latent_size, L = h->L;
  const auto& P = h->params;
  // ---- mapping: weight * (sqrt2/sqrt(in)) * lr_mult, bias * lr_mult   (networks_stylegan.py:507-518, lr_mult .01 :135)
  for (int i = 0; i < 8; ++i) {
    const std::string p = "mapping." + std::to_string(2 * i + 1);
    const HostTensor* w = need(P, p + ".weight", (size_t)Z * Z);
    const HostTensor* b = need(P, p + ".bias", Z);
    if (!w || !b) return -1;
    const float sc = (float)(std::sqrt(2.0) / std::sqrt((double)Z)) * 0.01f;
    std::vector<float> ws(w->v), bs(b->v);
    for (auto& x : ws) x *= sc;
    for (auto& x : bs) x *= 0.01f;
    cudaFree(h->d_map_w[i]); cudaFree(h->d_map_b[i]);
    h->d_map_w[i] = dev_upload(ws); h->d_map_b[i] = dev_upload(bs);
    if (!h->d_map_w[i] || !h->d_map_b[i]) return -2;
  }
  // ---- affine (styles): weight / sqrt(512), gain 1 (:244)
  {
    std::vector<float> aw((size_t)h->S_total * Z), ab(h->S_total);
    std::vector<int> ul(h->S_total);
    for (int l = 0; l < h->nlayers; ++l) {
      const int r = 2 + l / 2, C = h->nf(r);
      const std::string p = "net" + std::to_string(r) + ".adain" + std::to_string(1 + (l & 1)) + ".affine";
      const HostTensor* w = need(P, p + ".weight", (size_t)2 * C * Z);
      const HostTensor* b = need(P, p + ".bias", (size_t)2 * C);
      if (!w || !b) return -1;
      const float sc = (float)(1.0 / std::sqrt((double)Z));
      for (size_t i = 0; i < w->v.size(); ++i) aw[(size_t)h->style_off[l] * Z + i] = w->v[i] * sc;
      for (int i = 0; i < 2 * C; ++i) { ab[h->style_off[l] + i] = b->v[i]; ul[h->style_off[l] + i] = l; }
    }
    cudaFree(h->d_aff_w); cudaFree(h->d_aff_b); cudaFree(h->d_unit_layer);
    h->d_aff_w = dev_upload(aw); h->d_aff_b = dev_upload(ab); h->d_unit_layer = dev_upload(ul);
    if (!h->d_aff_w || !h->d_aff_b || !h->d_unit_layer) return -2;
  }
  {
    const HostTensor* avg = need(P, "latent_avg", Z);
    const HostTensor* psi = need(P, "truncation_psi", h->nlayers);
    if (!avg || !psi) return -1;
    cudaFree(h->d_latent_avg); cudaFree(h->d_psi);
    h->d_latent_avg = dev_upload(avg->v); h->d_psi = dev_upload(psi->v);
  }
  // ---- constant tensor -> blocked act_t [C/8][1][by][bx][8]
  {
    const int C = h->nf(2), by = h->cfg.base_scale_y, bx = h->cfg.base_scale_x;
    const HostTensor* c = need(P, "constant_tensor", (size_t)C * by * bx);
    if (!c) return -1;
    std::vector<act_t> cb((size_t)C * by * bx);
    for (int ch = 0; ch < C; ++ch)
      for (int p = 0; p < by * bx; ++p)
        cb[((size_t)(ch / 8) * by * bx + p) * 8 + (ch & 7)] = to_act(c->v[(size_t)ch * by * bx + p]);
    cudaFree(h->d_const);
    h->d_const = dev_upload(cb);
  }
  // ---- synthesis blocks
  for (auto& b : h->blocks) {
    cudaFree(b.conv1.wpack_dev); cudaFree(b.conv2.wpack_dev);
    cudaFree(b.conv1.taps_dev); cudaFree(b.conv2.taps_dev);
    cudaFree(b.conv1.wf32_dev); cudaFree(b.conv2.wf32_dev);
    cudaFree(b.ns1); cudaFree(b.b1); cudaFree(b.ns2); cudaFree(b.b2); cudaFree(b.wt);
  }
  h->blocks.clear();
  for (int r = 2; r <= L; ++r) {
    SynthBlock b;
    b.r = r; b.C = h->nf(r); b.Cin = r > 2 ? h->nf(r - 1) : b.C;
    h->hw(r, b.H, b.W);
    const std::string p = "net" + std::to_string(r);
    // AdaIN folded into the consumer convs for the channel-thin blocks (<= 64 channels, >= 64^2 pixels): per-sample
    // modulated weights are then a few KB to 150 KB per sample; the wide blocks keep the separate apply pass (their
    // weight sets are MBs per sample).  A block's output can only stay un-normalised if the next block's conv_1 supports
    // per-sample weights (the stacked-phase plans do, the phase-grid plan of cout >= 64 layers does not).
    auto thin = [&](int rr) { int hh, ww; h->hw(rr, hh, ww); return g_opt_fold_apply && h->nf(rr) <= 64 && hh * ww >= 64 * 64; };
    b.mod2 = thin(r);
    b.t2 = thin(r) && (r == L || h->nf(r + 1) < 64);
    b.mod1 = r > 2 && !h->blocks.empty() && h->blocks.back().t2;
    if (r > 2) {
      const bool deconv = r >= 7;                                            // networks_stylegan.py:154
      const int k = deconv ? 4 : 3;
      const HostTensor* w = need(P, p + ".block0.weight", (size_t)b.C * b.Cin * k * k);
      if (!w) return -1;
      const float std_ = (float)(std::sqrt(2.0) / std::sqrt((double)k * k * b.Cin));   // :399-403
      std::vector<float> ws(w->v);
      for (auto& x : ws) x *= std_;
      // thin deconv layers (phases stacked along the MMA N dimension anyway): fold the blur and the whole first-half
      // epilogue into the conv -- removes the separate blur/noise/bias/lrelu/stats pass over the largest tensors
      // (measured r01: 16-channel 1024^2 layer 1.01 -> 0.81 ms; at 32 channels the heavier epilogue cancels the gain,
      //  GSX_FOLD_MAXC raises the limit for experiments)
      const int fold_maxc = g_opt_fold_deconv_maxc;
      if (deconv && b.C <= fold_maxc && b.C < 64) {
        set_error("");
        plan_conv(b.conv1, DECONV4B, b.H / 2, b.W / 2, b.Cin, 0, b.C, 0, nullptr, /*aux: noise tile*/ 1, 0, b.mod1 ? 1 : 0);
        b.fold = !*gsx_last_error() && b.conv1.g.NB == 1 && b.conv1.g.up_cols;
      }
      if (b.fold) {
        std::vector<float> wt((size_t)16 * b.Cin * b.C);
        for (int ci = 0; ci < b.Cin; ++ci)
          for (int co = 0; co < b.C; ++co)
            for (int t = 0; t < 16; ++t) wt[((size_t)t * b.Cin + ci) * b.C + co] = ws[((size_t)ci * b.C + co) * 16 + t];
        b.wt = dev_upload(wt);
        if (!b.wt) return -2;
      } else {
        set_error("");
        plan_conv(b.conv1, deconv ? DECONV4 : UPCONV3, b.H / 2, b.W / 2, b.Cin, 0, b.C, 0, nullptr, 0, 0, b.mod1 ? 1 : 0);
      }
      if (*gsx_last_error()) return -1;
      if (!upload_conv(b.conv1, ws.data())) return -2;
    }
    {
      const HostTensor* w = need(P, p + ".block2.0.weight", (size_t)b.C * b.C * 9);
      if (!w) return -1;
      const float std_ = (float)(std::sqrt(2.0) / std::sqrt(9.0 * b.C));
      std::vector<float> ws(w->v);
      for (auto& x : ws) x *= std_;
      set_error("");
      // after a folded deconv+blur the block's first half lives in a conv-produced tensor: store it phase-planar
      // (the in-place AdaIN pass does not care) and let conv_2 use the space-to-depth plan with dense boxes
      // (measured: conv_2 0.67 -> 0.58 ms, the producer's split stores cost 0.03 ms; GSX_PLANAR_G=0 turns it off)
      const bool planar_in = b.fold && !(tune_env("GSX_PLANAR_G") && atoi(tune_env("GSX_PLANAR_G")) == 0) && !tune_env("GSX_NO_S2D");
      plan_conv(b.conv2, CONV3, b.H, b.W, b.C, 0, b.C, 0, nullptr, /*aux: noise tile*/ 1, planar_in ? 1 : 0, b.mod2 ? 1 : 0);
      if (*gsx_last_error()) return -1;
      if (planar_in) b.conv1.out_planar = 1;
      if (!upload_conv(b.conv2, ws.data())) return -2;
    }
    const HostTensor* n1 = need(P, p + ".block1.0.scale_factors", b.C);
    const HostTensor* b1 = need(P, p + ".block1.1.bias", b.C);
    const HostTensor* n2 = need(P, p + ".block2.1.scale_factors", b.C);
    const HostTensor* b2 = need(P, p + ".block2.2.bias", b.C);
    if (!n1 || !b1 || !n2 || !b2) return -1;
    b.ns1 = dev_upload(n1->v); b.b1 = dev_upload(b1->v); b.ns2 = dev_upload(n2->v); b.b2 = dev_upload(b2->v);
    h->blocks.push_back(b);
  }
  // ---- ToRGB: 1x1, gain 1 -> std = 1/sqrt(C) (:118-126)
  {
    const int C = h->nf(L), nc = h->cfg.channels;
    const std::string p = "to_rgb" + std::to_string(L) + ".0";
    const HostTensor* w = need(P, p + ".weight", (size_t)nc * C);
    const HostTensor* b = need(P, p + ".bias", nc);
    if (!w || !b) return -1;
    std::vector<float> ws(w->v);
    for (auto& x : ws) x *= (float)(1.0 / std::sqrt((double)C));
    cudaFree(h->d_wrgb); cudaFree(h->d_brgb);
    h->d_wrgb = dev_upload(ws); h->d_brgb = dev_upload(b->v);
  }
  if (!cuda_ok(cudaDeviceSynchronize(), "finalize")) return -2;
  set_error("");
  h->finalized = true;
  return 0;
}

extern "C" int gsx_synth_device_counter(gsx_synth* h, int enable, uint64_t start) {
  if (!h) { set_error("null handle"); return -1; }
  if (!enable) { cudaFree(h->d_counter); h->d_counter = nullptr; return 0; }
  if (!h->d_counter && !cuda_ok(cudaMalloc(&h->d_counter, sizeof(unsigned long long)), "counter")) return -2;
  const unsigned long long v = start;
  return cuda_ok(cudaMemcpy(h->d_counter, &v, sizeof(v), cudaMemcpyHostToDevice), "counter init") ? 0 : -2;
}

extern "C" int gsx_synth_workspace_bytes(const gsx_synth* h, int n, size_t* bytes) {
  if (!h || !bytes || n <= 0) { set_error("bad argument"); return -1; }
  *bytes = synth_layout(h, n, nullptr).total;
  return 0;
}
extern "C" int gsx_synth_num_layers(const gsx_synth* h) { return h ? h->nlayers : -1; }
extern "C" int gsx_synth_feature_shape(const gsx_synth* h, int level, int* c, int* hgt, int* wid) {
  if (!h || level < 0 || level > h->L - 2) { set_error("bad level"); return -1; }
  *c = h->nf(level + 2);
  h->hw(level + 2, *hgt, *wid);
  return 0;
}

extern "C" int gsx_synth_forward(gsx_synth* h, int N, const float* z_dev, const float* psi_host,
                                 const float* const* noise_dev, uint64_t seed, uint64_t first_sample,
                                 float* img_f32_dev, uint8_t* img_u8_dev, float* const* feats_f32_dev, void* ws,
                                 size_t ws_bytes, gsx_stream stream) {
  if (!h || !h->finalized) { set_error("generator not finalized"); return -1; }
  if (N <= 0 || !ws) { set_error("bad argument"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SynthWs w = synth_layout(h, N, ws);
  if (w.total > ws_bytes) { set_error("workspace too small"); return -1; }
  { int th, tw; h->hw(h->L, th, tw); pdl_set_for_work((double)N * th * tw); }
  const int Z = h->cfg.latent_size;
  set_error("");
  if (z_dev) {
    if (!cuda_ok(cudaMemcpyAsync(w.z, z_dev, (size_t)N * Z * sizeof(float), cudaMemcpyDeviceToDevice, st), "copy z")) return -2;
  } else {
    ProfScope ps("latents", 4.0 * N * Z, 0, st);
    launch_fill_latents(w.z, N, Z, seed, first_sample, st, h->d_counter); g_launches++;
  }
  const float* psi = h->d_psi;
  if (psi_host) {
    if (!cuda_ok(cudaMemcpyAsync(w.psi, psi_host, h->nlayers * sizeof(float), cudaMemcpyHostToDevice, st), "copy psi")) return -2;
    psi = w.psi;
  }
  bool forked = false;
  // noise planes: explicit inputs, or all of them from the Philox generator in one launch
  std::vector<const float*> noise(h->nlayers);
  {
    NoisePlanes pl{};
    bool any = false, all = true;
    for (int l = 0; l < h->nlayers; ++l) {
      int hh, ww;
      h->hw(2 + l / 2, hh, ww);
      pl.ptr[l] = w.noise[l];
      pl.elems[l] = (size_t)hh * ww;
      if (noise_dev && noise_dev[l]) { noise[l] = noise_dev[l]; all = false; }
      else { noise[l] = w.noise[l]; any = true; }
    }
    if (any && all && h->nlayers <= 24) {
      size_t tot = 0;
      for (int l = 0; l < h->nlayers; ++l) tot += pl.elems[l];
      // The noise planes depend on nothing but (seed, sample index): generate them on a side stream while the mapping
      // network (latency-bound, 128 CTAs) runs -- fork / join by events, so the caller's stream order (and a CUDA-graph
      // capture of it) still covers everything.  Per-launch profiling keeps the single stream.
      cudaStream_t ns = st;
      if (!g_prof_on) {
        if (!h->side) {
          cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking);
          cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
          cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
        }
        if (h->side && h->ev_fork && h->ev_join) {
          cudaEventRecord(h->ev_fork, st);
          cudaStreamWaitEvent(h->side, h->ev_fork, 0);
          ns = h->side;
          forked = true;
        }
      }
      ProfScope ps("noise", 4.0 * N * tot, 0, st);
      launch_fill_noise_all(pl, h->nlayers, N, seed, first_sample, ns, h->d_counter); g_launches++;
      if (forked) cudaEventRecord(h->ev_join, h->side);
    } else if (any) {
      for (int l = 0; l < h->nlayers; ++l) {
        if (noise[l] != w.noise[l]) continue;
        ProfScope ps("noise", 4.0 * N * pl.elems[l], 0, st);
        launch_fill_noise(w.noise[l], pl.elems[l], N, seed, first_sample, l, st, h->d_counter); g_launches++;
      }
    }
  }
  // mapping MLP + all style affines (truncation folded in): one cooperative launch
  {
    MapArgs m{};
    m.z = w.z; m.ya = w.wa; m.yb = w.wb;
    for (int i = 0; i < 8; ++i) { m.W[i] = h->d_map_w[i]; m.b[i] = h->d_map_b[i]; }
    m.Waff = h->d_aff_w; m.baff = h->d_aff_b; m.latent_avg = h->d_latent_avg; m.psi = psi; m.unit_layer = h->d_unit_layer;
    m.styles = w.styles; m.N = N; m.S = h->S_total;
    ProfScope ps("map+styles", 4.0 * (8.0 * Z * Z + (double)h->S_total * Z) + 4.0 * N * (Z + h->S_total),
                 2.0 * N * Z * (8.0 * Z + h->S_total), st);
    if (!launch_mapping(m, st)) { cuda_ok(cudaGetLastError(), "mapping launch"); return -2; }
    g_launches++;
  }
  if (forked) cudaStreamWaitEvent(st, h->ev_join, 0);
  h->last_n = N;
  if (h->d_counter) { launch_advance_counter(h->d_counter, (unsigned long long)N, st); g_launches++; }   // after its readers

  for (size_t bi = 0; bi < h->blocks.size(); ++bi) {
    const SynthBlock& b = h->blocks[bi];
    const int l1 = 2 * (int)bi, l2 = l1 + 1;
    const std::string tag = "g" + std::to_string(b.r) + ".";
    const double act_bytes = 2.0 * N * b.C * b.H * b.W, plane_bytes = 4.0 * N * b.H * b.W;
    float* st1 = w.partial[l1];
    float* st2 = w.partial[l2];
    // AdaIN 2 of the previous block folded into this block's conv_1: feat[bi-1] holds the un-normalised tensor
    const float* coef_in = b.mod1 ? w.coef[l1 - 1] : nullptr;
    const ModBufs* m1 = b.mod1 ? &w.mod1[bi] : nullptr;
    if (b.mod1) {
      ProfScope ps(tag + "modulate1", 0, 0, st);
      launch_modulate(b.conv1, coef_in, b.fold ? b.b1 : nullptr, N, m1->w, m1->bias_n, m1->bdelta, st); g_launches++;
    }
    Pass1Args p1{};
    p1.out = w.bufB; p1.C = b.C; p1.N = N; p1.H = b.H; p1.W = b.W;
    p1.nscale = b.ns1; p1.bias = b.b1; p1.noise = noise[l1]; p1.stats = st1;
    if (b.fold) {
      { ProfScope ps(tag + "border", 0, 0, st);
        launch_deconv_border(w.feat[bi - 1], b.wt, w.e_rows, w.e_cols, N, b.Cin, b.C, b.H / 2, b.W / 2, st, coef_in); g_launches++; }
      ConvEpi e{};
      e.out = w.bufB; e.Ho = b.H; e.Wo = b.W; e.up = 1; e.Cout = b.C;
      e.flags = EPI_LRELU | EPI_STATS;
      e.bias = b.b1; e.nscale = b.ns1; e.noise = noise[l1];
      e.stats = st1; e.stats_T = w.stats_T[l1];
      e.e_rows = w.e_rows; e.e_cols = w.e_cols;
      if (!run_conv(b.conv1, N, w.feat[bi - 1], nullptr, e, st, (tag + "deconv+blur").c_str(), m1)) return -2;
    } else if (b.r == 2) {
      p1.in = h->d_const; p1.in_broadcast = 1; p1.blur = 0;
    } else {
      ConvEpi e{};
      e.out = w.bufA; e.Ho = b.H; e.Wo = b.W; e.up = 1; e.flags = 0; e.Cout = b.C;
      if (!run_conv(b.conv1, N, w.feat[bi - 1], nullptr, e, st, (tag + (b.r >= 7 ? "deconv" : "upconv")).c_str(), m1)) return -2;
      p1.in = w.bufA; p1.in_broadcast = 0; p1.blur = 1;
    }
    if (!b.fold) { ProfScope ps(tag + "pass1", (b.r == 2 ? 1.0 : 2.0) * act_bytes + plane_bytes, 0, st); launch_pass1(p1, st); g_launches++; }
    // the coefficients of AdaIN 1 are only read by the apply pass of a block that does not fold it: computed there
    const bool inline1 = !b.mod2 && g_opt_inline_finalize;
    if (!inline1) {
      ProfScope ps(tag + "finalize", 0, 0, st);
      launch_finalize(st1, w.stats_T[l1], N, b.C, b.H * b.W, w.styles, h->S_total, h->style_off[l1], w.coef[l1], st);
      g_launches++;
    }
    const ModBufs* m2 = b.mod2 ? &w.mod2[bi] : nullptr;
    if (b.mod2) {
      // AdaIN 1 folded into conv_2: no pass over the tensor, only the per-sample weights / bias
      ProfScope ps(tag + "modulate2", 0, 0, st);
      launch_modulate(b.conv2, w.coef[l1], b.b2, N, m2->w, m2->bias_n, m2->bdelta, st); g_launches++;
    } else {
      ApplyArgs a1{};
      a1.in = w.bufB; a1.out = w.bufB; a1.C = b.C; a1.N = N; a1.H = b.H; a1.W = b.W;
      a1.coef = w.coef[l1];
      if (inline1) { a1.coef = nullptr; a1.partial = st1; a1.T = w.stats_T[l1]; a1.styles = w.styles; a1.style_stride = h->S_total; a1.style_off = h->style_off[l1]; }
      ProfScope ps(tag + "apply1", 2.0 * act_bytes, 0, st); launch_apply(a1, st); g_launches++;
    }

    act_t* t2buf = b.t2 ? w.feat[bi] : w.bufA;          // folded AdaIN 2: the un-normalised tensor IS the block's output
    ConvEpi e2{};
    e2.out = t2buf; e2.Ho = b.H; e2.Wo = b.W; e2.up = 0; e2.Cout = b.C;
    e2.bias = b.b2; e2.nscale = b.ns2; e2.noise = noise[l2];
    const bool fused_stats = b.conv2.g.NB == 1;
    e2.flags = EPI_LRELU | (fused_stats ? EPI_STATS : 0);
    e2.stats = fused_stats ? st2 : nullptr;
    e2.stats_T = w.stats_T[l2];
    if (!run_conv(b.conv2, N, w.bufB, nullptr, e2, st, (tag + "conv2").c_str(), m2)) return -2;
    if (!fused_stats) { ProfScope ps(tag + "stats", act_bytes, 0, st); launch_stats(t2buf, st2, b.C, N, b.H * b.W, st); g_launches++; }
    const bool last = b.r == h->L;
    // AdaIN 2 of a block that materialises its feature (and is not the ToRGB block): same
    const bool inline2 = !b.t2 && !last && g_opt_inline_finalize;
    if (!inline2) {
      ProfScope ps(tag + "finalize", 0, 0, st);
      launch_finalize(st2, w.stats_T[l2], N, b.C, b.H * b.W, w.styles, h->S_total, h->style_off[l2], w.coef[l2], st);
      g_launches++;
    }

    ApplyArgs a2{};
    a2.in = t2buf; a2.out = b.t2 ? nullptr : w.feat[bi]; a2.C = b.C; a2.N = N; a2.H = b.H; a2.W = b.W;
    a2.coef = w.coef[l2];
    if (inline2) { a2.coef = nullptr; a2.partial = st2; a2.T = w.stats_T[l2]; a2.styles = w.styles; a2.style_stride = h->S_total; a2.style_off = h->style_off[l2]; }
    a2.out_nchw_f32 = feats_f32_dev ? feats_f32_dev[bi] : nullptr;
    if (last) {
      a2.wrgb = h->d_wrgb; a2.brgb = h->d_brgb; a2.img_f32 = img_f32_dev; a2.img_u8 = img_u8_dev; a2.nc = h->cfg.channels;
    }
    if (!b.t2 || last || a2.out_nchw_f32) {
      // (with AdaIN 2 folded into the consumers this pass survives only as ToRGB on the last block -- a read of the
      //  tensor and a 3-byte write -- and as the fp32 NCHW feature output of the drop-in mode)
      double by = (b.t2 ? 1.0 : 2.0) * act_bytes;
      if (last) by += (img_u8_dev ? 1.0 : 0.0) * N * b.H * b.W * h->cfg.channels + (img_f32_dev ? 4.0 : 0.0) * N * b.H * b.W * h->cfg.channels;
      if (a2.out_nchw_f32) by += 2.0 * act_bytes;
      cudaStream_t on = st;
      if (last && b.t2 && !a2.out_nchw_f32 && t_defer_rgb && g_opt_defer_rgb && !g_prof_on && h->side) {
        if (!h->ev_rgb_fork) {
          cudaEventCreateWithFlags(&h->ev_rgb_fork, cudaEventDisableTiming);
          cudaEventCreateWithFlags(&h->ev_rgb_done, cudaEventDisableTiming);
        }
        if (h->ev_rgb_fork && h->ev_rgb_done && cudaEventRecord(h->ev_rgb_fork, st) == cudaSuccess &&
            cudaStreamWaitEvent(h->side, h->ev_rgb_fork, 0) == cudaSuccess)
          on = h->side;
      }
      ProfScope ps(tag + (last ? (b.t2 ? "rgb" : "apply2+rgb") : "apply2"), by, last ? 2.0 * N * b.H * b.W * b.C * h->cfg.channels : 0, on);
      launch_apply(a2, on); g_launches++;
      if (on != st) { cudaEventRecord(h->ev_rgb_done, on); h->rgb_pending = true; }
    }
  }
  if (!cuda_ok(cudaGetLastError(), "synth forward")) return -2;
  return 0;
}

extern "C" int gsx_synth_export_noise(gsx_synth* h, int n, int layer, float* out_dev, const void* ws, gsx_stream stream) {
  if (!h || layer < 0 || layer >= h->nlayers) { set_error("bad layer"); return -1; }
  SynthWs w = synth_layout(h, n, const_cast<void*>(ws));
  int hh, ww;
  h->hw(2 + layer / 2, hh, ww);
  return cuda_ok(cudaMemcpyAsync(out_dev, w.noise[layer], (size_t)n * hh * ww * sizeof(float), cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)), "export noise") ? 0 : -2;
}
extern "C" int gsx_synth_export_latents(gsx_synth* h, int n, float* out_dev, const void* ws, gsx_stream stream) {
  if (!h) { set_error("null handle"); return -1; }
  SynthWs w = synth_layout(h, n, const_cast<void*>(ws));
  return cuda_ok(cudaMemcpyAsync(out_dev, w.z, (size_t)n * h->cfg.latent_size * sizeof(float), cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)), "export z") ? 0 : -2;
}

// =============================================================================================
// decoder
// =============================================================================================
struct DecLevel {
  int H, W, cin, f, fnext;
  ConvLayer cvt, conv_a, conv_b, shortcut, final_;
  ConvLayer cvt_ps;                       // cvt with per-sample weights: the generator left this level's feature un-normalised
  bool has_cvt_ps = false;                //   (AdaIN folded into the consumers, modulate.cu)
  bool has_shortcut = false;
  float *b_cvt = nullptr, *b_a = nullptr, *b_b = nullptr, *b_sc = nullptr, *b_final = nullptr;
};

struct gsx_dec {
  gsx_dec_cfg cfg;
  int nf, num_classes;
  std::map<std::string, HostTensor> params;
  bool finalized = false;
  std::vector<DecLevel> levels;
  // side branches of a forward pass (gsx_dec_forward): one stream per level for the cvt block, one for the 1x1 shortcuts
  std::vector<cudaStream_t> cvt_streams;
  cudaStream_t sc_stream = nullptr;
  std::vector<cudaEvent_t> events;
};

struct DecWs {
  std::vector<act_t*> feat, c, a, sc, prev;     // prev[i] = input "prev" of level i (null at 0)
  std::vector<ModBufs> mod;                     // per level: per-sample operands of cvt_ps
  size_t total;
};

static DecWs dec_layout(const gsx_dec* d, int N, void* base, bool own_feats) {
  DecWs w;
  Arena ar(base);
  const int nf = d->nf;
  w.feat.assign(nf, nullptr); w.c.assign(nf, nullptr); w.a.assign(nf, nullptr); w.sc.assign(nf, nullptr);
  w.prev.assign(nf + 1, nullptr);
  for (int i = 0; i < nf; ++i) {
    const int H = d->cfg.base_y << i, W = d->cfg.base_x << i;
    const size_t plane = (size_t)N * H * W;
    if (own_feats) w.feat[i] = ar.take<act_t>(plane * d->cfg.in_channels[i]);
    w.c[i] = ar.take<act_t>(plane * d->cfg.features[i]);
    if (i < nf - 1) {
      w.a[i] = ar.take<act_t>(plane * 4 * d->cfg.features[i + 1]);
      w.sc[i] = ar.take<act_t>(plane * d->cfg.features[i + 1]);
      w.prev[i + 1] = ar.take<act_t>(plane * 4 * d->cfg.features[i + 1]);
    }
  }
  for (int i = 0; i < nf; ++i)
    w.mod.push_back((size_t)i < d->levels.size() && d->levels[i].has_cvt_ps ? take_mod(ar, d->levels[i].cvt_ps, N) : ModBufs());
  w.total = ar.off;
  return w;
}

extern "C" int gsx_dec_create(const gsx_dec_cfg* cfg, gsx_dec** out) {
  if (!cfg || !out || cfg->num_levels < 1 || cfg->num_levels > 16) { set_error("bad decoder config"); return -1; }
  int dev;
  if (!cuda_ok(cudaGetDevice(&dev), "cudaGetDevice (libgsx has no CPU fallback)")) return -2;
  for (int i = 0; i < cfg->num_levels; ++i)
    if (cfg->in_channels[i] % 16 || cfg->features[i] % 16) { set_error("decoder channels must be multiples of 16"); return -1; }
  if (cfg->features[cfg->num_levels] > 16) { set_error("at most 16 classes"); return -1; }
  gsx_dec* d = new gsx_dec();
  d->cfg = *cfg;
  d->nf = cfg->num_levels;
  d->num_classes = cfg->features[cfg->num_levels];
  *out = d;
  return 0;
}

static void free_level(DecLevel& l) {
  cudaFree(l.cvt.wpack_dev); cudaFree(l.conv_a.wpack_dev); cudaFree(l.conv_b.wpack_dev);
  cudaFree(l.shortcut.wpack_dev); cudaFree(l.final_.wpack_dev);
  cudaFree(l.cvt.taps_dev); cudaFree(l.conv_a.taps_dev); cudaFree(l.conv_b.taps_dev);
  cudaFree(l.shortcut.taps_dev); cudaFree(l.final_.taps_dev);
  cudaFree(l.cvt_ps.wpack_dev); cudaFree(l.cvt_ps.taps_dev); cudaFree(l.cvt_ps.wf32_dev);
  cudaFree(l.b_cvt); cudaFree(l.b_a); cudaFree(l.b_b); cudaFree(l.b_sc); cudaFree(l.b_final);
}

extern "C" void gsx_dec_destroy(gsx_dec* d) {
  if (!d) return;
  for (auto& l : d->levels) free_level(l);
  for (cudaStream_t st : d->cvt_streams) if (st) cudaStreamDestroy(st);
  if (d->sc_stream) cudaStreamDestroy(d->sc_stream);
  for (cudaEvent_t e : d->events) if (e) cudaEventDestroy(e);
  delete d;
}

extern "C" int gsx_dec_set_param(gsx_dec* d, const char* name, const float* data, const int64_t* shape, int ndim) {
  if (!d) { set_error("null handle"); return -1; }
  d->finalized = false;
  return store_param(d->params, name, data, shape, ndim);
}

// conv (+ optional inference BatchNorm fold): w' = w*g, b' = (b - mean)*g + beta, g = gamma/sqrt(var+1e-5)
static bool fold_conv_bn(const std::map<std::string, HostTensor>& P, const std::string& conv, const std::string& bn,
                         int cout, size_t per_out, std::vector<float>& w, std::vector<float>& b) {
  const HostTensor* cw = need(P, conv + ".weight", (size_t)cout * per_out);
  const HostTensor* cb = need(P, conv + ".bias", cout);
  if (!cw || !cb) return false;
  w = cw->v; b = cb->v;
  if (!bn.empty()) {
    const HostTensor* ga = need(P, bn + ".gamma", cout);
    const HostTensor* be = need(P, bn + ".beta", cout);
    const HostTensor* rm = need(P, bn + ".running_mean", cout);
    const HostTensor* rv = need(P, bn + ".running_var", cout);
    if (!ga || !be || !rm || !rv) return false;
    for (int co = 0; co < cout; ++co) {
      const float g = ga->v[co] / std::sqrt(rv->v[co] + 1e-5f);
      for (size_t i = 0; i < per_out; ++i) w[(size_t)co * per_out + i] *= g;
      b[co] = (b[co] - rm->v[co]) * g + be->v[co];
    }
  }
  return true;
}

extern "C" int gsx_dec_finalize(gsx_dec* d) {
  if (!d) { set_error("null handle"); return -1; }
  for (auto& l : d->levels) free_level(l);
  d->levels.clear();
  const auto& P = d->params;
  const int nf = d->nf;
  const bool bn = d->cfg.use_bn != 0;
  // The three thin tensors at the output resolution that are produced by a conv epilogue and consumed by a thin
  // 3x3 conv -- conv_a -> conv_b of the last res-block, conv_b -> final conv, top cvt -> final conv -- are stored
  // phase-planar, so that the consumers can use the space-to-depth plan with dense TMA boxes (plan.cpp).
  // GSX_PLANAR: 0 = off, 1 = final conv only, 2 = also the last conv_b (default).
  const int topH = d->cfg.base_y << (nf - 1), topW = d->cfg.base_x << (nf - 1);
  const int planar_env = tune_env("GSX_PLANAR") ? atoi(tune_env("GSX_PLANAR")) : 2;
  const bool planar_ok = nf >= 2 && topH % 2 == 0 && topW % 2 == 0 && topH >= 16 && topW >= 16 &&
                         d->cfg.features[nf - 1] <= 16 && !tune_env("GSX_NO_S2D");
  const bool planar_final = planar_ok && planar_env >= 1;
  const bool planar_cb = planar_ok && planar_env >= 2;
  for (int i = 0; i < nf; ++i) {
    DecLevel l;
    l.H = d->cfg.base_y << i; l.W = d->cfg.base_x << i;
    l.cin = d->cfg.in_channels[i]; l.f = d->cfg.features[i]; l.fnext = d->cfg.features[i + 1];
    std::vector<float> w, b;
    const std::string cv = "cvt_block_" + std::to_string(i);
    if (!fold_conv_bn(P, cv + ".0", bn ? cv + ".1" : "", l.f, (size_t)l.cin * 9, w, b)) return -1;
    set_error("");
    plan_conv(l.cvt, CONV3, l.H, l.W, l.cin, 0, l.f, 0, nullptr);
    if (*gsx_last_error()) return -1;
    l.cvt.out_planar = (i == nf - 1 && planar_final) ? 1 : 0;
    if (!upload_conv(l.cvt, w.data())) return -2;
    l.b_cvt = dev_upload(b);
    if (g_opt_fold_apply && l.cin <= 64 && l.H * l.W >= 64 * 64) {
      // the generator may hand this level over un-normalised (its AdaIN 2 folded into the consumers): same conv with
      // per-sample weights built by launch_modulate at forward time
      set_error("");
      plan_conv(l.cvt_ps, CONV3, l.H, l.W, l.cin, 0, l.f, 0, nullptr, 0, 0, /*per_sample*/ 1);
      if (!*gsx_last_error()) {
        l.cvt_ps.out_planar = l.cvt.out_planar;
        if (!upload_conv(l.cvt_ps, w.data())) return -2;
        l.has_cvt_ps = true;
      }
      set_error("");
    }
    const int c0 = i > 0 ? l.f : l.f, c1 = i > 0 ? l.f : 0;          // concat(prev, cvt) (networks_seg.py:108-109)
    const int cin_main = c0 + c1;
    if (i < nf - 1) {
      const std::string mb = "main_block_" + std::to_string(i) + ".1";
      const int j_b = bn ? 3 : 2;
      if (!fold_conv_bn(P, mb + ".base_layers.0", bn ? mb + ".base_layers.1" : "", l.fnext, (size_t)cin_main * 9, w, b)) return -1;
      plan_conv(l.conv_a, UPCONV3, l.H, l.W, c0, c1, l.fnext, 0, nullptr);
      if (*gsx_last_error()) return -1;
      const bool top_block = (i == nf - 2);
      l.conv_a.out_planar = (top_block && planar_cb) ? 1 : 0;
      if (!upload_conv(l.conv_a, w.data())) return -2;
      l.b_a = dev_upload(b);
      if (!fold_conv_bn(P, mb + ".base_layers." + std::to_string(j_b), bn ? mb + ".base_layers." + std::to_string(j_b + 1) : "",
                        l.fnext, (size_t)l.fnext * 9, w, b)) return -1;
      plan_conv(l.conv_b, CONV3, l.H * 2, l.W * 2, l.fnext, 0, l.fnext, 0, nullptr, /*aux: residual tile*/ 2,
                /*in_planar*/ (top_block && planar_cb) ? 1 : 0);
      if (*gsx_last_error()) return -1;
      l.conv_b.out_planar = (top_block && planar_final) ? 1 : 0;
      if (!upload_conv(l.conv_b, w.data())) return -2;
      l.b_b = dev_upload(b);
      l.has_shortcut = (l.fnext != cin_main);                        // networks_seg.py:35-41
      if (l.has_shortcut) {
        if (!fold_conv_bn(P, mb + ".shortcut.0", "", l.fnext, (size_t)cin_main, w, b)) return -1;
        plan_conv(l.shortcut, CONV1, l.H, l.W, c0, c1, l.fnext, 0, nullptr);
        if (*gsx_last_error()) return -1;
        if (!upload_conv(l.shortcut, w.data())) return -2;
        l.b_sc = dev_upload(b);
      } else if (c1 > 0) {
        set_error("identity shortcut over a concatenated input is not supported");
        return -1;
      }
    } else {
      const std::string mb = "main_block_" + std::to_string(i) + ".0";
      if (!fold_conv_bn(P, mb, "", l.fnext, (size_t)cin_main * 9, w, b)) return -1;
      plan_conv(l.final_, CONV3, l.H, l.W, c0, c1, l.fnext, l.fnext, nullptr, 0, /*in_planar*/ planar_final ? 1 : 0);
      if (*gsx_last_error()) return -1;
      if (!upload_conv(l.final_, w.data())) return -2;
      b.resize(16, 0.f);
      l.b_final = dev_upload(b);
    }
    d->levels.push_back(l);
  }
  if (!cuda_ok(cudaDeviceSynchronize(), "dec finalize")) return -2;
  set_error("");
  if (d->cvt_streams.empty()) {        // created once, up front: nothing is allocated while a forward pass is being captured
    d->cvt_streams.assign(nf, nullptr);
    bool ok = true;
    for (int i = 0; ok && i < nf; ++i) ok = cuda_ok(cudaStreamCreateWithFlags(&d->cvt_streams[i], cudaStreamNonBlocking), "stream");
    ok = ok && cuda_ok(cudaStreamCreateWithFlags(&d->sc_stream, cudaStreamNonBlocking), "stream");
    d->events.assign((size_t)3 * nf + 4, nullptr);
    for (size_t i = 0; ok && i < d->events.size(); ++i) ok = cuda_ok(cudaEventCreateWithFlags(&d->events[i], cudaEventDisableTiming), "event");
    if (!ok) return -2;
  }
  d->finalized = true;
  return 0;
}

extern "C" int gsx_dec_workspace_bytes(const gsx_dec* d, int n, size_t* bytes) {
  if (!d || !bytes || n <= 0) { set_error("bad argument"); return -1; }
  *bytes = dec_layout(d, n, nullptr, true).total;
  return 0;
}

extern "C" int gsx_dec_forward(gsx_dec* d, int N, const float* const* feats_f32_dev, const gsx_synth* synth,
                               const void* synth_ws, float* logits_dev, uint8_t* mask_dev, void* ws, size_t ws_bytes,
                               gsx_stream stream) {
  if (!d || !d->finalized) { set_error("decoder not finalized"); return -1; }
  if (N <= 0 || !ws || !mask_dev) { set_error("bad argument"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nf = d->nf;
  const bool own = feats_f32_dev != nullptr;
  DecWs w = dec_layout(d, N, ws, true);
  if (w.total > ws_bytes) { set_error("decoder workspace too small"); return -1; }
  pdl_set_for_work((double)N * d->levels[nf - 1].H * d->levels[nf - 1].W);
  std::vector<const act_t*> feat(nf);
  std::vector<const float*> feat_coef(nf, nullptr);      // per level: AdaIN coefficients still to be applied (un-normalised feature)
  if (own) {
    for (int i = 0; i < nf; ++i) {
      const DecLevel& l = d->levels[i];
      ProfScope ps("d" + std::to_string(i) + ".to_blocked", 6.0 * N * l.cin * l.H * l.W, 0, st);
      launch_nchw_to_blocked(feats_f32_dev[i], w.feat[i], l.cin, N, l.H * l.W, st); g_launches++;
      feat[i] = w.feat[i];
    }
  } else {
    if (!synth || !synth_ws) { set_error("no features given"); return -1; }
    if ((int)synth->blocks.size() != nf) { set_error("generator/decoder level mismatch"); return -1; }
    SynthWs sw = synth_layout(synth, N, const_cast<void*>(synth_ws));
    for (int i = 0; i < nf; ++i) {
      const DecLevel& l = d->levels[i];
      const SynthBlock& b = synth->blocks[i];
      if (b.C != l.cin || b.H != l.H || b.W != l.W) { set_error("generator/decoder feature shape mismatch"); return -1; }
      feat[i] = sw.feat[i];
      if (b.t2) {
        if (!l.has_cvt_ps) { set_error("the generator folds AdaIN at a level the decoder has no per-sample conv for"); return -1; }
        feat_coef[i] = sw.coef[2 * i + 1];
      }
    }
  }
  // Branches.  The cvt block of a level needs only that level's feature map, the 1x1 shortcut only the level's input; at the
  // levels whose convs fill a fraction of the GPU (all of them at batch 1; 4^2..32^2 at batch 32) they run on side streams
  // beside the main chain conv_a -> conv_b and join it by events (the same shape as the training step's branches,
  // train_step.cu).  Off while the per-layer table is being taken (its events time the caller's stream).
  const bool branches = g_opt_dec_branches && !g_prof_on && !d->cvt_streams.empty();
  auto is_small = [&](int i) { return branches && (double)N * d->levels[i].H * d->levels[i].W <= 1024.0 * g_opt_dec_branch_kpx; };
  size_t next_event = 0;
  bool ev_ok = true;
  auto record = [&](cudaStream_t on) -> cudaEvent_t {
    if (next_event >= d->events.size()) { ev_ok = false; return nullptr; }
    cudaEvent_t e = d->events[next_event++];
    ev_ok = ev_ok && cudaEventRecord(e, on) == cudaSuccess;
    return e;
  };
  auto wait = [&](cudaStream_t on, cudaEvent_t e) { if (e) ev_ok = ev_ok && cudaStreamWaitEvent(on, e, 0) == cudaSuccess; };
  auto run_cvt = [&](int i, cudaStream_t on) -> bool {
    const DecLevel& l = d->levels[i];
    ConvEpi e{};
    e.out = w.c[i]; e.Ho = l.H; e.Wo = l.W; e.flags = EPI_LRELU; e.Cout = l.f; e.bias = l.b_cvt;
    if (feat_coef[i]) {
      const ModBufs& m = w.mod[i];
      { ProfScope ps("d" + std::to_string(i) + ".modulate", 0, 0, on);
        launch_modulate(l.cvt_ps, feat_coef[i], l.b_cvt, N, m.w, m.bias_n, m.bdelta, on); g_launches++; }
      return run_conv(l.cvt_ps, N, feat[i], nullptr, e, on, ("d" + std::to_string(i) + ".cvt").c_str(), &m);
    }
    return run_conv(l.cvt, N, feat[i], nullptr, e, on, ("d" + std::to_string(i) + ".cvt").c_str());
  };
  std::vector<cudaEvent_t> ev_cvt(nf, nullptr);
  {
    cudaEvent_t start = nullptr;
    for (int i = 0; i < nf; ++i) {
      if (!is_small(i)) continue;
      if (!start) start = record(st);                    // the features (and everything earlier on the caller's stream)
      wait(d->cvt_streams[i], start);
      if (!run_cvt(i, d->cvt_streams[i])) return -2;
      ev_cvt[i] = record(d->cvt_streams[i]);
    }
  }
  for (int i = 0; i < nf; ++i) {
    const DecLevel& l = d->levels[i];
    if (ev_cvt[i]) wait(st, ev_cvt[i]);
    else if (!run_cvt(i, st)) return -2;
    const act_t* x0 = i > 0 ? w.prev[i] : w.c[i];
    const act_t* x1 = i > 0 ? w.c[i] : nullptr;
    if (i < nf - 1) {
      const act_t* sc = x0;
      cudaEvent_t ev_sc = nullptr;
      if (l.has_shortcut) {
        ConvEpi e{};
        e.out = w.sc[i]; e.Ho = l.H; e.Wo = l.W; e.flags = 0; e.Cout = l.fnext; e.bias = l.b_sc;
        cudaStream_t on = st;
        if (is_small(i)) { on = d->sc_stream; wait(on, record(st)); }       // x0 / x1 are complete on the caller's stream here
        if (!run_conv(l.shortcut, N, x0, x1, e, on, ("d" + std::to_string(i) + ".shortcut").c_str())) return -2;
        if (on != st) ev_sc = record(on);
        sc = w.sc[i];
      }
      {
        ConvEpi e{};
        e.out = w.a[i]; e.Ho = 2 * l.H; e.Wo = 2 * l.W; e.up = 1; e.flags = EPI_LRELU; e.Cout = l.fnext; e.bias = l.b_a;
        if (!run_conv(l.conv_a, N, x0, x1, e, st, ("d" + std::to_string(i) + ".conv_a").c_str())) return -2;
      }
      wait(st, ev_sc);
      {
        ConvEpi e{};
        e.out = w.prev[i + 1]; e.Ho = 2 * l.H; e.Wo = 2 * l.W; e.flags = EPI_LRELU; e.Cout = l.fnext; e.bias = l.b_b;
        e.addsrc = sc;
        if (!run_conv(l.conv_b, N, w.a[i], nullptr, e, st, ("d" + std::to_string(i) + ".conv_b").c_str())) return -2;
      }
    } else {
      ConvEpi e{};
      e.Ho = l.H; e.Wo = l.W; e.flags = EPI_ARGMAX; e.Cout = l.fnext; e.bias = l.b_final;
      e.mask = mask_dev; e.logits = logits_dev; e.num_classes = d->num_classes;
      if (!run_conv(l.final_, N, x0, x1, e, st, ("d" + std::to_string(i) + ".final+argmax").c_str())) return -2;
    }
  }
  if (!ev_ok) { set_error("dec forward: event record / wait failed"); return -2; }
  return cuda_ok(cudaGetLastError(), "dec forward") ? 0 : -2;
}

// generator + decoder on device buffers; the image pass runs beside the decoder and is joined here
static int generate_dev_impl(gsx_synth* s, gsx_dec* d, int n, const float* z_dev, const float* psi_host, uint64_t seed, uint64_t first_sample,
                             uint8_t* img_u8_dev, uint8_t* mask_dev, void* synth_ws, size_t synth_ws_bytes, void* dec_ws,
                             size_t dec_ws_bytes, cudaStream_t st) {
  t_defer_rgb = true;
  int rc = gsx_synth_forward(s, n, z_dev, psi_host, nullptr, seed, first_sample, nullptr, img_u8_dev, nullptr, synth_ws, synth_ws_bytes, st);
  t_defer_rgb = false;
  if (!rc) rc = gsx_dec_forward(d, n, nullptr, s, synth_ws, nullptr, mask_dev, dec_ws, dec_ws_bytes, st);
  if (s && s->rgb_pending) {                               // also on the error path: the side stream must rejoin the caller's
    s->rgb_pending = false;
    if (!cuda_ok(cudaStreamWaitEvent(st, s->ev_rgb_done, 0), "join image pass") && !rc) rc = -2;
  }
  return rc;
}

extern "C" int gsx_generate_dev(gsx_synth* s, gsx_dec* d, int n, const float* z_dev, const float* psi_host, uint64_t seed,
                                uint64_t first_sample, uint8_t* img_u8_dev, uint8_t* mask_dev, void* synth_ws,
                                size_t synth_ws_bytes, void* dec_ws, size_t dec_ws_bytes, gsx_stream stream) {
  if (!s || !d || !img_u8_dev || !mask_dev) { set_error("bad argument"); return -1; }
  return generate_dev_impl(s, d, n, z_dev, psi_host, seed, first_sample, img_u8_dev, mask_dev, synth_ws, synth_ws_bytes, dec_ws,
                           dec_ws_bytes, static_cast<cudaStream_t>(stream));
}

extern "C" int gsx_generate_host(gsx_synth* s, gsx_dec* d, int n, const float* z_host, const float* psi_host,
                                 uint64_t seed, uint64_t first_sample, uint8_t* img_u8_host, uint8_t* mask_host,
                                 void* synth_ws, size_t synth_ws_bytes, void* dec_ws, size_t dec_ws_bytes,
                                 void* stage_dev, size_t stage_bytes, gsx_stream stream, gsx_stream copy_stream,
                                 int slot) {
  if (!s || !d || !stage_dev || slot < 0 || slot > 1) { set_error("bad argument"); return -1; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaStream_t cs = copy_stream ? static_cast<cudaStream_t>(copy_stream) : st;
  int H, W;
  s->hw(s->L, H, W);
  const int Z = s->cfg.latent_size, nc = s->cfg.channels;
  const size_t zb = align_up((size_t)n * Z * sizeof(float), 1024), ib = align_up((size_t)n * H * W * nc, 1024),
               mb = align_up((size_t)n * H * W, 1024);
  const size_t per_slot = zb + ib + mb;
  if (stage_bytes < per_slot * (copy_stream ? 2 : 1)) { set_error("staging buffer too small"); return -1; }
  uint8_t* base = static_cast<uint8_t*>(stage_dev) + (copy_stream ? (size_t)slot * per_slot : 0);
  float* z_dev = reinterpret_cast<float*>(base);
  uint8_t* img_dev = base + zb;
  uint8_t* mask_dev = base + zb + ib;
  if (copy_stream) {
    for (int i = 0; i < 2; ++i) {
      if (!s->ev_done[i]) cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming);
      if (!s->ev_copied[i]) { cudaEventCreateWithFlags(&s->ev_copied[i], cudaEventDisableTiming); cudaEventRecord(s->ev_copied[i], cs); }
    }
    // the staging slot is free once its previous device-to-host copies have finished
    if (!cuda_ok(cudaStreamWaitEvent(st, s->ev_copied[slot], 0), "wait copied")) return -2;
  }
  if (z_host && !cuda_ok(cudaMemcpyAsync(z_dev, z_host, (size_t)n * Z * sizeof(float), cudaMemcpyHostToDevice, st), "H2D z"))
    return -2;
  int rc = generate_dev_impl(s, d, n, z_host ? z_dev : nullptr, psi_host, seed, first_sample, img_dev, mask_dev, synth_ws, synth_ws_bytes,
                             dec_ws, dec_ws_bytes, st);
  if (rc) return rc;
  if (copy_stream) {
    // device-to-host copies run on the copy stream and overlap the next step's kernels
    if (!cuda_ok(cudaEventRecord(s->ev_done[slot], st), "record done")) return -2;
    if (!cuda_ok(cudaStreamWaitEvent(cs, s->ev_done[slot], 0), "wait done")) return -2;
  }
  if (img_u8_host && mask_host == img_u8_host + ib) {
    // image and mask sit back to back on both sides (GeneratePipeline allocates them that way): one copy per step
    if (!cuda_ok(cudaMemcpyAsync(img_u8_host, img_dev, ib + (size_t)n * H * W, cudaMemcpyDeviceToHost, cs), "D2H img+mask")) return -2;
  } else {
    if (img_u8_host && !cuda_ok(cudaMemcpyAsync(img_u8_host, img_dev, (size_t)n * H * W * nc, cudaMemcpyDeviceToHost, cs), "D2H img"))
      return -2;
    if (mask_host && !cuda_ok(cudaMemcpyAsync(mask_host, mask_dev, (size_t)n * H * W, cudaMemcpyDeviceToHost, cs), "D2H mask")) return -2;
  }
  if (copy_stream && !cuda_ok(cudaEventRecord(s->ev_copied[slot], cs), "record copied")) return -2;
  return 0;
}

// =============================================================================================
// per-launch profile (bench.py's per-layer table)
// =============================================================================================
extern "C" int gsx_profile_enable(int on) {
  for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  g_prof.clear();
  g_prof_on = on != 0;
  return 0;
}

extern "C" int gsx_profile_dump(char* buf, size_t cap) {
  if (!buf || cap == 0) { set_error("bad argument"); return -1; }
  if (!cuda_ok(cudaDeviceSynchronize(), "profile dump")) return -2;
  size_t off = 0;
  buf[0] = 0;
  for (auto& r : g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    const int n = snprintf(buf + off, cap - off, "%s\t%.6f\t%.0f\t%.0f\t%.0f\t%s\n", r.label.c_str(), ms, r.bytes, r.flops, r.flops_exec,
                           r.kernel.c_str());
    if (n < 0 || (size_t)n >= cap - off) break;
    off += (size_t)n;
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  g_prof.clear();
  return (int)off;
}
