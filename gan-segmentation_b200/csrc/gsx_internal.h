// Internal declarations shared by the translation units of libgsx (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace gsx {

// 16-bit storage / tensor-core operand type, fixed per build (one .so per type, see Makefile):
//   GSX_FP16=1 (default build, libgsx.so): IEEE half -- 11-bit significand; needed to meet the image
//                tolerance of the parity tests (DESIGN.md "numerics").
//   GSX_FP16=0 (libgsx_act_t.so): bfloat16, the type BASELINE.json's north star names.
// tcgen05 kind::f16 runs both at the same rate; accumulation is fp32 either way.
#ifndef GSX_FP16
#define GSX_FP16 1
#endif
#if GSX_FP16
typedef __half act_t;
typedef __half2 act2_t;
#define GSX_DTYPE_NAME "fp16"
#else
typedef __nv_bfloat16 act_t;
typedef __nv_bfloat162 act2_t;
#define GSX_DTYPE_NAME "act_t"
#endif

__host__ __device__ inline act_t to_act(float v) {
#if GSX_FP16
  v = v > 65504.f ? 65504.f : (v < -65504.f ? -65504.f : v);      // saturate instead of overflowing to inf
  return __float2half_rn(v);
#else
  return __float2bfloat16(v);
#endif
}

// Tuning builds (make TUNING=1): the GSX_* environment switches of the planner / forward passes and the kernel's role-
// isolation switches (GSX_DBG) are live.  The release build ignores the environment and compiles the switches out.
#ifndef GSX_TUNING
#define GSX_TUNING 0
#endif
inline const char* tune_env(const char* name) { return GSX_TUNING ? std::getenv(name) : nullptr; }

static const int kConvHeaderBytes = 2048;    // shiftconv smem header: mbarriers + tmem slot

// ----------------------------------------------------------------------------------------------
// Activation layout in HBM ("blocked"): [C/8][N][H][W][8] act_t -- 8 channels of one pixel form one
// 16-byte vector, pixels of a (channel-block, sample) plane are contiguous.  One TMA box
// {x, y, n, cb} of this layout lands in shared memory exactly as the K-major, no-swizzle UMMA
// operand layout (row = pixel, 16 B = 8 channels), and epilogue stores are 16 B per thread with
// consecutive lanes on consecutive pixels.
// ----------------------------------------------------------------------------------------------

enum ConvMode {
  CONV3 = 0,        // 3x3, stride 1, pad 1                      (reference Conv2DW / nn.Conv2D)
  UPCONV3 = 1,      // nearest x2 then 3x3 pad 1, as 4 phases of 2x2 taps on the low-res input
  DECONV4 = 2,      // 4x4 stride-2 pad-1 transposed conv, as 4 phases of 2x2 taps
  CONV1 = 3,        // 1x1 (a res-block shortcut on the upsampled input == 1x1 at low res; the consumer upsamples)
  DECONV4B = 4      // DECONV4 followed by the 3x3 [1,2,1]^2/16 blur, folded into ONE up-conv: 4 phases of 3x3 low-res
                    //   taps (the same 9 input shifts the stacked-phase DECONV4 already issues, so no extra MMAs).
                    //   Exact in the interior; the 1-pixel output border gets a correction (the blur zero-pads the
                    //   CROPPED deconv output) from launch_deconv_border, subtracted in the epilogue.
};

enum EpiFlags {
  EPI_LRELU = 1,     // leaky ReLU 0.2
  EPI_STATS = 2,     // accumulate per-(n,c) sum / sum of squares (InstanceNorm statistics)
  EPI_ARGMAX = 4     // first-max argmax over num_classes columns -> uint8 mask (+ optional fp32 logits)
};

static const int kMaxSlots = 16;

struct ConvGeom {
  int H, W, N;                 // input spatial size, batch
  int TH, TW, NB;              // tile: rows, cols (input resolution), samples
  int BH, BW;                  // TMA box rows / cols = tile + 2 halo
  int tiles_x, tiles_y, tiles_n;
  int kch0;                    // k-chunks taken from source 0 (rest from source 1)
  int n_k;                     // k-chunks in total
  int CBK;                     // 8-channel blocks per k-chunk (even)
  int n_mtiles;                // 128-row MMA tiles per CTA
  int N_tile;                  // MMA N (output channels per CTA)
  int n_ntiles;
  int n_groups;                // always 1 (kept for the plan dump)
  int up_cols;                 // 1: the 4 output phases of an up-conv are column blocks of ONE accumulator
                               //    (N_tile = 4*cout_tile); every distinct input shift is a single MMA whose
                               //    weight tile is zero for the phases that do not use that shift
  int cout_tile;               // output channels per CTA (= N_tile unless up_cols)
  int mt_stride;               // positions between consecutive MMA tiles (128)
  int aux_kind;                // epilogue operand staged per tile by TMA into smem (2 buffers): 0 none,
                               //   1 = noise plane tile [NB][TH][TW] fp32, 2 = residual tile (blocked, half resolution)
  int aux_off, aux_bytes;      // smem offset of the 2 aux buffers, bytes per buffer (128-aligned)
  int aux_bytes_tx;            // bytes one TMA box delivers (the mbarrier transaction count)
  int aux_bw, aux_bh;          // residual box: columns / rows at half resolution
  int aux_up;                  // 1: the noise tile is at the OUTPUT resolution of an up-conv (2TH x 2TW)
  int aux_shift;               // residual tile: tile origin >> aux_shift = residual coordinates (1, or 0 in s2d mode)
  int s2d;                     // 1: 3x3 conv on the 2x2 space-to-depth grid (thin layers, see plan.cpp): H, W, TH, TW,
                               //    BH, BW count 2x2 pixel BLOCKS; a stage holds the 4 input phase planes; the 4 output
                               //    phases are column blocks of one accumulator (up_cols = 1)
  int varn;                    // 1: stacked-phase up-conv (nearest-x2 + 3x3 / 4x4-s2 transposed conv) whose MMAs only touch the
                               //    weight blocks of the phases a shift feeds: the 4 phase blocks sit in the accumulator in the
                               //    cyclic order (0,0) (0,1) (1,1) (1,0), so that the 1 / 2 / 4 phases of a corner / edge / centre
                               //    shift are one contiguous column range [slot_n0, slot_n0 + slot_n) x cout_tile (the edge shift
                               //    (1,0) wraps and takes all 4); the centre shift is issued first (it initialises every column)
  signed char slot_n0[kMaxSlots], slot_n[kMaxSlots];   // varn: first phase block / number of phase blocks per slot
  int n_slots;                 // filter taps per CTA
  int phase_grid;              // 1: blockIdx.z selects the phase (wide up-convs)
  int stages;
  int cb_stride_bytes;         // NB*BH*BW*16
  int a_stage_bytes, b_stage_bytes;
  int a_stage_stride;          // a_stage_bytes rounded up to 128 (TMA destination alignment)
  int plane_stride;            // s2d: bytes between the phase planes of a stage (128-aligned)
  int in_planar;               // s2d: the inputs are stored phase-planar ([C/8][N][4 phases][H/2][W/2][8]): dense plane boxes
  int dbg;                     // tuning only (env GSX_DBG): 1 skip epilogue work, 2 skip MMAs, 4 skip activation loads,
                               //   8 issuers do not wait for operands (with 4: pure MMA issue rate)
  int chan_off, chan_n;        // smem table of the per-channel epilogue operands: [2][chan_n] floats (bias, noise scale)
  int per_sample;              // 1: weights and bias differ per sample (the producer's AdaIN is folded into this conv, see
                               //    modulate.cu): the B operand is streamed per tile from wpack + n * wpack_n_stride, the bias
                               //    comes from e.bias_n per accumulator column (one private smem copy per epilogue warp)
  int bias_w_off, bias_cols;   // per-warp bias copies: smem offset, floats per copy (= n_ntiles * N_tile)
  int a_off;                   // byte offset of the A stages in dynamic smem (after header + stats slots + channel table + aux)
  int tmem_cols;               // allocated TMEM columns (power of two) = acc_bufs * n_groups*n_mtiles*N_tile rounded up
  int acc_bufs;                // 2: accumulators double buffered (MMA of tile i+1 overlaps epilogue of tile i)
  int b_resident;              // 1: the whole packed weight set is loaded once per CTA and stays in smem
  int epi_groups;              // G: epilogue warps per TMEM lane quarter (CTA has 2 + 4G warps)
  int ctas_per_sm;             // persistent grid = min(work items, SMs * ctas_per_sm)
  unsigned magic_box, magic_bw;  // ceil(2^32 / (BH*BW)), ceil(2^32 / BW): division by multiply-high
  int smem_bytes;
  int slot_shift[4][kMaxSlots];     // [phase or 0][slot] -> A offset in positions (16 B): dy*BW+dx inside the box (+ plane)
  signed char slot_group[kMaxSlots];
  signed char slot_first[kMaxSlots];  // first tap of its accumulator group (overwrite instead of accumulate)
};

struct ConvEpi {
  act_t* out;                   // blocked [Cout/8][N][Ho][Wo][8] (null in argmax mode)
  int Ho, Wo;                  // output spatial size (2H,2W for the phase modes)
  int up;                      // 1: phase modes (out pixel = 2*pos + phase)
  int flags;
  int Cout;                    // real output channels (N_tile*n_ntiles may be padded above it)
  const float* bias;           // [Cout] or null
  const float* nscale;         // [Cout] per-channel noise scale or null
  const float* noise;          // [N][Ho][Wo] fp32 or null
  float* stats;                // per-tile partial sums [N][stats_T][Cout][2] (sum, sumsq) or null; no atomics,
  int stats_T;                 //   so the InstanceNorm statistics are bit-reproducible (T = tiles per sample)
  const act_t* addsrc;          // blocked [Cout/8][N][Ho/2][Wo/2][8], added after activation, or null
  int out_planar;              // 1: store the output phase-planar (its only consumer is a space-to-depth conv)
  const float* e_rows;         // DECONV4B border corrections: [N][2 (top,bottom)][Wo][Cout] and
  const float* e_cols;         //   [N][2 (left,right)][Ho][Cout] fp32, subtracted before noise / bias; or null
  const float* bias_n;         // per_sample: [N][bias_cols] bias per accumulator column (layer bias + the folded AdaIN shift
                               //   through all taps), replaces `bias`
  const float* bdelta;         // per_sample: [N][9][bias_cols] added to the accumulators of rows on the image border (class =
                               //   3*rowclass + colclass, 0 first / 1 interior / 2 last): minus the taps that fall outside
  unsigned char* mask;         // [N][Ho][Wo]
  float* logits;               // [N][num_classes][Ho][Wo] or null
  int num_classes;
};

struct ConvParams {
  CUtensorMap tm[2];
  CUtensorMap tm_aux;          // noise (3-D fp32) or residual (4-D blocked) tensor map when g.aux_kind != 0
  CUtensorMap tm_pl[2][4];     // s2d mode: per source, the 4 input phase planes (py*2+px) as 5-D maps
  ConvGeom g;
  ConvEpi e;
  const act_t* wpack;
  size_t wpack_n_stride;       // per_sample: elements between the packed weight sets of consecutive samples (else 0)
  // per (phase, slot) tap table in device memory: {A shift in bytes, accumulator group, first-of-group, 0}.
  // (Indexing the by-value parameter arrays dynamically would make the compiler copy the whole parameter
  //  block to local memory and turn every field access of the MMA issue loop into a local load.)
  const int4* taps;
};

// A planned + packed convolution layer (host side).
struct ConvLayer {
  int mode = CONV3;
  int cin0 = 0, cin1 = 0, cout = 0;   // channels of the two concatenated sources, real out channels
  int H = 0, W = 0;                   // input spatial size the layer was planned for
  int out_planar = 0;                 // the layer's output goes to a phase-planar tensor (set by the owner of the graph)
  ConvGeom g{};                       // geometry with N-independent fields filled
  act_t* wpack_dev = nullptr;
  size_t wpack_elems = 0;
  float* wf32_dev = nullptr;          // per-sample layers: the packed weight stream in fp32 (modulate.cu scales it per sample)
  int4* taps_dev = nullptr;           // 4 * kMaxSlots entries
};
void build_tap_table(const ConvGeom& g, int4* out /* 4*kMaxSlots */);

struct PlanOverride {
  int TH, TW, NB, CBK, N_tile, stages, phase_grid;   // 0 / -1 = keep default
  int epi_groups, acc_bufs, max_mtiles;              // 0 = keep default
  int hstack;                                        // must be <= 0 (variant removed, see plan.cpp)
  int s2d;                                           // -1 = keep default, 0/1 force
};

// plan.cpp
extern int g_plan_varn;           // 1: variable-N MMAs for the stacked-phase up-convs (gsx_set_option("varn", 0/1))
extern int g_plan_epi_groups;     // epilogue warps per TMEM lane quarter planned for new layers (2; 4 = experiment)
void plan_conv(ConvLayer& L, int mode, int H, int W, int cin0, int cin1, int cout, int argmax_classes,
               const PlanOverride* ov, int aux_kind = 0, int in_planar = 0, int per_sample = 0);
void make_noise_tensormap(CUtensorMap* tm, const void* base, int N, int H, int W, int boxW, int boxH, int boxN);
void finish_geom_for_batch(ConvGeom& g, int N);
// weights: CONV3/UPCONV3 (Cout,Cin,3,3); DECONV4 (Cin,Cout,4,4); CONV1 (Cout,Cin,1,1); fp32, already
// scaled (wscale / BN folded).  Returns packed act_t host buffer in the order the kernel streams it.
void pack_conv_weights(const ConvLayer& L, const float* w, std::vector<act_t>& out, std::vector<float>* out_f32 = nullptr);
// gather form of the packing (4 source indices into the reference weight tensor per packed element, -1 = none)
bool pack_conv_sources(const ConvLayer& L, std::vector<int>& src);
// slot -> (sy, sx) in {0,1,2}^2 of the 3x3 shift grid (CONV3 / stacked-phase plans): row-major, or centre-first for varn plans
inline void slot_yx(const ConvGeom& g, int slot, int* sy, int* sx) {
  static const int order[9] = {4, 0, 1, 2, 3, 5, 6, 7, 8};
  const int s = g.varn ? order[slot] : slot;
  *sy = s / 3; *sx = s % 3;
}
// accumulator column block <-> output phase (py*2 + px) of a stacked-phase plan
__host__ __device__ inline int phase_block(int varn, int py, int px) { return varn ? (py ? 3 - px : px) : 2 * py + px; }
inline int block_phase(int varn, int blk) { static const int inv[4] = {0, 1, 3, 2}; return varn ? inv[blk] : blk; }
// tap offset (dy, dx) in {-1,0,1} of every slot of a CONV3 / s2d / stacked-phase up-conv plan (input-grid units)
void slot_offsets(const ConvLayer& L, int* dy, int* dx);
void make_act_tensormap(CUtensorMap* tm, const void* base, int C, int N, int H, int W, int boxW, int boxH, int boxN,
                        int boxCB);
// s2d mode: phase plane (py, px) of a blocked activation tensor, boxes in 2x2-block coordinates
void make_plane_tensormap(CUtensorMap* tm, const void* base, int C, int N, int H, int W, int py, int px, int boxW, int boxH,
                          int boxN, int boxCB);
// same for a tensor that is already stored phase-planar: a dense 4-D box of plane py*2+px
void make_planar_tensormap(CUtensorMap* tm, const void* base, int C, int N, int H, int W, int plane, int boxW, int boxH,
                           int boxN, int boxCB);
// fills p.tm[] (or p.tm_pl[][] in s2d mode) for the layer's input tensors
void make_input_tensormaps(ConvParams& p, const ConvLayer& L, int N, const void* x0, const void* x1);

// api.cu: builds the tensor maps of one planned conv for this batch / these buffers and launches it
bool run_conv_layer(const ConvLayer& L, int N, const act_t* x0, const act_t* x1, const ConvEpi& epi, cudaStream_t st,
                    const char* label);

// launchers (shiftconv.cu / elementwise.cu / styles.cu / modulate.cu)
void launch_shiftconv(const ConvParams& p, cudaStream_t st);
// Per-sample operands of a conv whose input's AdaIN (coef [N][Cin][2] = a, b) is folded into it: modulated 16-bit weight
// streams wout [N][L.wpack_elems], interior bias bias_n [N][bias_cols] (layer bias `bias` [cout] included, may be null) and
// border-class corrections bdelta [N][9][bias_cols].  One launch.
void launch_modulate(const ConvLayer& L, const float* coef, const float* bias, int N, act_t* wout, float* bias_n, float* bdelta,
                     cudaStream_t st);

// Border correction of the folded deconv + blur (DECONV4B).  x: blocked low-res input [Cin/8][N][H][W][8];
// wt: the (scaled) transposed-conv weights rearranged to [4][4][Cin][Cout] fp32.
void launch_deconv_border(const act_t* x, const float* wt, float* e_rows, float* e_cols, int N, int Cin, int Cout, int H,
                          int W, cudaStream_t st, const float* coef = nullptr /* [N][Cin][2]: x is a*t + b of the stored t */);

struct Pass1Args {            // blur? + noise + bias + lrelu + stats  (generator, first half of a block)
  const act_t* in; act_t* out;  // blocked; in may have sample stride 0 (constant tensor)
  int C, N, H, W;
  int blur;                   // 1: 3x3 [1,2,1]^2/16 zero-pad blur first
  int in_broadcast;           // 1: input has a single sample (constant tensor)
  const float* nscale; const float* bias; const float* noise;   // [C], [C], [N][H][W]
  float* stats;               // per-block partial sums [N][pass1_tiles(H,W)][C][2]
};
void launch_pass1(const Pass1Args& a, cudaStream_t st);
int pass1_tiles(int H, int W);

struct ApplyArgs {            // InstanceNorm + AdaIN: out = (t-mean)*rstd*(scale+1)+shift
  const act_t* in; act_t* out;  // blocked; out may be null (only the optional outputs below are written)
  int C, N, H, W;
  const float* coef;          // [N][C][2]: out = in * coef[..][0] + coef[..][1]  (from launch_finalize)
  // optional ToRGB fused on the un-rounded values (last layer): rgb = Wrgb[3][C] x + brgb
  const float* wrgb; const float* brgb; float* img_f32; unsigned char* img_u8; int nc;
  float* out_nchw_f32;        // optional fp32 NCHW copy of the feature (drop-in mode)
  // optional: the statistics finalize inlined (coef == null): the coefficients of the block's 8 channels are computed from
  // the per-tile partial sums in the kernel's prologue, in finalize_kernel's summation order -- one launch less per AdaIN
  const float* partial; int T; const float* styles; int style_stride, style_off;
};
void launch_apply(const ApplyArgs& a, cudaStream_t st);

void launch_stats(const act_t* in, float* stats_partial, int C, int N, int HW, cudaStream_t st, int tiles = 0 /* 0: stats_tiles(HW) */);
int stats_tiles(int HW);
// Sums the per-tile partials in a fixed order and turns them into the AdaIN coefficients
//   a = rstd * (scale + 1),  b = shift - mean * a      (networks_stylegan.py:254-262, eps 1e-5)
// styles == nullptr: writes the raw (sum, sumsq) instead (tests).
void launch_finalize(const float* partial, int T, int N, int C, int HW, const float* styles, int style_stride,
                     int style_off, float* coef, cudaStream_t st);
void launch_blocked_to_nchw(const act_t* in, float* out, int C, int N, int HW, cudaStream_t st);
void launch_nchw_to_blocked(const float* in, act_t* out, int C, int N, int HW, cudaStream_t st);
// first_dev != nullptr: the first sample index is read from device memory (CUDA-graph replays), not from the argument
void launch_fill_noise(float* out, size_t plane_elems, int N, uint64_t seed, uint64_t first_sample, int layer,
                       cudaStream_t st, const unsigned long long* first_dev = nullptr);
void launch_fill_latents(float* z, int N, int Z, uint64_t seed, uint64_t first_sample, cudaStream_t st,
                         const unsigned long long* first_dev = nullptr);
void launch_advance_counter(unsigned long long* counter, unsigned long long by, cudaStream_t st);
struct NoisePlanes { float* ptr[24]; size_t elems[24]; };   // per style layer: plane base, elements per sample (mult. of 4)
void launch_fill_noise_all(const NoisePlanes& pl, int nlayers, int N, uint64_t seed, uint64_t first_sample, cudaStream_t st,
                           const unsigned long long* first_dev = nullptr);

struct DenseArgs {            // y[n][u] = act( sum_k x'[n][k] W[u][k] + b[u] ), W pre-scaled fp32
  const float* x; const float* W; const float* b; float* y;
  int N, K, U;
  int lrelu;
  int pixelnorm;              // 1: x' = x * rsqrt(mean(x^2)+1e-8)  (first mapping layer)
  // styles mode: x' = avg*(1-psi[l]) + x*psi[l] with l = layer of unit u
  const float* latent_avg; const float* psi; const int* unit_layer;
};
void launch_dense(const DenseArgs& a, cudaStream_t st);
struct MapArgs {              // the whole mapping path (styles.cu: mapping_kernel); latent size 512
  const float* z;             // [N][512] latents
  const float* W[8]; const float* b[8];      // mapping layers, pre-scaled
  float *ya, *yb;             // [N][512] ping-pong buffers (layer 7 leaves the disentangled latents in yb)
  const float* Waff; const float* baff;      // [S][512], [S]: the AdaIN affines of all style layers
  const float* latent_avg; const float* psi; const int* unit_layer;
  float* styles;              // [N][S]
  int N, S;
};
bool launch_mapping(const MapArgs& a, cudaStream_t st);

// Launch with programmatic stream serialization (see ptx.cuh) when the current forward pass is small enough for it to
// pay (plan.cpp); GSX_NO_PDL=1: never, GSX_PDL=1: always.
// kind: 1 = shiftconv (one persistent CTA per SM holding most of the shared memory), 0 = everything else.
extern int g_pdl_mode;
void set_pdl_late_conv(int v);
void set_pdl_late_ew(int v);
bool pdl_enabled(int kind);
void pdl_set_for_work(double top_level_pixels);   // called at the top of the forward passes
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_kind(int kind, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                   Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled(kind) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  return launch_pdl_kind(0, kernel, grid, block, smem, st, args...);
}

// Scratch memory of the single-operator hooks (gsx_op_*): blocks are kept and reused across calls (every hook
// synchronises its stream before returning, so a block is free again when the call ends); cudaMalloc / cudaFree per
// call dominated the hook-based training step.  Thread-local; gsx_op_release_cache() returns everything.
void* pool_get(size_t bytes);
void pool_put(void* p);
void pool_release();

extern int g_wgrad_m64;
extern int g_wgrad_kxm;
// wgrad.cu: tensor-core weight gradient (see there).  scratch: wgrad_scratch_floats(...) floats.
size_t wgrad_scratch_floats(int K, int Cin, int Cout, int sms);
bool launch_wgrad(int K, int N, int H, int W, int Cin, int Cout, int cout_real, const act_t* x, const act_t* dy, float* dw,
                  int cin_off, int cin_total, float scale, float* scratch, cudaStream_t st);

void set_error(const std::string& msg);
bool cuda_ok(cudaError_t e, const char* what);

}  // namespace gsx
