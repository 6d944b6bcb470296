// AdaIN folded into the consumer convolution (reference networks_stylegan.py:250-264 followed by :24 / :16 / the
// decoder's cvt conv, networks_seg.py:68).
//
// The instance-norm + style modulation x' = a (.) t + b is affine per (sample, channel), so a conv that consumes it is
//     conv_W(x') = conv_{W diag(a_n)}(t)  +  sum over the taps that land inside the image of  W_tap b_n .
// For the channel-thin layers at the top resolutions the separate "apply" pass over the tensor (one read + one write at
// ~100 % of HBM bandwidth: perfect kernels doing avoidable work, 1.3 ms of the 11.2 ms FFHQ step in round 1) disappears:
// this file builds, per sample, the modulated 16-bit weight stream the conv kernel reads (a few KB to 150 KB) and the
// bias it adds -- one value per accumulator column for interior rows plus 8 corrections for rows on the image border
// (the shift b must not flow in through taps that fall on the zero padding).  The un-normalised tensor t is what stays
// in HBM; it is read once by each consumer and never rewritten.
#include "gsx_internal.h"
#include "ptx.cuh"

namespace gsx {

struct ModGeom {
  int N, Cin, elems, N_tile, n_slots, k16pc, n_k, CBK, bias_cols, cout_tile, cout;
  int dy[kMaxSlots], dx[kMaxSlots];
};

// packed element i -> input channel: the stream is [k-chunk][slot][k16 step][ (k>>3) x N_tile x 8 ]  (plan.cpp)
__device__ __forceinline__ int mod_channel(const ModGeom& g, int i) {
  const int tile = g.N_tile * 16;
  const int inner = i % tile, blk = i / tile;
  const int j = blk % g.k16pc;
  const int kc = blk / (g.k16pc * g.n_slots);
  const int k = (inner / (g.N_tile * 8)) * 8 + (inner & 7);
  return (kc * g.CBK + 2 * j) * 8 + k;
}

// grid (10, N): blocks 0..8 of a sample = one border class each: bias / correction per accumulator column as the sum, over
// the slots that class keeps (bias_n, class 4) or loses (bdelta), of  sum_ci Wf[slot][ci][col] * b[n][ci]  -- every block
// redoes the slot sums it needs (a few 10^4 MACs) instead of sharing them through a second pass; block 9 (and up) = the
// modulated weight stream  wout[n][i] = 16-bit( Wf[i] * a[n][channel(i)] ).
__global__ void __launch_bounds__(256) modulate_kernel(const float* __restrict__ wf, const float* coef,
                                                       const float* __restrict__ bias, act_t* __restrict__ wout,
                                                       float* __restrict__ bias_n, float* __restrict__ bdelta, const ModGeom g) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y;
  const float* cf = coef + (size_t)n * g.Cin * 2;
  if (blockIdx.x >= 9) {
    act_t* out = wout + (size_t)n * g.elems;
    for (int i = ((blockIdx.x - 9) * 256 + threadIdx.x) * 8; i < g.elems; i += (gridDim.x - 9) * 256 * 8) {
      // 8 consecutive elements = the 8 channels (k & 7) of one (row, k-half): one 16-byte store
      const int c0 = mod_channel(g, i);
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wf + i)), w1 = __ldg(reinterpret_cast<const float4*>(wf + i + 4));
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      __align__(16) act_t o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = to_act(wv[k] * ld_dep_f32(cf + (size_t)(c0 + k) * 2));       // coef: written by the finalize kernel before this one
      *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(o);
    }
    return;
  }
  extern __shared__ float bsh[];               // b[Cin], then partial sums [parts][2][bias_cols]
  float* part_s = bsh + g.Cin;
  for (int c = threadIdx.x; c < g.Cin; c += 256) bsh[c] = ld_dep_f32(cf + (size_t)c * 2 + 1);
  __syncthreads();
  const int cls = blockIdx.x, ry = cls / 3, rx = cls - 3 * ry;       // 0 first row / column, 1 interior, 2 last
  const int tile = g.N_tile * 16;
  // a column's sum runs over slots x k16 steps x 16 channels: split the (slot, k16 step) pairs over 256 / bias_cols
  // threads per column (a single thread per column is a ~600-deep chain of dependent global loads)
  const int parts = max(1, 256 / g.bias_cols), steps = g.n_slots * g.n_k * g.k16pc;
  for (int col0 = 0; col0 < g.bias_cols; col0 += 256) {
    const int col = col0 + threadIdx.x % min(256, g.bias_cols), part = threadIdx.x / min(256, g.bias_cols);
    float in = 0.f, out = 0.f;
    if (col < g.bias_cols && part < parts) {
      for (int it = part; it < steps; it += parts) {
        const int j = it % g.k16pc, kc = (it / g.k16pc) % g.n_k, sl = it / (g.k16pc * g.n_k);
        const bool outside = (ry == 0 && g.dy[sl] < 0) || (ry == 2 && g.dy[sl] > 0) || (rx == 0 && g.dx[sl] < 0) || (rx == 2 && g.dx[sl] > 0);
        if (cls != 4 && !outside) continue;          // border classes only need the slots they lose
        const float* wp = wf + (size_t)((kc * g.n_slots + sl) * g.k16pc + j) * tile + (size_t)col * 8;
        const int cb = (kc * g.CBK + 2 * j) * 8;
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(wp)), a1 = __ldg(reinterpret_cast<const float4*>(wp + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(wp + g.N_tile * 8)), b1 = __ldg(reinterpret_cast<const float4*>(wp + g.N_tile * 8 + 4));
        const float wv[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fmaf(wv[k], bsh[cb + k], acc);
        if (outside) out += acc;
        in += acc;
      }
      part_s[(part * 2) * g.bias_cols + col] = in;
      part_s[(part * 2 + 1) * g.bias_cols + col] = out;
    }
    __syncthreads();
    if (col < g.bias_cols && part == 0) {
      in = 0.f; out = 0.f;
      for (int q = 0; q < parts; ++q) { in += part_s[(q * 2) * g.bias_cols + col]; out += part_s[(q * 2 + 1) * g.bias_cols + col]; }   // fixed order
      const int ch = col % g.cout_tile;
      if (cls == 4) bias_n[(size_t)n * g.bias_cols + col] = in + ((bias && ch < g.cout) ? bias[ch] : 0.f);
      bdelta[((size_t)n * 9 + cls) * g.bias_cols + col] = cls == 4 ? 0.f : -out;
    }
    __syncthreads();
  }
}

void launch_modulate(const ConvLayer& L, const float* coef, const float* bias, int N, act_t* wout, float* bias_n, float* bdelta,
                     cudaStream_t st) {
  ModGeom g{};
  g.N = N; g.Cin = L.cin0 + L.cin1; g.elems = (int)L.wpack_elems; g.N_tile = L.g.N_tile; g.n_slots = L.g.n_slots;
  g.k16pc = L.g.CBK / 2; g.n_k = L.g.n_k; g.CBK = L.g.CBK; g.bias_cols = L.g.bias_cols; g.cout_tile = L.g.cout_tile; g.cout = L.cout;
  slot_offsets(L, g.dy, g.dx);
  const int blocks = std::max(1, std::min(16, (g.elems / 8 + 255) / 256));
  const size_t smem = (size_t)(g.Cin + 2 * 256 + 2 * g.bias_cols) * sizeof(float);
  launch_pdl(modulate_kernel, dim3(9 + blocks, N), dim3(256), smem, st, L.wf32_dev, coef, bias, wout, bias_n, bdelta, g);
}

}  // namespace gsx
