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

__global__ void __launch_bounds__(256) modulate_w_kernel(const float* __restrict__ wf, const float* __restrict__ coef,
                                                         act_t* __restrict__ wout, const ModGeom g) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.y;
  const float* cf = coef + (size_t)n * g.Cin * 2;
  act_t* out = wout + (size_t)n * g.elems;
  for (int i = (blockIdx.x * 256 + threadIdx.x) * 8; i < g.elems; i += gridDim.x * 256 * 8) {
    // 8 consecutive elements = the 8 channels (k & 7) of one (row, k-half): one 16-byte store
    const int c0 = mod_channel(g, i);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(wf + i)), w1 = __ldg(reinterpret_cast<const float4*>(wf + i + 4));
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    __align__(16) act_t o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = to_act(wv[k] * __ldg(cf + (size_t)(c0 + k) * 2));
    *reinterpret_cast<uint4*>(out + i) = *reinterpret_cast<const uint4*>(o);
  }
}

// grid (N): S[slot][col] = sum_ci Wf[slot][ci][col] * b[n][ci] in shared memory, then the interior bias and the 8 border
// corrections per column.
__global__ void __launch_bounds__(256) modulate_b_kernel(const float* __restrict__ wf, const float* __restrict__ coef,
                                                         const float* __restrict__ bias, float* __restrict__ bias_n,
                                                         float* __restrict__ bdelta, const ModGeom g) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float S[];                 // [n_slots][bias_cols], then b[Cin]
  float* bsh = S + g.n_slots * g.bias_cols;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < g.Cin; c += 256) bsh[c] = coef[((size_t)n * g.Cin + c) * 2 + 1];
  __syncthreads();
  const int tile = g.N_tile * 16;
  for (int e = threadIdx.x; e < g.n_slots * g.bias_cols; e += 256) {
    const int slot = e / g.bias_cols, col = e - slot * g.bias_cols;
    float acc = 0.f;
    for (int kc = 0; kc < g.n_k; ++kc)
      for (int j = 0; j < g.k16pc; ++j) {
        const float* wp = wf + (size_t)((kc * g.n_slots + slot) * g.k16pc + j) * tile + (size_t)col * 8;
        const int cb = (kc * g.CBK + 2 * j) * 8;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc = fmaf(__ldg(wp + (k >> 3) * (g.N_tile * 8) + (k & 7)), bsh[cb + k], acc);
      }
    S[e] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 9 * g.bias_cols; e += 256) {
    const int cls = e / g.bias_cols, col = e - cls * g.bias_cols;
    const int ry = cls / 3, rx = cls - 3 * ry;       // 0 first row / column, 1 interior, 2 last
    float in = 0.f, out = 0.f;
    for (int s = 0; s < g.n_slots; ++s) {
      const bool outside = (ry == 0 && g.dy[s] < 0) || (ry == 2 && g.dy[s] > 0) || (rx == 0 && g.dx[s] < 0) || (rx == 2 && g.dx[s] > 0);
      if (outside) out += S[s * g.bias_cols + col];
      in += S[s * g.bias_cols + col];
    }
    const int ch = col % g.cout_tile;
    if (cls == 4) bias_n[(size_t)n * g.bias_cols + col] = in + ((bias && ch < g.cout) ? bias[ch] : 0.f);
    bdelta[((size_t)n * 9 + cls) * g.bias_cols + col] = -out;
  }
}

void launch_modulate(const ConvLayer& L, const float* coef, const float* bias, int N, act_t* wout, float* bias_n, float* bdelta,
                     cudaStream_t st) {
  ModGeom g{};
  g.N = N; g.Cin = L.cin0 + L.cin1; g.elems = (int)L.wpack_elems; g.N_tile = L.g.N_tile; g.n_slots = L.g.n_slots;
  g.k16pc = L.g.CBK / 2; g.n_k = L.g.n_k; g.CBK = L.g.CBK; g.bias_cols = L.g.bias_cols; g.cout_tile = L.g.cout_tile; g.cout = L.cout;
  slot_offsets(L, g.dy, g.dx);
  const int blocks = std::max(1, std::min(64, (g.elems / 8 + 255) / 256));
  launch_pdl(modulate_w_kernel, dim3(blocks, N), dim3(256), 0, st, L.wf32_dev, coef, wout, g);
  const size_t smem = (size_t)(g.n_slots * g.bias_cols + g.Cin) * sizeof(float);
  launch_pdl(modulate_b_kernel, dim3(N), dim3(256), smem, st, L.wf32_dev, coef, bias, bias_n, bdelta, g);
}

}  // namespace gsx
