// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16, SS operands, no-swizzle K-major) as a function
// of N, of the A-operand start misalignment and of the A row-group stride (SBO).  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench tools/umma_bench.cu -I gan-segmentation_b200/csrc
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace gsx;

__global__ void __launch_bounds__(128, 1) bench(int N, int a_off_bytes, int lbo, int sbo, int iters, int taps, int tap_pitch,
                                                long long* out, int dcol_stride, int n_acc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  fence_proxy_async();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_16bit(128, (uint32_t)N, 0);
    const uint64_t a_hi = umma_desc_hi((uint32_t)lbo, (uint32_t)sbo);
    const uint64_t b_hi = umma_desc_hi((uint32_t)N * 16, 128);
    const uint32_t a_addr = smem_u32(smem) + (uint32_t)a_off_bytes;
    const uint32_t b_addr = smem_u32(smem) + 160 * 1024;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int t = 0; t < taps; ++t)
        umma_f16kind(tmem + (uint32_t)((i % n_acc) * dcol_stride), umma_desc(a_hi, a_addr + (uint32_t)(t * tap_pitch)), umma_desc(b_hi, b_addr), idesc,
                     t > 0);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 400, taps = 9;
  printf("cycles per MMA (M=128,K=16): rows = N, cols = config\n");
  struct Cfg { const char* name; int off, lbo, sbo, pitch; } cfgs[] = {
    {"aligned,pitch0", 0, 32768, 128, 0},
    {"aligned,pitch2048", 0, 32768, 128, 2048},
    {"off16,pitch0", 16, 32768, 128, 0},
    {"off32,pitch0", 32, 32768, 128, 0},
    {"off64,pitch0", 64, 32768, 128, 0},
    {"conv-like(pitch1056+16)", 0, 32768, 128, 1072},
    {"lbo=16(adjacent px)", 0, 16, 128, 0},
    {"sbo=256", 0, 32768, 256, 0},
  };
  for (int N : {16, 32, 64, 128, 256}) {
    printf("N=%3d:", N);
    for (auto& c : cfgs) {
      bench<<<148, 128, 200 * 1024>>>(N, c.off, c.lbo, c.sbo, iters, taps, c.pitch, d, N, 2);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" ERR(%s)", cudaGetErrorString(e)); return 1; }
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("  %s=%.1f", c.name, (double)cyc / (iters * taps));
    }
    printf("\n");
  }
  printf("accumulator-address effects (9 taps accumulate into one tile, tiles round-robin):\n");
  struct A { int N, dstride, nacc, off; } accs[] = {{16, 16, 16, 0}, {48, 48, 10, 0}, {48, 64, 8, 0}, {48, 48, 10, 2016}, {64, 64, 8, 0}, {96, 96, 5, 0}, {96, 128, 4, 0}, {32, 32, 16, 0}, {16, 16, 1, 0}};
  for (auto& a : accs) {
    bench<<<148, 128, 200 * 1024>>>(a.N, a.off, 32768, 128, iters, 3, 1056, d, a.dstride, a.nacc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" ERR(%s)\n", cudaGetErrorString(e)); return 1; }
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("  N=%d dcol_stride=%d n_acc=%d a_off=%d: %.1f cycles/MMA\n", a.N, a.dstride, a.nacc, a.off, (double)cyc / (iters * 3));
  }
  return 0;
}
