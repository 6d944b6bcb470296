// Microbenchmark 2: tensor-pipe cost of one tcgen05.mma (kind::f16, M=128, K=16, SS, no-swizzle K-major) versus N,
// with the issue pattern of the production kernel (one elected lane, 32-bit descriptor halves in registers, 8-fold
// unrolled), so that the pipe and not the issuing thread is what is timed.  Variants: conv-like shifted A starts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_bench2 tools/umma_bench2.cu -I gan-segmentation_b200/csrc
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace gsx;

__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int a_step16, int n_acc, int dstride, long long* out) {
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
  if (threadIdx.x < 32) {
    uint32_t idesc = umma_idesc_16bit(128, (uint32_t)N, 0);
    uint32_t a_hi = (uint32_t)(umma_desc_hi(32768, 128) >> 32);
    uint32_t b_hi = (uint32_t)(umma_desc_hi((uint32_t)N * 16, 128) >> 32);
    const uint32_t a_lbo = ((32768u >> 4) & 0x3FFF) << 16;
    const uint32_t b_lbo = ((((uint32_t)N * 16) >> 4) & 0x3FFF) << 16;
    uint32_t a_lo0 = ((smem_u32(smem) >> 4) & 0x3FFF) | a_lbo;
    uint32_t b_lo = (((smem_u32(smem) + 160 * 1024) >> 4) & 0x3FFF) | b_lbo;
    keep_in_reg(idesc); keep_in_reg(a_hi); keep_in_reg(b_hi); keep_in_reg(a_lo0); keep_in_reg(b_lo);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tmem + (uint32_t)((i % n_acc) * dstride);
#pragma unroll
        for (int t = 0; t < 8; ++t) umma_f16kind_lohi(d, a_lo0 + (uint32_t)(t * a_step16), a_hi, b_lo, b_hi, idesc, 1u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (blockIdx.x == 0 && t0 != 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 500;
  printf("cycles per MMA (M=128,K=16, elected-lane issue, unroll 8)\n");
  for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
    printf("N=%3d:", N);
    struct V { const char* name; int step16, nacc; } vs[] = {{"sameA,1acc", 0, 1}, {"sameA,2acc", 0, 2}, {"shift+1px", 1, 2}, {"shift+67px", 67, 2}, {"shift 128B", 8, 2}};
    for (auto& v : vs) {
      const int nacc = (v.nacc * N <= 512) ? v.nacc : 1;
      bench<<<148, 128, 200 * 1024>>>(N, iters, v.step16, nacc, N, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" ERR(%s)", cudaGetErrorString(e)); return 1; }
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("  %s=%.1f", v.name, (double)cyc / (iters * 8));
    }
    printf("\n");
  }
  return 0;
}
