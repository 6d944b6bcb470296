// Microbenchmark 3: tcgen05.mma (kind::f16, M=128, K=16) cost under production-like operand patterns:
// several accumulator tiles, a different B tile per MMA, non-power-of-two LBO, one or two issuing warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_bench3 tools/umma_bench3.cu -I gan-segmentation_b200/csrc
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace gsx;

__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int lbo, int n_acc, int n_b, int a_step16, int issuers,
                                                long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, (uint32_t)issuers); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  fence_proxy_async();
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp < issuers) {
    uint32_t idesc = umma_idesc_16bit(128, (uint32_t)N, 0);
    uint32_t a_hi = (uint32_t)(umma_desc_hi((uint32_t)lbo, 128) >> 32);
    uint32_t b_hi = (uint32_t)(umma_desc_hi((uint32_t)N * 16, 128) >> 32);
    const uint32_t a_lbo = (((uint32_t)lbo >> 4) & 0x3FFF) << 16;
    const uint32_t b_lbo = ((((uint32_t)N * 16) >> 4) & 0x3FFF) << 16;
    uint32_t a_lo0 = ((smem_u32(smem) >> 4) & 0x3FFF) | a_lbo;
    uint32_t b_lo0 = (((smem_u32(smem) + 120 * 1024) >> 4) & 0x3FFF) | b_lbo;
    uint32_t b_step = ((uint32_t)N * 32) >> 4;
    keep_in_reg(idesc); keep_in_reg(a_hi); keep_in_reg(b_hi); keep_in_reg(a_lo0); keep_in_reg(b_lo0); keep_in_reg(b_step);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      const int per = n_acc / issuers;                    // accumulator tiles per issuer
      for (int i = 0; i < iters; ++i) {
        const uint32_t b_lo = b_lo0 + (uint32_t)(i % n_b) * b_step;
        const uint32_t a_lo = a_lo0 + (uint32_t)((i % 9) * a_step16);
        for (int m = 0; m < per; ++m) {
          const int mt = warp * per + m;
          umma_f16kind_lohi(tmem + (uint32_t)(mt * N), a_lo + (uint32_t)(mt * 128), a_hi, b_lo, b_hi, idesc, 1u);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    t1 = clock64();
    if (blockIdx.x == 0 && t0 != 0 && warp == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  struct V { const char* name; int N, lbo, nacc, nb, astep, iss; } vs[] = {
    {"N64 base            ", 64, 32768, 4, 1, 0, 1},
    {"N64 lbo9504         ", 64, 9504, 4, 1, 0, 1},
    {"N64 lbo9504 9B      ", 64, 9504, 4, 9, 0, 1},
    {"N64 lbo9504 9B shift", 64, 9504, 4, 9, 1, 1},
    {"N64 ... 2 issuers   ", 64, 9504, 4, 9, 1, 2},
    {"N64 lbo34848 shift  ", 64, 34848, 4, 9, 67, 1},
    {"N16 16acc lbo34848  ", 16, 34848, 16, 9, 67, 1},
    {"N16 16acc 2 issuers ", 16, 34848, 16, 9, 67, 2},
    {"N128 4acc lbo9504   ", 128, 9504, 4, 9, 1, 1},
    {"N128 4acc 2 issuers ", 128, 9504, 4, 9, 1, 2},
    {"N128 2acc 2 issuers ", 128, 9504, 2, 9, 1, 2},
    {"N128 lbo 8192       ", 128, 8192, 4, 9, 1, 2},
    {"N128 lbo 9472(128al)", 128, 9472, 4, 9, 1, 2},
    {"N128 lbo 9600       ", 128, 9600, 4, 9, 1, 2},
  };
  for (auto& v : vs) {
    bench<<<148, 128, 200 * 1024>>>(v.N, iters, v.lbo, v.nacc, v.nb, v.astep, v.iss, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf(" ERR(%s)\n", cudaGetErrorString(e)); return 1; }
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("%s: %.1f cycles/MMA\n", v.name, (double)cyc / ((double)iters * v.nacc));
  }
  return 0;
}
