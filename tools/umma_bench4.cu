// Microbenchmark 4: which operand change between consecutive tcgen05.mma (kind::f16, M=128, K=16, N given) costs what.
// A 16-entry schedule {accumulator, A offset, B tile} is held in registers and replayed, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/umma_bench4 tools/umma_bench4.cu -I gan-segmentation_b200/csrc
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace gsx;

struct Sched { int d[16], a[16], b[16]; };     // accumulator index, A offset in 16-B units, B tile index

__global__ void __launch_bounds__(128, 1) bench(int N, int iters, int lbo, Sched sc, long long* out, int rnd, int commit_every = 0) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = rnd ? (0x38003800u + ((i * 2654435761u) >> 7 & 0x03ff03ffu)) : 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, commit_every < 0 ? 1 : 1000000); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  fence_proxy_async();
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32) {
    uint32_t idesc = umma_idesc_16bit(128, (uint32_t)N, 0);
    uint32_t a_hi = (uint32_t)(umma_desc_hi((uint32_t)lbo, 128) >> 32);
    uint32_t b_hi = (uint32_t)(umma_desc_hi((uint32_t)N * 16, 128) >> 32);
    const uint32_t a_lbo = (((uint32_t)lbo >> 4) & 0x3FFF) << 16;
    const uint32_t b_lbo = ((((uint32_t)N * 16) >> 4) & 0x3FFF) << 16;
    const uint32_t a_lo0 = ((smem_u32(smem) >> 4) & 0x3FFF) | a_lbo;
    const uint32_t b_lo0 = (((smem_u32(smem) + 120 * 1024) >> 4) & 0x3FFF) | b_lbo;
    uint32_t dd[16], aa[16], bb[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      dd[i] = tmem + (uint32_t)(sc.d[i] * N);
      aa[i] = a_lo0 + (uint32_t)sc.a[i];
      bb[i] = b_lo0 + (uint32_t)sc.b[i] * (((uint32_t)N * 32) >> 4);
      keep_in_reg(dd[i]); keep_in_reg(aa[i]); keep_in_reg(bb[i]);
    }
    keep_in_reg(idesc); keep_in_reg(a_hi); keep_in_reg(b_hi);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) umma_f16kind_lohi(dd[k], aa[k], a_hi, bb[k], b_hi, idesc, 1u);
        if (commit_every > 0 && (i % commit_every) == commit_every - 1) umma_commit(&bar2);
        if (commit_every < 0 && (i % (-commit_every)) == -commit_every - 1) { umma_commit(&bar2); umma_commit(&bar2); }
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
  const int iters = 400;
  for (int N : {16, 64, 128}) {
    const int nacc = 512 / N < 4 ? 512 / N : 4;
    struct V { const char* name; int dper, aper, astep, bper; } vs[] = {
      // period = number of consecutive MMAs sharing the operand (16 = never changes)
      {"all fixed                         ", 16, 16, 0, 16},
      {"D changes every MMA               ", 1, 16, 0, 16},
      {"A +2KB every MMA                  ", 16, 1, 128, 16},
      {"A +1px every MMA                  ", 16, 1, 1, 16},
      {"A +67px every MMA                 ", 16, 1, 67, 16},
      {"B changes every MMA               ", 16, 16, 0, 1},
      {"new order: D fixed, A+1px, B every", 16, 1, 1, 1},
      {"new order: D fixed, A+67px, B evry", 16, 1, 67, 1},
      {"old order: D,A+2KB every, B per 4 ", 1, 1, 128, 4},
      {"D per 4, A+67px, B every          ", 4, 1, 67, 1},
      {"D per 2, A+67px, B every          ", 2, 1, 67, 1},
    };
    for (auto& v : vs) {
      Sched sc;
      for (int i = 0; i < 16; ++i) {
        sc.d[i] = (i / v.dper) % nacc;
        sc.a[i] = ((i / v.aper) % 9) * v.astep;
        sc.b[i] = (i / v.bper) % 9;
      }
      bench<<<148, 128, 200 * 1024>>>(N, iters, 9504, sc, d, 0);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" ERR(%s)\n", cudaGetErrorString(e)); return 1; }
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("N=%3d %s: %.1f cycles/MMA\n", N, v.name, (double)cyc / ((double)iters * 16));
    }
  }
  // sustained run with non-trivial operand data: cycles (clock64) against wall time (events) = SM clock under MMA load
  for (int rnd = 0; rnd < 2; ++rnd)
    for (int N : {16, 64, 128, 256}) {
      Sched sc;
      const int nacc = 512 / N < 4 ? 512 / N : 4;
      for (int i = 0; i < 16; ++i) { sc.d[i] = (i / 4) % nacc; sc.a[i] = (i % 9) * 67; sc.b[i] = i % 9; }
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int it2 = 40000;
      bench<<<148, 128, 200 * 1024>>>(N, 2000, 9504, sc, d, rnd);
      cudaEventRecord(e0);
      bench<<<148, 128, 200 * 1024>>>(N, it2, 9504, sc, d, rnd);
      cudaEventRecord(e1);
      cudaDeviceSynchronize();
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      printf("sustained N=%3d data=%s: %.1f cycles/MMA, %.2f ms, SM clock %.0f MHz, %.1f ns/MMA\n", N, rnd ? "random" : "ones",
             (double)cyc / (it2 * 16.0), ms, cyc / (ms * 1e3), ms * 1e6 / (it2 * 16.0));
    }
  for (int ce : {0, 1, 2, 4, 8, -1, -2, -9}) {
    Sched sc;
    for (int i = 0; i < 16; ++i) { sc.d[i] = (i / 4) % 4; sc.a[i] = (i % 9) * 67; sc.b[i] = i % 9; }
    bench<<<148, 128, 200 * 1024>>>(64, 4000, 9504, sc, d, 1, ce);
    cudaDeviceSynchronize();
    long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    printf("N=64 commit every %2d x16 MMAs (negative: two commits completing a count-1 barrier): %.1f cycles/MMA\n", ce, (double)cyc / (4000 * 16.0));
  }
  return 0;
}
