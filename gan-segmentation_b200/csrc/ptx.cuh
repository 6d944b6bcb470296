// Thin inline-PTX wrappers for the sm_100a features the kernels use:
// mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (TMEM alloc / mma / commit / ld).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gsx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (error returned to the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}
// Same, sleeping between polls.  Every poll is a shared-memory access, and tcgen05.mma with SS operands is bound by
// shared-memory read bandwidth (M=128,K=16: (4 KB of A + N*32 B of B) / 128 B per cycle = 39 / 48 / 64 cycles at
// N = 16 / 64 / 128, tools/umma_bench4.cu), so spinning warps slow the tensor pipe down.  Used by the producer (runs
// stages ahead) and by the epilogue warps (a tile is thousands of cycles; the wake-up latency is noise).
template <int NS = 100>
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(NS);
    if (++spins > (1u << 22)) __trap();
  }
}

// ---------------------------------------------------------------- programmatic dependent launch
// Every kernel of the step is launched with cudaLaunchAttributeProgrammaticStreamSerialization (gsx_internal.h,
// launch_pdl): its CTAs may start -- barrier init, TMEM allocation, resident-weight loads, coefficient staging --
// while the previous kernel of the stream is still draining, and must execute pdl_wait() before the first access to
// memory an earlier kernel wrote (or still reads).  pdl_launch_dependents() at the top lets the next kernel do the same.
// Both are no-ops for launches without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 1-D bulk copy global -> shared (size multiple of 16 B, both addresses 16 B aligned).
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; 16-bit x 16-bit -> fp32 (kind::f16), issued by ONE thread.
__device__ __forceinline__ void umma_f16kind(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the descriptors given as (lo, hi) 32-bit halves: only the low word (start address) changes
// between the MMAs of a k-chunk, so the issue loop does 32-bit adds only.
__device__ __forceinline__ void umma_f16kind_lohi(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                                  uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier when all previously issued MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: lane i of the warp reads TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (canonical layout ((8,n),2):((16B,SBO),LBO)):
// rows of a core matrix are 16 B apart, 8-row groups SBO bytes apart, the two 8-element K halves
// LBO bytes apart.  Only 16-B start alignment is required, which is what lets one staged input
// tile serve every filter tap through a shifted start address.
__device__ __forceinline__ uint64_t umma_desc_hi(uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;  // descriptor version (Blackwell)
  return d;                              // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}
__device__ __forceinline__ uint64_t umma_desc(uint64_t hi, uint32_t smem_addr) {
  return hi | static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
}
// Instruction descriptor: D=F32, A=B=F16 (fmt 0) or BF16 (fmt 1), both K-major, N runtime.
__host__ __device__ inline uint32_t umma_idesc_16bit(uint32_t M, uint32_t N, uint32_t fmt) {
  uint32_t d = 0;
  d |= 1u << 4;           // c_format F32
  d |= fmt << 7;          // a_format
  d |= fmt << 10;         // b_format
  d |= (N >> 3) << 17;    // n_dim
  d |= (M >> 4) << 24;    // m_dim
  return d;
}

// 32 lanes x 32 consecutive fp32 columns (one wait per 32 columns instead of 16)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2)
// The conv epilogues are bound by the number of issued instructions (ncu r01: 24 thread-instructions per output value,
// 46 % issue-slot utilisation with 10 warps per SM); the two-wide fp32 forms halve the arithmetic part.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 pk2u(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// 32 values per lane x 32 lanes -> lane L returns value L summed over the warp: a halving butterfly, 31 shuffles in
// all (a plain xor-reduction of every value would take 160); fixed order, so results are bit-reproducible.
__device__ __forceinline__ float warp_transpose_reduce32(float (&vals)[32], int lane) {
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int half = 16 >> step;                      // values kept per lane after this step
    const bool upper = (lane & half) != 0;            // lane bit deciding which half is kept
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = upper ? vals[half + i] : vals[i];
      const float send = upper ? vals[i] : vals[half + i];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return vals[0];
}

// 16 values per lane x 32 lanes -> every lane L returns value (L & 15) summed over the warp (15 + 1 shuffles, fixed order)
__device__ __forceinline__ float warp_transpose_reduce16(float (&vals)[16], int lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int half = 8 >> step;
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = upper ? vals[half + i] : vals[i];
      const float send = upper ? vals[i] : vals[half + i];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return vals[0] + __shfl_xor_sync(0xffffffffu, vals[0], 16);
}

// ---------------------------------------------------------------- misc
// One lane of a converged warp; lets the compiler keep the tcgen05 / TMA issue sequence in uniform
// registers without wrapping every instruction in a per-thread loop (which `lane == 0` does).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// Keeps a loop-invariant value in a register (stops the compiler from re-loading it from the
// constant bank inside the single-thread MMA issue loop, where every load is on the critical path).
__device__ __forceinline__ void keep_in_reg(uint32_t& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void keep_in_reg(int& v) { asm volatile("" : "+r"(v)); }

// Loads of data that an EARLIER KERNEL OF THE STREAM produced (activations, statistics, AdaIN coefficients, per-sample
// weights and biases, noise).  Under programmatic dependent launch a kernel's lifetime overlaps its producer's -- and a
// chain of such launches has no ordinary kernel boundary in it -- so the non-coherent path (__ldg / ld.global.nc, also
// what the compiler picks for const __restrict__ pointers) must not be used for this data: it is outside the memory
// model, griddepcontrol.wait does not order it, and it returned lines the SM had cached before the producer's write
// (seen as the previous layer's statistics -> NaN features; tests/test_dropin_gpu.py::test_forward_is_independent_...
// and ::test_dependent_launch_does_not_change_results pin this).  These are ordinary weak loads (ld.global.ca), which
// pdl_wait() orders; as compiler intrinsics they are scheduled freely but never across pdl_wait()'s memory clobber.
// Measured at FFHQ batch 32: 10.21 ms (nc, wrong) -> 10.29 ms (weak); ld.global.cg (LDG.STRONG.GPU) instead: 10.50 ms.
// __ldg stays for what no kernel writes (weights, biases, tap tables).
__device__ __forceinline__ uint4 ld_dep_u4(const void* p) { return __ldca(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ float4 ld_dep_f4(const void* p) { return __ldca(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float2 ld_dep_f2(const void* p) { return __ldca(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float ld_dep_f32(const void* p) { return __ldca(reinterpret_cast<const float*>(p)); }
// 256-bit global store (sm_100: STG.E.ENL2.256); address 32-byte aligned
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace gsx
