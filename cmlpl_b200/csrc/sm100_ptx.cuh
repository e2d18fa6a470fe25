// sm_100a PTX wrappers shared by the tcgen05 kernels (patch_cnn_sm100.cu, head_sm100.cu):
// mbarrier, proxy fences, bulk async copies, TMEM alloc/ld, tcgen05.mma/commit, UMMA descriptors.
#pragma once
#include "common.cuh"

namespace cmlpl {

// ------------------------------------------------------------------ timeline tracing (debugging aid)
// `make trace` builds a second library (scripts/_trace/, never loaded by the package) with -DCMLPL_TRACE: CMLPL_TR stamps
// clock64 + a tag into a per-translation-unit device array for ONE CTA, CMLPL_TRACE_EXPORT(name) exports the function
// that copies it out (scripts/trace_kernel.py).  In the product build both expand to nothing.
#ifdef CMLPL_TRACE
static __device__ unsigned long long g_trace[6][4096];
__device__ __forceinline__ void trace_stamp(int role, uint32_t& n, uint32_t tag) {
  if (blockIdx.x == 5 && n < 2047) { g_trace[role][2 * n] = clock64(); g_trace[role][2 * n + 1] = tag; ++n; }
}
#define CMLPL_TR(role, n, tag) ::cmlpl::trace_stamp(role, n, tag)
#define CMLPL_TRACE_EXPORT(name)                                                                          \
  extern "C" int name(unsigned long long* host) {                                                         \
    return cudaMemcpyFromSymbol(host, ::cmlpl::g_trace, sizeof(::cmlpl::g_trace)) == cudaSuccess ? 0 : 1;  \
  }
#else
#define CMLPL_TR(role, n, tag) ((void)0)
#define CMLPL_TRACE_EXPORT(name)
#endif

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap (kernel error), never hang the GPU.  The bound is wall clock (10 s of
// %globaltimer, sampled every 4096 polls), not an iteration count: time slicing, a profiler replay, clock throttling
// or a TMEM allocation waiting for a co-resident kernel can legitimately stretch a wait by orders of magnitude.
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
#pragma unroll 1
  for (int it = 0; it < 4096; ++it)
    if (mbar_try_wait(bar, parity)) return;
  const uint64_t t0 = global_timer_ns();
#pragma unroll 1
  for (;;) {
#pragma unroll 1
    for (int it = 0; it < 4096; ++it)
      if (mbar_try_wait(bar, parity)) return;
    if (global_timer_ns() - t0 > 10000000000ull) break;
  }
  printf("cmlpl: mbarrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
  __trap();
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared (UBLKCP); completes `bytes` transaction bytes on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// weights -> shared memory as a handful of bulk copies in flight together, all completing on one mbarrier (the MMA
// issuer waits on it before its first tcgen05.mma); called by ONE thread right after it initialised the barrier.
// Replaces an LDG -> STS loop whose dependent round trips cost ~15 us at the head of every persistent kernel.
__device__ __forceinline__ void bulk_weights_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  mbar_arrive_expect_tx(bar, bytes);
  const unsigned char* g = static_cast<const unsigned char*>(src);
  for (uint32_t o = 0; o < bytes; o += 8192) {
    const uint32_t n = bytes - o < 8192 ? bytes - o : 8192;
    bulk_g2s(dst_smem + o, g + o, n, bar);
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
// 16-byte copy that writes zeros when src_bytes == 0 (out-of-map halo entries)
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, M=128 N=64 K=16, fp16 in / fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand in TENSOR MEMORY (M = 128 lanes = rows, K = 16 halves = 8 columns, element 2c in the low and
// 2c+1 in the high half of column c): the operand an epilogue has just produced never goes through shared memory
__device__ __forceinline__ void umma_f16_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// tcgen05.ld of 8 columns as four packed fp32 pairs (register pairs are adjacent, the packing is free)
__device__ __forceinline__ void tmem_ld8_x2(uint32_t taddr, unsigned long long* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) asm volatile("mov.b64 %0, {%1, %2};" : "=l"(v[i]) : "r"(r[2 * i]), "r"(r[2 * i + 1]));
}
// packed fp32 pairs (sm_100 FADD2 / FMUL2: two IEEE fp32 operations per issue slot, same rounding as the scalar ops)
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  unsigned long long d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d;
}
__device__ __forceinline__ void f2_unpack(unsigned long long a, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a));
}
__device__ __forceinline__ unsigned long long f2_relu(unsigned long long a) {
  float lo, hi; f2_unpack(a, lo, hi); return f2_pack(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
}
__device__ __forceinline__ unsigned long long f2_shfl_down1(unsigned long long a) {
  float lo, hi; f2_unpack(a, lo, hi);
  return f2_pack(__shfl_down_sync(0xffffffffu, lo, 1), __shfl_down_sync(0xffffffffu, hi, 1));
}
// named barrier among `count` threads (epilogue warps only; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> 16 TMEM columns of this warp's 32 lanes (scratch columns next to the accumulators)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, SWIZZLE_NONE, K-major (cute/arch/mma_sm100_desc.hpp):
//   [0,14) start>>4 | [16,30) LBO>>4 (stride between the two 16-B K-chunks of one MMA)
//   [32,46) SBO>>4 (stride between 8-row groups) | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((saddr & 0x3FFFF) >> 4) | (uint64_t(lbo_bytes >> 4) << 16) | (uint64_t(sbo_bytes >> 4) << 32) |
         (uint64_t(1) << 46);
}
// instruction descriptor kind::f16: c=f32 (bit4), a=b=f16 (0), K-major both, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) { return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24); }
constexpr uint32_t kIdesc = make_idesc_f16(128, 64);

__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}


}  // namespace cmlpl
