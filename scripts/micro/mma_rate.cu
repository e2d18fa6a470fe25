// Microbenchmark: cycles per tcgen05.mma (M=128, K=16, fp16, SS) as a function of N, and with
// the A operand either advancing (new 4 KB per MMA) or fixed.  nvcc -arch=sm_100a, run on B200.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../cmlpl_b200/csrc/sm100_ptx.cuh"
using namespace cmlpl;
namespace cmlpl { void set_error(const char*, ...) {} int sm_count() { return 148; } }

template <int N, int M = 128>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int a_step, int concurrent_lsu) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t tm; __shared__ unsigned long long bar;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tm), 512);
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    constexpr uint64_t kHi = (uint64_t(128 >> 4) | (uint64_t(1) << 14)) << 32;
    const uint32_t a_lo = ((sbase) >> 4) | (uint32_t(4096 >> 4) << 16);               // LBO 4096
    const uint32_t b_lo = ((sbase + 100 * 1024) >> 4) | (uint32_t(4096 >> 4) << 16);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t off = uint32_t((i % 16) * a_step) >> 4;
      umma_f16(0, kHi | (a_lo + off), kHi | (b_lo + ((i % 8) * 512 >> 4)), make_idesc_f16(M, N), 1);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, 99);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  } else if (concurrent_lsu && threadIdx.x >= 32) {
    // background shared-memory traffic from other warps (like loaders / epilogue)
    const uint32_t base = sbase + 160 * 1024;
    uint32_t acc = 0;
    for (int i = 0; i < concurrent_lsu; ++i) {
      uint32_t a, b, c, d;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
                   : "r"(base + ((threadIdx.x + i * 96) % 1024) * 16));
      acc ^= a ^ d;
    }
    if (acc == 12345) out[1] = acc;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N, int M = 128> void run(long long* d, int a_step, int lsu) {
  const int iters = 2000;
  cudaFuncSetAttribute(k<N, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  k<N, M><<<148, 128, 200 * 1024>>>(d, iters, a_step, lsu);
  cudaDeviceSynchronize();
  k<N, M><<<148, 128, 200 * 1024>>>(d, iters, a_step, lsu);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("M=%3d N=%3d a_step=%5d lsu=%5d : %.1f cycles/MMA  (%s)\n", M, N, a_step, lsu, double(h) / iters, cudaGetErrorString(e));
}
int main() {
  long long* d; cudaMalloc(&d, 64);
  for (int a_step : {0, 2048}) {
    run<64>(d, a_step, 0); run<128>(d, a_step, 0); run<192>(d, a_step, 0); run<256>(d, a_step, 0);
  }
  run<64>(d, 2048, 20000); run<192>(d, 2048, 20000);
  run<64, 64>(d, 2048, 0); run<128, 64>(d, 2048, 0); run<256, 64>(d, 2048, 0); run<32, 128>(d, 2048, 0); run<16, 128>(d, 2048, 0);
  return 0;
}
