// Probe (run on a B200): do the tcgen05 features the training kernels rely on behave as assumed?
//   T1  MN-major A and B operands in the chunk-planar no-swizzle layout [chunk of 8 ch][pos][8 halves] with an
//       arbitrary 16-byte-multiple start offset (tap shift along K = positions), M = 64:
//         D[co][ci] = sum_k dz[k + sa][co] * act[k + sb][ci]            (weight-gradient shape)
//       LBO = 128 B (next 8 positions), SBO = plane stride (next 8 channels); the swapped assignment is tried too.
//   T2  M = 64 accumulator placement: rows r -> TMEM lane (r % 16) + 32 * (r / 16); a second accumulator at lane
//       offset 16 of the same columns does not disturb the first.
//   T3  K-major A (positions x channels) with an MN-major B (weights [kchunk(ci)][co][8 ci] read as B[n=ci][k=co]):
//         D[pos][ci] = sum_co dz[pos][co] * W[co][ci]                    (data-gradient shape)
// nvcc -gencode arch=compute_100a,code=sm_100a -o mn_major_probe mn_major_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "../../cmlpl_b200/csrc/sm100_ptx.cuh"
using namespace cmlpl;
namespace cmlpl { void set_error(const char*, ...) {} int sm_count() { return 148; } }

constexpr int P = 96;            // positions per chunk plane
constexpr int CHB = P * 16;      // bytes per chunk plane

__host__ __device__ constexpr uint32_t idesc(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

// mode 0: T1 (lbo=128, sbo=CHB), mode 1: T1 with swapped lbo/sbo, mode 2: T3
__global__ void __launch_bounds__(128, 1) probe(const __half* dz, const __half* act, const __half* wgt, float* out,
                                                float* out2, int mode, int sa, int sb) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t tm; __shared__ unsigned long long bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // smem: dz planes @0 (8*CHB), act planes @16K, weights @32K ([8 kchunk(ci)][64 co][8 ci] = 8 KB)
  for (int i = tid; i < 8 * P * 8; i += 128) {
    reinterpret_cast<__half*>(smem)[i] = dz[i];
    reinterpret_cast<__half*>(smem + 16384)[i] = act[i];
  }
  for (int i = tid; i < 8 * 64 * 8; i += 128) reinterpret_cast<__half*>(smem + 32768)[i] = wgt[i];
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tm), 512);
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tm;
  if (tid == 0) {
    if (mode < 2) {
      const uint32_t lbo = mode == 0 ? 128 : CHB, sbo = mode == 0 ? CHB : 128;
      for (int acc = 0; acc < 2; ++acc) {          // second accumulator: lane offset 16, other shifts
        const int s_a = acc == 0 ? sa : sa + 1, s_b = acc == 0 ? sb : sb + 2;
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t a = make_desc(sbase + (s_a + ks * 16) * 16, lbo, sbo);
          const uint64_t b = make_desc(sbase + 16384 + (s_b + ks * 16) * 16, lbo, sbo);
          umma_f16(tmem + (uint32_t(acc * 16) << 16), a, b, idesc(64, 64, 1, 1), ks);
        }
      }
    } else {
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t a = make_desc(sbase + sa * 16 + ks * 2 * CHB, CHB, 128);            // K-major: LBO = plane stride
        const uint64_t b = make_desc(sbase + 32768 + ks * 16 * 16, 128, 64 * 16);          // MN-major: LBO = 8 co rows
        umma_f16(tmem, a, b, idesc(128, 64, 0, 1), ks);
      }
    }
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0, 1);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = v[j];
  }
  (void)out2;
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// cycles per MMA for M=64 N=64 MN-major operands (the weight-gradient instruction)
__global__ void __launch_bounds__(128, 1) rate(long long* out, int iters, int m, int a_mn, int b_mn) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  __shared__ uint32_t tm; __shared__ unsigned long long bar;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(&tm), 512);
  fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t a = make_desc(sbase + (i % 32) * 256, a_mn ? 128 : CHB, a_mn ? CHB : 128);
      const uint64_t b = make_desc(sbase + 16384 + (i % 16) * 256 + 16, b_mn ? 128 : CHB, b_mn ? CHB : 128);
      umma_f16(tm + (i % 4) * 64, a, b, idesc(m, 64, a_mn, b_mn), 1);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0, 99);
    out[0] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  std::vector<__half> dz(8 * P * 8), act(8 * P * 8), wg(8 * 64 * 8);
  std::vector<float> fdz(P * 64), fact(P * 64), fw(64 * 64);
  srand(7);
  auto rnd = [] { return float((rand() % 17) - 8) / 8.f; };
  for (int p = 0; p < P; ++p)
    for (int c = 0; c < 64; ++c) {
      fdz[p * 64 + c] = rnd(); fact[p * 64 + c] = rnd();
      dz[((c / 8) * P + p) * 8 + c % 8] = __float2half(fdz[p * 64 + c]);
      act[((c / 8) * P + p) * 8 + c % 8] = __float2half(fact[p * 64 + c]);
    }
  for (int co = 0; co < 64; ++co)
    for (int ci = 0; ci < 64; ++ci) {
      fw[co * 64 + ci] = rnd();
      wg[((ci / 8) * 64 + co) * 8 + ci % 8] = __float2half(fw[co * 64 + ci]);
    }
  __half *d_dz, *d_act, *d_w; float* d_out; long long* d_t;
  cudaMalloc(&d_dz, dz.size() * 2); cudaMalloc(&d_act, act.size() * 2); cudaMalloc(&d_w, wg.size() * 2);
  cudaMalloc(&d_out, 128 * 64 * 4); cudaMalloc(&d_t, 64);
  cudaMemcpy(d_dz, dz.data(), dz.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_act, act.data(), act.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(d_w, wg.data(), wg.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> out(128 * 64);
  const int sa = 3, sb = 5;
  for (int mode = 0; mode < 3; ++mode) {
    cudaMemset(d_out, 0, 128 * 64 * 4);
    probe<<<1, 128, 64 * 1024>>>(d_dz, d_act, d_w, d_out, nullptr, mode, sa, sb);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost);
    double err0 = 0, err1 = 0;
    if (mode < 2) {
      for (int acc = 0; acc < 2; ++acc) {
        const int s_a = acc == 0 ? sa : sa + 1, s_b = acc == 0 ? sb : sb + 2;
        for (int co = 0; co < 64; ++co)
          for (int ci = 0; ci < 64; ++ci) {
            double ref = 0;
            for (int k = 0; k < 32; ++k) ref += double(fdz[(k + s_a) * 64 + co]) * fact[(k + s_b) * 64 + ci];
            const int lane = (co % 16) + 32 * (co / 16) + acc * 16;
            const double d = fabs(out[lane * 64 + ci] - ref);
            (acc == 0 ? err0 : err1) = fmax(acc == 0 ? err0 : err1, d);
          }
      }
      printf("T1/T2 mode %d (%s): max|err| acc0 = %.4g, acc1(lane+16) = %.4g   [%s]\n", mode,
             mode == 0 ? "LBO=128,SBO=plane" : "LBO=plane,SBO=128", err0, err1, cudaGetErrorString(e));
    } else {
      for (int p = 0; p < 80; ++p)
        for (int ci = 0; ci < 64; ++ci) {
          double ref = 0;
          for (int co = 0; co < 64; ++co) ref += double(fdz[(p + sa) * 64 + co]) * fw[co * 64 + ci];
          err0 = fmax(err0, fabs(out[p * 64 + ci] - ref));
        }
      printf("T3 K-major A x MN-major B: max|err| = %.4g   [%s]\n", err0, cudaGetErrorString(e));
    }
  }
  for (int m : {64, 128})
    for (int mn = 0; mn < 2; ++mn) {
      rate<<<148, 128, 64 * 1024>>>(d_t, 2000, m, mn, mn);
      cudaDeviceSynchronize();
      rate<<<148, 128, 64 * 1024>>>(d_t, 2000, m, mn, mn);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0; cudaMemcpy(&h, d_t, 8, cudaMemcpyDeviceToHost);
      printf("rate M=%d N=64 %s: %.1f cycles/MMA [%s]\n", m, mn ? "MN-major" : "K-major", double(h) / 2000, cudaGetErrorString(e));
    }
  return 0;
}
