// Shared helpers for libcmlpl_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cmlpl.h"

namespace cmlpl {

void set_error(const char* fmt, ...);

#define CMLPL_CHECK_ARG(cond, ...)                     \
  do {                                                 \
    if (!(cond)) {                                     \
      ::cmlpl::set_error(__VA_ARGS__);                 \
      return CMLPL_ERR_ARG;                            \
    }                                                  \
  } while (0)

#define CMLPL_CHECK_LAUNCH(name)                                                      \
  do {                                                                                \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ::cmlpl::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return CMLPL_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

#define CMLPL_CUDA(call)                                                              \
  do {                                                                                \
    cudaError_t e__ = (call);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::cmlpl::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return CMLPL_ERR_CUDA;                                                          \
    }                                                                                 \
  } while (0)

// opt a kernel into `bytes` of dynamic shared memory; the attribute call is made once per (call site, device) and
// again only if a larger size is requested -- it costs microseconds of host time that a 1 ms step cannot hide
#define CMLPL_MAX_DYN_SMEM(kernel, bytes)                                                             \
  do {                                                                                                \
    static int set__[64];                                                                             \
    int dev__ = 0;                                                                                    \
    if (cudaGetDevice(&dev__) != cudaSuccess || dev__ < 0 || dev__ >= 64) dev__ = 0;                  \
    if (int(bytes) > set__[dev__]) {                                                                  \
      CMLPL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes))); \
      set__[dev__] = int(bytes);                                                                      \
    }                                                                                                 \
  } while (0)

// number of SMs of the current device (cached)
int sm_count();

// a1 index map (tools/hyper_tools.py:35-55 == symmetric padding): single reflection.
__host__ __device__ __forceinline__ int mirror_index(int o, int n) {
  return o < 0 ? -o - 1 : (o >= n ? 2 * n - 1 - o : o);
}

// window of pixel r: rows r+lo .. r+lo+w-1.  even w (ExtractPatches): lo=-w/2;
// odd w (ExtractPatches_for_base): lo=-(w-1)/2.   Both equal -(w/2) in integer division.
__host__ __device__ __forceinline__ int window_lo(int w) { return -(w / 2); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- class bookkeeping of the dense (scene-level) conv2 / pool / classifier path (conv2_scene_sm100.cu) ----
// pooled row/column index that represents conv2 class k (0: i=0, 1: i=1, 2: i=2..7, 3: i=8, 4: i=9) of a
// 10-wide pooled window, and the PM border class (top/mid/bot) of pooled index i
__host__ __device__ constexpr int rep_of(int k) { return k == 0 ? 0 : (k == 1 ? 1 : (k == 2 ? 4 : (k == 3 ? 8 : 9))); }
__host__ __device__ constexpr int cls_of(int i) { return i == 0 ? 0 : (i == 9 ? 2 : 1); }
// block b = Alpha*3 + Beta: number of I / J values, first map index, Y classes per (Alpha,u)
__host__ __device__ constexpr int blk_n(int cls) { return cls == 1 ? 3 : 1; }
__host__ __device__ constexpr int blk_first(int cls) { return cls == 0 ? 0 : (cls == 1 ? 1 : 4); }
__host__ __device__ constexpr int blk_start(int b) {
  return b == 0 ? 0 : b == 1 ? 1 : b == 2 ? 4 : b == 3 ? 5 : b == 4 ? 8 : b == 5 ? 17 : b == 6 ? 20 : b == 7 ? 21 : 24;
}

// ---- packed BaseNet2 weights (layout shared by pack.cu and the scene kernels) ----
// All offsets in bytes from the start of the packed buffer, 256-B aligned.
struct PackedLayout {
  size_t w1;      // f16 [3 dx][8 kchunks][192 rows = (2-dy)*64+n][8]   conv1, UMMA no-swizzle K-major B operand
  size_t w2;      // f16 same for conv2
  size_t b1;      // f32 [64]
  size_t b2;      // f32 [64]
  size_t w0;      // f32 [60][64]  conv0 weight transposed (ci-major) for the per-pixel map
  size_t b0;      // f32 [64]
  size_t wspe;    // f32 [1024][B]  (as in the state dict)
  size_t bspe;    // f32 [1024]
  size_t wc_conv; // f32 [C][P*64]   classifier columns of the conv features, permuted to (pos, ch)
  size_t wc_spe;  // f32 [C][1024]
  size_t bc;      // f32 [C] (padded to 16)
  size_t w1s;     // f16 [4 n-tiles][KC][256][8]  feat_spe weight, UMMA B operand tiles (K padded to 16)
  size_t wc16;    // f16 [(P*8 + 128) k-chunks][16 classes][8]  classifier over [conv(pos,ch) | spectral]
  size_t wcq;     // f16 per (Alpha,Beta) block [8 k-chunks][N = maps*16 rows = map*16+cls][8]: 0.25 * conv classifier
                  //     columns of pooled cell (I,J), per border block (Al,Be), rows J-major (w = 20 only; pool2_cls_kernel)
  size_t total;
  int conv_pos;   // P = (w/4)^2 pooled positions
  int kc_spe_in;  // KC = ceil(B/16)*2: 16-byte K-chunks of the spectral input
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

__host__ __device__ inline PackedLayout packed_layout(int B, int C, int w) {
  PackedLayout L;
  int p = (w / 2) / 2;
  L.conv_pos = p * p;
  size_t o = 0;
  L.w1 = o; o = align256(o + 9 * 8 * 64 * 8 * 2);
  L.w2 = o; o = align256(o + 9 * 8 * 64 * 8 * 2);
  L.b1 = o; o = align256(o + 64 * 4);
  L.b2 = o; o = align256(o + 64 * 4);
  L.w0 = o; o = align256(o + 60 * 64 * 4);
  L.b0 = o; o = align256(o + 64 * 4);
  L.wspe = o; o = align256(o + size_t(1024) * B * 4);
  L.bspe = o; o = align256(o + 1024 * 4);
  L.wc_conv = o; o = align256(o + size_t(C) * L.conv_pos * 64 * 4);
  L.wc_spe = o; o = align256(o + size_t(C) * 1024 * 4);
  L.bc = o; o = align256(o + size_t(C > 16 ? C : 16) * 4);
  L.kc_spe_in = ((B + 15) / 16) * 2;
  L.w1s = o; o = align256(o + size_t(4) * L.kc_spe_in * 256 * 16);
  L.wc16 = o; o = align256(o + size_t(L.conv_pos * 8 + 128) * 16 * 16);
  L.wcq = o; o = align256(o + size_t(400) * 128);
  L.total = o;
  return L;
}

}  // namespace cmlpl
