// Per-pixel conv2 stage of the shared-conv1 scene path: for every pixel, gather its 10x10 pooled
// conv1 window from the scene-level pooled maps (conv1_scene_sm100.cu), run conv2 (+bias, +residual,
// ReLU, 2x2 avg-pool; tools/models.py:137-140) with zero padding at the PATCH border, write the 5x5x64
// pooled features as UMMA tiles for the head GEMM.
//
// Same machinery as patch_cnn_sm100.cu (no-swizzle K-major planes, taps = descriptor offsets, parity
// planes so that pooling stays in-lane, taps sharing an A operand fused into N=128 tcgen05.mma), plus:
//  * TWO pixels per accumulator tile pair: the plane rows of pixel A and pixel B are interleaved
//    (plane row 2k+p = row k of pixel p), so a dy shift is 2 plane rows and one 128-row tile holds
//    the 5x12 output positions of both pixels (120 of 128 rows instead of 60);
//  * three activation-plane stages and two TMEM stages, so the gather runs two pairs ahead of the
//    MMAs (it slows down while the tensor core saturates shared memory) and the epilogue of pair i-1
//    overlaps the MMAs of pair i.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "conv_pair_issue.cuh"

namespace cmlpl {

namespace pc2 {
constexpr int W = 20, H2 = 10, PW2 = 12, P = 25;
constexpr int ROWP = 2 * PW2;                          // entries per pixel-row step (two interleaved pixels)
constexpr int ENT = 1 + 12 * PW2 + 2;                  // 147 entries per chunk plane (6 plane rows x 2 pixels)
constexpr int CH = ENT * 16;
constexpr int PLANE = 8 * CH;
constexpr int ABYTES = 2 * PLANE;                      // both parities of a pixel pair
constexpr int WBYTES = 3 * 8 * 192 * 16;
constexpr int NITEM = 100 * 8;                         // (pooled cell, chunk) items per pixel
constexpr int NBUF = 3;                                // activation-plane stages (the gather runs two pairs ahead)
constexpr int S_W = 0, S_A = WBYTES, S_TAB = S_A + NBUF * ABYTES, S_BIAS = S_TAB + NITEM * 16;
constexpr int S_BAR = (S_BIAS + 256 + 7) / 8 * 8, S_TMEM = S_BAR + 128;
// the tile reads up to entry 127 + ROWP + 2 of the last chunk plane of the second buffer
constexpr int S_END = S_TMEM + 16;
constexpr int SMEM = (S_END + 127) / 128 * 128;
constexpr int kEpi = 256, kLoad = 128, kThreads = kEpi + kLoad + 32;
constexpr int kMmaWarp = (kEpi + kLoad) / 32;
enum { A_FULL0 = 0, A_EMPTY0 = NBUF, D_FULL0 = 2 * NBUF, D_EMPTY0 = 2 * NBUF + 2 };
}  // namespace pc2

// pm: f16 [9][PR][PC][64] pooled conv1 maps (one 128-byte line per cell); p2t: UMMA tiles [ceil(n/128)][200][128][8]
__global__ void __launch_bounds__(pc2::kThreads, 1)
patch_conv2_kernel(const __half* __restrict__ pm, int cols, int band_rows, const unsigned char* __restrict__ w2p,
                   const float* __restrict__ b2g, __half* __restrict__ p2t, long long* __restrict__ trace) {
  using namespace pc2;
#define PC2_TRACE(slot)                                                                                   \
  do {                                                                                                    \
    if (trace && blockIdx.x == 0 && jj < 64) trace[jj * 16 + (slot)] = clock64();                        \
  } while (0)
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + S_BAR;
  float* sbias = reinterpret_cast<float*>(smem + S_BIAS);
  const int PR = band_rows + W - 1, PC = cols + W - 1;
  const int64_t plane = int64_t(PR) * PC;
  const int64_t npix = int64_t(band_rows) * cols;
  const int64_t per = ((npix + gridDim.x - 1) / gridDim.x + 1) & ~int64_t(1);   // even, so pairs never straddle CTAs
  const int64_t p_begin = blockIdx.x * per;
  const int64_t p_end = (p_begin + per < npix) ? p_begin + per : npix;

  {  // weights, zeroed activation planes, bias, gather table
    const uint4* gw = reinterpret_cast<const uint4*>(w2p);
    uint4* sw = reinterpret_cast<uint4*>(smem + S_W);
    for (int i = tid; i < WBYTES / 16; i += kThreads) sw[i] = __ldg(gw + i);
    uint4* z = reinterpret_cast<uint4*>(smem + S_A);
    for (int i = tid; i < NBUF * ABYTES / 16; i += kThreads) z[i] = make_uint4(0, 0, 0, 0);
    // table[cell*8 + ch] = (global offset in halves relative to the pixel's window origin, smem offset of
    // pixel 0): the 8 chunks of a cell are one 128-byte line
    for (int it = tid; it < NITEM; it += kThreads) {
      const int ch = it & 7, cell = it >> 3;
      const int i = cell / H2, j = cell - i * H2;
      const int A = i == 0 ? 0 : (i == H2 - 1 ? 2 : 1), B = j == 0 ? 0 : (j == H2 - 1 ? 2 : 1);
      const int64_t goff = (int64_t(A * 3 + B) * plane + int64_t(2 * i) * PC + 2 * j) * 64 + ch * 8;
      const int q = i & 1, prow = (i + q) >> 1;              // even plane: i/2 ; odd plane: (i+1)/2
      const uint32_t soff = q * PLANE + ch * CH + (1 + (2 * prow) * PW2 + j) * 16;
      *reinterpret_cast<int64_t*>(smem + S_TAB + it * 16) = goff;
      *reinterpret_cast<uint32_t*>(smem + S_TAB + it * 16 + 8) = soff;
    }
  }
  if (tid < 64) sbias[tid] = b2g[tid];
  if (tid == 0) {
    for (int b = 0; b < NBUF; ++b) { mbar_init(bars + 8 * (A_FULL0 + b), kLoad); mbar_init(bars + 8 * (A_EMPTY0 + b), 1 + kEpi); }
    for (int b = 0; b < 2; ++b) { mbar_init(bars + 8 * (D_FULL0 + b), 1); mbar_init(bars + 8 * (D_EMPTY0 + b), kEpi); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(sbase + S_TMEM, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + S_TMEM);
  const int dbg = trace ? int(trace[1023]) : 0;            // diagnostics only: bit0 skip gather, bit1 skip epilogue work

  if (warp >= 8 && warp < kMmaWarp) {
    // ================================================================ loaders
    const int lt = tid - kEpi;
    uint32_t jj = 0;
    for (int64_t p = p_begin; p < p_end; p += 2, ++jj) {
      const uint32_t buf = jj % NBUF, ph = (jj / NBUF) & 1;
      mbar_wait(bars + 8 * (A_EMPTY0 + buf), ph ^ 1, 51);
      if (lt == 0) PC2_TRACE(0);
#pragma unroll 1
      for (int px = 0; px < 2; ++px) {
        const int64_t pp = (p + px < p_end) ? p + px : p;      // odd tail: pixel A twice, second result discarded
        const int rb = int(pp / cols), c = int(pp - int64_t(rb) * cols);
        const __half* src0 = pm + (int64_t(rb) * PC + c) * 64;
        const uint32_t dst0 = sbase + S_A + buf * ABYTES + px * (PW2 * 16);
#pragma unroll
        for (int it = lt; it < ((dbg & 1) ? 0 : NITEM); it += kLoad) {
          const uint4 rec = *reinterpret_cast<const uint4*>(smem + S_TAB + it * 16);
          const int64_t goff = int64_t((uint64_t(rec.y) << 32) | rec.x);
          cp_async16(dst0 + rec.z, src0 + goff);
        }
      }
      // software pipeline: keep this pair's copies in flight while completing the previous pair's
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (jj > 0) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        fence_proxy_async();
        mbar_arrive(bars + 8 * (A_FULL0 + (jj - 1) % NBUF));
      }
      if (lt == 0) PC2_TRACE(1);
    }
    if (jj > 0) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_proxy_async();
      mbar_arrive(bars + 8 * (A_FULL0 + (jj - 1) % NBUF));
    }
  } else if (warp == kMmaWarp) {
    // ================================================================ MMA issuer
    if (tmem != 0) { printf("patch_conv2: unexpected TMEM base %u\n", tmem); __trap(); }
    const uint32_t w_lo = ((sbase + S_W) >> 4) | (uint32_t(kWLbo >> 4) << 16);
    uint32_t jj = 0;
    for (int64_t p = p_begin; p < p_end; p += 2, ++jj) {
      const uint32_t buf = jj % NBUF, ph = (jj / NBUF) & 1, stage = jj & 1, dph = (jj >> 1) & 1;
      const uint32_t a_lo = ((sbase + S_A + buf * ABYTES + 16) >> 4) | (uint32_t(CH >> 4) << 16);
      mbar_wait(bars + 8 * (A_FULL0 + buf), ph, 52);
      if (lane == 0) PC2_TRACE(2);
      mbar_wait(bars + 8 * (D_EMPTY0 + stage), dph ^ 1, 53);
      if (lane == 0) PC2_TRACE(3);
      tc_fence_after();
      if (elect_one_sync()) {
        issue_conv_pair<ROWP, CH, PLANE>(stage * 128, a_lo, w_lo);
        umma_commit(bars + 8 * (D_FULL0 + stage));
        umma_commit(bars + 8 * (A_EMPTY0 + buf));
      }
      __syncwarp();
      if (lane == 0) PC2_TRACE(4);
    }
  } else {
    // ================================================================ epilogue (warps 0-7)
    const int L = (warp & 3) * 32 + lane, chalf = warp >> 2;
    const uint32_t lane_addr = (uint32_t((warp & 3) * 32) << 16) + chalf * 32;
    const float* bias2 = sbias + chalf * 32;
    const int rho = L / PW2, x = L - rho * PW2;               // interleaved plane row, column
    const int i = rho >> 1, px = rho & 1;
    const bool in_tile = rho < 2 * (H2 / 2) && x < H2;
    uint32_t jj = 0;
    for (int64_t p = p_begin; p < p_end; p += 2, ++jj) {
      const uint32_t buf = jj % NBUF, ph = (jj / NBUF) & 1, stage = jj & 1, dph = (jj >> 1) & 1;
      const int64_t pix = p + px;
      const bool valid = in_tile && pix < p_end;
      const bool writer = valid && ((x & 1) == 0);
      // residual = pooled conv1 output (models.py:137,139): rows 2i (even plane row i) and 2i+1 (odd plane row i+1)
      uint4 re[4], ro[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { re[k] = make_uint4(0, 0, 0, 0); ro[k] = make_uint4(0, 0, 0, 0); }
      if (tid == 0) PC2_TRACE(5);
      mbar_wait(bars + 8 * (A_FULL0 + buf), ph, 54);
      if (in_tile) {
        const unsigned char* ab = smem + S_A + buf * ABYTES + (chalf * 4) * CH;
        const unsigned char* rE = ab + (1 + (2 * i + px) * PW2 + x) * 16;
        const unsigned char* rO = ab + PLANE + (1 + (2 * (i + 1) + px) * PW2 + x) * 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          re[k] = *reinterpret_cast<const uint4*>(rE + k * CH);
          ro[k] = *reinterpret_cast<const uint4*>(rO + k * CH);
        }
      }
      {
        uint32_t sink = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) sink |= re[k].x ^ ro[k].w;
        asm volatile("" ::"r"(sink) : "memory");
        mbar_arrive(bars + 8 * (A_EMPTY0 + buf));             // this thread's residual reads are done
      }
      const int pos = i * (H2 / 2) + (x >> 1);
      __half* dst = p2t + (((pix >> 7) * (P * 8) + pos * 8 + chalf * 4) * 128 + (pix & 127)) * 8;
      if (tid == 0) PC2_TRACE(6);
      mbar_wait(bars + 8 * (D_FULL0 + stage), dph, 55);
      if (tid == 0) PC2_TRACE(7);
      tc_fence_after();
#pragma unroll
      for (int hg = 0; hg < 2; ++hg) {
        float e[16], o[16];
        tmem_ld16(lane_addr + stage * 128 + hg * 16, e);
        tmem_ld16(lane_addr + stage * 128 + 64 + hg * 16, o);
        tmem_ld_wait();
        if (hg == 1) { tc_fence_before(); mbar_arrive(bars + 8 * (D_EMPTY0 + stage)); }
        if (dbg & 2) continue;
        const __half2* he = reinterpret_cast<const __half2*>(&re[hg * 2]);
        const __half2* ho = reinterpret_cast<const __half2*>(&ro[hg * 2]);
        float pooled[16];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float2 fe = __half22float2(he[k]), fo = __half22float2(ho[k]);
          const float b0 = bias2[hg * 16 + 2 * k], b1 = bias2[hg * 16 + 2 * k + 1];
          pooled[2 * k] = fmaxf(e[2 * k] + b0 + fe.x, 0.f) + fmaxf(o[2 * k] + b0 + fo.x, 0.f);
          pooled[2 * k + 1] = fmaxf(e[2 * k + 1] + b1 + fe.y, 0.f) + fmaxf(o[2 * k + 1] + b1 + fo.y, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) pooled[k] = (pooled[k] + __shfl_xor_sync(0xffffffffu, pooled[k], 1)) * 0.25f;
        if (writer) {
          __half2 hv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) hv[k] = __floats2half2_rn(pooled[2 * k], pooled[2 * k + 1]);
          uint4* d4 = reinterpret_cast<uint4*>(dst) + hg * 2 * 128;
          __stcs(d4, *reinterpret_cast<uint4*>(&hv[0]));
          __stcs(d4 + 128, *reinterpret_cast<uint4*>(&hv[4]));
        }
      }
      if (tid == 0) PC2_TRACE(8);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

}  // namespace cmlpl

using namespace cmlpl;

static int launch_patch_conv2(const void* pm, int cols, int w, int band_rows, const void* packed, void* p2t,
                              long long* trace, cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(pm && packed && p2t, "patch_conv2: null pointer");
  CMLPL_CHECK_ARG(w == 20 && cols > 0 && band_rows > 0, "patch_conv2: bad dims (w must be 20)");
  const PackedLayout L = packed_layout(1, 1, w);
  const unsigned char* pk = static_cast<const unsigned char*>(packed);
  CMLPL_MAX_DYN_SMEM(patch_conv2_kernel, pc2::SMEM);
  const int64_t npix = int64_t(band_rows) * cols;
  int64_t grid = sm_count();
  if (grid > (npix + 1) / 2) grid = (npix + 1) / 2;
  patch_conv2_kernel<<<int(grid), pc2::kThreads, pc2::SMEM, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(pm), cols, band_rows, pk + L.w2, reinterpret_cast<const float*>(pk + L.b2),
      static_cast<__half*>(p2t), trace);
  CMLPL_CHECK_LAUNCH("patch_conv2");
  return CMLPL_OK;
}

extern "C" int cmlpl_patch_conv2_f16_tiled(const void* pm, int cols, int w, int band_rows, const void* packed,
                                           void* p2t, cmlpl_stream_t stream) {
  return launch_patch_conv2(pm, cols, w, band_rows, packed, p2t, nullptr, stream);
}

// Diagnostics: CTA 0 writes clock64() stamps of its first 64 pixel pairs to trace i64 [64][16].
extern "C" int cmlpl_debug_patch_conv2_trace(const void* pm, int cols, int w, int band_rows, const void* packed,
                                             void* p2t, long long* trace, cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(trace, "patch_conv2_trace: null trace buffer");
  return launch_patch_conv2(pm, cols, w, band_rows, packed, p2t, trace, stream);
}
