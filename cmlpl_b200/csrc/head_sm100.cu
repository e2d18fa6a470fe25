// Tensor-core head of the scene path (sm_100a): the spectral branch relu(feat_spe(x))
// (tools/models.py:142-143) and the classifier over [conv features | spectral features] + argmax
// (models.py:144,150; hyper_tools.py:426) as two tcgen05 GEMMs with fp16 operands / fp32 accumulate.
//
// Every A operand lives in HBM already in the UMMA "no-swizzle K-major" tile layout
//     [m-tile][K/8 chunks][128 rows][8 halves]
// so that a K-slab of a tile is one contiguous block: the loader is a single thread issuing
// cp.async.bulk (UBLKCP) copies that complete on an mbarrier, and a chunk needs no repacking
// before tcgen05.mma reads it.  x16_tile_kernel writes the spectra in this layout,
// spectral_hidden_kernel writes the 1024 hidden features in it, and patch_cnn writes the pooled
// conv features in it (p2_tiled).
//   spectral_hidden_kernel: H = relu(X . Wspe^T + b)    M=128 x N=256 tiles, K = B (padded to 16)
//   head_kernel           : logits = [P2 | H] . Wc^T + bc, argmax     M=128 x N=16, K = 1600 + 1024
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cmlpl {

// ------------------------------------------------------------------ spectra -> fp16 tiles
__global__ void x16_tile_kernel(const float* __restrict__ X, int64_t n, int B, int KC, int64_t total,
                                __half* __restrict__ out) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int row = int(t & 127);
    const int64_t r = t >> 7;
    const int kc = int(r % KC);
    const int64_t p = (r / KC) * 128 + row;
    __half2 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = kc * 8 + 2 * e;
      const float a = (p < n && k < B) ? __ldg(X + p * B + k) : 0.f;
      const float b = (p < n && k + 1 < B) ? __ldg(X + p * B + k + 1) : 0.f;
      h[e] = __floats2half2_rn(a, b);
    }
    *reinterpret_cast<uint4*>(out + t * 8) = *reinterpret_cast<uint4*>(h);
  }
}

// same from the raw cube: x16 = (raw - mu) * inv_sigma  (featureNormalize(X, 1), hyper_tools.py:292)
template <typename T>
__global__ void x16_tile_raw_kernel(const T* __restrict__ X, int64_t n, int B, int KC, int64_t total,
                                    const float* __restrict__ mu, const float* __restrict__ inv_sigma,
                                    __half* __restrict__ out) {
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    const int row = int(t & 127);
    const int64_t r = t >> 7;
    const int kc = int(r % KC);
    const int64_t p = (r / KC) * 128 + row;
    __half2 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = kc * 8 + 2 * e;
      const float a = (p < n && k < B) ? (float(X[p * B + k]) - __ldg(mu + k)) * __ldg(inv_sigma + k) : 0.f;
      const float b = (p < n && k + 1 < B) ? (float(X[p * B + k + 1]) - __ldg(mu + k + 1)) * __ldg(inv_sigma + k + 1) : 0.f;
      h[e] = __floats2half2_rn(a, b);
    }
    *reinterpret_cast<uint4*>(out + t * 8) = *reinterpret_cast<uint4*>(h);
  }
}

// ------------------------------------------------------------------ H = relu(X . W^T + b)
constexpr int kHidThreads = 320;   // warps 0-7 epilogue, warp 8 loader, warp 9 MMA issuer

__global__ void __launch_bounds__(kHidThreads, 1)
spectral_hidden_kernel(const __half* __restrict__ x16, int64_t mtiles, int KC, const __half* __restrict__ w1t,
                       const float* __restrict__ bspe, __half* __restrict__ h16) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t wbytes = uint32_t(KC) * 4096, abytes = uint32_t(KC) * 2048;
  const uint32_t S_W = 0, S_A = wbytes, S_BIAS = S_A + 2 * abytes, S_BAR = S_BIAS + 1024, S_TMEM = S_BAR + 64;
  const uint32_t bars = sbase + S_BAR;
  enum { A_FULL0 = 0, A_FULL1, A_EMPTY0, A_EMPTY1, D_FULL0, D_FULL1, D_EMPTY0, D_EMPTY1 };
  const int ntile = blockIdx.x & 3;
  const int64_t mt0 = blockIdx.x >> 2, mstep = gridDim.x >> 2;

  {  // weights of this N tile (already in UMMA layout) and its bias slice
    const uint4* g = reinterpret_cast<const uint4*>(w1t) + size_t(ntile) * (wbytes / 16);
    uint4* s = reinterpret_cast<uint4*>(smem + S_W);
    for (uint32_t i = tid; i < wbytes / 16; i += kHidThreads) s[i] = __ldg(g + i);
    float* sb = reinterpret_cast<float*>(smem + S_BIAS);
    if (tid < 256) sb[tid] = bspe[ntile * 256 + tid];
  }
  if (tid == 0) {
    mbar_init(bars + 8 * A_FULL0, 1); mbar_init(bars + 8 * A_FULL1, 1);
    mbar_init(bars + 8 * A_EMPTY0, 1); mbar_init(bars + 8 * A_EMPTY1, 1);
    mbar_init(bars + 8 * D_FULL0, 1); mbar_init(bars + 8 * D_FULL1, 1);
    mbar_init(bars + 8 * D_EMPTY0, 256); mbar_init(bars + 8 * D_EMPTY1, 256);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc(sbase + S_TMEM, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + S_TMEM);

  if (warp == 8) {
    if (lane == 0) {
      uint32_t j = 0;
      for (int64_t mt = mt0; mt < mtiles; mt += mstep, ++j) {
        const uint32_t s = j & 1, ph = (j >> 1) & 1;
        mbar_wait(bars + 8 * (A_EMPTY0 + s), ph ^ 1, 21);
        mbar_arrive_expect_tx(bars + 8 * (A_FULL0 + s), abytes);
        bulk_g2s(sbase + S_A + s * abytes, x16 + mt * int64_t(KC) * 1024, abytes, bars + 8 * (A_FULL0 + s));
      }
    }
  } else if (warp == 9) {
    if (tmem != 0) { printf("spectral_hidden: unexpected TMEM base %u\n", tmem); __trap(); }
    constexpr uint64_t kHi = (uint64_t(128 >> 4) | (uint64_t(1) << 14)) << 32;
    constexpr uint32_t idesc = make_idesc_f16(128, 256);
    uint32_t j = 0;
    for (int64_t mt = mt0; mt < mtiles; mt += mstep, ++j) {
      const uint32_t s = j & 1, ph = (j >> 1) & 1;
      mbar_wait(bars + 8 * (A_FULL0 + s), ph, 22);
      mbar_wait(bars + 8 * (D_EMPTY0 + s), ph ^ 1, 23);
      tc_fence_after();
      if (elect_one_sync()) {
        uint32_t a_lo = ((sbase + S_A + s * abytes) >> 4) | (uint32_t(2048 >> 4) << 16);
        uint32_t b_lo = ((sbase + S_W) >> 4) | (uint32_t(4096 >> 4) << 16);
        for (int ks = 0; ks < KC / 2; ++ks) {
          umma_f16(s * 256, kHi | a_lo, kHi | b_lo, idesc, ks ? 1u : 0u);
          a_lo += 4096 >> 4; b_lo += 8192 >> 4;
        }
        umma_commit(bars + 8 * (A_EMPTY0 + s));
        umma_commit(bars + 8 * (D_FULL0 + s));
      }
      __syncwarp();
    }
  } else {
    const int L = (warp & 3) * 32 + lane, chalf = warp >> 2;
    const uint32_t lane_addr = (uint32_t((warp & 3) * 32) << 16) + chalf * 128;
    const float* sb = reinterpret_cast<const float*>(smem + S_BIAS) + chalf * 128;
    uint32_t j = 0;
    for (int64_t mt = mt0; mt < mtiles; mt += mstep, ++j) {
      const uint32_t s = j & 1, ph = (j >> 1) & 1;
      __half* dst = h16 + ((mt * 128 + (ntile * 32 + chalf * 16)) * 128 + L) * 8;
      mbar_wait(bars + 8 * (D_FULL0 + s), ph, 24);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < 8; g += 2) {
        float v0[16], v1[16];
        tmem_ld16(lane_addr + s * 256 + g * 16, v0);
        tmem_ld16(lane_addr + s * 256 + g * 16 + 16, v1);
        tmem_ld_wait();
        if (g == 6) { tc_fence_before(); mbar_arrive(bars + 8 * (D_EMPTY0 + s)); }
        __half2 h[16];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          h[e] = __floats2half2_rn(fmaxf(v0[2 * e] + sb[g * 16 + 2 * e], 0.f), fmaxf(v0[2 * e + 1] + sb[g * 16 + 2 * e + 1], 0.f));
          h[8 + e] = __floats2half2_rn(fmaxf(v1[2 * e] + sb[g * 16 + 16 + 2 * e], 0.f),
                                        fmaxf(v1[2 * e + 1] + sb[g * 16 + 16 + 2 * e + 1], 0.f));
        }
        const uint4* hv = reinterpret_cast<const uint4*>(h);
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(dst + (g * 2 + q) * 1024) = hv[q];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------ logits = [P2 | H] . Wc^T + bc ; argmax
constexpr int kHeadThreads = 192;   // warps 0-3 epilogue, warp 4 loader, warp 5 MMA issuer
constexpr int kHeadStages = 6, kHeadChunkKc = 8, kHeadChunkBytes = kHeadChunkKc * 2048;

__global__ void __launch_bounds__(kHeadThreads, 1)
head_kernel(const __half* __restrict__ p2t, int kc_conv, const __half* __restrict__ h16, int kc_spe, int64_t mtiles,
            int64_t n, int C, const __half* __restrict__ wc16, const float* __restrict__ bc,
            uint8_t* __restrict__ labels, float* __restrict__ logits) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const int kc_all = kc_conv + kc_spe;
  const uint32_t wbytes = uint32_t(kc_all) * 256;
  const uint32_t S_W = 0, S_A = (wbytes + 127) / 128 * 128, S_BAR = S_A + kHeadStages * kHeadChunkBytes, S_TMEM = S_BAR + 256;
  const uint32_t bars = sbase + S_BAR;
  // barriers: full[0..5], empty[6..11], d_full[12,13], d_empty[14,15]
  {
    const uint4* g = reinterpret_cast<const uint4*>(wc16);
    uint4* s = reinterpret_cast<uint4*>(smem + S_W);
    for (uint32_t i = tid; i < wbytes / 16; i += kHeadThreads) s[i] = __ldg(g + i);
  }
  if (tid == 0) {
    for (int i = 0; i < kHeadStages; ++i) { mbar_init(bars + 8 * i, 1); mbar_init(bars + 8 * (kHeadStages + i), 1); }
    mbar_init(bars + 8 * 12, 1); mbar_init(bars + 8 * 13, 1);
    mbar_init(bars + 8 * 14, 128); mbar_init(bars + 8 * 15, 128);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc(sbase + S_TMEM, 32);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + S_TMEM);
  const int nch_conv = kc_conv / kHeadChunkKc, nch = nch_conv + kc_spe / kHeadChunkKc;

  if (warp == 4) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t mt = blockIdx.x; mt < mtiles; mt += gridDim.x) {
        for (int c = 0; c < nch; ++c, ++it) {
          const uint32_t s = it % kHeadStages, ph = (it / kHeadStages) & 1;
          mbar_wait(bars + 8 * (kHeadStages + s), ph ^ 1, 31);
          mbar_arrive_expect_tx(bars + 8 * s, kHeadChunkBytes);
          const __half* src = c < nch_conv ? p2t + (mt * kc_conv + int64_t(c) * kHeadChunkKc) * 1024
                                           : h16 + (mt * kc_spe + int64_t(c - nch_conv) * kHeadChunkKc) * 1024;
          bulk_g2s(sbase + S_A + s * kHeadChunkBytes, src, kHeadChunkBytes, bars + 8 * s);
        }
      }
    }
  } else if (warp == 5) {
    constexpr uint64_t kHi = (uint64_t(128 >> 4) | (uint64_t(1) << 14)) << 32;
    constexpr uint32_t idesc = make_idesc_f16(128, 16);
    uint32_t it = 0, tj = 0;
    for (int64_t mt = blockIdx.x; mt < mtiles; mt += gridDim.x, ++tj) {
      const uint32_t acc = tj & 1, aph = (tj >> 1) & 1;
      mbar_wait(bars + 8 * (14 + acc), aph ^ 1, 32);
      tc_fence_after();
      for (int c = 0; c < nch; ++c, ++it) {
        const uint32_t s = it % kHeadStages, ph = (it / kHeadStages) & 1;
        mbar_wait(bars + 8 * s, ph, 33);
        tc_fence_after();
        if (elect_one_sync()) {
          uint32_t a_lo = ((sbase + S_A + s * kHeadChunkBytes) >> 4) | (uint32_t(2048 >> 4) << 16);
          uint32_t b_lo = ((sbase + S_W + uint32_t(c) * kHeadChunkKc * 256) >> 4) | (uint32_t(256 >> 4) << 16);
#pragma unroll
          for (int ks = 0; ks < kHeadChunkKc / 2; ++ks) {
            umma_f16(tmem + acc * 16, kHi | a_lo, kHi | b_lo, idesc, (c | ks) ? 1u : 0u);
            a_lo += 4096 >> 4; b_lo += 512 >> 4;
          }
          umma_commit(bars + 8 * (kHeadStages + s));
          if (c == nch - 1) umma_commit(bars + 8 * (12 + acc));
        }
        __syncwarp();
      }
    }
  } else {
    const int L = warp * 32 + lane;
    uint32_t tj = 0;
    for (int64_t mt = blockIdx.x; mt < mtiles; mt += gridDim.x, ++tj) {
      const uint32_t acc = tj & 1, aph = (tj >> 1) & 1;
      mbar_wait(bars + 8 * (12 + acc), aph, 34);
      tc_fence_after();
      float v[16];
      tmem_ld16(tmem + (uint32_t(warp * 32) << 16) + acc * 16, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bars + 8 * (14 + acc));
      const int64_t p = mt * 128 + L;
      if (p < n) {
        float best = -INFINITY; int arg = 0;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < C) {
            const float z = v[c] + __ldg(bc + c);
            if (logits) logits[p * C + c] = z;
            if (z > best) { best = z; arg = c; }        // strict '>': first index wins ties (hyper_tools.py:426)
          }
        }
        labels[p] = uint8_t(arg);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}


// ------------------------------------------------------------------ spectral logits: Wc_spe . relu(Wspe . x + b), hidden never in HBM
// spectral_logits_kernel fuses the two GEMMs of the spectral branch (models.py:142-143 and the spectral columns of
// models.py:150): a CTA owns one quarter (256) of the 1024 hidden features and walks pixel tiles of 128.
//   MMA1 (M=128, N=128, K=B)  hidden half-tile -> TMEM (ring of 3 accumulator slots)
//   epilogue                   bias + ReLU -> fp16, written back IN PLACE into the slot it came from (tcgen05.st): a thread
//                              packs the 64 fp32 columns it has read into the first 32 of them, i.e. the half-tile sits in
//                              columns [0,32) and [64,96) of the slot as a K-major A operand
//   MMA2 (M=128, N=16, K=128)  classifier columns of those hidden features, A operand FROM TENSOR MEMORY, accumulated over
//                              both halves -> TMEM; its commit hands the slot back to MMA1
//   readout                    part[quarter][pixel][16] f32: the head adds the 4 quarter partials (53 MB instead of the
//                              425 MB hidden tensor written and read back)
// The hidden features touch neither HBM nor shared memory, whose space goes to the input-tile ring.
namespace spl {
constexpr int kEpi = 512, kThreads = kEpi + 96;      // warps 0-15 epilogue, warp 16 loader, warp 17 MMA1 issuer (+ TMEM), warp 18 MMA2 issuer
constexpr int kLoadWarp = kEpi / 32, kMmaWarp = kLoadWarp + 1, kMma2Warp = kLoadWarp + 2;
constexpr int WCBYTES = 32 * 256;                    // classifier columns of this quarter: 32 k-chunks x 16 classes x 16 B
constexpr int D2COL = 384;                           // TMEM: MMA1 ring at 0/128/256, 4 stages x 2 logits accumulators at 384..511
enum { A_FULL0 = 0, A_EMPTY0 = 4, D1_FULL0 = 8, D1_EMPTY0 = 11, H_FULL0 = 14, L_FULL0 = 20, L_EMPTY0 = 24, W_FULL = 28 };   // nA <= 4, 4 logits stages
}  // namespace spl

__global__ void __launch_bounds__(spl::kThreads, 1)
spectral_logits_kernel(const __half* __restrict__ x16, int64_t mtiles, int KC, int nA,
                       const __half* __restrict__ w1t, const float* __restrict__ bspe,
                       const __half* __restrict__ wc_spe16, float* __restrict__ part) {
  using namespace spl;
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t wbytes = uint32_t(KC) * 4096, abytes = uint32_t(KC) * 2048;
  const uint32_t S_W = 0, S_A = wbytes, S_WC = S_A + uint32_t(nA) * abytes, S_BIAS = S_WC + WCBYTES, S_BAR = S_BIAS + 1024,
                 S_TMEM = S_BAR + 256;
  const uint32_t bars = sbase + S_BAR;
  const int ntile = blockIdx.x & 3;
  const int64_t mt0 = blockIdx.x >> 2, mstep = gridDim.x >> 2;
  const int64_t my_tiles = mt0 < mtiles ? (mtiles - mt0 + mstep - 1) / mstep : 0;
  const uint32_t U = uint32_t(2 * my_tiles);           // units = (tile, hidden half)

  {
    float* sb = reinterpret_cast<float*>(smem + S_BIAS);
    if (tid < 256) sb[tid] = bspe[ntile * 256 + tid];
  }
  if (tid == 0) {
    mbar_init(bars + 8 * W_FULL, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(bars + 8 * W_FULL, wbytes + WCBYTES);   // this quarter's feat_spe rows and classifier columns
    const unsigned char* gw = reinterpret_cast<const unsigned char*>(w1t) + size_t(ntile) * wbytes;
    for (uint32_t o = 0; o < wbytes; o += 8192) bulk_g2s(sbase + S_W + o, gw + o, wbytes - o < 8192 ? wbytes - o : 8192, bars + 8 * W_FULL);
    bulk_g2s(sbase + S_WC, reinterpret_cast<const unsigned char*>(wc_spe16) + size_t(ntile) * WCBYTES, WCBYTES, bars + 8 * W_FULL);
    for (int i = 0; i < 4; ++i) { mbar_init(bars + 8 * (A_FULL0 + i), 1); mbar_init(bars + 8 * (A_EMPTY0 + i), 1); }
    // two epilogue groups of eight warps take alternate units; a slot goes MMA1 -> (D1_FULL) -> epilogue group ->
    // (H_FULL) -> MMA2 -> (D1_EMPTY) -> MMA1
    for (int i = 0; i < 3; ++i) { mbar_init(bars + 8 * (D1_FULL0 + i), 1); mbar_init(bars + 8 * (D1_EMPTY0 + i), 1); }
    for (int i = 0; i < 3; ++i) mbar_init(bars + 8 * (H_FULL0 + i), kEpi / 2);
    for (int i = 0; i < 4; ++i) { mbar_init(bars + 8 * (L_FULL0 + i), 1); mbar_init(bars + 8 * (L_EMPTY0 + i), 128); }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(sbase + S_TMEM, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + S_TMEM);

  if (warp == kLoadWarp) {
    // ================================================================ loader: one bulk copy per pixel tile
    if (lane == 0) {
      [[maybe_unused]] uint32_t ntr = 0;
      for (uint32_t j = 0; j < uint32_t(my_tiles); ++j) {
        const uint32_t s = j % uint32_t(nA), ph = (j / uint32_t(nA)) & 1;
        mbar_wait(bars + 8 * (A_EMPTY0 + s), ph ^ 1, 91);
        CMLPL_TR(4, ntr, j);
        mbar_arrive_expect_tx(bars + 8 * (A_FULL0 + s), abytes);
        // one bulk copy moves ~6-7 GB/s however large it is: the tile goes as KC/2 copies of 4 KB in flight together
        const __half* src = x16 + (mt0 + int64_t(j) * mstep) * int64_t(KC) * 1024;
        for (int c = 0; c < KC / 2; ++c)
          bulk_g2s(sbase + S_A + s * abytes + c * 4096, src + c * 2048, 4096, bars + 8 * (A_FULL0 + s));
      }
    }
  } else if (warp == kMmaWarp) {
    // ================================================================ MMA1 issuer
    if (tmem != 0) { printf("spectral_logits: unexpected TMEM base %u\n", tmem); __trap(); }
    constexpr uint64_t kHi = (uint64_t(128 >> 4) | (uint64_t(1) << 14)) << 32;
    constexpr uint32_t idesc1 = make_idesc_f16(128, 128);
    // gated by the input tile and by the slot's previous MMA2; runs up to the full ring (3 units) ahead
    mbar_wait(bars + 8 * W_FULL, 0, 90);                 // weights have landed
    [[maybe_unused]] uint32_t ntr = 0;
    for (uint32_t u = 0; u < U; ++u) {
      const uint32_t ti = u >> 1, hh = u & 1, d = u % 3, s = ti % uint32_t(nA);
      if (lane == 0) CMLPL_TR(0, ntr, u * 4);
      if (hh == 0) mbar_wait(bars + 8 * (A_FULL0 + s), (ti / uint32_t(nA)) & 1, 92);
      if (lane == 0) CMLPL_TR(0, ntr, u * 4 + 1);
      mbar_wait(bars + 8 * (D1_EMPTY0 + d), ((u / 3) & 1) ^ 1, 93);
      tc_fence_after();
      if (lane == 0) CMLPL_TR(0, ntr, u * 4 + 2);
      if (elect_one_sync()) {
        uint32_t a_lo = ((sbase + S_A + s * abytes) >> 4) | (uint32_t(2048 >> 4) << 16);
        uint32_t b_lo = ((sbase + S_W + hh * 2048) >> 4) | (uint32_t(4096 >> 4) << 16);
        for (int ks = 0; ks < KC / 2; ++ks) {
          umma_f16(d * 128, kHi | a_lo, kHi | b_lo, idesc1, ks ? 1u : 0u);
          a_lo += 4096 >> 4; b_lo += 8192 >> 4;
        }
        umma_commit(bars + 8 * (D1_FULL0 + d));
        if (hh == 1) umma_commit(bars + 8 * (A_EMPTY0 + s));
      }
      __syncwarp();
      if (lane == 0) CMLPL_TR(0, ntr, u * 4 + 3);
    }
  } else if (warp == kMma2Warp) {
    // ================================================================ MMA2 issuer: classifier columns over the hidden half-tiles (A from TMEM)
    constexpr uint64_t kHi = (uint64_t(128 >> 4) | (uint64_t(1) << 14)) << 32;
    constexpr uint32_t idesc2 = make_idesc_f16(128, 16);
    mbar_wait(bars + 8 * W_FULL, 0, 90);
    [[maybe_unused]] uint32_t ntr = 0;
    for (uint32_t u = 0; u < U; ++u) {
      const uint32_t ti = u >> 1, hh = u & 1, d = u % 3, ls = ti & 3;
      if (lane == 0) CMLPL_TR(1, ntr, u * 4);
      mbar_wait(bars + 8 * (H_FULL0 + d), (u / 3) & 1, 94);
      if (hh == 0) mbar_wait(bars + 8 * (L_EMPTY0 + ls), ((ti >> 2) & 1) ^ 1, 95);
      tc_fence_after();
      if (lane == 0) CMLPL_TR(1, ntr, u * 4 + 2);
      if (elect_one_sync()) {
        uint32_t b_lo = ((sbase + S_WC + hh * 16 * 256) >> 4) | (uint32_t(256 >> 4) << 16);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          // k-step ks = hidden features 16 ks .. 16 ks + 15 of the half-tile = 8 packed columns; two independent
          // accumulators (even / odd k-steps), the readout adds them
          umma_f16_ta(D2COL + ls * 32 + (ks & 1) * 16, d * 128 + (ks >> 2) * 64 + (ks & 3) * 8, kHi | b_lo, idesc2,
                      (hh | (ks >> 1)) ? 1u : 0u);
          b_lo += 512 >> 4;
        }
        umma_commit(bars + 8 * (D1_EMPTY0 + d));
        if (hh == 1) umma_commit(bars + 8 * (L_FULL0 + ls));
      }
      __syncwarp();
      if (lane == 0) CMLPL_TR(1, ntr, u * 4 + 3);
    }
  } else {
    // ================================================================ epilogue (warps 0-15)
    // two groups of eight warps, group 0 takes the first hidden halves (even units), group 1 the second halves; a thread
    // owns one pixel row x 64 columns of the slot
    const int grp = warp >> 3, q = warp & 3, ch = (warp >> 2) & 1, L = q * 32 + lane;
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const float* sb = reinterpret_cast<const float*>(smem + S_BIAS);
    auto readout = [&](uint32_t ti) {                  // partial logits of tile ti (four warps: one lane quarter each)
      const uint32_t ls = ti & 3;
      mbar_wait(bars + 8 * (L_FULL0 + ls), (ti >> 2) & 1, 96);
      tc_fence_after();
      float v[16], w1[16];
      tmem_ld16(lane_addr + D2COL + ls * 32, v);
      tmem_ld16(lane_addr + D2COL + ls * 32 + 16, w1);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bars + 8 * (L_EMPTY0 + ls));
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] += w1[c];
      float4* dst = reinterpret_cast<float4*>(part) + ((int64_t(ntile) * mtiles + (mt0 + int64_t(ti) * mstep)) * 128 + L) * 4;
#pragma unroll
      for (int k = 0; k < 4; ++k) dst[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    };
    [[maybe_unused]] uint32_t ntr = 0;
    const bool tracer = (warp & 7) == 0 && lane == 0;
    for (uint32_t u = uint32_t(grp); u < U; u += 2) {
      const uint32_t ti = u >> 1, hh = u & 1, d = u % 3;
      if (hh == 1 && ch == 0 && ti > 1) readout(ti - 2);      // deferred by two tiles (4 logits stages): its MMA2 has long completed
      if (tracer) CMLPL_TR(2 + grp, ntr, u * 4);
      mbar_wait(bars + 8 * (D1_FULL0 + d), (u / 3) & 1, 97);
      tc_fence_after();
      if (tracer) CMLPL_TR(2 + grp, ntr, u * 4 + 1);
      const float* bb = sb + hh * 128 + ch * 64;
      const uint32_t col0 = lane_addr + d * 128 + ch * 64;
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        float v0[16], v1[16];
        tmem_ld16(col0 + g * 32, v0);
        tmem_ld16(col0 + g * 32 + 16, v1);
        tmem_ld_wait();
        uint32_t h[16];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const __half2 a = __floats2half2_rn(fmaxf(v0[2 * e] + bb[g * 32 + 2 * e], 0.f), fmaxf(v0[2 * e + 1] + bb[g * 32 + 2 * e + 1], 0.f));
          const __half2 b = __floats2half2_rn(fmaxf(v1[2 * e] + bb[g * 32 + 16 + 2 * e], 0.f),
                                              fmaxf(v1[2 * e + 1] + bb[g * 32 + 16 + 2 * e + 1], 0.f));
          h[e] = *reinterpret_cast<const uint32_t*>(&a);
          h[8 + e] = *reinterpret_cast<const uint32_t*>(&b);
        }
        tmem_st16u(col0 + g * 16, h);                    // packed pair c of the chunk -> column c: trails this thread's own reads
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bars + 8 * (H_FULL0 + d));
      if (tracer) CMLPL_TR(2 + grp, ntr, u * 4 + 3);
    }
    if (grp == 1 && ch == 0) {
      if (my_tiles > 1) readout(uint32_t(my_tiles - 2));
      if (my_tiles > 0) readout(uint32_t(my_tiles - 1));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// ------------------------------------------------------------------ head of the dense path: sums + argmax
// logits(p) = bc + sum over the 4 hidden quarters of part[q][p] + the 5 gathered conv partials M[I][r'+2I, c'] of pixel p
// (row maps of pool2_cls_kernel); first index wins ties (hyper_tools.py:426).  One thread per pixel:
// adjacent lanes read adjacent 16-byte quads of the two parity planes of a row.
__global__ void __launch_bounds__(256)
head_sum_kernel(const float* __restrict__ part, int64_t mtiles, const float* __restrict__ lmap, int cols, int PR2, int PC2,
                int64_t n, int C, int ncell /* pooled rows: 5 (w = 20) or 2 (w = 11) */, const float* __restrict__ bc,
                uint8_t* __restrict__ labels, float* __restrict__ logits) {
  const int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (p >= n) return;
  const int nq = (C + 3) >> 2;                         // class quads that hold real classes
  float v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = 0.f;
  const int r = int(p / cols), c0 = int(p - int64_t(r) * cols);
  const int64_t psz = int64_t(PR2) * PC2;
  const float4* base = reinterpret_cast<const float4*>(lmap) + int64_t((r & 1) * 2 + (c0 & 1)) * 20 * psz +
                       int64_t(r >> 1) * PC2 + (c0 >> 1);
#pragma unroll
  for (int I = 0; I < 5; ++I) {
    if (I >= ncell) break;
    const float4* q = base + int64_t(I * 4) * psz + (2 * I) * PC2;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nq) {
        const float4 t = __ldg(q + int64_t(k) * psz);
        v[4 * k] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
      }
    }
  }
#pragma unroll
  for (int qd = 0; qd < 4; ++qd) {
    const float4* s = reinterpret_cast<const float4*>(part) + (int64_t(qd) * mtiles * 128 + p) * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (k < nq) {
        const float4 t = __ldg(s + k);
        v[4 * k] += t.x; v[4 * k + 1] += t.y; v[4 * k + 2] += t.z; v[4 * k + 3] += t.w;
      }
    }
  }
  float best = -INFINITY; int arg = 0;
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    if (c < C) {
      const float z = v[c] + __ldg(bc + c);
      if (logits) logits[p * C + c] = z;
      if (z > best) { best = z; arg = c; }
    }
  }
  labels[p] = uint8_t(arg);
}

}  // namespace cmlpl

using namespace cmlpl;

CMLPL_TRACE_EXPORT(cmlpl_debug_spl_trace)

// shared-memory plan of spectral_logits_kernel for KC input k-chunks: as many input tiles in flight as fit (<= 4)
static bool spectral_logits_plan(int KC, int* nA, size_t* smem) {
  for (int a = 4; a >= 1; --a) {
    const size_t b = size_t(KC) * 4096 + size_t(a) * KC * 2048 + spl::WCBYTES + 1024 + 256 + 64;
    if (b <= 232448) { *nA = a; *smem = b; return true; }
  }
  return false;
}

static int spectral_hidden_impl(const void* x, int dtype /* -1 = preprocessed f32 spectra */, int64_t n, int num_features,
                                int num_classes, int w, const float* mu, const float* inv_sigma, const void* packed,
                                void* x16, void* h16, cmlpl_stream_t stream, bool fused_logits = false) {
  CMLPL_CHECK_ARG(x && packed && x16 && h16, "spectral_hidden_tc: null pointer");
  CMLPL_CHECK_ARG(n > 0 && num_features > 0, "spectral_hidden_tc: bad dims");
  const PackedLayout L = packed_layout(num_features, num_classes, w);
  CMLPL_CHECK_ARG(L.kc_spe_in <= (fused_logits ? 28 : 26), "spectral_hidden_tc: %d bands exceed what the tensor-core tile supports",
                  num_features);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int KC = L.kc_spe_in;
  const int64_t mtiles = (n + 127) / 128, total = mtiles * KC * 128;
  int64_t g = (total + 255) / 256; const int64_t cap = int64_t(sm_count()) * 16; if (g > cap) g = cap;
  if (dtype < 0)
    x16_tile_kernel<<<int(g), 256, 0, s>>>(static_cast<const float*>(x), n, num_features, KC, total, static_cast<__half*>(x16));
  else if (dtype == 0)
    x16_tile_raw_kernel<uint16_t><<<int(g), 256, 0, s>>>(static_cast<const uint16_t*>(x), n, num_features, KC, total, mu, inv_sigma, static_cast<__half*>(x16));
  else
    x16_tile_raw_kernel<float><<<int(g), 256, 0, s>>>(static_cast<const float*>(x), n, num_features, KC, total, mu, inv_sigma, static_cast<__half*>(x16));
  CMLPL_CHECK_LAUNCH("x16_tile");
  if (fused_logits) {
    int nA = 0; size_t fsmem = 0;
    CMLPL_CHECK_ARG(spectral_logits_plan(KC, &nA, &fsmem), "spectral_logits_tc: %d bands do not fit shared memory", num_features);
    CMLPL_CHECK_ARG(num_classes <= 16, "spectral_logits_tc: needs <= 16 classes, got %d", num_classes);
    CMLPL_MAX_DYN_SMEM(spectral_logits_kernel, int(fsmem));
    int fgrid = sm_count() / 4 * 4;
    if (fgrid > mtiles * 4) fgrid = int(mtiles * 4);
    const unsigned char* fpk = static_cast<const unsigned char*>(packed);
    spectral_logits_kernel<<<fgrid, spl::kThreads, fsmem, s>>>(
        static_cast<const __half*>(x16), mtiles, KC, nA, reinterpret_cast<const __half*>(fpk + L.w1s),
        reinterpret_cast<const float*>(fpk + L.bspe),
        reinterpret_cast<const __half*>(fpk + L.wc16 + size_t(L.conv_pos) * 8 * 256), static_cast<float*>(h16));
    CMLPL_CHECK_LAUNCH("spectral_logits");
    return CMLPL_OK;
  }
  const size_t smem = size_t(KC) * 8192 + 1024 + 64 + 64;
  CMLPL_MAX_DYN_SMEM(spectral_hidden_kernel, int(smem));
  int grid = sm_count() / 4 * 4;
  if (grid > mtiles * 4) grid = int(mtiles * 4);
  const unsigned char* pk = static_cast<const unsigned char*>(packed);
  spectral_hidden_kernel<<<grid, kHidThreads, smem, s>>>(static_cast<const __half*>(x16), mtiles, KC,
                                                         reinterpret_cast<const __half*>(pk + L.w1s),
                                                         reinterpret_cast<const float*>(pk + L.bspe),
                                                         static_cast<__half*>(h16));
  CMLPL_CHECK_LAUNCH("spectral_hidden");
  return CMLPL_OK;
}

extern "C" int cmlpl_spectral_hidden_tc(const float* spectra, int64_t n, int num_features, int num_classes, int w,
                                        const void* packed, void* x16, void* h16, cmlpl_stream_t stream) {
  return spectral_hidden_impl(spectra, -1, n, num_features, num_classes, w, nullptr, nullptr, packed, x16, h16, stream);
}

extern "C" int cmlpl_spectral_hidden_raw_tc(const void* raw, int dtype, int64_t n, int num_features, int num_classes, int w,
                                            const float* mu, const float* inv_sigma, const void* packed, void* x16,
                                            void* h16, cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG((dtype == 0 || dtype == 1) && mu && inv_sigma, "spectral_hidden_raw_tc: bad args");
  return spectral_hidden_impl(raw, dtype, n, num_features, num_classes, w, mu, inv_sigma, packed, x16, h16, stream);
}

extern "C" int cmlpl_head_tc(const void* p2t, const void* h16, int64_t n, int num_features, int num_classes, int w,
                             const void* packed, uint8_t* labels, float* logits, cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(p2t && h16 && packed && labels, "head_tc: null pointer");
  CMLPL_CHECK_ARG(n > 0 && num_classes > 0 && num_classes <= 16, "head_tc: needs 1..16 classes, got %d", num_classes);
  const PackedLayout L = packed_layout(num_features, num_classes, w);
  const int kc_conv = L.conv_pos * 8, kc_spe = 128;
  CMLPL_CHECK_ARG(kc_conv % kHeadChunkKc == 0, "head_tc: conv feature width %d not a multiple of 64", kc_conv * 8);
  const size_t wbytes = size_t(kc_conv + kc_spe) * 256;
  const size_t smem = (wbytes + 127) / 128 * 128 + size_t(kHeadStages) * kHeadChunkBytes + 256 + 64;
  CMLPL_CHECK_ARG(smem <= 227 * 1024, "head_tc: shared memory %zu too large", smem);
  CMLPL_MAX_DYN_SMEM(head_kernel, int(smem));
  const int64_t mtiles = (n + 127) / 128;
  int grid = sm_count(); if (grid > mtiles) grid = int(mtiles);
  const unsigned char* pk = static_cast<const unsigned char*>(packed);
  head_kernel<<<grid, kHeadThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __half*>(p2t), kc_conv, static_cast<const __half*>(h16), kc_spe, mtiles, n, num_classes,
      reinterpret_cast<const __half*>(pk + L.wc16), reinterpret_cast<const float*>(pk + L.bc), labels, logits);
  CMLPL_CHECK_LAUNCH("head");
  return CMLPL_OK;
}

// Fused spectral branch of the dense path: part f32 [4 hidden quarters][ceil(n/128)*128][16] partial spectral logits
// (spectral_logits_kernel; the 1024 hidden features stay in shared memory).  x16 is scratch for the fp16 input tiles.
extern "C" int cmlpl_spectral_logits_tc(const float* spectra, int64_t n, int num_features, int num_classes, int w,
                                        const void* packed, void* x16, float* part, cmlpl_stream_t stream) {
  return spectral_hidden_impl(spectra, -1, n, num_features, num_classes, w, nullptr, nullptr, packed, x16, part, stream, true);
}

extern "C" int cmlpl_spectral_logits_raw_tc(const void* raw, int dtype, int64_t n, int num_features, int num_classes, int w,
                                            const float* mu, const float* inv_sigma, const void* packed, void* x16,
                                            float* part, cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG((dtype == 0 || dtype == 1) && mu && inv_sigma, "spectral_logits_raw_tc: bad args");
  return spectral_hidden_impl(raw, dtype, n, num_features, num_classes, w, mu, inv_sigma, packed, x16, part, stream, true);
}

// Head of the dense path without a GEMM: quarter partials of cmlpl_spectral_logits_tc + the 5 gathered conv partials
// per pixel from lmap (cmlpl_pool2_cls_f16) + bias, argmax.  Pixels are the band's raster order.
extern "C" int cmlpl_head_sum_lmap(const float* part, const float* lmap, int cols, int band_rows, int num_features,
                                   int num_classes, int w, const void* packed, uint8_t* labels, float* logits,
                                   cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(part && lmap && packed && labels, "head_sum_lmap: null pointer");
  CMLPL_CHECK_ARG((w == 20 || w == 11) && cols > 0 && band_rows > 0, "head_sum_lmap: bad dims (w must be 20 or 11)");
  CMLPL_CHECK_ARG(num_classes > 0 && num_classes <= 16, "head_sum_lmap: needs 1..16 classes, got %d", num_classes);
  const PackedLayout L = packed_layout(num_features, num_classes, w);
  const int64_t n = int64_t(band_rows) * cols, mtiles = (n + 127) / 128;
  const unsigned char* pk = static_cast<const unsigned char*>(packed);
  head_sum_kernel<<<unsigned((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      part, mtiles, lmap, cols, (band_rows + w) / 2, (cols + w) / 2, n, num_classes, w == 20 ? 5 : 2,
      reinterpret_cast<const float*>(pk + L.bc), labels, logits);
  CMLPL_CHECK_LAUNCH("head_sum");
  return CMLPL_OK;
}
