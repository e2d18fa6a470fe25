// Full-scene inference (tools/hyper_tools.py:416-437 test_whole) without materialised
// patches.  Stages (one launch each, all on the caller's stream):
//   1. conv0_map     conv0 (1x1, models.py:102,132) once per scene pixel -> mirrored,
//                    halo-padded fp16 map F0pad[8 chunks][(rows+w-1)][(cols+w-1)][8 channels]; a pixel's patch
//                    is then a plain w x w window of this map (hyper_tools.py:35-55,226-243)
//   2. spectral_head relu(feat_spe(x)) (models.py:142-143) and its classifier columns
//   3. conv1_scene   conv1 + residual + ReLU once per scene position in 9 patch-border classes and the
//                    pooled maps as parity planes (conv1_scene_sm100.cu); conv2_scene: conv2 + residual +
//                    ReLU once per position in 25 classes, border classes pooled in its epilogue; pool2_cls: rest of
//                    the avg-pool + conv columns of the classifier -> 25 class-partial maps (conv2_scene_sm100.cu)
//   4. head          spectral columns of the classifier (models.py:150) + the pixel's 25 gathered conv
//                    partials + bias, argmax (hyper_tools.py:426, first index wins ties)
//   (> 16 classes / > 224 bands: the all-per-pixel patch_cnn_sm100.cu kernel + CUDA-core classify instead)
#include <mutex>

#include "common.cuh"
#include "gemm_core.cuh"

namespace cmlpl {

// ------------------------------------------------------------------ conv0 map
// tcgen05 kernel of conv0_sm100.cu: plain fp16 operands for the 60-channel PCA cube (z-scored, well conditioned); for
// the RAW cube both operands are split into fp16 hi + lo parts (3 MMAs per K-step), because folding the PCA projection
// into conv0 makes the contraction ill-conditioned in fp16 -- the noise components of the PCA are differences of band
// values ~10x larger than the result (measured: with single fp16 operands the B = 200 raw parity test misses the 1e-3
// logit bar).  PCA projection, both z-scores and conv0 are per-pixel affine maps, so they fold into one
// F0 = Wf . (x - mu) + bf with Wf = W0 . (U/s)^T [64 x B].  Output: mirrored halo baked in, chunk-planar fp16
// [8 chunks][prow_n][pcol_n][8].
template <typename T, bool kVec4, bool kSplit>
int launch_conv0_tc(const T* in, int K, int scene_rows, int cols, int slab_row0, int w, int band_row0, int prow_n, int pcol_n,
                    const float* wt, const float* bias, const float* mu, const float* inv_sigma, __half* f0pad, cudaStream_t s);

// ------------------------------------------------------------------ classifier + argmax
// one warp per pixel; lanes stride over the K = P*64 pooled features in 8-half chunks.
template <int CMAX>
__global__ void __launch_bounds__(256)
classify_kernel(const __half* __restrict__ p2, const float* __restrict__ spe, const float* __restrict__ part,
                int64_t part_stride, int64_t n, int K, int C,
                const float* __restrict__ wc, const float* __restrict__ bc,
                uint8_t* __restrict__ labels, float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = blockIdx.x * int64_t(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = int64_t(gridDim.x) * (blockDim.x >> 5);
  const int chunks = K >> 3;
  for (int64_t p = warp; p < n; p += nwarps) {
    float acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) acc[c] = 0.f;
    const uint4* src = reinterpret_cast<const uint4*>(p2 + p * K);
    for (int ch = lane; ch < chunks; ch += 32) {
      const uint4 raw = __ldcs(src + ch);
      const __half2* h = reinterpret_cast<const __half2*>(&raw);
      float x[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = __half22float2(h[j]); x[2 * j] = f.x; x[2 * j + 1] = f.y; }
#pragma unroll
      for (int c = 0; c < CMAX; ++c) {
        if (c < C) {
          const float4 wa = __ldg(reinterpret_cast<const float4*>(wc + int64_t(c) * K + ch * 8));
          const float4 wb = __ldg(reinterpret_cast<const float4*>(wc + int64_t(c) * K + ch * 8 + 4));
          float s = acc[c];
          s = fmaf(x[0], wa.x, s); s = fmaf(x[1], wa.y, s); s = fmaf(x[2], wa.z, s); s = fmaf(x[3], wa.w, s);
          s = fmaf(x[4], wb.x, s); s = fmaf(x[5], wb.y, s); s = fmaf(x[6], wb.z, s); s = fmaf(x[7], wb.w, s);
          acc[c] = s;
        }
      }
    }
    float best = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
      if (c < C) {
        float v = warp_sum(acc[c]);
        v += (spe ? spe[p * C + c] : 0.f) + bc[c];
        if (part)      // spectral_logits_kernel: four hidden-quarter partials f32 [4][part_stride][16]
          v += (part[p * 16 + c] + part[(part_stride + p) * 16 + c]) +
               (part[(2 * part_stride + p) * 16 + c] + part[(3 * part_stride + p) * 16 + c]);
        if (logits && lane == 0) logits[p * C + c] = v;
        if (v > best) { best = v; arg = c; }   // strict '>' : first index wins ties
      }
    }
    if (lane == 0) labels[p] = uint8_t(arg);
  }
}

// argmax over dense fp32 logits (batch-mode test_whole path, hyper_tools.py:426)
__global__ void argmax_kernel(const float* __restrict__ logits, int64_t n, int C, uint8_t* __restrict__ labels) {
  for (int64_t p = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; p < n; p += int64_t(gridDim.x) * blockDim.x) {
    float best = -INFINITY; int arg = 0;
    for (int c = 0; c < C; ++c) { const float v = logits[p * C + c]; if (v > best) { best = v; arg = c; } }
    labels[p] = uint8_t(arg);
  }
}

// Tensor-core head (head_sm100.cu) applies for <= 16 classes and <= 208 bands; otherwise the
// CUDA-core spectral_head + classify kernels above run (same C ABI, same results contract).
static bool use_tc_head(int B, int C) { return C <= 16 && ((B + 15) / 16) * 2 <= 28; }   // <= 224 bands (spectral_logits_kernel)

// 1 (default): the exact-compute-sharing kernels where they apply; 0: the per-pixel tensor-core kernels everywhere (the
// independent implementation the tests compare against).  Per host thread; the workspace layout follows the mode.
static thread_local int g_scene_path_mode = 1;

struct SceneWs {
  size_t f0pad, p2, spe, hidden, x16, h16, g, pm, yq, lmap, total;
  int64_t chunk;
  bool tc, dense;
};
// tensor-core path (<= 16 classes, <= 224 bands): conv0 map, spectral tiles, conv1 variants (fp32 scratch), pooled
// parity planes, the 25 conv2 variants, the 25 class-partial maps.  CUDA-core path: conv0 map, per-pixel pooled
// features, spectral logits, hidden chunk.
static SceneWs scene_ws(int band_rows, int cols, int B, int C, int w) {
  SceneWs s;
  const int64_t n = int64_t(band_rows) * cols;
  const int64_t mtiles = (n + 127) / 128;
  const int P = ((w / 2) / 2) * ((w / 2) / 2);
  s.tc = use_tc_head(B, C);
  // the exact-compute-sharing kernels: w = 20 (9 / 25 border classes) and w = 11, whose 4 / 9 classes and 2x2 pooled cells
  // are a subset of them (conv1 rows 0..9 and conv2 rows 0..3 of an 11-window are top / mid class only)
  s.dense = s.tc && (w == 20 || w == 11) && g_scene_path_mode == 1;
  size_t o = 0;
  s.f0pad = o; o = align256(o + size_t(band_rows + w - 1) * (cols + w - 1) * 64 * 2);
  s.p2 = s.spe = s.hidden = s.x16 = s.h16 = s.g = s.pm = s.yq = s.lmap = 0;
  s.chunk = n < 16384 ? n : 16384;
  if (s.tc) {
    s.x16 = o; o = align256(o + size_t(mtiles) * (((B + 15) / 16) * 2) * 2048);
    s.h16 = o; o = align256(o + size_t(4) * mtiles * 128 * 64);   // partial spectral logits f32 [4 quarters][mtiles*128][16]
  }
  if (s.dense) {
    const size_t qpos = size_t(4) * ((band_rows + w) / 2) * ((cols + w) / 2);     // positions of the 4 parity planes
    s.g = o;                                               // (fp32 conv1 variants: not materialised any more)
    s.pm = o; o = align256(o + qpos * 9 * 64 * 2);         // pooled variants, f16 parity planes [9][4][8][PR2][PC2][8]
    s.yq = o; o = align256(o + qpos * 9 * 64 * 2);         // half-pooled conv2 maps, f16 [9][4][8][PR2][PC2][8]
    s.lmap = o; o = align256(o + qpos * 5 * 16 * 4);       // class partials per pooled row, f32 [4][5][4][PR2][PC2][4]
  } else {
    s.p2 = o; o = align256(o + size_t(n) * P * 64 * 2);     // per-pixel pooled conv features (patch_cnn_sm100.cu)
    if (!s.tc) {
      s.spe = o; o = align256(o + size_t(n) * C * 4);
      s.hidden = o; o = align256(o + size_t(s.chunk) * 1024 * 4);
    }
  }
  s.total = o;
  return s;
}

}  // namespace cmlpl

using namespace cmlpl;

extern "C" int cmlpl_set_scene_path_mode(int mode) {
  CMLPL_CHECK_ARG(mode == 0 || mode == 1, "set_scene_path_mode: 0 (per-pixel kernels) or 1 (exact compute sharing, default)");
  g_scene_path_mode = mode;
  return CMLPL_OK;
}

extern "C" size_t cmlpl_scene_workspace_bytes(int band_rows, int cols, int num_features, int num_classes, int w) {
  if (band_rows <= 0 || cols <= 0 || num_classes <= 0 || num_features <= 0 || w < 4) return 0;
  return scene_ws(band_rows, cols, num_features, num_classes, w).total;
}

extern "C" int cmlpl_scene_workspace_layout(int band_rows, int cols, int num_features, int num_classes, int w,
                                            size_t* offsets) {
  CMLPL_CHECK_ARG(offsets && band_rows > 0 && cols > 0 && num_classes > 0 && num_features > 0 && w >= 4,
                  "scene_workspace_layout: bad args");
  const SceneWs s = scene_ws(band_rows, cols, num_features, num_classes, w);
  const size_t v[12] = {s.f0pad, s.x16, s.h16, s.g, s.pm, s.yq, s.lmap, s.p2, s.spe, s.hidden, s.total, size_t(s.dense)};
  for (int i = 0; i < 12; ++i) offsets[i] = v[i];
  return CMLPL_OK;
}

// The spectral branch (fp16 tiles of the spectra + the two spectral GEMMs) does not depend on the conv tower until the
// head: it runs on a side stream forked off the caller's, so that its conversion kernel shares the SMs with conv0 and
// its tail with the head of conv1_pool; the head waits for both.  One side stream and two events per device; host threads
// that drive the same device from several streams take turns (the launches are microseconds).
struct SideStream { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; std::mutex mu; };   // mu: one fork/join sequence at a time
static SideStream* side_stream() {
  static SideStream per_dev[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& x = per_dev[dev];
  std::lock_guard<std::mutex> lock(x.mu);
  if (!x.join) {                                             // `join` is created last: set only when all three exist
    if (!x.s && cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) != cudaSuccess) { x.s = nullptr; return nullptr; }
    if (!x.fork && cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) != cudaSuccess) { x.fork = nullptr; return nullptr; }
    if (cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) != cudaSuccess) { x.join = nullptr; return nullptr; }
  }
  return &x;
}

// conv1 (9 border classes) -> pooled parity planes -> conv2 (25 classes) -> pool + conv classifier columns ->
// spectral classifier columns + gathered conv partials + argmax: everything after conv0 / the spectral GEMM
static int dense_tail(unsigned char* wsb, const SceneWs& ws, int cols, int w, int band_rows, int num_features,
                      int num_classes, const void* packed, uint8_t* labels, float* logits, cmlpl_stream_t stream,
                      cudaEvent_t spectral_done = nullptr) {
  int rc = cmlpl_conv1_pool_planes_f16(wsb + ws.f0pad, cols, w, band_rows, packed, wsb + ws.pm, stream);
  if (rc != CMLPL_OK) return rc;
  rc = cmlpl_conv2_scene_f16(wsb + ws.pm, cols, w, band_rows, packed, wsb + ws.yq, stream);
  if (rc != CMLPL_OK) return rc;
  rc = cmlpl_pool2_cls_f16(wsb + ws.yq, cols, w, band_rows, num_features, num_classes, packed,
                           reinterpret_cast<float*>(wsb + ws.lmap), stream);
  if (rc != CMLPL_OK) return rc;
  if (spectral_done) CMLPL_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), spectral_done, 0));
  return cmlpl_head_sum_lmap(reinterpret_cast<const float*>(wsb + ws.h16), reinterpret_cast<const float*>(wsb + ws.lmap), cols,
                             band_rows, num_features, num_classes, w, packed, labels, logits, stream);
}

extern "C" int cmlpl_conv0_map_f16(const float* cube, int scene_rows, int cols, int slab_row0, int slab_rows, int w,
                                   int band_row0, int band_rows, const void* packed, void* f0pad,
                                   cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(cube && packed && f0pad, "conv0_map: null pointer");
  CMLPL_CHECK_ARG(scene_rows > 0 && cols > 0 && w >= 2 && band_rows > 0, "conv0_map: bad dims");
  CMLPL_CHECK_ARG(w / 2 <= scene_rows && w / 2 <= cols, "conv0_map: window larger than the scene");
  CMLPL_CHECK_ARG(band_row0 >= 0 && band_row0 + band_rows <= scene_rows, "conv0_map: band outside the scene");
  CMLPL_CHECK_ARG(reinterpret_cast<uintptr_t>(cube) % 16 == 0, "conv0_map: cube must be 16-byte aligned");
  // the slab must hold every (mirrored) source row of the band's halo
  const int a = band_row0 + window_lo(w), b = band_row0 + band_rows - 1 + window_lo(w) + w - 1;
  int rmin = a < 0 ? 0 : a, rmax = b >= scene_rows ? scene_rows - 1 : b;
  if (a < 0 && -a - 1 > rmax) rmax = -a - 1;                          // rows -1..a reflect to 0..-a-1
  if (b >= scene_rows && 2 * scene_rows - 1 - b < rmin) rmin = 2 * scene_rows - 1 - b;
  CMLPL_CHECK_ARG(rmin >= slab_row0 && rmax < slab_row0 + slab_rows,
                  "conv0_map: slab rows [%d,%d) do not cover the band's halo [%d,%d]", slab_row0,
                  slab_row0 + slab_rows, rmin, rmax);
  const PackedLayout L = packed_layout(1, 1, w);  // w0/b0 offsets do not depend on B, C
  const unsigned char* pk = static_cast<const unsigned char*>(packed);
  const int prow_n = band_rows + w - 1, pcol_n = cols + w - 1;
  return launch_conv0_tc<float, true, false>(cube, 60, scene_rows, cols, slab_row0, w, band_row0, prow_n, pcol_n,
                                      reinterpret_cast<const float*>(pk + L.w0), reinterpret_cast<const float*>(pk + L.b0), nullptr,
                                      nullptr, static_cast<__half*>(f0pad), static_cast<cudaStream_t>(stream));
}

extern "C" int cmlpl_spectral_head_f32(const float* spectra, int64_t n, int num_features, int num_classes, int w,
                                       const void* packed, float* hidden, int64_t chunk, float* spe_logits,
                                       cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(spectra && packed && hidden && spe_logits, "spectral_head: null pointer");
  CMLPL_CHECK_ARG(n >= 0 && chunk > 0 && num_features > 0 && num_classes > 0, "spectral_head: bad dims");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned char* pk = static_cast<const unsigned char*>(packed);
  const PackedLayout L = packed_layout(num_features, num_classes, w);
  const float* wspe = reinterpret_cast<const float*>(pk + L.wspe);
  const float* bspe = reinterpret_cast<const float*>(pk + L.bspe);
  const float* wc_spe = reinterpret_cast<const float*>(pk + L.wc_spe);
  for (int64_t s0 = 0; s0 < n; s0 += chunk) {
    const int m = int((n - s0) < chunk ? (n - s0) : chunk);
    // hidden = relu(X . Wspe^T + bspe)         (models.py:142-143)
    int rc = launch_gemm(m, 1024, num_features, 1,
                         StridedA{spectra + s0 * num_features, num_features, 1},
                         StridedB{wspe, 1, num_features},
                         StridedC{hidden, 1024, 1, bspe, 1.f, 0.f, 1}, s, "spectral_hidden");
    if (rc != CMLPL_OK) return rc;
    // spe_logits = hidden . Wc_spe^T           (spectral columns of models.py:150)
    rc = launch_gemm(m, num_classes, 1024, 1, StridedA{hidden, 1024, 1}, StridedB{wc_spe, 1, 1024},
                     StridedC{spe_logits + s0 * num_classes, num_classes, 1, nullptr, 1.f, 0.f, 0}, s,
                     "spectral_logits");
    if (rc != CMLPL_OK) return rc;
  }
  return CMLPL_OK;
}

static int classify_launch(const void* p2, const float* spe_logits, const float* part, int64_t part_stride, int64_t n,
                           int num_features, int num_classes, int w, const void* packed, uint8_t* labels, float* logits,
                           cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(p2 && packed && labels, "classify: null pointer");
  CMLPL_CHECK_ARG(n >= 0 && num_classes > 0 && num_classes <= 32 && w >= 4, "classify: bad dims (C=%d w=%d)",
                  num_classes, w);
  if (n == 0) return CMLPL_OK;
  const unsigned char* pk = static_cast<const unsigned char*>(packed);
  const PackedLayout L = packed_layout(num_features, num_classes, w);
  const int K = L.conv_pos * 64;
  const float* wc = reinterpret_cast<const float*>(pk + L.wc_conv);
  const float* bc = reinterpret_cast<const float*>(pk + L.bc);
  int64_t grid = (n + 7) / 8;
  const int64_t cap = int64_t(sm_count()) * 8;
  if (grid > cap) grid = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (num_classes <= 16)
    classify_kernel<16><<<int(grid), 256, 0, s>>>(static_cast<const __half*>(p2), spe_logits, part, part_stride, n, K,
                                                  num_classes, wc, bc, labels, logits);
  else
    classify_kernel<32><<<int(grid), 256, 0, s>>>(static_cast<const __half*>(p2), spe_logits, part, part_stride, n, K,
                                                  num_classes, wc, bc, labels, logits);
  CMLPL_CHECK_LAUNCH("classify");
  return CMLPL_OK;
}

extern "C" int cmlpl_classify_f16(const void* p2, const float* spe_logits, int64_t n, int num_features,
                                  int num_classes, int w, const void* packed, uint8_t* labels, float* logits,
                                  cmlpl_stream_t stream) {
  return classify_launch(p2, spe_logits, nullptr, 0, n, num_features, num_classes, w, packed, labels, logits, stream);
}

extern "C" int cmlpl_argmax_u8(const float* logits, int64_t n, int num_classes, uint8_t* labels, cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(logits && labels && n >= 0 && num_classes > 0 && num_classes <= 255, "argmax: bad args");
  if (n == 0) return CMLPL_OK;
  int64_t grid = (n + 255) / 256;
  const int64_t cap = int64_t(sm_count()) * 8;
  if (grid > cap) grid = cap;
  argmax_kernel<<<int(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, n, num_classes, labels);
  CMLPL_CHECK_LAUNCH("argmax");
  return CMLPL_OK;
}

extern "C" int cmlpl_scene_infer(const float* cube, int scene_rows, int cols, int slab_row0, int slab_rows,
                                 const float* spectra, int num_features, int num_classes, int w, int band_row0,
                                 int band_rows, const void* packed, void* workspace, size_t workspace_bytes,
                                 uint8_t* labels, float* logits, cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(cube && spectra && packed && workspace && labels, "scene_infer: null pointer");
  // w = 20: the reference's BaseNet2 (classifier hard-wired to 2624 inputs, tools/models.py:127).  w = 11: the odd-window
  // variant BASELINE configs[4] names (ExtractPatches_for_base windows, hyper_tools.py:300-317; pooled 5 -> 2, classifier
  // over 64*2*2 + 1024 = 1280 inputs) -- a documented extension; its border classes and pooled cells are a subset of the
  // w = 20 ones, so the same scene-level kernels run; patch_cnn_kernel<11> per pixel behind path mode 0.
  CMLPL_CHECK_ARG(w == 20 || w == 11, "scene_infer: w=%d unsupported (20, or 11 for the odd-window variant)", w);
  CMLPL_CHECK_ARG(band_rows > 0 && cols > 0 && num_classes > 0 && num_classes <= 32, "scene_infer: bad dims");
  const SceneWs ws = scene_ws(band_rows, cols, num_features, num_classes, w);
  CMLPL_CHECK_ARG(workspace_bytes >= ws.total, "scene_infer: workspace %zu < required %zu", workspace_bytes, ws.total);
  CMLPL_CHECK_ARG(reinterpret_cast<uintptr_t>(workspace) % 256 == 0, "scene_infer: workspace must be 256-byte aligned");
  unsigned char* wsb = static_cast<unsigned char*>(workspace);
  const int64_t n = int64_t(band_rows) * cols;
  int rc = CMLPL_OK;
  if (ws.tc && ws.dense) {
    SideStream* sd = side_stream();
    CMLPL_CHECK_ARG(sd, "scene_infer: cannot create the side stream");
    std::lock_guard<std::mutex> lock(sd->mu);
    cudaStream_t ms = static_cast<cudaStream_t>(stream);
    CMLPL_CUDA(cudaEventRecord(sd->fork, ms));
    CMLPL_CUDA(cudaStreamWaitEvent(sd->s, sd->fork, 0));
    rc = cmlpl_spectral_logits_tc(spectra, n, num_features, num_classes, w, packed, wsb + ws.x16,
                                  reinterpret_cast<float*>(wsb + ws.h16), sd->s);
    if (rc != CMLPL_OK) return rc;
    CMLPL_CUDA(cudaEventRecord(sd->join, sd->s));
    rc = cmlpl_conv0_map_f16(cube, scene_rows, cols, slab_row0, slab_rows, w, band_row0, band_rows, packed, wsb + ws.f0pad, stream);
    if (rc != CMLPL_OK) return rc;
    return dense_tail(wsb, ws, cols, w, band_rows, num_features, num_classes, packed, labels, logits, stream, sd->join);
  }
  rc = cmlpl_conv0_map_f16(cube, scene_rows, cols, slab_row0, slab_rows, w, band_row0, band_rows, packed,
                           wsb + ws.f0pad, stream);
  if (rc != CMLPL_OK) return rc;
  if (ws.tc) {
    rc = cmlpl_spectral_logits_tc(spectra, n, num_features, num_classes, w, packed, wsb + ws.x16,
                                  reinterpret_cast<float*>(wsb + ws.h16), stream);
    if (rc != CMLPL_OK) return rc;
    rc = cmlpl_patch_cnn_f16(wsb + ws.f0pad, cols, w, band_rows, packed, wsb + ws.p2, stream);
    if (rc != CMLPL_OK) return rc;
    return classify_launch(wsb + ws.p2, nullptr, reinterpret_cast<const float*>(wsb + ws.h16), ((n + 127) / 128) * 128, n,
                           num_features, num_classes, w, packed, labels, logits, stream);
  }
  rc = cmlpl_spectral_head_f32(spectra, n, num_features, num_classes, w, packed,
                               reinterpret_cast<float*>(wsb + ws.hidden), ws.chunk,
                               reinterpret_cast<float*>(wsb + ws.spe), stream);
  if (rc != CMLPL_OK) return rc;
  rc = cmlpl_patch_cnn_f16(wsb + ws.f0pad, cols, w, band_rows, packed, wsb + ws.p2, stream);
  if (rc != CMLPL_OK) return rc;
  return cmlpl_classify_f16(wsb + ws.p2, reinterpret_cast<const float*>(wsb + ws.spe), n, num_features, num_classes,
                            w, packed, labels, logits, stream);
}


// ---------------------------------------------------------------------------------------------
// Scene inference from the RAW cube (uint16 / float32): the preprocessing of
// tools/hyper_tools.py:285-292 is folded into the first kernels (no PCA cube / z-scored spectra in HBM).
//   raw       [slab_rows*cols, B]  raw scene rows slab_row0..  (dtype 0 = uint16, 1 = float32)
//   wf f32 [B][64], bf f32 [64]    conv0 folded with the PCA projection and both z-scores
//   mu, inv_sigma f32 [B]          band means / 1/std (z-score of the spectral branch)
extern "C" int cmlpl_scene_infer_raw(const void* raw, int dtype, int scene_rows, int cols, int slab_row0, int slab_rows,
                                     int num_features, int num_classes, int w, int band_row0, int band_rows,
                                     const float* wf, const float* bf, const float* mu, const float* inv_sigma,
                                     const void* packed, void* workspace, size_t workspace_bytes, uint8_t* labels,
                                     float* logits, cmlpl_stream_t stream) {
  CMLPL_CHECK_ARG(raw && wf && bf && mu && inv_sigma && packed && workspace && labels, "scene_infer_raw: null pointer");
  CMLPL_CHECK_ARG(dtype == 0 || dtype == 1, "scene_infer_raw: dtype must be 0 (uint16) or 1 (float32)");
  CMLPL_CHECK_ARG(w == 20 || w == 11, "scene_infer_raw: w=%d unsupported (20, or 11 for the odd-window variant)", w);
  CMLPL_CHECK_ARG(band_rows > 0 && cols > 0 && num_classes > 0 && num_features > 0, "scene_infer_raw: bad dims");
  CMLPL_CHECK_ARG(use_tc_head(num_features, num_classes), "scene_infer_raw: needs <= 16 classes and <= 224 bands");
  CMLPL_CHECK_ARG(w / 2 <= scene_rows && w / 2 <= cols, "scene_infer_raw: window larger than the scene");
  CMLPL_CHECK_ARG(band_row0 >= 0 && band_row0 + band_rows <= scene_rows, "scene_infer_raw: band outside the scene");
  const int a = band_row0 + window_lo(w), b = band_row0 + band_rows - 1 + window_lo(w) + w - 1;
  int rmin = a < 0 ? 0 : a, rmax = b >= scene_rows ? scene_rows - 1 : b;
  if (a < 0 && -a - 1 > rmax) rmax = -a - 1;
  if (b >= scene_rows && 2 * scene_rows - 1 - b < rmin) rmin = 2 * scene_rows - 1 - b;
  CMLPL_CHECK_ARG(rmin >= slab_row0 && rmax < slab_row0 + slab_rows && band_row0 >= slab_row0,
                  "scene_infer_raw: slab rows [%d,%d) do not cover the band's halo [%d,%d]", slab_row0,
                  slab_row0 + slab_rows, rmin, rmax);
  const SceneWs ws = scene_ws(band_rows, cols, num_features, num_classes, w);
  CMLPL_CHECK_ARG(workspace_bytes >= ws.total, "scene_infer_raw: workspace %zu < required %zu", workspace_bytes, ws.total);
  unsigned char* wsb = static_cast<unsigned char*>(workspace);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t n = int64_t(band_rows) * cols;
  const int prow_n = band_rows + w - 1, pcol_n = cols + w - 1;
  auto conv0_raw = [&]() {
    __half* f0 = reinterpret_cast<__half*>(wsb + ws.f0pad);
    return dtype == 0
        ? launch_conv0_tc<uint16_t, false, true>(static_cast<const uint16_t*>(raw), num_features, scene_rows, cols, slab_row0, w,
                                                 band_row0, prow_n, pcol_n, wf, bf, mu, inv_sigma, f0, s)
        : launch_conv0_tc<float, false, true>(static_cast<const float*>(raw), num_features, scene_rows, cols, slab_row0, w,
                                              band_row0, prow_n, pcol_n, wf, bf, mu, inv_sigma, f0, s);
  };
  const size_t esz = dtype == 0 ? 2 : 4;
  const void* band_raw = static_cast<const unsigned char*>(raw) + size_t(band_row0 - slab_row0) * cols * num_features * esz;
  int rc = CMLPL_OK;
  if (ws.dense) {                                          // spectral branch beside the conv tower (see side_stream())
    SideStream* sd = side_stream();
    CMLPL_CHECK_ARG(sd, "scene_infer_raw: cannot create the side stream");
    std::lock_guard<std::mutex> lock(sd->mu);
    CMLPL_CUDA(cudaEventRecord(sd->fork, s));
    CMLPL_CUDA(cudaStreamWaitEvent(sd->s, sd->fork, 0));
    rc = cmlpl_spectral_logits_raw_tc(band_raw, dtype, n, num_features, num_classes, w, mu, inv_sigma, packed,
                                      wsb + ws.x16, reinterpret_cast<float*>(wsb + ws.h16), sd->s);
    if (rc != CMLPL_OK) return rc;
    CMLPL_CUDA(cudaEventRecord(sd->join, sd->s));
    rc = conv0_raw();
    if (rc != CMLPL_OK) return rc;
    return dense_tail(wsb, ws, cols, w, band_rows, num_features, num_classes, packed, labels, logits, stream, sd->join);
  }
  rc = conv0_raw();
  if (rc != CMLPL_OK) return rc;
  rc = cmlpl_spectral_logits_raw_tc(band_raw, dtype, n, num_features, num_classes, w, mu, inv_sigma, packed,
                                    wsb + ws.x16, reinterpret_cast<float*>(wsb + ws.h16), stream);
  if (rc != CMLPL_OK) return rc;
  rc = cmlpl_patch_cnn_f16(wsb + ws.f0pad, cols, w, band_rows, packed, wsb + ws.p2, stream);
  if (rc != CMLPL_OK) return rc;
  return classify_launch(wsb + ws.p2, nullptr, reinterpret_cast<const float*>(wsb + ws.h16), ((n + 127) / 128) * 128, n,
                         num_features, num_classes, w, packed, labels, logits, stream);
}
