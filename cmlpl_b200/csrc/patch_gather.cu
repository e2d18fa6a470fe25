// Materialising patch gather.
//   tools/hyper_tools.py:35-55   MirrowCut            (symmetric mirror padding)
//   tools/hyper_tools.py:226-243 ExtractPatches       (even w, window r-w/2 .. r+w/2-1)
//   tools/hyper_tools.py:300-317 ExtractPatches_for_base (odd w, window r-(w-1)/2 .. r+(w-1)/2)
//   train.py:157                 x + randn*noise      (optional fused noise add)
//
// HBM-bound: 4*F*w*w bytes written per pixel (96 000 B at F=60, w=20); the 400-fold
// overlapping window reads are served by L2.  One CTA stages one pixel's window
// (channels-last in the cube) into shared memory with coalesced float4 reads, then
// streams it out channel-major ([F][w][w]) with coalesced float4 stores -- the
// transposition happens in shared memory (row stride w*w+1 words to spread banks).
#include "common.cuh"

namespace cmlpl {

template <bool VEC>
__global__ void __launch_bounds__(256, 2)
patch_gather_kernel(const float* __restrict__ cube, int scene_rows, int cols, int feat,
                    int slab_row0, int w, const int64_t* __restrict__ idx, int64_t first,
                    int64_t n, const float* __restrict__ noise, float noise_scale,
                    float* __restrict__ out) {
  extern __shared__ float tile[];  // [feat][w*w + 1]
  const int ww = w * w;
  const int stride = ww + 1;
  const int lo = window_lo(w);
  const int tid = threadIdx.x;

  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const int64_t pix = idx ? idx[p] : first + p;
    const int r = int(pix / cols), c = int(pix % cols);

    if (VEC) {
      const int f4n = feat >> 2;
      const int total = ww * f4n;
      // latency-bound on L2: keep U independent 16-byte loads in flight per thread before the
      // (dependent) transposing stores
      constexpr int U = 8;
      for (int base = tid; base < total; base += blockDim.x * U) {
        float4 v[U];
        int dst[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int i = base + u * blockDim.x;
          dst[u] = -1;
          if (i < total) {
            const int pos = i / f4n, f4 = i - pos * f4n;
            const int y = pos / w, x = pos - y * w;
            const int sr = mirror_index(r + lo + y, scene_rows) - slab_row0;
            const int sc = mirror_index(c + lo + x, cols);
            v[u] = __ldg(reinterpret_cast<const float4*>(cube + (int64_t(sr) * cols + sc) * feat) + f4);
            dst[u] = (f4 * 4) * stride + pos;
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (dst[u] >= 0) {
            float* t = tile + dst[u];
            t[0] = v[u].x; t[stride] = v[u].y; t[2 * stride] = v[u].z; t[3 * stride] = v[u].w;
          }
        }
      }
      __syncthreads();
      float4* o4 = reinterpret_cast<float4*>(out + p * int64_t(feat) * ww);
      const float4* n4 = noise ? reinterpret_cast<const float4*>(noise + p * int64_t(feat) * ww) : nullptr;
      const int total4 = (feat * ww) >> 2;
      for (int i = tid; i < total4; i += blockDim.x) {
        const int e = i << 2;
        const int f = e / ww, q = e - f * ww;
        const float* t = tile + f * stride + q;
        float4 v = make_float4(t[0], t[1], t[2], t[3]);
        if (n4) {
          const float4 z = __ldcs(n4 + i);
          // mul then add, two roundings like torch's `x + randn*noise` (never an FMA)
          v.x = __fadd_rn(v.x, __fmul_rn(z.x, noise_scale)); v.y = __fadd_rn(v.y, __fmul_rn(z.y, noise_scale));
          v.z = __fadd_rn(v.z, __fmul_rn(z.z, noise_scale)); v.w = __fadd_rn(v.w, __fmul_rn(z.w, noise_scale));
        }
        __stcs(o4 + i, v);
      }
    } else {
      const int total = ww * feat;
      for (int i = tid; i < total; i += blockDim.x) {
        const int pos = i / feat, f = i - pos * feat;
        const int y = pos / w, x = pos - y * w;
        const int sr = mirror_index(r + lo + y, scene_rows) - slab_row0;
        const int sc = mirror_index(c + lo + x, cols);
        tile[f * stride + pos] = __ldg(cube + (int64_t(sr) * cols + sc) * feat + f);
      }
      __syncthreads();
      float* o = out + p * int64_t(feat) * ww;
      const float* nz = noise ? noise + p * int64_t(feat) * ww : nullptr;
      for (int i = tid; i < total; i += blockDim.x) {
        const int f = i / ww, q = i - f * ww;
        float v = tile[f * stride + q];
        if (nz) v = __fadd_rn(v, __fmul_rn(__ldcs(nz + i), noise_scale));
        __stcs(o + i, v);
      }
    }
    __syncthreads();
  }
}


// Fast path (feat % 4 == 0, w*w % 4 == 0, 16-byte aligned): no shared memory.  A thread owns a
// 4-position x 4-channel block: four 16-byte loads (one per source pixel, 15 lanes cover a pixel's
// 240 contiguous bytes), a 4x4 transpose in registers, four 16-byte streaming stores (one per
// channel).  Adjacent lanes take adjacent position groups of the same channels, so every store
// instruction fills whole 32-byte sectors.  ~70 instructions per 256 bytes moved and 2 CTAs x 256
// threads of pure loads/stores in flight per SM.
__device__ __forceinline__ float4 ldg_keep(const float4* p, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
  return v;
}

template <int Q>                                             // groups (16-byte stores) that adjacent lanes write side by side
__global__ void __launch_bounds__(256)
patch_gather_reg_kernel(const float* __restrict__ cube, int scene_rows, int cols, int feat, int slab_row0, int w,
                        const int64_t* __restrict__ idx, int64_t first, int64_t n, const float* __restrict__ noise,
                        float noise_scale, float* __restrict__ out) {
  const int ww = w * w, f4n = feat >> 2, groups = ww >> 2;
  const int lo = window_lo(w);
  const int items = groups * f4n;
  const int pair_items = Q * f4n;
  // the cube is re-read ~w*w times while the output streams through L2 once: keep the cube's lines
  // (evict_last), let the stores go first (st.global.cs)
  uint64_t keep;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
  constexpr int U = 4;                                       // items (4 loads each) in flight per thread
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const int64_t pix = idx ? idx[p] : first + p;
    const int r = int(pix / cols), c = int(pix % cols);
    float* obase = out + p * int64_t(feat) * ww;
    const float* nbase = noise ? noise + p * int64_t(feat) * ww : nullptr;
    for (int t0 = threadIdx.x; t0 < items; t0 += blockDim.x * U) {
      float4 v[U][4];
      int f4s[U], gs[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int t = t0 + u * blockDim.x;
        const int gp = t / pair_items, rem = t - gp * pair_items;
        // groups come in runs of Q (adjacent lanes -> adjacent 16-byte stores = 16*Q contiguous bytes per channel); a
        // group count that is not a multiple of Q leaves a last short run whose items belong to its `groups % Q` groups
        const int full = groups / Q, last = groups - full * Q;
        const bool tail = gp == full;
        f4s[u] = tail ? rem / (last > 0 ? last : 1) : rem / Q;
        gs[u] = (t < items) ? gp * Q + (tail ? rem - f4s[u] * last : rem - f4s[u] * Q) : groups;
        if (gs[u] < groups) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int pos = gs[u] * 4 + k;
            const int y = pos / w, x = pos - y * w;
            const int sr = mirror_index(r + lo + y, scene_rows) - slab_row0;
            const int sc = mirror_index(c + lo + x, cols);
            v[u][k] = ldg_keep(reinterpret_cast<const float4*>(cube + (int64_t(sr) * cols + sc) * feat) + f4s[u], keep);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (gs[u] >= groups) continue;
        float4 o[4];
        o[0] = make_float4(v[u][0].x, v[u][1].x, v[u][2].x, v[u][3].x);
        o[1] = make_float4(v[u][0].y, v[u][1].y, v[u][2].y, v[u][3].y);
        o[2] = make_float4(v[u][0].z, v[u][1].z, v[u][2].z, v[u][3].z);
        o[3] = make_float4(v[u][0].w, v[u][1].w, v[u][2].w, v[u][3].w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int64_t off = int64_t(f4s[u] * 4 + j) * ww + gs[u] * 4;
          if (nbase) {
            const float4 z = __ldcs(reinterpret_cast<const float4*>(nbase + off));
            // mul then add, two roundings like torch's `x + randn*noise` (never an FMA)
            o[j].x = __fadd_rn(o[j].x, __fmul_rn(z.x, noise_scale)); o[j].y = __fadd_rn(o[j].y, __fmul_rn(z.y, noise_scale));
            o[j].z = __fadd_rn(o[j].z, __fmul_rn(z.z, noise_scale)); o[j].w = __fadd_rn(o[j].w, __fmul_rn(z.w, noise_scale));
          }
          __stcs(reinterpret_cast<float4*>(obase + off), o[j]);
        }
      }
    }
  }
}

}  // namespace cmlpl

extern "C" int cmlpl_patch_gather_f32(const float* cube, int scene_rows, int cols, int feat,
                                      int slab_row0, int slab_rows, int w, int odd_mode,
                                      const int64_t* idx, int64_t first, int64_t n,
                                      const float* noise, float noise_scale, float* out,
                                      cmlpl_stream_t stream) {
  using namespace cmlpl;
  CMLPL_CHECK_ARG(cube && out, "patch_gather: null pointer");
  CMLPL_CHECK_ARG(scene_rows > 0 && cols > 0 && feat > 0 && w > 0, "patch_gather: bad dims");
  // the reference raises ValueError for the wrong parity (hyper_tools.py:240 shape mismatch)
  CMLPL_CHECK_ARG(odd_mode ? (w % 2 == 1) : (w % 2 == 0),
                  "patch_gather: w=%d has the wrong parity for %s", w,
                  odd_mode ? "ExtractPatches_for_base" : "ExtractPatches");
  CMLPL_CHECK_ARG(w / 2 <= scene_rows && w / 2 <= cols, "patch_gather: window larger than the scene");
  CMLPL_CHECK_ARG(slab_row0 >= 0 && slab_rows > 0 && slab_row0 + slab_rows <= scene_rows,
                  "patch_gather: slab [%d,%d) outside the scene", slab_row0, slab_row0 + slab_rows);
  if (n <= 0) return CMLPL_OK;
  const size_t smem = size_t(feat) * (w * w + 1) * sizeof(float);
  CMLPL_CHECK_ARG(smem <= 227 * 1024, "patch_gather: window of %zu bytes does not fit shared memory", smem);
  const bool vec = (feat % 4 == 0) && ((w * w) % 4 == 0) &&
                   (reinterpret_cast<uintptr_t>(cube) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(out) % 16 == 0) &&
                   (!noise || reinterpret_cast<uintptr_t>(noise) % 16 == 0);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec) {
    int64_t grid = int64_t(sm_count()) * 8;
    if (grid > n) grid = n;
    // runs of 4 groups: 64 contiguous bytes per channel and warp store (measured: 56 / 63 % of the HBM copy peak on random /
    // raster pixels, against 54 / 59 % with pairs and 54 / 61 % with runs of 8)
    patch_gather_reg_kernel<4><<<int(grid), 256, 0, st>>>(cube, scene_rows, cols, feat, slab_row0, w, idx, first, n, noise,
                                                          noise_scale, out);
  } else {
    auto kern = patch_gather_kernel<false>;
    CMLPL_MAX_DYN_SMEM(kern, int(smem));
    const int ctas_per_sm = smem * 2 <= 227 * 1024 ? 2 : 1;
    int64_t grid = int64_t(sm_count()) * ctas_per_sm * 4;
    if (grid > n) grid = n;
    kern<<<int(grid), 256, smem, st>>>(cube, scene_rows, cols, feat, slab_row0, w, idx, first, n, noise, noise_scale, out);
  }
  CMLPL_CHECK_LAUNCH("patch_gather");
  return CMLPL_OK;
}
