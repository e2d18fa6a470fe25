/*
 * cmlpl.h -- C ABI of libcmlpl_sm100.so: the B200 (sm_100a) kernels behind the CMLPL
 * hot path (patch gather -> BaseNet2 -> losses -> full-scene inference -> OA/AA/kappa).
 *
 * The reference (liuli33/CMLPL) has no FFI: its "API" is Python names.  Each entry
 * point below names the reference code it replaces (file:line under the reference
 * root); cmlpl_b200/ binds them with ctypes and re-exposes the reference's Python
 * names on top (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer; the library allocates nothing persistent;
 *   - all launches are asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value 0 = ok, <0 = error; cmlpl_last_error() gives the message
 *     (thread-local);
 *   - tensors are dense row-major with the shapes written in the comments.
 */
#ifndef CMLPL_H_
#define CMLPL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMLPL_OK 0
#define CMLPL_ERR_ARG (-1)
#define CMLPL_ERR_CUDA (-2)
#define CMLPL_ERR_UNSUPPORTED (-3)

typedef void* cmlpl_stream_t;

int cmlpl_version(void);
const char* cmlpl_last_error(void);
/* 1 if the device the current context runs on is sm_100 (B200), else 0 (or <0). */
int cmlpl_device_ok(void);

/* ------------------------------------------------------------------ patches --
 * tools/hyper_tools.py:35-55 (MirrowCut), :226-243 (ExtractPatches, even w) and
 * :300-317 (ExtractPatches_for_base, odd w; odd_mode=1).
 * cube   f32 [slab_rows, cols, feat]   channels-last slab = scene rows
 *                                       [slab_row0, slab_row0+slab_rows) of a scene with
 *                                       scene_rows rows (band sharding; mirror padding is
 *                                       applied at true scene edges only)
 * idx    i64 [n] raster pixel indices r*cols+c, or NULL for first, first+1, ...
 * noise  f32 [n, feat, w, w] or NULL;  out = patch + noise*noise_scale (train.py:157)
 * out    f32 [n, feat, w, w]
 */
int cmlpl_patch_gather_f32(const float* cube, int scene_rows, int cols, int feat,
                           int slab_row0, int slab_rows, int w, int odd_mode,
                           const int64_t* idx, int64_t first, int64_t n,
                           const float* noise, float noise_scale,
                           float* out, cmlpl_stream_t stream);

/* ------------------------------------------------------------ fp32 building --
 * Dense fp32 building blocks of BaseNet2 (tools/models.py:97-152) in batch mode,
 * NCHW like the reference; used by training forward/backward (a5, a12).
 */

/* C[M,N] = act( alpha * op(A)[M,K] . op(B)[K,N] + bias[N] + beta*C ) with arbitrary
 * element strides (so any transpose is a stride choice).  act: 0 none, 1 relu.
 * nn.Linear forward / dgrad / wgrad (models.py:142,150). */
int cmlpl_sgemm_f32(int M, int N, int K, float alpha,
                    const float* A, int64_t a_rs, int64_t a_cs,
                    const float* B, int64_t b_rs, int64_t b_cs,
                    const float* bias, float beta,
                    float* C, int64_t c_rs, int64_t c_cs, int act,
                    cmlpl_stream_t stream);

/* y[b,co,h,w] = act( conv_kxk(x, wgt, pad=k/2) + bias + (res ? res : 0) ), k in {1,3}.
 * models.py:132-135,138-139.   x f32 [b,ci,h,w], wgt f32 [co,ci,k,k], res/y f32 [b,co,h,w].
 * transpose_w=1 computes the data gradient: uses wgt[ci_out... ] flipped/transposed, i.e.
 * y = conv(x, flip(wgt)^T) with x = dL/dy [b,co,h,w] -> y = dL/dx [b,ci,h,w]. */
int cmlpl_conv2d_f32(const float* x, const float* wgt, const float* bias, const float* res,
                     float* y, int b, int ci, int co, int h, int w, int k, int act,
                     int transpose_w, cmlpl_stream_t stream);

/* dW[co,ci,k,k] = sum_{b,y,x} dy[b,co,y,x] * x[b,ci,y+ky-p,x+kx-p];  db[co] = sum dy. */
int cmlpl_conv2d_wgrad_f32(const float* x, const float* dy, float* dw, float* db,
                           int b, int ci, int co, int h, int w, int k, cmlpl_stream_t stream);

/* 2x2/2 average pool forward / backward (models.py:136,140), NCHW. */
int cmlpl_avgpool2_f32(const float* x, float* y, int64_t planes, int h, int w, cmlpl_stream_t stream);
int cmlpl_avgpool2_bwd_f32(const float* dy, float* dx, int64_t planes, int h, int w, cmlpl_stream_t stream);

/* relu backward with optional residual fan-out: dx = dy * (y > 0). */
int cmlpl_relu_bwd_f32(const float* y, const float* dy, float* dx, int64_t n, cmlpl_stream_t stream);
/* column sums: out[n] = sum_m x[m,n] (bias gradients of nn.Linear). */
int cmlpl_colsum_f32(const float* x, float* out, int64_t m, int64_t n, cmlpl_stream_t stream);

/* models.py:87-90 Normalize: y = x / sqrt(sum x^2) per row (no epsilon), and its backward
 * dx = (dy - y * sum(dy*y)) / norm. */
int cmlpl_l2norm_f32(const float* x, float* y, float* norm, int64_t rows, int64_t cols, cmlpl_stream_t stream);
int cmlpl_l2norm_bwd_f32(const float* y, const float* norm, const float* dy, float* dx,
                         int64_t rows, int64_t cols, cmlpl_stream_t stream);

/* ------------------------------------------------------------ scene inference --
 * tools/hyper_tools.py:416-437 (test_whole) over a whole scene / row band without ever
 * materialising patches: conv0 is evaluated once per scene pixel into a mirrored,
 * halo-padded fp16 map; conv1 (+residual, ReLU, 2x2 avg-pool), conv2 and the pool + conv columns of
 * the classifier are evaluated ONCE PER SCENE POSITION in the 9 / 25 classes of how a patch border can
 * cut them (exact compute sharing; persistent tcgen05 kernels fed by TMA tile loads); the spectral
 * branch is one fused two-GEMM kernel; a sum head adds each pixel's 25 gathered conv partials to its
 * spectral partials and takes the argmax.  > 16 classes or > 224 bands fall back to the all-per-pixel
 * tcgen05 kernel (patch window staged in shared memory) and CUDA-core heads.
 *
 * Packed weights: cmlpl_pack_basenet2() converts the reference state_dict tensors
 * (models.py:102-127: conv0/1/2.{weight,bias}, feat_spe.*, classifier.*) into the
 * kernel layouts inside a caller-owned buffer of cmlpl_packed_bytes() bytes.
 */
size_t cmlpl_packed_bytes(int num_features, int num_classes, int w);
int cmlpl_pack_basenet2(const float* conv0_w, const float* conv0_b,
                        const float* conv1_w, const float* conv1_b,
                        const float* conv2_w, const float* conv2_b,
                        const float* spe_w, const float* spe_b,
                        const float* cls_w, const float* cls_b,
                        int num_features, int num_classes, int w,
                        void* packed, cmlpl_stream_t stream);

/* Which kernels cmlpl_scene_infer / cmlpl_scene_infer_raw run on the tensor-core path (<= 16 classes, <= 224 bands), per
 * calling host thread: 1 (default) = the scene-level kernels with exact compute sharing (w = 20 and w = 11), 0 = the
 * per-pixel kernels (patch_cnn_kernel<w> per pixel; the independent implementation the tests compare the default against,
 * models.py:133-150 per patch).  The workspace size and layout follow the mode: set it before sizing the workspace. */
int cmlpl_set_scene_path_mode(int mode);

/* Workspace needed to infer `band_rows` rows of a scene with `cols` columns. */
size_t cmlpl_scene_workspace_bytes(int band_rows, int cols, int num_features, int num_classes, int w);

/* Byte offsets of the workspace regions (for stage-level profiling): offsets[12] =
 * {f0pad, x16, h16 (partial spectral logits f32 [4][ceil(n/128)*128][16] on the dense path), g (unused, zero-sized),
 *  pmq, yq, lmap, p2, spe_logits, hidden, total, uses_tensor_core_path}. */
int cmlpl_scene_workspace_layout(int band_rows, int cols, int num_features, int num_classes, int w,
                                 size_t* offsets);

/* cube     f32 [slab_rows, cols, 60]  PCA cube slab (rows slab_row0.. of the scene)
 * spectra  f32 [band_rows*cols, num_features]   rows of X for the band's pixels
 * labels   u8  [band_rows*cols]  argmax (first index on ties, hyper_tools.py:426)
 * logits   f32 [band_rows*cols, num_classes] or NULL
 * Band = scene rows [band_row0, band_row0+band_rows). */
int cmlpl_scene_infer(const float* cube, int scene_rows, int cols, int slab_row0, int slab_rows,
                      const float* spectra, int num_features, int num_classes, int w,
                      int band_row0, int band_rows, const void* packed,
                      void* workspace, size_t workspace_bytes,
                      uint8_t* labels, float* logits, cmlpl_stream_t stream);

/* The same from the RAW cube (dtype 0 = uint16, 1 = float32; <= 16 classes, <= 224 bands): the PCA
 * projection and both z-scores of tools/hyper_tools.py:285-292 are folded into conv0
 * (wf f32 [B][64] = (W0 . (U/s)^T)^T, bf f32 [64] = b0 - W0 . m/s) and into the fp16 conversion of the
 * spectral branch (mu, inv_sigma f32 [B]); neither the PCA cube nor the z-scored spectra touch HBM.
 * raw [slab_rows*cols, B] holds scene rows slab_row0.. (halo included, like `cube` above). */
int cmlpl_scene_infer_raw(const void* raw, int dtype, int scene_rows, int cols, int slab_row0, int slab_rows,
                          int num_features, int num_classes, int w, int band_row0, int band_rows,
                          const float* wf, const float* bf, const float* mu, const float* inv_sigma,
                          const void* packed, void* workspace, size_t workspace_bytes,
                          uint8_t* labels, float* logits, cmlpl_stream_t stream);
int cmlpl_spectral_hidden_raw_tc(const void* raw, int dtype, int64_t n, int num_features, int num_classes, int w,
                                 const float* mu, const float* inv_sigma, const void* packed, void* x16,
                                 void* h16, cmlpl_stream_t stream);

/* Individual stages of cmlpl_scene_infer (exposed for tests / profiling). */
int cmlpl_conv0_map_f16(const float* cube, int scene_rows, int cols, int slab_row0, int slab_rows,
                        int w, int band_row0, int band_rows, const void* packed,
                        void* f0pad /* f16 [8, band_rows+w-1, cols+w-1, 8]: chunk-planar (8 channels per 16-B chunk) */, cmlpl_stream_t stream);
int cmlpl_patch_cnn_f16(const void* f0pad, int cols, int w, int band_rows, const void* packed,
                        void* p2 /* f16 [band_rows*cols, (w/4)^2, 64] */, cmlpl_stream_t stream);
int cmlpl_spectral_head_f32(const float* spectra, int64_t n, int num_features, int num_classes, int w,
                            const void* packed, float* hidden /* f32 [chunk,1024] scratch */,
                            int64_t chunk, float* spe_logits /* f32 [n, C] */, cmlpl_stream_t stream);
int cmlpl_classify_f16(const void* p2, const float* spe_logits /* may be NULL */, int64_t n,
                       int num_features, int num_classes, int w, const void* packed,
                       uint8_t* labels, float* logits /* may be NULL */, cmlpl_stream_t stream);
/* Tensor-core head (<= 16 classes, <= 208 bands), fp16 operands / fp32 accumulate.  A operands are
 * "UMMA tiles": f16 [ceil(n/128)][K/8][128 rows][8], pixel p = tile p/128, row p%128.
 *   cmlpl_patch_cnn_f16_tiled : as cmlpl_patch_cnn_f16 but p2t is tiled with K = (w/4)^2*64 (pos, ch)
 *   cmlpl_spectral_hidden_tc  : x16 (scratch, K = B padded to 16) and h16 = relu(feat_spe(x)) tiles, K = 1024
 *   cmlpl_head_tc             : classifier over [p2t | h16] + bias, argmax -> labels (and logits) */
int cmlpl_patch_cnn_f16_tiled(const void* f0pad, int cols, int w, int band_rows, const void* packed,
                              void* p2t, cmlpl_stream_t stream);
int cmlpl_spectral_hidden_tc(const float* spectra, int64_t n, int num_features, int num_classes, int w,
                             const void* packed, void* x16, void* h16, cmlpl_stream_t stream);
int cmlpl_head_tc(const void* p2t, const void* h16, int64_t n, int num_features, int num_classes, int w,
                  const void* packed, uint8_t* labels, float* logits, cmlpl_stream_t stream);
/* Exact compute sharing of conv1: conv1 + bias + residual + ReLU evaluated once per
 * position of the conv0 map in 3x3 = 9 patch-border classes, then the 2x2 average pools of every
 * top-left position in the 9 pooled border classes (tools/models.py:133-136 for all patches at once).
 *   f0pad f16 [8][PR][PC][8] (PR = band_rows+w-1, PC = cols+w-1), g f32 [9][PR*PC][64] scratch,
 *   pm f16 [9][PR][PC][64]: pooled maps, variant = A*3+B (A,B in top/mid/bot, left/mid/right). */
int cmlpl_conv1_scene_f16(const void* f0pad, int cols, int w, int band_rows, const void* packed, float* g,
                          void* pm, cmlpl_stream_t stream);
/* Dense (scene-level) rest of the conv tower, exact compute sharing of conv2 / pool / classifier
 * (tools/models.py:137-150 for all patches at once; csrc/conv2_scene_sm100.cu).  PR2 = (band_rows+w)/2,
 * PC2 = (cols+w)/2; "planes" = the 4 parity planes (pr&1)*2+(pc&1) of the padded map, y' = pr>>1, x' = pc>>1.
 *   cmlpl_conv1_scene_variants_f32: g f32 [9][PR*PC][64] only (no pooling)
 *   cmlpl_conv1_scene_planes_f16  : + pooled maps as planes   pmq f16 [9][4][8 chunks][PR2][PC2][8]
 *   cmlpl_conv2_scene_f16         : conv2+bias+residual+ReLU in 25 border classes (rho, kap in {i=0, 1, 2..7, 8, 9}); the
 *                                   border halves of the 2x2 pool (rho 0 + rho 1 one row down, rho 3 + rho 4, same for
 *                                   kap) are averaged in the epilogue:  yq f16 [9 = Al*3+Be][4][8][PR2][PC2][8],
 *                                   yq[Al][Be][y',x'] = mean over the border partners of Y[rho][kap][y'+u, x'+v]
 *   cmlpl_pool2_cls_f16           : rest of the 2x2 avg-pool (middle classes) + conv columns of the classifier per
 *                                   pooled cell (I,J), summed over the pooled columns J of a pooled row I:
 *                                   lmap f32 [4][5 = I][4 class quads][PR2][PC2][4],
 *                                   lmap[I][y',x'] = sum_J L[I][J][y', x'+2J] */
int cmlpl_conv1_scene_variants_f32(const void* f0pad, int cols, int w, int band_rows, const void* packed, float* g,
                                   cmlpl_stream_t stream);
int cmlpl_conv1_scene_planes_f16(const void* f0pad, int cols, int w, int band_rows, const void* packed, float* g,
                                 void* pmq, cmlpl_stream_t stream);
/* cmlpl_conv1_pool_planes_f16: the same pooled planes from ONE kernel (conv1 variants pooled in the epilogue through a
 * shared-memory / shuffle exchange; no fp32 scratch in HBM) -- what cmlpl_scene_infer runs. */
int cmlpl_conv1_pool_planes_f16(const void* f0pad, int cols, int w, int band_rows, const void* packed, void* pmq,
                                cmlpl_stream_t stream);
int cmlpl_conv2_scene_f16(const void* pmq, int cols, int w, int band_rows, const void* packed, void* yq,
                          cmlpl_stream_t stream);
int cmlpl_pool2_cls_f16(const void* yq, int cols, int w, int band_rows, int num_features, int num_classes,
                        const void* packed, float* lmap, cmlpl_stream_t stream);
/* Fused spectral branch + sum head (what cmlpl_scene_infer runs since the hidden features stopped going through HBM):
 *   cmlpl_spectral_logits_tc     : part f32 [4 hidden quarters][ceil(n/128)*128][16] = Wc_spe . relu(Wspe . x + b) per
 *                                  quarter of the 1024 hidden features (tools/models.py:142-143,150), <= 224 bands
 *   cmlpl_spectral_logits_raw_tc : the same from the raw cube (z-score folded into the fp16 conversion)
 *   cmlpl_head_sum_lmap          : 4 quarter partials + the 5 gathered conv partials + bias, argmax */
int cmlpl_spectral_logits_tc(const float* spectra, int64_t n, int num_features, int num_classes, int w,
                             const void* packed, void* x16, float* part, cmlpl_stream_t stream);
int cmlpl_spectral_logits_raw_tc(const void* raw, int dtype, int64_t n, int num_features, int num_classes, int w,
                                 const float* mu, const float* inv_sigma, const void* packed, void* x16, float* part,
                                 cmlpl_stream_t stream);
int cmlpl_head_sum_lmap(const float* part, const float* lmap, int cols, int band_rows, int num_features,
                        int num_classes, int w, const void* packed, uint8_t* labels, float* logits,
                        cmlpl_stream_t stream);
/* Per-pixel conv2 stage on the pooled conv1 maps of cmlpl_conv1_scene_f16 (models.py:137-140 per patch):
 * pm f16 [9][PR][PC][64] -> p2t UMMA tiles [ceil(n/128)][(w/4)^2*8][128][8] for cmlpl_head_tc. */
int cmlpl_patch_conv2_f16_tiled(const void* pm, int cols, int w, int band_rows, const void* packed,
                                void* p2t, cmlpl_stream_t stream);
int cmlpl_debug_patch_conv2_trace(const void* pm, int cols, int w, int band_rows, const void* packed,
                                  void* p2t, long long* trace, cmlpl_stream_t stream);
/* Diagnostics: cmlpl_patch_cnn_f16 with CTA 0 writing clock64() stamps of its first 64 patches
 * (16 slots each: loader / MMA issuer / epilogue protocol points) to trace i64 [64,16]. */
int cmlpl_debug_patch_cnn_trace(const void* f0pad, int cols, int w, int band_rows, const void* packed,
                                void* p2, long long* trace, cmlpl_stream_t stream);
/* hyper_tools.py:426 torch.max(outputs, 1): first index on ties.  logits f32 [n, C] -> u8 [n]. */
int cmlpl_argmax_u8(const float* logits, int64_t n, int num_classes, uint8_t* labels,
                    cmlpl_stream_t stream);

/* ------------------------------------------------------------- preprocessing --
 * tools/hyper_tools.py:8-32 (featureNormalize, PCANorm) as called at :289-292, on device.
 * x: raw scene [n, B], dtype 0 = uint16, 1 = float32.
 * fit  : mean f64 [B] = column means, gram f64 [B,B] = sum_n (x-mean)(x-mean)^T, upper 32x32 tiles only
 *        (mirror on the host; np.cov = gram/(n-1), np.std^2 = diag(gram)/n).  The B x B SVD stays on the host.
 * apply: spectra f32 [n,B] = (x-mu)*inv_sigma (may be NULL);  cube f32 [n,npc] = (x-mu).Us - shift with
 *        Us f32 [B,npc] = U[:, :npc]/s and shift = m/s (m, s = mean/std of the projected data). */
int cmlpl_preprocess_fit_f64(const void* x, int dtype, int64_t n, int B, double* mean, double* gram,
                             cmlpl_stream_t stream);
int cmlpl_preprocess_apply(const void* x, int dtype, int64_t n, int B, int npc, const float* mu,
                           const float* inv_sigma, const float* Us, const float* shift, float* cube,
                           float* spectra, cmlpl_stream_t stream);

/* ------------------------------------------------------------------ metrics --
 * tools/hyper_tools.py:208-223 (CalAccuracy): integer confusion matrix
 * cm[label, pred] += 1 for pairs with both in [0, C).  pred u8 [n], label i64 [n],
 * cm i64 [C, C] (accumulated into; zero it first).  OA/AA/kappa are float64 host
 * arithmetic on these counts. */
int cmlpl_confusion_i64(const uint8_t* pred, const int64_t* label, int64_t n, int num_classes,
                        int64_t* cm, cmlpl_stream_t stream);

/* ------------------------------------------------------------------- losses --
 * train.py:191-265 and tools/models.py:14-39.  All fp32, rows = samples; `loss` is a device
 * scalar that is ACCUMULATED into (zero it first).
 */
/* Hard-label CE (train.py:191-192) or soft/masked CE (train.py:239-242):
 *   loss_i = -sum_c log_softmax(z_i)_c * t_ic * m_i ;  *loss += scale * mean_i loss_i
 *   dz_ic  = scale/rows * m_i * (softmax(z_i)_c * sum_c t_ic - t_ic)
 * labels i64 [rows] (hard) XOR probs f32 [rows, C] (soft); mask f32 [rows] or NULL; dlogits may be NULL. */
int cmlpl_ce_fwd_bwd_f32(const float* logits, const int64_t* labels, const float* probs,
                         const float* mask, int64_t rows, int C, float scale,
                         float* loss, float* dlogits, cmlpl_stream_t stream);

/* S = A . B^T with A f32 [M, K], B f32 [N, K] (K contiguous, K % 64 == 0) on tcgen05: operands rounded to fp16, fp32
 * accumulation; C f32 [M, N].  The similarity matrices of train.py:213,246 and tools/models.py:27.
 * cmlpl_set_loss_gemm_mode(1) makes cmlpl_bank_smooth_f32 / cmlpl_graph_contrast_f32 / cmlpl_ntxent_f32 use it for
 * their similarity GEMM (default 0: fp32 CUDA-core tiles, the 1e-5 parity path).  The setting is per host thread. */
int cmlpl_sim_nt_tc_f32(const float* A, const float* B, int M, int N, int K, float* C, cmlpl_stream_t stream);
int cmlpl_set_loss_gemm_mode(int mode);

/* trian_CCT.py:76-84 softmax_js_loss: *loss += scale * 0.5 * (kl_div(log_softmax(z), M, 'mean') +
 * kl_div(log(t + 1e-5), M, 'mean')) with M = (softmax(z) + t)/2 and 'mean' over all rows*C elements;
 * dlogits (may be NULL) = scale * dL/dz, targets are constants.  logits, targets f32 [rows, C]. */
int cmlpl_softmax_js_f32(const float* logits, const float* targets, int64_t rows, int C, float scale,
                         float* loss, float* dlogits, cmlpl_stream_t stream);

/* loss_helper.py:247-248: entropy_i = -sum_c p_ic log(p_ic + eps), p = softmax(logits_i).  f32 [rows]. */
int cmlpl_softmax_entropy_f32(const float* logits, int64_t rows, int C, float eps, float* entropy,
                              cmlpl_stream_t stream);

/* train.py:203,213-215,220-222: probs_orig = softmax(logits); if smooth: A = exp(feats.Q^T/T)
 * row-normalised, probs = alpha*probs_orig + (1-alpha)*A.Qp; mask = max(probs) >= thr.
 * feats f32 [rows, dim], queue_feats f32 [queue, dim], queue_probs f32 [queue, C],
 * work f32 [rows, queue] scratch (only when smooth).  probs_orig may be NULL. */
int cmlpl_bank_smooth_f32(const float* logits, const float* feats, const float* queue_feats,
                          const float* queue_probs, int64_t rows, int C, int dim, int64_t queue,
                          float alpha, float T, int smooth, float thr, float* work,
                          float* probs_orig, float* probs, float* mask, cmlpl_stream_t stream);

/* train.py:246-265: pseudo-label-graph contrastive loss, forward + gradient.
 * f_row, f_col f32 [n, dim]; p1 (rows), p (cols) f32 [n, C].
 *   L = mean_i( -sum_j log(sp_ij) Q_ij + sum_j log(sp_ij + 1) Qn_ij ),  sp = row-softmax(f_row.f_col^T / T)
 * grad_side 0 -> dfeat = scale*dL/df_row (loss_contrast, :260-262); 1 -> scale*dL/df_col (loss_contrast1,
 * :263-265).  work f32 [3*n*n] scratch.  *loss += scale*L.  dfeat f32 [n, dim] may be NULL. */
int cmlpl_graph_contrast_f32(const float* f_row, const float* f_col, const float* p1, const float* p,
                             int64_t n, int dim, int C, float T, int grad_side, float scale,
                             float* work, float* loss, float* dfeat, cmlpl_stream_t stream);

/* tools/models.py:22-39 ContrastiveLoss (NT-Xent) on z = [normalize(emb_i); normalize(emb_j)]
 * f32 [2*bs, dim] (unit rows; use cmlpl_l2norm_f32 / _bwd around it).  *loss += L;
 * dz f32 [2*bs, dim] = dL/dz or NULL.  work f32 [2*(2bs)^2] scratch. */
int cmlpl_ntxent_f32(const float* z, int64_t bs, int dim, float T, float* work, float* loss,
                     float* dz, cmlpl_stream_t stream);

/* torch.optim.Adam (train.py:131-132,268,272; default betas/eps) over a list of tensors, one launch
 * per 16 tensors.  *_host are HOST arrays of n_tensors device pointers / element counts; step is the
 * 1-based step count.  Tensors whose grad pointer is NULL are skipped (grad=None). */
int cmlpl_adam_multi_f32(int n_tensors, float* const* p_host, const float* const* g_host,
                         float* const* m_host, float* const* v_host, const int64_t* numel_host,
                         float lr, float beta1, float beta2, float eps, int step,
                         cmlpl_stream_t stream);


/* ------------------------------------------------------------- collectives --
 * The two exchanges at the end of a row-band sharded scene inference, for hosts that do not bring their
 * own process group (the Python layer of this repo uses torch.distributed): NCCL over NVLink, one communicator per
 * process / device, resolved at run time from libnccl.so.2.
 *   cmlpl_comm_unique_id        rank 0 fills id128 (128 bytes), the host hands it to the other ranks
 *   cmlpl_comm_init             collective over all ranks; *out is owned by the caller until cmlpl_comm_destroy
 *   cmlpl_comm_allgather_labels out u8 [world * per_rank] <- every rank's local u8 [per_rank] (bands padded to the
 *                               common height, hyper_tools.py:426-431 label map in raster order)
 *   cmlpl_comm_allreduce_confusion  cm i64 [count] summed in place (hyper_tools.py:208-223 counts) */
typedef struct cmlpl_comm cmlpl_comm;
int cmlpl_comm_unique_id(void* id128);
int cmlpl_comm_init(int rank, int world, const void* id128, cmlpl_comm** out);
int cmlpl_comm_allgather_labels(cmlpl_comm* comm, const uint8_t* local, int64_t per_rank, uint8_t* out,
                                cmlpl_stream_t stream);
int cmlpl_comm_allreduce_confusion(cmlpl_comm* comm, int64_t* cm, int count, cmlpl_stream_t stream);
int cmlpl_comm_destroy(cmlpl_comm* comm);

/* ------------------------------------------------------- fused training step --
 * One mutual-learning step of train.py:150-272 for BOTH BaseNet2 peers as ~17 kernel launches on `stream`
 * (CUDA-graph capturable: every per-step scalar lives in the device-side cmlpl_train_params block).
 * Contractions run on tcgen05 (fp16 operands = TF32's 11-bit significand, fp32 accumulate in TMEM): conv0 / conv1 /
 * conv2 forward (tools/models.py:132-140), their data gradients and weight gradients (train.py:267,271); gradient
 * operands are scaled by a power of two chosen on device from max|dL/dcat| so they sit in fp16's normal range.
 * Bar: |d| <= 1e-3 * max|ref| against the fp32 path (cmlpl_conv2d_f32 & co. stay the 1e-5 reference).
 *
 * Row order inside a net: [bs labelled ; btu unlabelled] (train.py:174,184); nb = bs + btu.
 * Tensor index order in cmlpl_train_net arrays: 0 conv0.weight [64,60,1,1], 1 conv0.bias, 2 conv1.weight
 * [64,64,3,3], 3 conv1.bias, 4 conv2.weight, 5 conv2.bias, 6 feat_spe.weight [1024,B], 7 feat_spe.bias,
 * 8 classifier.weight [C,2624], 9 classifier.bias (state-dict layouts, tools/models.py:102-127).
 */
#define CMLPL_TRAIN_TENSORS 10
typedef struct cmlpl_train_net {
  float* p[CMLPL_TRAIN_TENSORS];   /* parameters, updated in place by the Adam phase */
  float* g[CMLPL_TRAIN_TENSORS];   /* gradients of total_loss (train.py:266/270), overwritten every step */
  float* m[CMLPL_TRAIN_TENSORS];   /* Adam exp_avg */
  float* v[CMLPL_TRAIN_TENSORS];   /* Adam exp_avg_sq */
  float* queue_feats;              /* f32 [queue, 1024]  memory bank this net's targets are smoothed with (train.py:138-145) */
  float* queue_probs;              /* f32 [queue, C] */
} cmlpl_train_net;

typedef struct cmlpl_train_params {   /* lives in DEVICE memory; the host refreshes it before every step */
  float noise_scale;      /* train.py:157 args.noise (ignored for inputs passed with cube == NULL) */
  float dropout_p;        /* tools/models.py:147; used only when drop_mask == NULL */
  float temperature;      /* train.py:213,246 */
  float alpha;            /* train.py:215 */
  float adap_thr;         /* args.thr * exp(-0.5 (epoch/num_epochs)^2), train.py:147-148,221 */
  float lr, beta1, beta2, eps, bc1, bc2_sqrt;   /* Adam: bias corrections 1-beta1^t, sqrt(1-beta2^t) */
  int smooth;             /* epoch > 0 or batch_index > queue_batch, train.py:212 */
  int queue_ptr[2];       /* write positions of the two banks for THIS step (train.py:232-237, host keeps the quirk) */
  unsigned long long seed, offset;   /* device Philox stream for noise / dropout when no tensors are injected */
  float grad_amax;        /* scratch: max |dL/dcat[:, :1600]| of this step (written by the head backward) */
  int pad_;
} cmlpl_train_params;

typedef struct cmlpl_train_io {
  int bs, btu, bands, classes, w, queue;      /* w must be 20, classes <= 16, bands <= 256 */
  /* ---- inputs */
  const float* cube; int scene_rows, cols;    /* f32 [R, cols, 60] PCA cube, or NULL: patch_noise then HOLDS the input patches */
  const int64_t* pix;                         /* i64 [nb] raster pixel per row (both nets read the same pixels) */
  const float* patch_noise;                   /* f32 [2, nb, 60, w, w] N(0,1) draws (train.py:157-182) or NULL -> Philox */
  const float* spectra;                       /* f32 [rows, B] z-scored spectra */
  const int64_t* spec_row;                    /* i64 [nb] row of `spectra` per sample; NULL: row pix[i] (cube mode) or the
                                                 assembled inputs f32 [2, nb, B] themselves (cube == NULL) */
  const float* spec_noise;                    /* f32 [2, nb, B] or NULL -> Philox (none when cube == NULL) */
  const float* drop_mask;                     /* f32 [2, nb, 2624] inverted-dropout mask (0 or 1/(1-p)) or NULL -> Philox */
  const int64_t* labels;                      /* i64 [bs] 0-based (train.py:159) */
  cmlpl_train_net net[2];
  cmlpl_train_params* params;                 /* device */
  /* ---- outputs (all device) */
  float* logits;      /* f32 [2, nb, C] */
  float* feat;        /* f32 [2, nb, 1024] unit rows (models.py:145-146) */
  float* probs;       /* f32 [2, btu, C]: [0] = probs (target of net 0, from net 1), [1] = probs1 (train.py:203-217) */
  float* mask;        /* f32 [2, btu]:    [0] = mask, [1] = masks (train.py:222,228) */
  float* hist;        /* f32 [12]: lc, total, cls, con, acc(net1 on labelled), total1, cls1, con1, lc1, 3 spare */
  float* grad_flat; size_t grad_flat_bytes;   /* optional: one block holding every g[] tensor of both nets (cleared with ONE
                                                 memset instead of one per tensor) */
  void* work; size_t work_bytes;              /* cmlpl_train_workspace_bytes() */
} cmlpl_train_io;

size_t cmlpl_train_workspace_bytes(int bs, int btu, int bands, int classes, int queue);
/* Byte offsets of the workspace regions (tests / profiling): offsets[20] = {x16, a0, p1, m1, m2, cat, dmask, ynoisy,
 * norm, dlogits, dfeat, dcat, dhp, dz1, da0, S, G, dG, probs_orig, total}.  Per-sample activations x16 / a0 / dz1 / da0
 * are f16 [2*nb][8 chunks][400][8], p1 f16 [2*nb][8][100][8]; m1 / m2 are ReLU masks u32 [2*nb][positions][2] (bit c%32
 * of word c/32); cat / dmask / dcat f32 [2*nb][2624]; dz1 / da0 carry the power-of-two gradient scale
 * 2^(12 - exponent(grad_amax)). */
int cmlpl_train_workspace_layout(int bs, int btu, int bands, int classes, int queue, size_t* offsets);
/* phases: bit 0 forward (gather+noise -> logits, feat), bit 1 losses (+ bank update, dlogits, dfeat),
 * bit 2 backward (all gradients), bit 3 Adam.  15 = the whole step. */
int cmlpl_train_step(const cmlpl_train_io* io, int phases, cmlpl_stream_t stream);
/* number of kernel launches cmlpl_train_step(io, phases) enqueues (for bench.py's gpu_launches) */
int cmlpl_train_step_launches(int phases);

#ifdef __cplusplus
}
#endif
#endif /* CMLPL_H_ */
