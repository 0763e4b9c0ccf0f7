/*
 * disvae_b200.h -- C ABI of libdisvae_b200.so: hand-written sm_100a kernels for the
 * disvae training hot path (BASELINE.json:north_star).
 *
 * The reference (YannDubs/disentangling-vae) is pure Python on PyTorch and has NO
 * native/FFI layer; every arithmetic op it runs is an ATen call.  Each entry point
 * below therefore cites the reference *call site* (file:line relative to the
 * reference checkout) whose ATen work it replaces.  The Python binding a maintainer
 * adds is a ctypes stub -- see INTEGRATION.md.
 *
 * Conventions
 *  - plain C: raw DEVICE pointers + explicit sizes; no torch types; fp32 everywhere.
 *  - the caller owns and allocates every buffer including workspaces
 *    (dv_*_workspace_bytes() say how much); the callee never allocates, frees or retains.
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant
 *    across streams, CUDA-graph capturable (no host sync, no allocation, no host RNG).
 *  - return value: DV_OK (0) or a negative DvStatus.  Nothing throws.
 *  - activations between conv layers are NHWC ("pixel-major": 32 channels = one 128-byte
 *    line per pixel).  Images at the model boundary (input x, reconstruction) are NCHW,
 *    as the reference's callers expect (utils/visualize.py:219-222).
 *  - "lo"/"hi": every Burgess conv layer (k=4, s=2, p=1) links a low-resolution tensor
 *    lo[B,H,W,32] and a high-resolution tensor hi[B,2H,2W,CH], CH in {1,3,32}, through a
 *    weight w[32][CH][4][4].  That is the memory layout of BOTH nn.Conv2d.weight
 *    [Cout=32,Cin=CH,4,4] and nn.ConvTranspose2d.weight [Cin=32,Cout=CH,4,4]
 *    (SURVEY.md trap T14), so three kernels cover all conv work:
 *       down : hi -> lo   (Conv2d forward;           ConvTranspose2d input-gradient)
 *       up   : lo -> hi   (ConvTranspose2d forward;  Conv2d input-gradient)
 *       wgrad: lo x hi -> dw (both weight-gradients)
 */
#ifndef DISVAE_B200_H
#define DISVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum DvStatus {
  DV_OK = 0,
  DV_ERR_BAD_SHAPE = -1,     /* unsupported size / channel count */
  DV_ERR_BAD_ARG = -2,       /* null pointer, bad enum */
  DV_ERR_WORKSPACE = -3,     /* workspace too small */
  DV_ERR_CUDA = -4,          /* a CUDA runtime call failed; see dv_last_cuda_error() */
  DV_ERR_ARCH = -5           /* device is not sm_100 */
} DvStatus;

enum { DV_ACT_NONE = 0, DV_ACT_RELU = 1, DV_ACT_SIGMOID = 2, DV_ACT_LEAKY = 3 };
enum { DV_DIST_BERNOULLI = 0, DV_DIST_GAUSSIAN = 1, DV_DIST_LAPLACE = 2 };

/* ---- library probes ------------------------------------------------------------- */
int dv_version(void);                 /* 10000*major + 100*minor + patch */
int dv_built_arch(void);              /* 100 == compiled for sm_100a */
const char* dv_status_string(int status);
int dv_last_cuda_error(void);         /* cudaError_t of the last failing call on this thread */
int dv_device_check(void);            /* DV_OK if the current device is compute capability 10.x */
/* how many kernels this library has launched in this process (for bench.py "gpu_launches") */
long long dv_launch_count(void);

/* ---- convolutions ---------------------------------------------------------------
 * Replaces: nn.Conv2d forward  (disvae/models/encoders.py:73-77)   -> dv_conv_down
 *           nn.ConvTranspose2d forward (disvae/models/decoders.py:77-82) -> dv_conv_up
 *           their autograd backward (disvae/training.py:157, aten::convolution_backward)
 *           -> dv_conv_up / dv_conv_down (input grads), dv_conv_wgrad (+ bias grads).
 * w is the torch weight tensor itself, [32][CH][4][4] contiguous; w_packed is produced by
 * dv_conv_pack_weights (layout private to the library).  B images, lo is H x W.
 * hi_nchw != 0: hi is [B,CH,2H,2W] (model boundary; required for CH in {1,3});
 * hi_nchw == 0: hi is [B,2H,2W,CH] (required for CH == 32).
 * mask (optional, same shape/layout as the OUTPUT): out *= (mask > 0) -- the ReLU backward
 * of the layer that produced `mask`, fused into this epilogue.
 */
size_t dv_conv_packed_floats(int CH);                         /* size of w_packed in floats */
int dv_conv_pack_weights(const float* w, float* w_packed, int CH, void* stream);
/* the same for n layers in ONE launch: w[i], w_packed[i] (dv_conv_packed_floats(CH[i]) floats each) and CH are HOST
 * arrays of device pointers / channel counts (the conv layers of an encoder or decoder node, whose weights only
 * change in the optimizer step). */
int dv_conv_pack_multi(int n, const void* const* w, void* const* w_packed, const int* CH, void* stream);
/* lo = act(down(hi) + bias) * [mask>0];  bias may be NULL.  act in {NONE, RELU}.
 * colsum_out (optional, [32]): sum of the stored `lo` over all pixels, accumulated in the kernel's
 * epilogue.  In the backward pass `lo` is the gradient reaching the previous ConvTranspose2d's
 * output, so this IS that layer's bias gradient (no extra pass over the tensor).
 * colsum_workspace: dv_channel_sum_workspace_bytes() bytes, required with colsum_out.
 * ReLU masks as bits (both optional, 32-channel NHWC outputs only, one 32-bit word per OUTPUT pixel, bit c =
 * channel c):  relu_bits_out receives [out > 0] of the stored tensor from the same epilogue -- kept by the
 * caller, it is the mask of the backward pass through the ReLU that follows this layer;  mask_bits is that
 * word form of `mask` (which must be passed as well: the CUDA-core fallbacks read the floats): the backward
 * epilogue then reads 4 bytes per pixel instead of 128 (autograd's threshold_backward, fused and compressed). */
int dv_conv_down(const float* hi, const float* w_packed, const float* bias, const float* mask,
                 float* lo, int B, int H, int W, int CH, int hi_nchw, int act, float* colsum_out,
                 void* colsum_workspace, const unsigned* mask_bits, unsigned* relu_bits_out, void* stream);
/* hi = act(up(lo) + bias) * [mask>0];  bias[CH] may be NULL.  act in {NONE, RELU, SIGMOID}.
 * mask_bits / relu_bits_out as above (CH == 32 only). */
int dv_conv_up(const float* lo, const float* w_packed, const float* bias, const float* mask,
               float* hi, int B, int H, int W, int CH, int hi_nchw, int act, const unsigned* mask_bits,
               unsigned* relu_bits_out, void* stream);
/* dw[32][CH][4][4] = sum_pixels lo (x) patch(hi);  dbias_lo[32] (optional) = sum_pixels lo.
 * Deterministic split-K: partials go to `workspace`, reduced in a fixed order. */
size_t dv_conv_wgrad_workspace_bytes(int B, int H, int W, int CH);
int dv_conv_wgrad(const float* lo, const float* hi, float* dw, float* dbias_lo, void* workspace,
                  size_t workspace_bytes, int B, int H, int W, int CH, int hi_nchw, void* stream);
/* out[C] = sum over all pixels of x; x is [rows, C] pixel-major (nchw == 0) or
 * [B, C, HW] (nchw != 0, rows = B, hw given).  Bias gradients of the ConvTranspose2d layers. */
size_t dv_channel_sum_workspace_bytes(void);
int dv_channel_sum(const float* x, float* out, long long rows, int C, int nchw, int hw,
                   void* workspace, void* stream);
/* [B,32,4,4] <-> [B,4,4,32] re-ordering at the conv/linear seam (encoders.py:80, decoders.py:74) */
int dv_flat_transpose(const float* src, float* dst, int B, int C, int S, int to_nhwc, void* stream);
/* g = dy * act'(y) for y = act(.) : sigmoid (decoders.py:82) backward.  n elements. */
int dv_act_bwd(const float* dy, const float* y, float* g, long long n, int act, float slope, void* stream);

/* ---- fully connected -------------------------------------------------------------
 * Replaces nn.Linear + activation: encoders.py:81-86, decoders.py:71-73, discriminator.py:63-68.
 * x[M,K], w[N,K] (torch layout), y[M,N].  slope is the LeakyReLU negative slope.
 */
/* Shapes whose activation rows are 16-byte pitched run on the tensor cores (tcgen05, 3xTF32) and need
 * a scratch buffer for the hi/lo split weight planes; the query returns 0 when the FFMA path is used
 * (workspace may then be NULL). */
size_t dv_linear_fwd_workspace_bytes(int M, int N, int K);
size_t dv_linear_dgrad_workspace_bytes(int M, int N, int K);
int dv_linear_fwd(const float* x, const float* w, const float* bias, float* y, int M, int N, int K,
                  int act, float slope, void* workspace, void* stream);
/* dx[M,K] = (g[M,N] . w[N,K]) * act'(mask_src[M,K]); mask_src is the POST-activation output
 * of the previous layer (NULL: no mask); act in {NONE, RELU, LEAKY}. */
int dv_linear_dgrad(const float* g, const float* w, const float* mask_src, float* dx, int M, int N,
                    int K, int act, float slope, void* workspace, void* stream);
/* dw[N,K] = g^T . x ; dbias[N] = column sums of g (may be NULL).  Small N*K problems are split over
 * the batch (deterministic split-K); workspace may be NULL when the query returns 0. */
size_t dv_linear_wgrad_workspace_bytes(int M, int N, int K);
int dv_linear_wgrad(const float* g, const float* x, float* dw, float* dbias, int M, int N, int K,
                    void* workspace, void* stream);
/* Pre-packed weights: dv_linear_pack_multi splits n weight matrices (w[i]: [N[i], K[i]], HOST arrays of device
 * pointers / sizes) into the tcgen05 hi/lo operand planes of BOTH directions in one launch, into caller-owned
 * buffers of dv_linear_packed_floats(N, K) floats (16-byte aligned); dv_linear_fwd_packed / dv_linear_dgrad_packed are
 * dv_linear_fwd / dv_linear_dgrad on those planes (w is still passed: shapes that run on the CUDA cores read it).
 * One pack launch per network node per step instead of one per layer per direction. */
size_t dv_linear_packed_floats(int N, int K);
int dv_linear_pack_multi(int n, const void* const* w, void* const* packed, const int* N, const int* K, void* stream);
int dv_linear_fwd_packed(const float* x, const float* w, const float* packed, const float* bias, float* y,
                         int M, int N, int K, int act, float slope, void* stream);
int dv_linear_dgrad_packed(const float* g, const float* w, const float* packed, const float* mask_src, float* dx,
                           int M, int N, int K, int act, float slope, void* stream);

/* ---- reparameterised sampling ------------------------------------------------------
 * Replaces VAE.reparameterize (disvae/models/vae.py:65-68): z = mu + exp(0.5*logvar)*eps.
 * mu/logvar are read with an element stride (`ld`), rows are `row_stride` apart, so the
 * interleaved encoder output (encoders.py:86-87) can be consumed in place.
 * eps_in != NULL: use the caller's noise (parity tests).  eps_in == NULL: Philox4x32-10 +
 * Box-Muller on the device, keyed by (seed, *offset_dev + element index); the kernel then
 * advances *offset_dev by B*D (graph-replay safe).  eps_out (optional) receives the noise.
 */
int dv_reparam_fwd(const float* mu, const float* logvar, int ld, int row_stride, const float* eps_in,
                   unsigned long long seed, unsigned long long* offset_dev, float* z, float* eps_out,
                   int B, int D, void* stream);
/* g_mu = g_z ; g_logvar = g_z * eps * 0.5 * exp(0.5*logvar)  (contiguous [B,D] outputs) */
int dv_reparam_bwd(const float* g_z, const float* logvar, int ld, int row_stride, const float* eps,
                   float* g_mu, float* g_logvar, int B, int D, void* stream);

/* ---- fused reconstruction loss + analytic KL ---------------------------------------
 * Replaces _reconstruction_loss (losses.py:394-449) and _kl_normal_loss (losses.py:452-480)
 * in ONE launch.  out[0] = reconstruction loss (sum / B, all three distributions, traps
 * T8/T9), out[1] = total KL, out[2+d] = per-dimension KL (the kl_loss_<d> log entries).
 * recon/data: n_img_elems = C*H*W per image, any layout (same for both).
 */
/* workspace: zero on first use (the kernel leaves its counter at zero). */
size_t dv_vae_loss_workspace_bytes(int B, long long n_img_elems);
int dv_vae_loss_fwd(const float* recon, const float* data, long long n_img_elems, int B, int dist,
                    const float* mu, const float* logvar, int ld, int row_stride, int D,
                    float* out, void* workspace, void* stream);
/* upstream = device float[2]: d loss / d out[0], d loss / d out[1].  g_recon like recon;
 * g_mu, g_logvar contiguous [B,D].  Any output pointer may be NULL. */
int dv_vae_loss_bwd(const float* recon, const float* data, long long n_img_elems, int B, int dist,
                    const float* mu, const float* logvar, int ld, int row_stride, int D,
                    const float* fwd_out, const float* upstream, float* g_recon, float* g_mu,
                    float* g_logvar, void* stream);

/* ---- beta-TCVAE log-density decomposition --------------------------------------------
 * Replaces _get_log_pz_qz_prodzi_qzCx (losses.py:523-544) + matrix_log_density_gaussian /
 * log_importance_weight_matrix (disvae/utils/math.py:8-73) + the three means at
 * losses.py:369-373.  Nothing B x B (x D) is ever materialised; the importance-weight
 * matrix is evaluated analytically from its column structure (trap T3), D-fold in log_qz
 * (trap T2).  is_mss == 0 reproduces the reference's is_mss=False branch (trap T4).
 * rowstats is a structure of arrays [4 + D][B]: rows log_pz, log_qz, log_prod_qzi, log_q_zCx,
 * then the per-dimension logsumexp P[d][i] (kept for the backward).  terms[3] = mi, tc, dw_kl.
 * `workspace` must be 16-byte aligned and its first 64 bytes zero on first use (the kernel
 * leaves them zero); it holds the per-column parameters the backward re-reads, so pass the
 * SAME workspace (untouched) and rowstats to dv_btcvae_bwd.
 */
size_t dv_btcvae_workspace_bytes(int B, int D);
int dv_btcvae_fwd(const float* z, const float* mu, const float* logvar, int ld, int row_stride,
                  int B, int D, long long n_data, int is_mss, float* rowstats, float* terms,
                  void* workspace, void* stream);
/* g_terms = device float[3] (d loss / d mi, tc, dw_kl); outputs contiguous [B,D] (any may be NULL). */
int dv_btcvae_bwd(int B, int D, long long n_data, int is_mss, const float* rowstats,
                  const void* workspace, const float* g_terms, float* g_z, float* g_mu,
                  float* g_logvar, void* stream);
/* Row-window form (the estimator over a batch all-gathered from several GPUs, each rank
 * evaluating its own rows of the B x B matrix -- losses.py:523-544 semantics of the GLOBAL batch).
 * z / mu / logvar hold all B rows; only rows [row0, row0 + nrows) are evaluated: their rowstats entries are
 * written (global row index), terms = means over the window.  Backward: g_z is [nrows, D] (the window's rows);
 * g_mu / g_logvar are [B, D] PARTIAL sums over the window's rows for every column -- summing them over the
 * windows of all ranks (reduce-scatter) gives the gradient of the mean-over-ranks loss.
 * dv_btcvae_fwd / dv_btcvae_bwd are the (0, B) window. */
int dv_btcvae_fwd_rows(const float* z, const float* mu, const float* logvar, int ld, int row_stride,
                       int B, int D, int row0, int nrows, long long n_data, int is_mss, float* rowstats,
                       float* terms, void* workspace, void* stream);
int dv_btcvae_bwd_rows(int B, int D, int row0, int nrows, long long n_data, int is_mss,
                       const float* rowstats, const void* workspace, const float* g_terms, float* g_z,
                       float* g_mu, float* g_logvar, void* stream);

/* ---- input pipeline and step glue -----------------------------
 * dv_u8_to_f32: dst[i] = src[i] / 255 (true division) -- torchvision ToTensor on the device, so host batches can be
 * uploaded as bytes (training.py:150; utils/datasets.py:182,247,364-367).  Both pointers 16-byte aligned.
 */
int dv_u8_to_f32(const unsigned char* src, float* dst, long long n, void* stream);
/* loss[0] = sum_{i<na} coef_a[i]*a[i] + sum_{j<nb} coef_b[j]*b[j]; a, b device vectors, coef_* HOST arrays (<= 8 each,
 * passed to the kernel by value).  The scalar combinations of losses.py:151 (rec + anneal*beta*kl), :199-200 is not
 * covered (|kl - C|), :381-382 (rec + alpha*mi + beta*tc + anneal*gamma*dw_kl).  Backward: g_a[0..na_total) (zeros past
 * na), g_b[0..nb) from the upstream scalar g[0]. */
int dv_loss_combine_fwd(const float* a, const float* coef_a, int na, const float* b, const float* coef_b, int nb,
                        float* loss, void* stream);
int dv_loss_combine_bwd(const float* g, const float* coef_a, int na, int na_total, const float* coef_b, int nb,
                        float* g_a, float* g_b, void* stream);
/* g = dy * act'(y) over an NCHW tensor [B, C <= 4, hw] fused with chansum[c] = sum_{b,hw} g (the bias gradient of the
 * ConvTranspose2d that produced y; decoders.py:82).  workspace: dv_channel_sum_workspace_bytes(). */
int dv_act_bwd_chansum(const float* dy, const float* y, float* g, int B, int C, int hw, int act, float slope,
                       float* chansum, void* workspace, void* stream);

/* ---- disentanglement metrics: marginal-entropy estimator ---------------------
 * Replaces Evaluator._estimate_latent_entropies (disvae/evaluate.py:233-297), the inner loop of the MIG / AAM metrics
 * (:119-161, :299-317): for S samples zs[D][S] (row d = samples of latent dimension d) and the N posteriors
 * mean/logvar[n][d] (element (n,d) at n*row_stride + d*ld),
 *   H[d] = -(1/S) sum_s ( -log N + logsumexp_n log N(zs[d][s]; mean[n][d], exp(logvar[n][d])) ).
 * logq_out (optional, [D][S]) receives the per-sample log q(z).  Deterministic (fixed merge order).
 */
size_t dv_latent_entropy_workspace_bytes(int N, int D, int S);
int dv_latent_entropy(const float* zs, const float* mean, const float* logvar, int ld, int row_stride,
                      int N, int D, int S, float* H, float* logq_out, void* workspace, void* stream);

/* ---- FactorVAE pieces ---------------------------------------------------------------------
 * dv_permute_dims replaces _permute_dims (losses.py:483-508): out[b][d] = z[perm[d][b]][d].
 * perms != NULL: int64 [D][B] (the reference's CPU randperm stream, trap T7).  perms == NULL:
 * per-dimension Philox-keyed random permutation generated on the device (B <= 4096).
 */
int dv_permute_dims(const float* z, const long long* perms, unsigned long long seed,
                    unsigned long long* offset_dev, float* out, int B, int D, void* stream);
/* tc[0] = mean(d_z[:,0] - d_z[:,1])  (losses.py:265) */
int dv_factor_tc_fwd(const float* d_z, int h, float* tc, void* stream);
int dv_factor_tc_bwd(const float* upstream, int h, float* g_d_z, void* stream);
/* out[0] = 0.5*(CE(d_z, 0) + CE(d_perm, 1))  (losses.py:293-295) */
int dv_factor_ce_fwd(const float* d_z, const float* d_perm, int h, float* out, void* stream);
int dv_factor_ce_bwd(const float* d_z, const float* d_perm, const float* upstream, int h,
                     float* g_d_z, float* g_d_perm, void* stream);

/* ---- optimiser -------------------------------------------------------
 * torch.optim.Adam semantics (main.py:208, losses.py:238): eps outside the sqrt, no weight
 * decay, bias correction from *step_dev (float, incremented by the kernel).  Operates on one
 * flat fp32 buffer so a whole model is one launch.  betas are doubles: 1-beta and the bias corrections are
 * evaluated in fp64 like torch's Python-side arithmetic.  grad_scale multiplies the gradient
 * (1/world_size after a sum-allreduce). */
int dv_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                 float* step_dev, long long n, float lr, double beta1, double beta2, float eps,
                 float grad_scale, void* stream);

/* Multi-tensor form: `count` (<= dv_adam_multi_max_tensors()) parameter tensors updated by ONE launch.
 * The arrays are HOST arrays of device pointers / element counts; they are copied into kernel-parameter
 * space, nothing is retained.  Same arithmetic and step counter semantics as dv_adam_step. */
int dv_adam_multi_max_tensors(void);
int dv_adam_multi(int count, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const long long* numel, float* step_dev, float lr, double beta1,
                  double beta2, float eps, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DISVAE_B200_H */
