// Fused reconstruction-loss + analytic-KL kernel, reparameterised sampling, Adam.
// Reference call sites: disvae/models/losses.py:394-449 (_reconstruction_loss),
// losses.py:452-480 (_kl_normal_loss), disvae/models/vae.py:65-68 (reparameterize),
// main.py:208 / losses.py:238 (optim.Adam).
#include "dv_common.cuh"

namespace dv {

constexpr int kLossBlocks = 2 * kNumSMs;
constexpr int kLossThreads = 256;
// workspace layout (floats): [0] counter (as unsigned), [1..kLossBlocks] recon partials,
// [1+kLossBlocks .. +1024) per-dimension KL
constexpr int kLossWsFloats = 1 + kLossBlocks + 1024;

__device__ __forceinline__ float block_sum(float v, float* red /*>= 8 floats*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.f;
    t = warp_sum(t);
  }
  return t;   // valid in warp 0
}

__device__ __forceinline__ float recon_elem(float r, float x, int dist) {
  if (dist == DV_DIST_BERNOULLI) {
    // aten binary_cross_entropy: (t - 1) * max(log1p(-i), -100) - t * max(log(i), -100)   (trap T8)
    const float l1 = fmaxf(log1pf(-r), -100.f), l0 = fmaxf(logf(r), -100.f);
    return (x - 1.f) * l1 - x * l0;
  } else if (dist == DV_DIST_GAUSSIAN) {
    const float d = r * 255.f - x * 255.f;
    return d * d;
  } else {
    return fabsf(r - x);
  }
}

__global__ void __launch_bounds__(kLossThreads)
vae_loss_fwd_kernel(const float* __restrict__ recon, const float* __restrict__ data, long long n, int B, int dist,
                    const float* __restrict__ mu, const float* __restrict__ logvar, int ld, int row_stride, int D,
                    float* __restrict__ out, float* __restrict__ ws) {
  __shared__ float red[8];
  __shared__ bool is_last;
  unsigned* counter = reinterpret_cast<unsigned*>(ws);
  float* partial = ws + 1;
  float* klp = ws + 1 + kLossBlocks;

  // (1) reconstruction partial sum of this block
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* r4 = reinterpret_cast<const float4*>(recon);
  const float4* x4 = reinterpret_cast<const float4*>(data);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 r = __ldg(r4 + i), x = __ldg(x4 + i);
    s += recon_elem(r.x, x.x, dist) + recon_elem(r.y, x.y, dist) + recon_elem(r.z, x.z, dist) + recon_elem(r.w, x.w, dist);
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += recon_elem(recon[i], data[i], dist);
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;

  // (2) block d (< D) also reduces latent dimension d over the batch: 0.5*(-1 - lv + mu^2 + e^lv)
  for (int d = blockIdx.x; d < D; d += gridDim.x) {
    float k = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
      const float m = mu[(long long)b * row_stride + (long long)d * ld];
      const float lv = logvar[(long long)b * row_stride + (long long)d * ld];
      k += -1.f - lv + m * m + expf(lv);
    }
    k = block_sum(k, red);
    if (threadIdx.x == 0) klp[d] = 0.5f * k / (float)B;
  }

  // (3) last block to finish combines everything in a fixed order
  __threadfence();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  if (threadIdx.x < 32) {
    float t = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) t += partial[i];
    t = warp_sum(t);
    float kl = 0.f;
    for (int d = threadIdx.x; d < D; d += 32) { const float v = klp[d]; out[2 + d] = v; kl += v; }
    kl = warp_sum(kl);
    if (threadIdx.x == 0) {
      float loss = t;
      if (dist == DV_DIST_GAUSSIAN) loss = loss / 255.f;
      else if (dist == DV_DIST_LAPLACE) { loss = loss * 3.f; loss = loss * (loss != 0.f ? 1.f : 0.f); }
      out[0] = loss / (float)B;
      out[1] = kl;
      *counter = 0u;
    }
  }
}

__global__ void __launch_bounds__(256)
vae_loss_bwd_kernel(const float* __restrict__ recon, const float* __restrict__ data, long long n, int B, int dist,
                    const float* __restrict__ mu, const float* __restrict__ logvar, int ld, int row_stride, int D,
                    const float* __restrict__ fwd_out, const float* __restrict__ upstream,
                    float* __restrict__ g_recon, float* __restrict__ g_mu, float* __restrict__ g_logvar) {
  const float g_rec = upstream[0] / (float)B;
  const float g_kl = upstream[1] / (float)B;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  if (g_recon) {
    float scale = g_rec;
    if (dist == DV_DIST_LAPLACE) scale = (fwd_out[0] != 0.f) ? g_rec * 3.f : 0.f;
    for (long long i = tid; i < n; i += nthreads) {
      const float r = recon[i], x = data[i];
      float g;
      if (dist == DV_DIST_BERNOULLI) g = scale * (r - x) / fmaxf((1.f - r) * r, 1e-12f);   // aten bce backward
      else if (dist == DV_DIST_GAUSSIAN) g = scale * (2.f * (r * 255.f - x * 255.f));       // 255 (chain) / 255 (norm)
      else g = scale * ((r > x) ? 1.f : ((r < x) ? -1.f : 0.f));
      g_recon[i] = g;
    }
  }
  if (g_mu || g_logvar) {
    const long long nz = (long long)B * D;
    for (long long i = tid; i < nz; i += nthreads) {
      const int b = (int)(i / D), d = (int)(i % D);
      const float m = mu[(long long)b * row_stride + (long long)d * ld];
      const float lv = logvar[(long long)b * row_stride + (long long)d * ld];
      if (g_mu) g_mu[i] = g_kl * m;
      if (g_logvar) g_logvar[i] = g_kl * 0.5f * (expf(lv) - 1.f);
    }
  }
}

// ---- reparameterisation -----------------------------------------------------------
__global__ void reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ logvar, int ld, int row_stride,
                                   const float* __restrict__ eps_in, unsigned long long seed,
                                   const unsigned long long* __restrict__ offset_dev, float* __restrict__ z,
                                   float* __restrict__ eps_out, int B, int D) {
  const long long n = (long long)B * D;
  const unsigned long long off = (eps_in == nullptr && offset_dev) ? *offset_dev : 0ull;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / D), d = (int)(i % D);
    float e;
    if (eps_in) e = eps_in[i];
    else {
      const unsigned long long c = off + (unsigned long long)i;
      const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      const float u1 = u32_to_unit_open(r.x), u2 = u32_to_unit_open(r.y);
      e = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);          // Box-Muller
    }
    const float m = mu[(long long)b * row_stride + (long long)d * ld];
    const float lv = logvar[(long long)b * row_stride + (long long)d * ld];
    z[i] = m + expf(0.5f * lv) * e;
    if (eps_out) eps_out[i] = e;
  }
}
__global__ void advance_offset_kernel(unsigned long long* offset_dev, unsigned long long by) { *offset_dev += by; }

__global__ void reparam_bwd_kernel(const float* __restrict__ g_z, const float* __restrict__ logvar, int ld, int row_stride,
                                   const float* __restrict__ eps, float* __restrict__ g_mu, float* __restrict__ g_logvar,
                                   int B, int D) {
  const long long n = (long long)B * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / D), d = (int)(i % D);
    const float g = g_z[i];
    const float lv = logvar[(long long)b * row_stride + (long long)d * ld];
    if (g_mu) g_mu[i] = g;
    if (g_logvar) g_logvar[i] = g * eps[i] * (0.5f * expf(0.5f * lv));
  }
}

// ---- Adam (torch.optim.Adam, amsgrad=False, weight_decay=0, maximize=False) ---------
// torch evaluates 1-beta and the bias corrections in Python doubles; 1-0.999 in fp32 would be off by 1.3e-5
struct AdamCoef { float omb1, omb2, b2, step_size, bc2_sqrt; };
__device__ __forceinline__ AdamCoef adam_coef(const float* step_dev, float lr, double b1, double b2) {
  const double step = (double)(*step_dev) + 1.0;
  const double bc1 = 1.0 - pow(b1, step), bc2 = 1.0 - pow(b2, step);
  AdamCoef c;
  c.omb1 = (float)(1.0 - b1); c.omb2 = (float)(1.0 - b2); c.b2 = (float)b2;
  c.step_size = (float)((double)lr / bc1);
  c.bc2_sqrt = (float)sqrt(bc2);
  return c;
}
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            const float* __restrict__ step_dev, long long n, float lr, double b1d, double b2d, float eps, float gscale) {
  const AdamCoef c = adam_coef(step_dev, lr, b1d, b2d);
  const float b2 = c.b2, step_size = c.step_size, bc2_sqrt = c.bc2_sqrt;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gscale;
    const float mi = m[i] + (gi - m[i]) * c.omb1;              // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * b2 + c.omb2 * (gi * gi);           // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}
__global__ void bump_step_kernel(float* step_dev) { *step_dev += 1.f; }

// ---- multi-tensor Adam: every parameter tensor of a model in ONE launch -------------------------
// The tensor table travels by value in kernel-parameter space (no device-side table to upload, so the
// per-step gradient tensors handed out by autograd can be consumed where they are).
constexpr int kAdamMaxTensors = 48;
constexpr int kAdamChunk = 4096;                 // elements per block-iteration
struct AdamTable {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  int chunk_begin[kAdamMaxTensors + 1];          // prefix sum of ceil(n / kAdamChunk)
  int n[kAdamMaxTensors];
  int count;
};
__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ AdamTable t, const float* __restrict__ step_dev, float lr, double b1d, double b2d,
                  float eps, float gscale) {
  const AdamCoef c = adam_coef(step_dev, lr, b1d, b2d);
  const float b2 = c.b2, step_size = c.step_size, bc2_sqrt = c.bc2_sqrt;
  const int total_chunks = t.chunk_begin[t.count];
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int ti = 0;
    while (chunk >= t.chunk_begin[ti + 1]) ++ti;               // <= 48 entries, block-uniform
    const int base = (chunk - t.chunk_begin[ti]) * kAdamChunk;
    const int end = min(t.n[ti], base + kAdamChunk);
    float* p = t.p[ti]; const float* g = t.g[ti]; float* m = t.m[ti]; float* v = t.v[ti];
    for (int i = base + threadIdx.x; i < end; i += blockDim.x) {
      const float gi = g[i] * gscale;
      const float mi = m[i] + (gi - m[i]) * c.omb1;
      const float vi = v[i] * b2 + c.omb2 * (gi * gi);
      m[i] = mi; v[i] = vi;
      p[i] = p[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
  }
}

static int grid_for(long long n, int per_block, int max_blocks) {
  long long g = (n + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

}  // namespace dv

using namespace dv;

extern "C" {

size_t dv_vae_loss_workspace_bytes(int, long long) { return (size_t)kLossWsFloats * sizeof(float); }

int dv_vae_loss_fwd(const float* recon, const float* data, long long n_img_elems, int B, int dist,
                    const float* mu, const float* logvar, int ld, int row_stride, int D,
                    float* out, void* workspace, void* stream) {
  if (!recon || !data || !mu || !logvar || !out || !workspace) return DV_ERR_BAD_ARG;
  if (B <= 0 || n_img_elems <= 0 || D <= 0 || D > 1024) return DV_ERR_BAD_SHAPE;
  if (dist < DV_DIST_BERNOULLI || dist > DV_DIST_LAPLACE) return DV_ERR_BAD_ARG;
  if (((uintptr_t)recon | (uintptr_t)data) & 15) return DV_ERR_BAD_ARG;
  const long long n = n_img_elems * B;
  int grid = grid_for(n / 4, kLossThreads * 4, kLossBlocks);
  vae_loss_fwd_kernel<<<grid, kLossThreads, 0, as_stream(stream)>>>(recon, data, n, B, dist, mu, logvar, ld, row_stride, D,
                                                                  out, reinterpret_cast<float*>(workspace));
  return check_launch();
}

int dv_vae_loss_bwd(const float* recon, const float* data, long long n_img_elems, int B, int dist,
                    const float* mu, const float* logvar, int ld, int row_stride, int D,
                    const float* fwd_out, const float* upstream, float* g_recon, float* g_mu, float* g_logvar, void* stream) {
  if (!recon || !data || !mu || !logvar || !fwd_out || !upstream) return DV_ERR_BAD_ARG;
  if (B <= 0 || n_img_elems <= 0 || D <= 0) return DV_ERR_BAD_SHAPE;
  const long long n = n_img_elems * B;
  vae_loss_bwd_kernel<<<grid_for(n, 1024, 8 * kNumSMs), 256, 0, as_stream(stream)>>>(
      recon, data, n, B, dist, mu, logvar, ld, row_stride, D, fwd_out, upstream, g_recon, g_mu, g_logvar);
  return check_launch();
}

int dv_reparam_fwd(const float* mu, const float* logvar, int ld, int row_stride, const float* eps_in,
                   unsigned long long seed, unsigned long long* offset_dev, float* z, float* eps_out,
                   int B, int D, void* stream) {
  if (!mu || !logvar || !z) return DV_ERR_BAD_ARG;
  if (!eps_in && !offset_dev) return DV_ERR_BAD_ARG;
  if (B <= 0 || D <= 0) return DV_ERR_BAD_SHAPE;
  const long long n = (long long)B * D;
  reparam_fwd_kernel<<<grid_for(n, 256, 2 * kNumSMs), 256, 0, as_stream(stream)>>>(mu, logvar, ld, row_stride, eps_in, seed,
                                                                                 offset_dev, z, eps_out, B, D);
  int rc = check_launch();
  if (rc != DV_OK || eps_in) return rc;
  advance_offset_kernel<<<1, 1, 0, as_stream(stream)>>>(offset_dev, (unsigned long long)n);
  return check_launch();
}

int dv_reparam_bwd(const float* g_z, const float* logvar, int ld, int row_stride, const float* eps,
                   float* g_mu, float* g_logvar, int B, int D, void* stream) {
  if (!g_z || !logvar || !eps) return DV_ERR_BAD_ARG;
  if (B <= 0 || D <= 0) return DV_ERR_BAD_SHAPE;
  const long long n = (long long)B * D;
  reparam_bwd_kernel<<<grid_for(n, 256, 2 * kNumSMs), 256, 0, as_stream(stream)>>>(g_z, logvar, ld, row_stride, eps, g_mu,
                                                                                 g_logvar, B, D);
  return check_launch();
}

int dv_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step_dev, long long n,
                 float lr, double beta1, double beta2, float eps, float grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !step_dev) return DV_ERR_BAD_ARG;
  if (n <= 0) return DV_ERR_BAD_SHAPE;
  adam_kernel<<<grid_for(n, 1024, 4 * kNumSMs), 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, step_dev, n, lr,
                                                                           beta1, beta2, eps, grad_scale);
  int rc = check_launch();
  if (rc != DV_OK) return rc;
  bump_step_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev);
  return check_launch();
}

int dv_adam_multi_max_tensors(void) { return kAdamMaxTensors; }

int dv_adam_multi(int count, float* const* params, const float* const* grads, float* const* exp_avg,
                  float* const* exp_avg_sq, const long long* numel, float* step_dev, float lr, double beta1,
                  double beta2, float eps, float grad_scale, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || !step_dev) return DV_ERR_BAD_ARG;
  if (count <= 0 || count > kAdamMaxTensors) return DV_ERR_BAD_SHAPE;
  AdamTable t;
  t.count = count;
  t.chunk_begin[0] = 0;
  for (int i = 0; i < count; ++i) {
    if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i] || numel[i] <= 0 || numel[i] >= (1LL << 31)) return DV_ERR_BAD_ARG;
    t.p[i] = params[i]; t.g[i] = grads[i]; t.m[i] = exp_avg[i]; t.v[i] = exp_avg_sq[i];
    t.n[i] = (int)numel[i];
    t.chunk_begin[i + 1] = t.chunk_begin[i] + (int)((numel[i] + kAdamChunk - 1) / kAdamChunk);
  }
  int grid = t.chunk_begin[count];
  if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
  adam_multi_kernel<<<grid, 256, 0, as_stream(stream)>>>(t, step_dev, lr, beta1, beta2, eps, grad_scale);
  int rc = check_launch();
  if (rc != DV_OK) return rc;
  bump_step_kernel<<<1, 1, 0, as_stream(stream)>>>(step_dev);
  return check_launch();
}

}  // extern "C"
