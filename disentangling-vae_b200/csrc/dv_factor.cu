// FactorVAE pieces: per-dimension batch permutation, TC estimate from the discriminator logits,
// two-class cross entropy.  Reference: disvae/models/losses.py:483-508 (_permute_dims),
// :265 (tc_loss), :291-295 (d_tc_loss).
#include "dv_common.cuh"

namespace dv {

constexpr int kPermMaxB = 4096;

// perms given: out[b][d] = z[perm[d][b]][d]
__global__ void permute_given_kernel(const float* __restrict__ z, const long long* __restrict__ perms,
                                     float* __restrict__ out, int B, int D) {
  const long long n = (long long)B * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / D), d = (int)(i % D);
    const long long src = perms[(long long)d * B + b];
    out[i] = z[src * D + d];
  }
}

// perms generated on device: block d sorts (philox key, index) pairs -> uniform random permutation
__global__ void __launch_bounds__(512)
permute_philox_kernel(const float* __restrict__ z, unsigned long long seed, const unsigned long long* __restrict__ offset_dev,
                      float* __restrict__ out, int B, int D, int npow2) {
  extern __shared__ unsigned long long keys[];
  const int d = blockIdx.x;
  const unsigned long long off = *offset_dev + (unsigned long long)d * (unsigned long long)B;
  for (int b = threadIdx.x; b < npow2; b += blockDim.x) {
    unsigned long long k = ~0ull;
    if (b < B) {
      const unsigned long long c = off + (unsigned long long)b;
      const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x5eedu, 0u),
                                    make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
      k = ((unsigned long long)r.x << 32) | (unsigned long long)(uint32_t)b;
    }
    keys[b] = k;
  }
  __syncthreads();
  for (int size = 2; size <= npow2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = threadIdx.x; t < (npow2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const unsigned long long a = keys[lo], b2 = keys[hi];
        if ((a > b2) == up) { keys[lo] = b2; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int src = (int)(keys[b] & 0xffffffffull);
    out[(long long)b * D + d] = z[(long long)src * D + d];
  }
}
__global__ void advance_offset2_kernel(unsigned long long* offset_dev, unsigned long long by) { *offset_dev += by; }

__global__ void __launch_bounds__(256) factor_tc_fwd_kernel(const float* __restrict__ d_z, int h, float* __restrict__ tc) {
  __shared__ float red[8];
  float s = 0.f;
  for (int b = threadIdx.x; b < h; b += blockDim.x) s += d_z[2 * b] - d_z[2 * b + 1];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; tc[0] = t / (float)h; }
}
__global__ void factor_tc_bwd_kernel(const float* __restrict__ upstream, int h, float* __restrict__ g) {
  const float u = upstream[0] / (float)h;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < h; b += gridDim.x * blockDim.x) { g[2 * b] = u; g[2 * b + 1] = -u; }
}

__device__ __forceinline__ float nll2(float x0, float x1, int target) {
  const float m = fmaxf(x0, x1);
  const float lse = m + logf(expf(x0 - m) + expf(x1 - m));
  return lse - (target == 0 ? x0 : x1);
}
__global__ void __launch_bounds__(256)
factor_ce_fwd_kernel(const float* __restrict__ d_z, const float* __restrict__ d_perm, int h, float* __restrict__ out) {
  __shared__ float red[2][8];
  float s0 = 0.f, s1 = 0.f;
  for (int b = threadIdx.x; b < h; b += blockDim.x) {
    s0 += nll2(d_z[2 * b], d_z[2 * b + 1], 0);
    s1 += nll2(d_perm[2 * b], d_perm[2 * b + 1], 1);
  }
  s0 = warp_sum(s0); s1 = warp_sum(s1);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float t0 = 0.f, t1 = 0.f;
    for (int w = 0; w < 8; ++w) { t0 += red[0][w]; t1 += red[1][w]; }
    out[0] = 0.5f * (t0 / (float)h + t1 / (float)h);
  }
}
__global__ void factor_ce_bwd_kernel(const float* __restrict__ d_z, const float* __restrict__ d_perm,
                                     const float* __restrict__ upstream, int h, float* __restrict__ g_z, float* __restrict__ g_p) {
  const float u = upstream[0] * 0.5f / (float)h;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < h; b += gridDim.x * blockDim.x) {
    {
      const float x0 = d_z[2 * b], x1 = d_z[2 * b + 1], m = fmaxf(x0, x1);
      const float e0 = expf(x0 - m), e1 = expf(x1 - m), inv = 1.f / (e0 + e1);
      if (g_z) { g_z[2 * b] = u * (e0 * inv - 1.f); g_z[2 * b + 1] = u * (e1 * inv); }
    }
    {
      const float x0 = d_perm[2 * b], x1 = d_perm[2 * b + 1], m = fmaxf(x0, x1);
      const float e0 = expf(x0 - m), e1 = expf(x1 - m), inv = 1.f / (e0 + e1);
      if (g_p) { g_p[2 * b] = u * (e0 * inv); g_p[2 * b + 1] = u * (e1 * inv - 1.f); }
    }
  }
}

}  // namespace dv

using namespace dv;

extern "C" {

int dv_permute_dims(const float* z, const long long* perms, unsigned long long seed, unsigned long long* offset_dev,
                    float* out, int B, int D, void* stream) {
  if (!z || !out) return DV_ERR_BAD_ARG;
  if (B <= 0 || D <= 0) return DV_ERR_BAD_SHAPE;
  cudaStream_t st = as_stream(stream);
  if (perms) {
    const long long n = (long long)B * D;
    int grid = (int)((n + 255) / 256); if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
    permute_given_kernel<<<grid, 256, 0, st>>>(z, perms, out, B, D);
    return check_launch();
  }
  if (!offset_dev) return DV_ERR_BAD_ARG;
  if (B > kPermMaxB) return DV_ERR_BAD_SHAPE;
  int npow2 = 2; while (npow2 < B) npow2 <<= 1;
  permute_philox_kernel<<<D, 512, npow2 * sizeof(unsigned long long), st>>>(z, seed, offset_dev, out, B, D, npow2);
  int rc = check_launch();
  if (rc != DV_OK) return rc;
  advance_offset2_kernel<<<1, 1, 0, st>>>(offset_dev, (unsigned long long)B * D);
  return check_launch();
}

int dv_factor_tc_fwd(const float* d_z, int h, float* tc, void* stream) {
  if (!d_z || !tc) return DV_ERR_BAD_ARG;
  if (h <= 0) return DV_ERR_BAD_SHAPE;
  factor_tc_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(d_z, h, tc);
  return check_launch();
}
int dv_factor_tc_bwd(const float* upstream, int h, float* g_d_z, void* stream) {
  if (!upstream || !g_d_z) return DV_ERR_BAD_ARG;
  if (h <= 0) return DV_ERR_BAD_SHAPE;
  factor_tc_bwd_kernel<<<(h + 255) / 256, 256, 0, as_stream(stream)>>>(upstream, h, g_d_z);
  return check_launch();
}
int dv_factor_ce_fwd(const float* d_z, const float* d_perm, int h, float* out, void* stream) {
  if (!d_z || !d_perm || !out) return DV_ERR_BAD_ARG;
  if (h <= 0) return DV_ERR_BAD_SHAPE;
  factor_ce_fwd_kernel<<<1, 256, 0, as_stream(stream)>>>(d_z, d_perm, h, out);
  return check_launch();
}
int dv_factor_ce_bwd(const float* d_z, const float* d_perm, const float* upstream, int h, float* g_d_z, float* g_d_perm, void* stream) {
  if (!d_z || !d_perm || !upstream) return DV_ERR_BAD_ARG;
  if (h <= 0) return DV_ERR_BAD_SHAPE;
  factor_ce_bwd_kernel<<<(h + 255) / 256, 256, 0, as_stream(stream)>>>(d_z, d_perm, upstream, h, g_d_z, g_d_perm);
  return check_launch();
}

}  // extern "C"
