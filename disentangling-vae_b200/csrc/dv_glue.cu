// Small fused "glue" kernels that replace chains of tiny framework launches on the training step
// (VERDICT r1 #8 "launch and glue diet") and the uint8 input path.
//
//   dv_u8_to_f32          uint8 image batch -> fp32 / 255  (torchvision ToTensor semantics, utils/datasets.py:182,247,
//                         364-367: `img.float().div(255)`), so a host batch travels over PCIe as bytes (4x less H2D)
//   dv_loss_combine_*     loss = sum_i ca[i]*a[i] + sum_j cb[j]*b[j] for two short device vectors (the fused loss kernel's
//                         (rec, kl, ...) and the beta-TCVAE (mi, tc, dw_kl)): losses.py:151, 199-200, 381-382 as ONE
//                         launch forward and ONE backward instead of ~10 scalar mul/add/select kernels and their
//                         zero-filled gradient buffers
//   dv_act_bwd_chansum    ConvTranspose2d output layer backward prologue: g = dy * act'(y) (decoders.py:82 sigmoid) fused
//                         with the per-channel sum of g (that layer's bias gradient) -- one pass instead of two
#include "dv_common.cuh"

namespace dv {

__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, long long n) {
  const long long n16 = n >> 4;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = s4[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float4 o;
      o.x = (float)(w[k] & 0xffu) / 255.0f;                    // true division, like Tensor.div(255)
      o.y = (float)((w[k] >> 8) & 0xffu) / 255.0f;
      o.z = (float)((w[k] >> 16) & 0xffu) / 255.0f;
      o.w = (float)(w[k] >> 24) / 255.0f;
      d4[4 * i + k] = o;
    }
  }
  for (long long i = (n16 << 4) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (float)src[i] / 255.0f;
}

struct Coefs { float a[8]; float b[8]; };

__global__ void loss_combine_fwd_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, Coefs c,
                                        float* __restrict__ loss) {
  if (threadIdx.x != 0) return;
  float s = 0.f;
  for (int i = 0; i < na; ++i) s += c.a[i] * a[i];             // fixed order: rec first, like rec + (...) in the reference
  float t = 0.f;
  for (int j = 0; j < nb; ++j) t += c.b[j] * b[j];
  loss[0] = s + t;
}

// g_a has na_total entries (the producing node's full output, e.g. 2 + latent_dim): zeros beyond the na weighted ones
__global__ void loss_combine_bwd_kernel(const float* __restrict__ g, int na, int na_total, int nb, Coefs c,
                                        float* __restrict__ g_a, float* __restrict__ g_b) {
  const float gv = g[0];
  for (int i = threadIdx.x; i < na_total; i += blockDim.x) g_a[i] = i < na ? gv * c.a[i] : 0.f;
  if (g_b)
    for (int j = threadIdx.x; j < nb; j += blockDim.x) g_b[j] = gv * c.b[j];
}

// one block walks whole (image, channel) planes: plane p = b*C + c; per-channel partial sums -> partial[block][32]
__global__ void __launch_bounds__(256)
act_bwd_chansum_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ g, int planes, int C,
                       int hw, int act, float slope, float* __restrict__ partial) {
  __shared__ float red[8][4];
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  const int hw4 = hw >> 2;
  for (int p = blockIdx.x; p < planes; p += gridDim.x) {
    const int c = p % C;
    const float4* dy4 = reinterpret_cast<const float4*>(dy + (long long)p * hw);
    const float4* y4 = reinterpret_cast<const float4*>(y + (long long)p * hw);
    float4* g4 = reinterpret_cast<float4*>(g + (long long)p * hw);
    float s = 0.f;
    for (int i = threadIdx.x; i < hw4; i += blockDim.x) {
      const float4 d = dy4[i], yv = y4[i];
      float4 r;
      if (act == DV_ACT_SIGMOID) {                             // aten sigmoid_backward: grad * (1 - y) * y
        r.x = d.x * ((1.f - yv.x) * yv.x); r.y = d.y * ((1.f - yv.y) * yv.y);
        r.z = d.z * ((1.f - yv.z) * yv.z); r.w = d.w * ((1.f - yv.w) * yv.w);
      } else if (act == DV_ACT_RELU) {
        r.x = yv.x > 0.f ? d.x : 0.f; r.y = yv.y > 0.f ? d.y : 0.f; r.z = yv.z > 0.f ? d.z : 0.f; r.w = yv.w > 0.f ? d.w : 0.f;
      } else if (act == DV_ACT_LEAKY) {
        r.x = yv.x > 0.f ? d.x : d.x * slope; r.y = yv.y > 0.f ? d.y : d.y * slope;
        r.z = yv.z > 0.f ? d.z : d.z * slope; r.w = yv.w > 0.f ? d.w : d.w * slope;
      } else r = d;
      g4[i] = r;
      s += (r.x + r.y) + (r.z + r.w);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) if (k == c) cs[k] += s;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float v = warp_sum(cs[k]);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = 0.f;
    if (threadIdx.x < 4)
      for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    partial[blockIdx.x * 32 + threadIdx.x] = t;                // channel_sum_final_kernel layout: [block][32]
  }
}

__global__ void __launch_bounds__(1024)
chansum_final32_kernel(const float* __restrict__ partial, float* __restrict__ out, int nblocks, int C) {
  __shared__ float red[32][33];
  const int c = threadIdx.x & 31, w = threadIdx.x >> 5;
  float s = 0.f;
  if (c < C)
    for (int b = w; b < nblocks; b += 32) s += partial[b * 32 + c];
  red[w][c] = s;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) t += red[k][c];
    out[c] = t;
  }
}

}  // namespace dv

using namespace dv;

extern "C" {

int dv_u8_to_f32(const unsigned char* src, float* dst, long long n, void* stream) {
  if (!src || !dst) return DV_ERR_BAD_ARG;
  if (n <= 0) return DV_ERR_BAD_SHAPE;
  if (((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) return DV_ERR_BAD_ARG;
  long long blocks = ((n >> 4) + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  u8_to_f32_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(src, dst, n);
  return check_launch();
}

int dv_loss_combine_fwd(const float* a, const float* coef_a, int na, const float* b, const float* coef_b, int nb, float* loss,
                        void* stream) {
  if (!a || !coef_a || !loss || na < 1 || na > 8 || nb < 0 || nb > 8 || (nb > 0 && (!b || !coef_b))) return DV_ERR_BAD_ARG;
  Coefs c = {};
  for (int i = 0; i < na; ++i) c.a[i] = coef_a[i];             // HOST arrays: the coefficients travel by value
  for (int j = 0; j < nb; ++j) c.b[j] = coef_b[j];
  loss_combine_fwd_kernel<<<1, 32, 0, as_stream(stream)>>>(a, na, b, nb, c, loss);
  return check_launch();
}

int dv_loss_combine_bwd(const float* g, const float* coef_a, int na, int na_total, const float* coef_b, int nb, float* g_a,
                        float* g_b, void* stream) {
  if (!g || !coef_a || !g_a || na < 1 || na > 8 || na_total < na || nb < 0 || nb > 8 || (nb > 0 && !coef_b)) return DV_ERR_BAD_ARG;
  Coefs c = {};
  for (int i = 0; i < na; ++i) c.a[i] = coef_a[i];
  for (int j = 0; j < nb; ++j) c.b[j] = coef_b[j];
  loss_combine_bwd_kernel<<<1, 128, 0, as_stream(stream)>>>(g, na, na_total, nb, c, g_a, g_b);
  return check_launch();
}

int dv_act_bwd_chansum(const float* dy, const float* y, float* g, int B, int C, int hw, int act, float slope, float* chansum,
                       void* workspace, void* stream) {
  if (!dy || !y || !g || !chansum || !workspace) return DV_ERR_BAD_ARG;
  if (B < 1 || C < 1 || C > 4 || hw < 4 || (hw & 3)) return DV_ERR_BAD_SHAPE;
  const int planes = B * C;
  const int grid = planes < 296 ? planes : 296;                // <= dv_channel_sum_workspace_bytes() / 128 partial rows
  float* partial = reinterpret_cast<float*>(workspace);
  cudaStream_t st = as_stream(stream);
  act_bwd_chansum_kernel<<<grid, 256, 0, st>>>(dy, y, g, planes, C, hw, act, slope, partial);
  int rc = check_launch();
  if (rc != DV_OK) return rc;
  chansum_final32_kernel<<<1, 1024, 0, st>>>(partial, chansum, grid, C);
  return check_launch();
}

}  // extern "C"
