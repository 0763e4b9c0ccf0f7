// Marginal-entropy estimator of the MIG / AAM disentanglement metrics.
//
// Reference: disvae/evaluate.py:233-297 (_estimate_latent_entropies), called once for H(z_j) over the whole dataset
// and once per factor value for H(z_j | v_k) (:299-317).  For S samples z[d][s] of latent dimension d and the N
// posteriors q(z_d | x_n) = N(mean[n][d], exp(logvar[n][d])):
//
//   log q(z[d][s]) = -log N + logsumexp_n ( -0.5 (log 2pi + logvar[n][d]) - 0.5 (z[d][s] - mean[n][d])^2 exp(-logvar[n][d]) )
//   H[d]           = -(1/S) sum_s log q(z[d][s])
//
// -- the same pairwise Gaussian log-density pattern as the beta-TCVAE kernel (dv_btcvae.cu) with N != S and no sum
// over dimensions.  The reference materialises [N, D, 10] tensors 1000 times per call ("slow", README.md:51); here
// nothing but the O(N*D) inputs and O(D*S) partial states touches memory.
//
// Mapping: grid = (sample tiles of 512, D, n-splits).  A block stages chunks of 256 posteriors of ITS dimension in
// shared memory as {c*log2e, 0.5*exp(-lv)*log2e, mean} (computed on the fly), every thread owns two samples and runs
// a branch-free online logsumexp in the log2 domain (one MUFU.EX2 per (n, s) pair; LDS.128 broadcast per n).  The
// n-splits' (max, sum) states are merged in a fixed order by entropy_finalize_kernel, which also forms the mean over
// samples with a fixed tree -> deterministic.
#include "dv_common.cuh"

namespace dv {
namespace ent {

constexpr float kLog2Pi = 1.8378770664093453f;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kThreads = 256;
constexpr int kSPT = 2;                       // samples per thread
constexpr int kTile = kThreads * kSPT;        // samples per block
constexpr int kChunk = 256;                   // posteriors staged per iteration

__device__ __forceinline__ float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kThreads)
entropy_partial_kernel(const float* __restrict__ zs, const float* __restrict__ mean, const float* __restrict__ logvar,
                       int ld, int row_stride, int N, int D, int S, int nsplit, float2* __restrict__ part) {
  __shared__ float4 sp[kChunk];
  const int d = blockIdx.y, split = blockIdx.z;
  const int s0 = blockIdx.x * kTile + threadIdx.x;
  float zv[kSPT], m[kSPT], a[kSPT];
#pragma unroll
  for (int u = 0; u < kSPT; ++u) {
    const int s = s0 + u * kThreads;
    zv[u] = (s < S) ? zs[(long long)d * S + s] : 0.f;
    m[u] = -INFINITY; a[u] = 0.f;
  }
  const int per = (N + nsplit - 1) / nsplit;
  const int n_begin = split * per, n_end = min(N, n_begin + per);
  for (int n0 = n_begin; n0 < n_end; n0 += kChunk) {
    const int cnt = min(kChunk, n_end - n0);
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      const long long off = (long long)(n0 + threadIdx.x) * row_stride + (long long)d * ld;
      const float mu = mean[off], lv = logvar[off];
      sp[threadIdx.x] = make_float4(-0.5f * (kLog2Pi + lv) * kLog2e, 0.5f * expf(-lv) * kLog2e, mu, 0.f);
    }
    __syncthreads();
    // two-level accumulation: an online (max, sum) over THIS chunk, merged into the running state afterwards -- a
    // single running fp32 sum over 10^5..10^6 terms loses ~sqrt(n)*eps (1e-4 in log q at dSprites size)
    float mc[kSPT], ac[kSPT];
#pragma unroll
    for (int u = 0; u < kSPT; ++u) { mc[u] = -INFINITY; ac[u] = 0.f; }
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const float4 p = sp[k];
#pragma unroll
      for (int u = 0; u < kSPT; ++u) {
        const float t = zv[u] - p.z;
        const float v = fmaf(-p.y, t * t, p.x);
        const float dd = v - mc[u];
        const float e = ex2a(-fabsf(dd));                      // first element: dd = +inf -> e = 0 -> a = 1, m = v
        const bool up = dd > 0.f;
        ac[u] = up ? fmaf(ac[u], e, 1.f) : ac[u] + e;
        mc[u] = up ? v : mc[u];
      }
    }
#pragma unroll
    for (int u = 0; u < kSPT; ++u) {
      if (a[u] == 0.f) { m[u] = mc[u]; a[u] = ac[u]; }
      else {
        const float nm = fmaxf(m[u], mc[u]);
        a[u] = a[u] * exp2f(m[u] - nm) + ac[u] * exp2f(mc[u] - nm);
        m[u] = nm;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kSPT; ++u) {
    const int s = s0 + u * kThreads;
    if (s < S) part[((long long)split * D + d) * S + s] = make_float2(m[u], a[u]);
  }
}

// H[d] = -(1/S) sum_s ( -log N + ln2 * (m + log2 a) ) with the n-splits merged in order; one block per dimension.
__global__ void __launch_bounds__(256)
entropy_finalize_kernel(const float2* __restrict__ part, int N, int D, int S, int nsplit, float* __restrict__ H,
                        float* __restrict__ logq_out) {
  __shared__ float red[8];
  const int d = blockIdx.x;
  const float logN = logf((float)N);
  float acc = 0.f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    float2 st = part[(long long)d * S + s];
    float m = st.x, a = st.y;
    for (int sp = 1; sp < nsplit; ++sp) {
      st = part[((long long)sp * D + d) * S + s];
      if (st.y == 0.f) continue;
      if (a == 0.f) { m = st.x; a = st.y; continue; }
      const float nm = fmaxf(m, st.x);
      a = a * exp2f(m - nm) + st.y * exp2f(st.x - nm);
      m = nm;
    }
    const float lq = -logN + (m + log2f(a)) * kLn2;            // evaluate.py:284
    if (logq_out) logq_out[(long long)d * S + s] = lq;
    acc += -lq;                                                // evaluate.py:287
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    H[d] = t / (float)S;                                       // evaluate.py:291
  }
}

static int pick_nsplit(int N, int D, int S) {
  const int tiles = (S + kTile - 1) / kTile;
  int ns = (4 * kNumSMs + tiles * D - 1) / (tiles * D);        // ~4 blocks per SM in flight
  const int max_by_n = (N + 4 * kChunk - 1) / (4 * kChunk);    // at least 4 chunks per split
  if (ns > max_by_n) ns = max_by_n;
  if (ns < 1) ns = 1;
  if (ns > 64) ns = 64;
  return ns;
}

}  // namespace ent
}  // namespace dv

using namespace dv;

extern "C" {

size_t dv_latent_entropy_workspace_bytes(int N, int D, int S) {
  return (size_t)ent::pick_nsplit(N, D, S) * D * S * sizeof(float2);
}

int dv_latent_entropy(const float* zs, const float* mean, const float* logvar, int ld, int row_stride, int N, int D, int S,
                      float* H, float* logq_out, void* workspace, void* stream) {
  if (!zs || !mean || !logvar || !H || !workspace) return DV_ERR_BAD_ARG;
  if (N < 1 || D < 1 || S < 1 || D > 65535) return DV_ERR_BAD_SHAPE;
  cudaStream_t st = as_stream(stream);
  const int ns = ent::pick_nsplit(N, D, S);
  float2* part = reinterpret_cast<float2*>(workspace);
  dim3 grid((S + ent::kTile - 1) / ent::kTile, D, ns);
  ent::entropy_partial_kernel<<<grid, ent::kThreads, 0, st>>>(zs, mean, logvar, ld, row_stride, N, D, S, ns, part);
  int rc = check_launch();
  if (rc != DV_OK) return rc;
  ent::entropy_finalize_kernel<<<D, 256, 0, st>>>(part, N, D, S, ns, H, logq_out);
  return check_launch();
}

}  // extern "C"
