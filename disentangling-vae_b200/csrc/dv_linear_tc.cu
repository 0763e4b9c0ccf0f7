// tcgen05 GEMM for the fully connected layers (encoder/decoder MLPs and the FactorVAE discriminator):
//     C[M, Nout] = epilogue( A[M, R] . Bw[Nout, R]^T )          (both operands K-major = R contiguous)
// used for the forward pass (A = activations, Bw = weight [N, K], bias + ReLU/LeakyReLU epilogue) and for the
// input gradient (A = upstream gradient [M, N], Bw = weight transposed [K, N], activation-gradient mask epilogue).
// Replaces nn.Linear + activation (disvae/models/encoders.py:81-86, decoders.py:71-73, discriminator.py:63-68)
// and the dgrad half of their autograd backward.
//
// Error-compensated 3xTF32 like the convolutions: the weight is split once by a pack kernel into tf32-exact hi
// and residual lo planes (plus the transposed copy for dgrad); the activation tile is split on the fly by the
// split warps into TMEM.  Per 32-wide K block and K=8 slice: one MMA  a_hi x [b_hi | b_lo]  (N = 2*BN) and one
// a_lo x b_hi (N = BN); the two accumulator halves are added in the epilogue.
//   CTA tile 128 x 64, grid = ceil(M/128) x ceil(Nout/64); 6-stage TMA ring of {A raw 16 KB, B hi 8 KB, B lo 8 KB};
//   warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 epilogue, warps 8-15 two split groups.
//   Row / column / K tails are handled by TMA out-of-bounds zero fill and guarded stores.
#include "dv_common.cuh"
#include "dv_ptx.cuh"
#include <cstdlib>
#include <cstring>

namespace dv {
namespace ltc {

using namespace ptx;

constexpr int kBM = 128, kBN = 64, kBK = 32;
constexpr int kThreads = 512;
constexpr int kStages = 6;
constexpr int kATile = kBM * 128;                 // 16 KB raw activation tile
constexpr int kBTile = kBN * 128;                 // 8 KB per weight plane tile
constexpr int kStageBytes = kATile + 2 * kBTile;  // 32 KB
constexpr int kAStages = 4;                       // TMEM A stages (hi 32 | lo 32 columns each)
constexpr int kACol0 = 256;
constexpr uint32_t kHiMask = 0xFFFFE000u;

struct Barriers {
  uint64_t full[kStages], a_consumed[kStages], b_consumed[kStages];
  uint64_t a_ready[kAStages], a_empty[kAStages];
  uint64_t acc_full;
  uint32_t tmem_base;
};
constexpr int kSmemBytes = kStages * kStageBytes + 1024 + 512;
static_assert(sizeof(Barriers) <= 512, "barriers");

struct Epilogue {
  const float* bias;       // [Nout] or null
  const float* mask_src;   // [M, Nout] post-activation output of the previous layer, or null
  int act;                 // DV_ACT_* (forward activation, or which activation's gradient masks)
  float slope;
};

__global__ void __launch_bounds__(kThreads, 1)
linear_nt_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_bhi,
                    const __grid_constant__ CUtensorMap tmap_blo, float* __restrict__ C, int M, int Nout, int R,
                    Epilogue ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Barriers* bars = reinterpret_cast<Barriers*>(smem + kStages * kStageBytes);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform role index
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
  const int nkb = (R + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->a_consumed[s], 128); mbar_init(&bars->b_consumed[s], 1); }
    for (int s = 0; s < kAStages; ++s) { mbar_init(&bars->a_ready[s], 128); mbar_init(&bars->a_empty[s], 1); }
    mbar_init(&bars->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (bars->tmem_base != 0u) __trap();                       // 1 CTA/SM owns the whole tensor memory
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0 && elect_one()) {
    // ---- TMA producer: {A raw, B hi, B lo} of one K block per stage ----
    prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_bhi); prefetch_tmap(&tmap_blo);
    for (int kb = 0; kb < nkb; ++kb) {
      const int stage = kb % kStages;
      const uint32_t par = ((kb / kStages) & 1u) ^ 1u;
      mbar_wait(&bars->a_consumed[stage], par);              // split warps are done with the raw A tile
      mbar_wait(&bars->b_consumed[stage], par);              // MMAs that read the B tiles have completed
      uint8_t* st = smem + stage * kStageBytes;
      mbar_arrive_expect_tx(&bars->full[stage], kStageBytes);
      tma_load_2d(st, &tmap_a, &bars->full[stage], kb * kBK, m0);
      tma_load_2d(st + kATile, &tmap_bhi, &bars->full[stage], kb * kBK, n0);
      tma_load_2d(st + kATile + kBTile, &tmap_blo, &bars->full[stage], kb * kBK, n0);
    }
  } else if (warp == 1 && elect_one()) {      // ONE elected lane runs the whole issue loop (barrier waits included):
                                              // ptxas then keeps every MMA operand in uniform registers (back-to-back UTCHMMA)
    // ---- MMA issuer (whole warp, one elected lane issues) ----
    constexpr uint32_t idesc2 = umma_idesc_tf32(128, 2 * kBN), idesc1 = umma_idesc_tf32(128, kBN);
    for (int kb = 0; kb < nkb; ++kb) {
      const int stage = kb % kStages, as = kb % kAStages;
      mbar_wait(&bars->full[stage], (kb / kStages) & 1u);    // B tiles landed (A raw too)
      mbar_wait(&bars->a_ready[as], (kb / kAStages) & 1u);   // A hi/lo in TMEM
      tc_fence_after_sync();
      const uint32_t a_hi = tmem_base + kACol0 + as * 64, a_lo = a_hi + 32;
      const uint64_t b_d = umma_desc_sw128_kmajor(smem_u32(smem + stage * kStageBytes + kATile));   // [b_hi (64 rows) | b_lo (64 rows)]
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        // two accumulation chains (even / odd K slices): the tensor core truncates when it accumulates, shorter
        // chains keep the result fp32-grade.  Chain c: cols [128c, 128c+64) hi*hi + lo*hi, [128c+64, 128c+128) hi*lo.
        const uint32_t d = tmem_base + (k4 & 1) * 128;
        umma_tf32_ts_1t(d, a_hi + 8 * k4, b_d + 2 * k4, idesc2, (kb | (k4 >> 1)) != 0);
        umma_tf32_ts_1t(d, a_lo + 8 * k4, b_d + 2 * k4, idesc1, 1);
      }
      umma_commit_1t(&bars->a_empty[as]);
      umma_commit_1t(&bars->b_consumed[stage]);
    }
    umma_commit_1t(&bars->acc_full);
  } else if (warp >= 4 && warp < 8) {
    // ---- epilogue: thread = output row ----
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
    mbar_wait(&bars->acc_full, 0);
    tc_fence_after_sync();
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      uint32_t r0[32], r1[32];
      float acc[32];
      tmem_ld_32x32b_x32(taddr + half * 32, r0);             // chain 0: hi*hi + lo*hi
      tmem_ld_32x32b_x32(taddr + kBN + half * 32, r1);       //          hi*lo
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = __uint_as_float(r0[c]) + __uint_as_float(r1[c]);
      tmem_ld_32x32b_x32(taddr + 128 + half * 32, r0);       // chain 1
      tmem_ld_32x32b_x32(taddr + 128 + kBN + half * 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] += __uint_as_float(r0[c]) + __uint_as_float(r1[c]);
      if (m < M) {
        const int nb = n0 + half * 32;
        float* crow = C + (long long)m * Nout;
        const float* mrow = ep.mask_src ? ep.mask_src + (long long)m * Nout : nullptr;
        if ((Nout & 3) == 0) {
          // 16-byte path: each thread writes 128 contiguous bytes of its row
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const int n = nb + c4 * 4;
            if (n < Nout) {
              float v[4] = {acc[c4 * 4], acc[c4 * 4 + 1], acc[c4 * 4 + 2], acc[c4 * 4 + 3]};
              if (mrow) {
                const float4 y = *reinterpret_cast<const float4*>(mrow + n);
                const float yy[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (ep.act == DV_ACT_RELU) v[e] = yy[e] > 0.f ? v[e] : 0.f;
                  else if (ep.act == DV_ACT_LEAKY) v[e] = yy[e] > 0.f ? v[e] : v[e] * ep.slope;
                }
              } else {
                if (ep.bias) {
                  const float4 bb = *reinterpret_cast<const float4*>(ep.bias + n);
                  v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = apply_act(v[e], ep.act, ep.slope);
              }
              *reinterpret_cast<float4*>(crow + n) = make_float4(v[0], v[1], v[2], v[3]);
            }
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int n = nb + c;
            if (n < Nout) {
              float v = acc[c];
              if (mrow) {
                const float y = mrow[n];
                if (ep.act == DV_ACT_RELU) v = y > 0.f ? v : 0.f;
                else if (ep.act == DV_ACT_LEAKY) v = y > 0.f ? v : v * ep.slope;
              } else {
                if (ep.bias) v += ep.bias[n];
                v = apply_act(v, ep.act, ep.slope);
              }
              crow[n] = v;
            }
          }
        }
      }
    }
  } else if (warp >= 8) {
    // ---- split warps: raw A tile -> hi/lo planes in TMEM (two groups on alternate K blocks) ----
    const int q = warp & 3, grp = (warp - 8) >> 2;
    const int row = q * 32 + lane;
    int prev_stage = -1, prev_as = 0;                         // K block whose TMEM stores are still in flight
    for (int kb = grp; kb < nkb; kb += 2) {
      const int stage = kb % kStages, as = kb % kAStages;
      mbar_wait(&bars->full[stage], (kb / kStages) & 1u);
      const uint8_t* raw = smem + stage * kStageBytes;
      uint32_t h[32], l[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = lds128(raw + row * 128 + ((c ^ (row & 7)) << 4));
        const uint32_t vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t hb = vv[e] & kHiMask;
          h[c * 4 + e] = hb;
          l[c * 4 + e] = __float_as_uint(__uint_as_float(vv[e]) - __uint_as_float(hb));
        }
      }
      // software pipeline: the stores of the previous K block overlapped this block's loads and split; publish
      // it (and release its raw tile -- only after tcgen05.wait::st, when its reads were certainly consumed) now
      if (prev_stage >= 0) {
        tmem_st_wait();
        mbar_arrive(&bars->a_consumed[prev_stage]);
        tc_fence_before_sync();
        mbar_arrive(&bars->a_ready[prev_as]);
      }
      mbar_wait(&bars->a_empty[as], ((kb / kAStages) & 1u) ^ 1u);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kACol0 + as * 64;
      tmem_st_32x32b_x32(taddr, h);
      tmem_st_32x32b_x32(taddr + 32, l);
      prev_stage = stage; prev_as = as;
    }
    if (prev_stage >= 0) {
      tmem_st_wait();
      mbar_arrive(&bars->a_consumed[prev_stage]);
      tc_fence_before_sync();
      mbar_arrive(&bars->a_ready[prev_as]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

// w[N][K] -> hi/lo planes.  transpose == 0: [N][Kp] (forward; Kp = K rounded up to 4 floats, TMA pitch is 16-byte);
//                            transpose == 1: [K][Np] (the transposed copy the input-gradient GEMM reads).
__global__ void linear_pack_kernel(const float* __restrict__ w, float* __restrict__ p_hi, float* __restrict__ p_lo,
                                   int N, int K, int pitch, int transpose) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    const float v = (n < N && k < K) ? w[(long long)n * K + k] : 0.f;
    if (transpose) {
      tile[r][tx] = v;
    } else if (n < N && k < pitch) {
      const float hi = __uint_as_float(__float_as_uint(v) & kHiMask);
      p_hi[(long long)n * pitch + k] = hi;
      p_lo[(long long)n * pitch + k] = v - hi;
    }
  }
  if (!transpose) return;
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + tx;
    if (k < K && n < pitch) {
      const float v = tile[tx][r];
      const float hi = __uint_as_float(__float_as_uint(v) & kHiMask);
      p_hi[(long long)k * pitch + n] = hi;
      p_lo[(long long)k * pitch + n] = v - hi;
    }
  }
}

// Multi-tensor pack: every weight matrix of a network node (encoder MLP, decoder MLP, the 6-layer discriminator) in
// ONE launch, both layouts at once -- packed[i] = [fwd hi | fwd lo | transposed hi | transposed lo] with pitches
// round4(K) / round4(N).  Replaces one linear_pack_kernel launch per layer per direction (12 per VAE step, 22 per
// FactorVAE step).
constexpr int kPackMax = 8;
struct PackTable {
  const float* w[kPackMax];
  float* dst[kPackMax];
  int N[kPackMax], K[kPackMax], tiles_k[kPackMax], tile0[kPackMax + 1];
  int n;
};
__global__ void linear_pack_multi_kernel(PackTable t) {
  __shared__ float tile[32][33];
  int m = 0;
  while (m + 1 < t.n && (int)blockIdx.x >= t.tile0[m + 1]) ++m;
  const int local = blockIdx.x - t.tile0[m];
  const int N = t.N[m], K = t.K[m];
  const int Kp = (K + 3) & ~3, Np = (N + 3) & ~3;
  const int k0 = (local % t.tiles_k[m]) * 32, n0 = (local / t.tiles_k[m]) * 32;
  const float* __restrict__ w = t.w[m];
  float* f_hi = t.dst[m];
  float* f_lo = f_hi + (size_t)N * Kp;
  float* t_hi = f_lo + (size_t)N * Kp;
  float* t_lo = t_hi + (size_t)K * Np;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    const float v = (n < N && k < K) ? w[(long long)n * K + k] : 0.f;
    tile[r][tx] = v;
    if (n < N && k < Kp) {
      const float hi = __uint_as_float(__float_as_uint(v) & kHiMask);
      f_hi[(long long)n * Kp + k] = hi;
      f_lo[(long long)n * Kp + k] = v - hi;
    }
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + tx;
    if (k < K && n < Np) {
      const float v = tile[tx][r];
      const float hi = __uint_as_float(__float_as_uint(v) & kHiMask);
      t_hi[(long long)k * Np + n] = hi;
      t_lo[(long long)k * Np + n] = v - hi;
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
// row-major [rows][cols] fp32 with a row pitch of `pitch` floats; box = {32 cols, box_rows}
static bool make_2d(CUtensorMap* m, const float* base, long long rows, long long cols, long long pitch, int box_rows,
                    CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)pitch * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline int round4(int v) { return (v + 3) & ~3; }

static int pack(const float* w, float* p_hi, float* p_lo, int N, int K, int transpose, cudaStream_t st) {
  const int pitch = transpose ? round4(N) : round4(K);
  dim3 grid((round4(K) + 31) / 32, (round4(N) + 31) / 32);
  linear_pack_kernel<<<grid, 256, 0, st>>>(w, p_hi, p_lo, N, K, pitch, transpose);
  return check_launch();
}

// C[M,Nout] = epi(A[M,R] . B[Nout,R]^T);  b_hi/b_lo are packed planes with row pitch round4(R)
static int launch_nt(const float* A, long long a_pitch, const float* b_hi, const float* b_lo, float* C, int M, int Nout, int R,
                     Epilogue ep, cudaStream_t st) {
  CUtensorMap ta, tbh, tbl;
  if (!make_2d(&ta, A, M, R, a_pitch, kBM)) return DV_ERR_CUDA;
  if (!make_2d(&tbh, b_hi, Nout, R, round4(R), kBN)) return DV_ERR_CUDA;
  if (!make_2d(&tbl, b_lo, Nout, R, round4(R), kBN)) return DV_ERR_CUDA;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(linear_nt_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) {
      g_last_cuda_error = (int)cudaGetLastError();
      return DV_ERR_CUDA;
    }
    attr = true;
  }
  dim3 grid((M + kBM - 1) / kBM, (Nout + kBN - 1) / kBN);
  linear_nt_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(ta, tbh, tbl, C, M, Nout, R, ep);
  return check_launch();
}

// ------------------------------------------------------------------------------------------
// weight gradient:  dW[n][k] = sum_m G[m][n] * X[m][k]   (reduction over the batch rows)
// Both operands come from TMA as [128 batch rows][32 features] tiles, i.e. with the reduction index along the
// smem rows -> MN-major tcgen05 operands (32-byte-atom 128B swizzle), no transposition anywhere.  One MMA
// (M=128, N=64, K=8 rows):  A = [Xa_hi | Xb_hi | Xa_lo | Xb_lo] (two 32-wide k groups of X, hi/lo planes),
// B = [G_hi | G_lo] (one 32-wide n group): D holds all four hi/lo cross products.  A CTA owns one n group,
// up to 16 k groups (8 pair accumulators x 64 columns = all of TMEM) and one slice of the batch (deterministic
// split-K: partials to the workspace, reduced in a fixed order).  Same scheme as the conv weight gradient.
// ------------------------------------------------------------------------------------------
constexpr int kWgStages = 2;
constexpr int kWgStageBytes = 4 * kATile;              // Xa_hi, Xb_hi, Xa_lo, Xb_lo
constexpr int kWgGBytes = 2 * kATile;                  // G_hi, G_lo
struct WgBarriers {
  uint64_t raw_full[kWgStages], ready[kWgStages], empty[kWgStages];
  uint64_t g_raw_full[2], g_ready[2], g_empty[2];
  uint64_t acc_full;
  uint32_t tmem_base;
  float lscr[128][4];                                    // column sums of G (bias gradient), per split thread
};
constexpr int kWgSmemBytes = kWgStages * kWgStageBytes + 2 * kWgGBytes + 1024 + 3072;
static_assert(sizeof(WgBarriers) <= 3072, "barriers");
static_assert(kWgSmemBytes <= 232448, "smem");

__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}
// residual plane lo = v - trunc_tf32(v).  The hi operand is the RAW fp32 tile itself: kind::tf32 reads only the upper
// 19 bits of each 32-bit element, which is exactly the truncation the mask performs (the fp64-accuracy tests hold this).
__device__ __forceinline__ void split_lo_only(const uint4* raw, uint4* lo4, int t) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int idx = t + 128 * k;
    const uint4 v = raw[idx];
    uint4 l;
    l.x = __float_as_uint(__uint_as_float(v.x) - __uint_as_float(v.x & kHiMask));
    l.y = __float_as_uint(__uint_as_float(v.y) - __uint_as_float(v.y & kHiMask));
    l.z = __float_as_uint(__uint_as_float(v.z) - __uint_as_float(v.z & kHiMask));
    l.w = __float_as_uint(__uint_as_float(v.w) - __uint_as_float(v.w & kHiMask));
    lo4[idx] = l;
  }
}

struct WgGeom {
  int M, N, K;
  int m_tiles, tiles_per_split;
};

__global__ void __launch_bounds__(kThreads, 1)
linear_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_g,
                       float* __restrict__ out_base, long long split_stride, float* __restrict__ dbias_base,
                       long long dbias_stride, WgGeom g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* St = smem;                                        // [stage][Xa_hi|Xb_hi|Xa_lo|Xb_lo]
  uint8_t* Gs = smem + kWgStages * kWgStageBytes;            // [buf][G_hi|G_lo]
  WgBarriers* bars = reinterpret_cast<WgBarriers*>(Gs + 2 * kWgGBytes);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform role index
  const int ng = blockIdx.x, kc = blockIdx.y, sp = blockIdx.z;
  const int kgroups = (g.K + 31) / 32;
  const int my_groups = min(16, kgroups - kc * 16);
  const int npairs = (my_groups + 1) / 2;
  const int t_begin = sp * g.tiles_per_split;
  const int t_end = min(g.m_tiles, t_begin + g.tiles_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&bars->raw_full[s], 1); mbar_init(&bars->ready[s], 128); mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->g_raw_full[s], 1); mbar_init(&bars->g_ready[s], 128); mbar_init(&bars->g_empty[s], 1); }
    mbar_init(&bars->acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (bars->tmem_base != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmap_x); prefetch_tmap(&tmap_g);
    int stage = 0; uint32_t phase = 0; int gb = 0; uint32_t gphase = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      const int m0 = tile * 128;
      mbar_wait(&bars->g_empty[gb], gphase ^ 1);
      mbar_arrive_expect_tx(&bars->g_raw_full[gb], kATile);
      tma_load_2d(Gs + gb * kWgGBytes, &tmap_g, &bars->g_raw_full[gb], ng * 32, m0);
      if (++gb == 2) { gb = 0; gphase ^= 1; }
      for (int pr = 0; pr < npairs; ++pr) {
        mbar_wait(&bars->empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&bars->raw_full[stage], 2 * kATile);
#pragma unroll
        for (int h = 0; h < 2; ++h)                          // a k group past the end is all out-of-bounds: zero filled
          tma_load_2d(St + stage * kWgStageBytes + h * kATile, &tmap_x, &bars->raw_full[stage], (kc * 16 + 2 * pr + h) * 32, m0);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && elect_one()) {      // ONE elected lane runs the whole issue loop (barrier waits included):
                                              // ptxas then keeps every MMA operand in uniform registers (back-to-back UTCHMMA)
    constexpr uint32_t idesc = umma_idesc_tf32(128, 64) | (1u << 15) | (1u << 16);   // both operands MN-major
    int stage = 0; uint32_t phase = 0; int gb = 0; uint32_t gphase = 0;
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait(&bars->g_ready[gb], gphase);
      tc_fence_after_sync();
      const uint32_t g_addr = smem_u32(Gs + gb * kWgGBytes);
      for (int pr = 0; pr < npairs; ++pr) {
        mbar_wait(&bars->ready[stage], phase);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(St + stage * kWgStageBytes);
        const uint32_t d = tmem_base + pr * 64;
#pragma unroll 4
        for (int kk = 0; kk < 16; ++kk)                      // 8 batch rows per MMA
          umma_tf32_ss_1t(d, umma_desc_sw128_mnmajor(a_addr + kk * 1024, kATile),
                       umma_desc_sw128_mnmajor(g_addr + kk * 1024, kATile), idesc, (tile != t_begin) || (kk != 0));
        umma_commit_1t(&bars->empty[stage]);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
      umma_commit_1t(&bars->g_empty[gb]);
      if (++gb == 2) { gb = 0; gphase ^= 1; }
    }
    umma_commit_1t(&bars->acc_full);
  } else if (warp >= 8 && warp < 12) {
    const int t = threadIdx.x - 256;
    int stage = 0; uint32_t phase = 0; int gb = 0; uint32_t gphase = 0;
    float ls[4] = {0.f, 0.f, 0.f, 0.f};                       // this thread's 16-byte chunk of every G row it touches
    const bool want_bias = dbias_base != nullptr && kc == 0;   // (every k chunk sees the same G tiles)
    for (int tile = t_begin; tile < t_end; ++tile) {
      mbar_wait(&bars->g_raw_full[gb], gphase);
      if (want_bias) {
        const uint4* graw = reinterpret_cast<const uint4*>(Gs + gb * kWgGBytes);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint4 u = graw[t + 128 * k];
          ls[0] += __uint_as_float(u.x); ls[1] += __uint_as_float(u.y); ls[2] += __uint_as_float(u.z); ls[3] += __uint_as_float(u.w);
        }
      }
      split_lo_only(reinterpret_cast<const uint4*>(Gs + gb * kWgGBytes), reinterpret_cast<uint4*>(Gs + gb * kWgGBytes + kATile), t);
      fence_proxy_async_smem();
      mbar_arrive(&bars->g_ready[gb]);
      if (++gb == 2) { gb = 0; gphase ^= 1; }
      for (int pr = 0; pr < npairs; ++pr) {
        mbar_wait(&bars->raw_full[stage], phase);
        uint8_t* base = St + stage * kWgStageBytes;
        split_lo_only(reinterpret_cast<const uint4*>(base), reinterpret_cast<uint4*>(base + 2 * kATile), t);
        split_lo_only(reinterpret_cast<const uint4*>(base + kATile), reinterpret_cast<uint4*>(base + 3 * kATile), t);
        fence_proxy_async_smem();
        mbar_arrive(&bars->ready[stage]);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
    // bias gradient = column sums of G: under the 32-byte-atom swizzle thread t always sees logical 16-byte chunk
    // `quad` of a row (quad = ((t&7)>>1 ^ (t>>3)&3) << 1 | t&1); 16 threads share a chunk, summed in a fixed order
#pragma unroll
    for (int e = 0; e < 4; ++e) bars->lscr[t][e] = ls[e];
    asm volatile("bar.sync 2, 128;" ::: "memory");
    if (want_bias && t < 32) {
      const int want = t >> 2, e = t & 3;
      float acc = 0.f;
      for (int u = 0; u < 128; ++u)
        if ((((((u & 7) >> 1) ^ ((u >> 3) & 3)) << 1) | (u & 1)) == want) acc += bars->lscr[u][e];
      const int n = ng * 32 + t;
      if (n < g.N) dbias_base[(long long)sp * dbias_stride + n] = acc;
    }
  }

  // ---- epilogue (once per CTA): TMEM -> fold the four hi/lo quadrants -> dW (or this split's partial) ----
  if (warp >= 4 && warp < 8) {
    const int q = warp & 3;
    const int r = q * 32 + lane;
    mbar_wait(&bars->acc_full, 0);
    tc_fence_after_sync();
    float* red = reinterpret_cast<float*>(St);               // all MMAs have completed: stage buffers are free
    float* out = out_base + (long long)sp * split_stride;
    for (int pr = 0; pr < npairs; ++pr) {
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + pr * 64;
      uint32_t r0[32], r1[32];
      tmem_ld_32x32b_x32(taddr, r0);
      tmem_ld_32x32b_x32(taddr + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int cl = 0; cl < 32; ++cl) red[r * 33 + cl] = __uint_as_float(r0[cl]) + __uint_as_float(r1[cl]);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (r < 64) {
        const int k = (kc * 16 + 2 * pr + (r >> 5)) * 32 + (r & 31);
        if (k < g.K) {
#pragma unroll 4
          for (int cl = 0; cl < 32; ++cl) {
            const int n = ng * 32 + cl;
            if (n < g.N) out[(long long)n * g.K + k] = red[r * 33 + cl] + red[(r + 64) * 33 + cl];
          }
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

static bool enabled() {
  static const int v = [] { const char* e = getenv("DV_LINEAR_IMPL"); return (e && strcmp(e, "ffma") == 0) ? 0 : 1; }();
  return v == 1;
}
// the tensor-core path needs 16-byte pitched activation rows for TMA: K % 4 == 0 (fwd) / N % 4 == 0 (dgrad)
size_t fwd_workspace_bytes(int M, int N, int K) {
  if (!enabled() || K % 4 != 0 || K < 32) return 0;
  return (size_t)2 * N * K * sizeof(float);
}
size_t dgrad_workspace_bytes(int M, int N, int K) {
  if (!enabled() || N % 4 != 0 || N < 32) return 0;
  return (size_t)2 * K * N * sizeof(float);
}

int fwd(const float* x, const float* w, const float* bias, float* y, int M, int N, int K, int act, float slope, float* ws,
        cudaStream_t st) {
  float* hi = ws;
  float* lo = ws + (size_t)N * K;
  int rc = pack(w, hi, lo, N, K, 0, st);
  if (rc != DV_OK) return rc;
  Epilogue ep{bias, nullptr, act, slope};
  return launch_nt(x, K, hi, lo, y, M, N, K, ep, st);
}
int dgrad(const float* g, const float* w, const float* mask_src, float* dx, int M, int N, int K, int act, float slope, float* ws,
          cudaStream_t st) {
  float* hi = ws;
  float* lo = ws + (size_t)K * N;
  int rc = pack(w, hi, lo, N, K, 1, st);
  if (rc != DV_OK) return rc;
  Epilogue ep{nullptr, mask_src, mask_src ? act : DV_ACT_NONE, slope};
  return launch_nt(g, N, hi, lo, dx, M, K, N, ep, st);
}


size_t packed_floats(int N, int K) { return (size_t)2 * N * round4(K) + (size_t)2 * K * round4(N); }

int pack_multi(int n, const float* const* w, float* const* packed, const int* N, const int* K, cudaStream_t st) {
  for (int base = 0; base < n; base += kPackMax) {
    PackTable t;
    t.n = n - base < kPackMax ? n - base : kPackMax;
    int tiles = 0;
    for (int i = 0; i < t.n; ++i) {
      t.w[i] = w[base + i]; t.dst[i] = packed[base + i]; t.N[i] = N[base + i]; t.K[i] = K[base + i];
      t.tiles_k[i] = (round4(K[base + i]) + 31) / 32;
      t.tile0[i] = tiles;
      tiles += t.tiles_k[i] * ((round4(N[base + i]) + 31) / 32);
    }
    t.tile0[t.n] = tiles;
    linear_pack_multi_kernel<<<tiles, 256, 0, st>>>(t);
    int rc = check_launch();
    if (rc != DV_OK) return rc;
  }
  return DV_OK;
}

// the same two GEMMs on planes that dv_linear_pack_multi produced
int fwd_packed(const float* x, const float* packed, const float* bias, float* y, int M, int N, int K, int act, float slope,
               cudaStream_t st) {
  const float* hi = packed;
  const float* lo = packed + (size_t)N * round4(K);
  Epilogue ep{bias, nullptr, act, slope};
  return launch_nt(x, K, hi, lo, y, M, N, K, ep, st);
}
int dgrad_packed(const float* g, const float* packed, const float* mask_src, float* dx, int M, int N, int K, int act, float slope,
                 cudaStream_t st) {
  const float* hi = packed + (size_t)2 * N * round4(K);
  const float* lo = hi + (size_t)K * round4(N);
  Epilogue ep{nullptr, mask_src, mask_src ? act : DV_ACT_NONE, slope};
  return launch_nt(g, N, hi, lo, dx, M, K, N, ep, st);
}

static void wgrad_plan(int M, int N, int K, int* S, int* tiles_per_split) {
  const int m_tiles = (M + 127) / 128;
  const int base = ((N + 31) / 32) * (((K + 31) / 32 + 15) / 16);
  int want = (kNumSMs + base - 1) / base;
  if (want > m_tiles) want = m_tiles;
  if (want < 1) want = 1;
  const int per = (m_tiles + want - 1) / want;
  *tiles_per_split = per;
  *S = (m_tiles + per - 1) / per;                             // no empty split
}
bool wgrad_ok(int M, int N, int K) { return enabled() && N % 4 == 0 && K % 4 == 0 && K >= 32 && M >= 32; }
size_t wgrad_workspace_bytes(int M, int N, int K) {
  int S, per;
  wgrad_plan(M, N, K, &S, &per);
  return S > 1 ? ((size_t)S * N * K + (size_t)S * N) * sizeof(float) : 0;     // dW partials, then dbias partials
}
// returns the number of splits written (1: dw is final) through *nsplit
// dbias != NULL: the column sums of g come out of the same kernel (partials at ws + S*N*K when S > 1)
int wgrad(const float* g, const float* x, float* dw, float* dbias, float* ws, int M, int N, int K, int* nsplit, cudaStream_t st) {
  int S, per;
  wgrad_plan(M, N, K, &S, &per);
  CUtensorMap tx, tg;
  if (!make_2d(&tx, x, M, K, K, 128, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return DV_ERR_CUDA;
  if (!make_2d(&tg, g, M, N, N, 128, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return DV_ERR_CUDA;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(linear_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes) != cudaSuccess) {
      g_last_cuda_error = (int)cudaGetLastError();
      return DV_ERR_CUDA;
    }
    attr = true;
  }
  WgGeom geo{M, N, K, (M + 127) / 128, per};
  dim3 grid((N + 31) / 32, ((K + 31) / 32 + 15) / 16, S);
  float* dbias_base = !dbias ? nullptr : (S > 1 ? ws + (size_t)S * N * K : dbias);
  linear_wgrad_tc_kernel<<<grid, kThreads, kWgSmemBytes, st>>>(tx, tg, S > 1 ? ws : dw, (long long)N * K, dbias_base, (long long)N, geo);
  *nsplit = S;
  return check_launch();
}

}  // namespace ltc
}  // namespace dv
