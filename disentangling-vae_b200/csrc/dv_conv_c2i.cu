// "up" convolution (ConvTranspose2d k4/s2/p1 forward) for the few-channel output layer, col2im form.
// Replaces convT3 + sigmoid of the Burgess decoder (disvae/models/decoders.py:58-59, 82): lo[B,H,W,32] (NHWC) ->
// hi[B,CH,2H,2W] (NCHW), CH in {1,3}.
//
// With CH this small the GEMM that matters is  D[position][tap*CH + c] = lo[position][0..31] . w[0..31][c][tap]
// (N = 16*CH outputs per input position, no reduction over neighbours): ONE operand tile per 128 positions, no
// shifted copies.  The transposed convolution's scatter  hi(2i-1+kh, 2j-1+kw) += D[(i,j)][kh,kw]  is done as a
// gather in the epilogue, from shared memory: every output pixel sums at most four D entries, adds the bias and
// applies the activation.  A tile is RT = 128/W whole image rows; a CTA walks the tiles of an image top to bottom
// and carries the last D row of the previous tile, so each tile emits 2*RT finished output rows and nothing is
// computed twice.  (The halo kernel in dv_conv_tc.cu pays 9 shifted operand tiles per 128 positions instead.)
//
// 3xTF32: a_hi x [b_hi | b_lo] (N = 2*16*CH) + a_lo x b_hi (N = 16*CH) per K=8 slice, operand A split on the fly
// into TMEM by the split warps, weights (<= 12 KB) resident in shared memory.
//   warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 epilogue, warps 8-15 two split groups.
#include "dv_common.cuh"
#include "dv_ptx.cuh"

namespace dv {
namespace c2i {

using namespace ptx;

constexpr int kThreads = 512;
constexpr int kStages = 6;                        // raw lo tiles (16 KB each)
constexpr int kATile = 128 * 128;
constexpr int kAStages = 4;                       // TMEM A stages (hi 32 | lo 32 columns)
constexpr int kACol0 = 256;
constexpr int kAccCols = 128;                     // two accumulator stages at columns 0 / 128
constexpr uint32_t kHiMask = 0xFFFFE000u;

template <int CH> struct Cfg {
  static constexpr int kNT = 16 * CH;                              // D columns that matter (tap*CH + c)
  static constexpr int kBRows = (2 * kNT + 63) / 64 * 64;          // packed weight rows (hi rows, lo rows, zero padding)
  static constexpr int kBBytes = kBRows * 128;
  static constexpr int kNTP = kNT + 1;                             // smem row pitch of the D staging (odd: conflict free)
  static constexpr int kStageRows = 160;                           // carry row (<= 32 positions) + 128 positions
  static constexpr int kDBytes = kStageRows * kNTP * 4;
  static constexpr int kSmem = kBBytes + kStages * kATile + ((kDBytes + 127) / 128) * 128 + 1024 + 512;
};

struct Barriers {
  uint64_t full[kStages], consumed[kStages];
  uint64_t a_ready[kAStages], a_empty[kAStages];
  uint64_t b_full;
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
  float bias[4];
};
static_assert(sizeof(Barriers) <= 512, "barriers");

struct Geom {
  int B, H, W, RT, tiles_per_img;
};

template <int CH>
__global__ void __launch_bounds__(kThreads, 1)
conv_up_c2i_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                   const float* __restrict__ bias, float* __restrict__ hi_out, Geom g, int act) {
  using C = Cfg<CH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Bs = smem;
  uint8_t* Raw = smem + C::kBBytes;
  float* Ds = reinterpret_cast<float*>(Raw + kStages * kATile);
  Barriers* bars = reinterpret_cast<Barriers*>(reinterpret_cast<uint8_t*>(Ds) + ((C::kDBytes + 127) / 128) * 128);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // provably warp-uniform role index
  // this CTA's images: blockIdx.x, blockIdx.x + gridDim.x, ...; tiles of an image in order
  const int n_img = (g.B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_seq = n_img * g.tiles_per_img;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->consumed[s], 128); }
    for (int s = 0; s < kAStages; ++s) { mbar_init(&bars->a_ready[s], 128); mbar_init(&bars->a_empty[s], 1); }
    mbar_init(&bars->b_full, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&bars->acc_full[a], 1); mbar_init(&bars->acc_empty[a], 128); }
    fence_mbar_init();
  }
  if (threadIdx.x < 4) bars->bias[threadIdx.x] = (bias && (int)threadIdx.x < CH) ? bias[threadIdx.x] : 0.f;
  if (warp == 2) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (bars->tmem_base != 0u) __trap();
  constexpr uint32_t tmem_base = 0u;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmap_a); prefetch_tmap(&tmap_b);
    mbar_arrive_expect_tx(&bars->b_full, C::kBBytes);
    for (int h = 0; h < C::kBRows / 64; ++h) tma_load_2d(Bs + h * 8192, &tmap_b, &bars->b_full, 0, h * 64);
    for (int s = 0; s < n_seq; ++s) {
      const int stage = s % kStages;
      const int img = blockIdx.x + (s / g.tiles_per_img) * gridDim.x;
      const int i0 = (s % g.tiles_per_img) * g.RT;
      mbar_wait(&bars->consumed[stage], ((s / kStages) & 1u) ^ 1u);
      mbar_arrive_expect_tx(&bars->full[stage], kATile);
      tma_load_2d(Raw + stage * kATile, &tmap_a, &bars->full[stage], 0, (img * g.H + i0) * g.W);
    }
  } else if (warp == 1 && elect_one()) {      // ONE elected lane runs the whole issue loop (barrier waits included):
                                              // ptxas then keeps every MMA operand in uniform registers (back-to-back UTCHMMA)
    constexpr uint32_t idesc2 = umma_idesc_tf32(128, 2 * C::kNT), idesc1 = umma_idesc_tf32(128, C::kNT);
    mbar_wait(&bars->b_full, 0);
    const uint64_t b_d = umma_desc_sw128_kmajor(smem_u32(Bs));         // rows [0,NT) hi, [NT,2NT) lo
    for (int s = 0; s < n_seq; ++s) {
      const int as = s % kAStages, acc = s & 1;
      mbar_wait(&bars->acc_empty[acc], ((s >> 1) & 1u) ^ 1u);
      mbar_wait(&bars->a_ready[as], (s / kAStages) & 1u);
      tc_fence_after_sync();
      const uint32_t a_hi = tmem_base + kACol0 + as * 64, a_lo = a_hi + 32;
      const uint32_t d = tmem_base + acc * kAccCols;
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        umma_tf32_ts_1t(d, a_hi + 8 * k4, b_d + 2 * k4, idesc2, k4 != 0);      // cols [0,NT) hi*hi, [NT,2NT) hi*lo
        umma_tf32_ts_1t(d, a_lo + 8 * k4, b_d + 2 * k4, idesc1, 1);            // cols [0,NT) += lo*hi
      }
      umma_commit_1t(&bars->a_empty[as]);
      umma_commit_1t(&bars->acc_full[acc]);
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- epilogue: D -> smem, then gather the finished output rows ----
    const int q = warp & 3;
    const int r = q * 32 + lane;                                  // position inside the tile
    const int W = g.W, H = g.H, RT = g.RT;
    const int W2 = 2 * W, H2 = 2 * H;
    for (int s = 0; s < n_seq; ++s) {
      const int acc = s & 1;
      const int img = blockIdx.x + (s / g.tiles_per_img) * gridDim.x;
      const int t_in_img = s % g.tiles_per_img;
      const int i0 = t_in_img * RT;
      mbar_wait(&bars->acc_full[acc], (s >> 1) & 1u);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * kAccCols;
      float v[C::kNT];
      if (CH == 1) {
        uint32_t c0[32];
        tmem_ld_32x32b_x32(taddr, c0);
        tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < 16; ++n) v[n] = __uint_as_float(c0[n]) + __uint_as_float(c0[16 + n]);
      } else {
        uint32_t c0[32], c1[32], c2[32];
        tmem_ld_32x32b_x32(taddr, c0);
        tmem_ld_32x32b_x32(taddr + 32, c1);
        tmem_ld_32x32b_x32(taddr + 64, c2);
        tmem_ld_wait();
        // columns [0,48) + [48,96)
#pragma unroll
        for (int n = 0; n < 48; ++n) {
          const float a = __uint_as_float(n < 32 ? c0[n] : c1[n - 32]);
          const int m = n + 48;
          const float b = __uint_as_float(m < 64 ? c1[m - 32] : c2[m - 64]);
          v[n] = a + b;
        }
      }
      tc_fence_before_sync();
      mbar_arrive(&bars->acc_empty[acc]);                          // accumulator stage is free again
      // stage the tile's D rows behind the carry row (positions [0,W) = image row i0-1)
      if (t_in_img == 0 && r < W) {
#pragma unroll
        for (int n = 0; n < C::kNT; ++n) Ds[r * C::kNTP + n] = 0.f;     // row -1 does not exist
      }
      {
        float* dst = Ds + (W + r) * C::kNTP;
#pragma unroll
        for (int n = 0; n < C::kNT; ++n) dst[n] = v[n];
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // output rows y = 2*i0 - 1 + yy, yy in [0, 2RT) (+ the image's last row from its last tile)
      const bool last = (t_in_img == g.tiles_per_img - 1);
      const int n_rows = 2 * RT + (last ? 1 : 0);
      const int total = n_rows * W2;
      for (int idx = r; idx < total; idx += 128) {
        const int yy = idx / W2, x = idx - yy * W2;
        const int y = 2 * i0 - 1 + yy;
        if (y < 0) continue;
        const int ia = (y + 1) >> 1, kha = y + 1 - 2 * ia;          // rows ia (kh = kha) and ia-1 (kh = kha+2)
        const int ja = (x + 1) >> 1, kwa = x + 1 - 2 * ja;
        const int la = ia - i0 + 1;                                  // staged row index (0 = carry)
        const bool va = ia < H && la <= RT, ca = ja < W, cb = ja >= 1;
        float o[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) o[c] = bars->bias[c];
        if (va) {
          const float* row = Ds + (la * W) * C::kNTP;
          if (ca) {
#pragma unroll
            for (int c = 0; c < CH; ++c) o[c] += row[ja * C::kNTP + (kha * 4 + kwa) * CH + c];
          }
          if (cb) {
#pragma unroll
            for (int c = 0; c < CH; ++c) o[c] += row[(ja - 1) * C::kNTP + (kha * 4 + kwa + 2) * CH + c];
          }
        }
        {
          const float* row = Ds + ((la - 1) * W) * C::kNTP;           // image row ia-1 (zeros above the image)
          if (ca) {
#pragma unroll
            for (int c = 0; c < CH; ++c) o[c] += row[ja * C::kNTP + ((kha + 2) * 4 + kwa) * CH + c];
          }
          if (cb) {
#pragma unroll
            for (int c = 0; c < CH; ++c) o[c] += row[(ja - 1) * C::kNTP + ((kha + 2) * 4 + kwa + 2) * CH + c];
          }
        }
#pragma unroll
        for (int c = 0; c < CH; ++c)
          hi_out[(((long long)img * CH + c) * H2 + y) * W2 + x] = apply_act(o[c], act, 0.f);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (!last && r >= 128 - W) {                                   // last row of this tile becomes the carry
        float* dst = Ds + (r - (128 - W)) * C::kNTP;
#pragma unroll
        for (int n = 0; n < C::kNT; ++n) dst[n] = v[n];
      }
      // (the next tile's first bar.sync orders the carry write before its reads; an image's last tile writes no
      //  carry: the next image zeroes it, possibly from other threads)
    }
  } else if (warp >= 8) {
    // ---- split warps: raw lo tile -> hi/lo planes in TMEM (two groups on alternate tiles) ----
    const int q = warp & 3, grp = (warp - 8) >> 2;
    const int row = q * 32 + lane;
    int prev_stage = -1, prev_as = 0;                         // tile whose TMEM stores are still in flight
    for (int s = grp; s < n_seq; s += 2) {
      const int stage = s % kStages, as = s % kAStages;
      mbar_wait(&bars->full[stage], (s / kStages) & 1u);
      const uint8_t* raw = Raw + stage * kATile;
      uint32_t h[32], l[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 vv4 = lds128(raw + row * 128 + ((c ^ (row & 7)) << 4));
        const uint32_t vv[4] = {vv4.x, vv4.y, vv4.z, vv4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t hb = vv[e] & kHiMask;
          h[c * 4 + e] = hb;
          l[c * 4 + e] = __float_as_uint(__uint_as_float(vv[e]) - __uint_as_float(hb));
        }
      }
      // software pipeline: the stores of the previous tile overlapped this tile's loads and split; publish it
      // (and release its raw stage -- only after tcgen05.wait::st, when its reads were certainly consumed) now
      if (prev_stage >= 0) {
        tmem_st_wait();
        mbar_arrive(&bars->consumed[prev_stage]);
        tc_fence_before_sync();
        mbar_arrive(&bars->a_ready[prev_as]);
      }
      mbar_wait(&bars->a_empty[as], ((s / kAStages) & 1u) ^ 1u);
      tc_fence_after_sync();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + kACol0 + as * 64;
      tmem_st_32x32b_x32(taddr, h);
      tmem_st_32x32b_x32(taddr + 32, l);
      prev_stage = stage; prev_as = as;
    }
    if (prev_stage >= 0) {
      tmem_st_wait();
      mbar_arrive(&bars->consumed[prev_stage]);
      tc_fence_before_sync();
      mbar_arrive(&bars->a_ready[prev_as]);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

// w[cl][c][kh][kw] (ConvTranspose2d layout [in=32][out=CH][4][4]) -> rows n = tap*CH + c: hi plane rows [0,NT),
// lo plane rows [NT,2NT), zero rows up to a multiple of 64; 32 floats (cl) per row.
__global__ void conv_pack_c2i_kernel(const float* __restrict__ w, float* __restrict__ wp, int CH, int rows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * 32) return;
  const int row = idx >> 5, cl = idx & 31;
  const int NT = 16 * CH;
  float out = 0.f;
  if (row < 2 * NT) {
    const int n = row < NT ? row : row - NT;
    const int tap = n / CH, c = n - tap * CH;
    const float v = w[(cl * CH + c) * 16 + tap];
    const float hi = __uint_as_float(__float_as_uint(v) & kHiMask);
    out = row < NT ? hi : v - hi;
  }
  wp[idx] = out;
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
static bool make_rows_tmap(CUtensorMap* m, const float* base, long long rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {32, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t packed_floats(int CH) { return CH == 1 ? (size_t)Cfg<1>::kBRows * 32 : (size_t)Cfg<3>::kBRows * 32; }

int pack(const float* w, float* wp, int CH, cudaStream_t st) {
  const int rows = CH == 1 ? Cfg<1>::kBRows : Cfg<3>::kBRows;
  conv_pack_c2i_kernel<<<(rows * 32 + 255) / 256, 256, 0, st>>>(w, wp, CH, rows);
  return check_launch();
}

bool shape_ok(int B, int H, int W, int CH) {
  return (CH == 1 || CH == 3) && (W == 16 || W == 32) && H % (128 / W) == 0 && (long long)B * H * W < (1ll << 31);
}

template <int CH>
static int launch(const float* lo, const float* wp, const float* bias, float* hi, int B, int H, int W, int act, cudaStream_t st) {
  using C = Cfg<CH>;
  CUtensorMap ta, tb;
  if (!make_rows_tmap(&ta, lo, (long long)B * H * W, 128)) return DV_ERR_CUDA;
  if (!make_rows_tmap(&tb, wp, C::kBRows, 64)) return DV_ERR_CUDA;
  static bool attr = false;
  if (!attr) {
    if (cudaFuncSetAttribute(conv_up_c2i_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem) != cudaSuccess) {
      g_last_cuda_error = (int)cudaGetLastError();
      return DV_ERR_CUDA;
    }
    attr = true;
  }
  Geom g{B, H, W, 128 / W, H / (128 / W)};
  const int grid = B < kNumSMs ? B : kNumSMs;
  conv_up_c2i_kernel<CH><<<grid, kThreads, C::kSmem, st>>>(ta, tb, bias, hi, g, act);
  return check_launch();
}

int conv_up(const float* lo, const float* wp, const float* bias, float* hi, int B, int H, int W, int CH, int act, cudaStream_t st) {
  if (CH == 1) return launch<1>(lo, wp, bias, hi, B, H, W, act, st);
  return launch<3>(lo, wp, bias, hi, B, H, W, act, st);
}

}  // namespace c2i
}  // namespace dv
