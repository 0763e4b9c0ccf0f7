// Microbenchmark: issue rate / throughput of tcgen05.mma kind::tf32 (M = 128, K = 8) as a function of N, operand source
// (A in tensor memory or shared memory) and accumulator reuse.  One CTA per SM, one issuing thread, NMMA back-to-back
// MMAs, one commit; clocks measured around issue + completion.  Build on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -I disentangling-vae_b200/csrc -o /tmp/mma_rate scripts/micro/mma_rate.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "dv_ptx.cuh"
using namespace dv::ptx;

__device__ __forceinline__ uint64_t desc_mn(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(512 >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}

// mode 0: TS (A in TMEM), B K-major smem;  mode 1: SS, both K-major;  mode 2: TS, B MN-major
template <int N, int MODE, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int nmma) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && elect_one()) {
    constexpr uint32_t idesc = umma_idesc_tf32(128, N) | (MODE == 2 ? (1u << 16) : 0u);
    const uint32_t a_s = smem_u32(smem), b_s = smem_u32(smem + 32 * 1024);
    const long long t0 = clock64();
    for (int i = 0; i < nmma; ++i) {
      const uint32_t d = tmem + (i % NACC) * N;
      const int k4 = i & 3;
      if (MODE == 1) umma_tf32_ss_1t(d, umma_desc_sw128_kmajor(a_s) + 2 * k4, umma_desc_sw128_kmajor(b_s) + 2 * k4, idesc, i >= NACC);
      else if (MODE == 0) umma_tf32_ts_1t(d, tmem + 448 + 8 * k4, umma_desc_sw128_kmajor(b_s) + 2 * k4, idesc, i >= NACC);
      else umma_tf32_ts_1t(d, tmem + 448 + 8 * k4, desc_mn(b_s + k4 * 1024, 16384), idesc, i >= NACC);
    }
    const long long t1 = clock64();
    umma_commit_1t(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { tc_fence_after_sync(); tmem_dealloc(tmem, 512); }
}

template <int N, int MODE, int NACC>
static void run(const char* what, long long* d_out, int grid) {
  const int nmma = 4096, smem = 65 * 1024 + 1024;
  cudaFuncSetAttribute(rate_kernel<N, MODE, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[2] = {0, 0};
  for (int rep = 0; rep < 2; ++rep) {
    rate_kernel<N, MODE, NACC><<<grid, 128, smem>>>(d_out, nmma);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", what, cudaGetErrorString(e)); return; }
  }
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-34s N=%3d acc=%d grid=%3d : issue %6.1f clk/MMA, complete %6.1f clk/MMA  (%.2f clk per N column, %.0f MAC/clk/SM)\n", what, N, NACC,
         grid, (double)h[0] / nmma, (double)h[1] / nmma, (double)h[1] / nmma / N, 128.0 * N * 8 / ((double)h[1] / nmma));
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 16);
  for (int grid : {1, 148}) {
    run<32, 0, 1>("TS  B k-major, same accumulator", d_out, grid);
    run<64, 0, 1>("TS  B k-major, same accumulator", d_out, grid);
    run<128, 0, 1>("TS  B k-major, same accumulator", d_out, grid);
    run<256, 0, 1>("TS  B k-major, same accumulator", d_out, grid);
    run<32, 0, 4>("TS  B k-major, 4 accumulators", d_out, grid);
    run<64, 0, 4>("TS  B k-major, 4 accumulators", d_out, grid);
    run<32, 2, 1>("TS  B mn-major, same accumulator", d_out, grid);
    run<64, 2, 1>("TS  B mn-major, same accumulator", d_out, grid);
    run<32, 1, 1>("SS  k-major, same accumulator", d_out, grid);
    run<64, 1, 1>("SS  k-major, same accumulator", d_out, grid);
    run<256, 1, 1>("SS  k-major, same accumulator", d_out, grid);
  }
  return 0;
}
