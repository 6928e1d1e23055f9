// Micro-benchmark: issue rate / latency of tcgen05.mma (kind::f16, M=128) for the operand layouts and
// shapes conv_tc.cu uses.  One CTA per SM; lane 0 of warp 0 issues NMMA instructions back to back
// and waits for the commit; cycles per MMA are reported for the slowest and the median CTA.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include "../driving-dirty_b200/csrc/umma.cuh"

struct Cfg { int N; int nacc; int layout; int a_bytes_step; int mn_major; const char* name; };

__global__ void __launch_bounds__(128) rate_kernel(long long* out, int N, int nacc, int layout, int nmma, int a_step, int mn, int uniform_mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
  if (threadIdx.x < 32) umma::tmem_alloc(&tbase, 512);
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const int warp_idx = uniform_mode == 3 ? __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0) : (int)(threadIdx.x >> 5);
  if (warp_idx == 0) {
    // whole warp runs the (uniform) loop; only the elected lane issues -> operands stay in uniform registers
    const bool leader = (uniform_mode ? umma::elect_one() : (threadIdx.x == 0));
    if (!uniform_mode && threadIdx.x != 0) goto done;
    const uint32_t idesc = umma::make_idesc_bf16(128, N, mn, mn);
    const uint32_t tb = uniform_mode >= 2 ? __shfl_sync(0xffffffffu, tbase, 0) : tbase;
    const uint32_t sa0 = umma::smem_u32(smem);
    const uint32_t sa = uniform_mode >= 2 ? __shfl_sync(0xffffffffu, sa0, 0) : sa0, sb = sa + 32 * 1024;
    uint32_t lo_a, hi_a, lo_b, hi_b;
    if (layout == 0) {          // SWIZZLE_NONE K-major, conv geometry (or MN-major wgrad geometry)
      lo_a = umma::desc_lo(sa, mn ? 128 : 2176); hi_a = umma::desc_hi(mn ? 2176 : 128);
      lo_b = umma::desc_lo(sb, mn ? 128 : 512);  hi_b = umma::desc_hi(mn ? 2048 : 128);
    } else {                    // SWIZZLE_128B K-major: 8 rows x 128 B atoms, SBO = 1024
      lo_a = umma::desc_lo(sa, 16); hi_a = umma::desc_hi(1024) | (2u << 29);
      lo_b = umma::desc_lo(sb, 16); hi_b = umma::desc_hi(1024) | (2u << 29);
    }
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < nmma; ++i) {
      const uint32_t d = tb + (i & (nacc - 1)) * N, a = lo_a + (((i & 7) * a_step) >> 4);
      if (leader) umma::mma_bf16_lohi(d, a, hi_a, lo_b, hi_b, idesc, 1u);
    }
    long long t1 = clock64();
    if (leader) umma::mma_commit(&bar);
    umma::mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (leader) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
done:
  umma::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tbase, 512);
}

int main() {
  const int nmma = 2048, grid = 148;
  long long* d; cudaMalloc(&d, grid * 2 * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  Cfg cfgs[] = {
    {32, 1, 0, 0, 0, "noswz K-major N=32  1 acc  same A"},
    {32, 1, 0, 16, 0, "noswz K-major N=32  1 acc  A shifted 16B steps"},
    {32, 4, 0, 16, 0, "noswz K-major N=32  4 acc"},
    {64, 1, 0, 16, 0, "noswz K-major N=64  1 acc"},
    {96, 1, 0, 16, 0, "noswz K-major N=96  1 acc"},
    {128, 1, 0, 16, 0, "noswz K-major N=128 1 acc"},
    {256, 1, 0, 16, 0, "noswz K-major N=256 1 acc"},
    {32, 1, 1, 0, 0, "swz128 K-major N=32  1 acc"},
    {32, 4, 1, 0, 0, "swz128 K-major N=32  4 acc"},
    {64, 1, 1, 0, 0, "swz128 K-major N=64  1 acc"},
    {128, 1, 1, 0, 0, "swz128 K-major N=128 1 acc"},
    {256, 1, 1, 0, 0, "swz128 K-major N=256 1 acc"},
    {64, 1, 0, 256, 1, "noswz MN-major N=64 1 acc (wgrad)"},
    {64, 4, 0, 256, 1, "noswz MN-major N=64 4 acc (wgrad)"},
    {32, 1, 0, 256, 1, "noswz MN-major N=32 1 acc"},
  };
  for (int um = 0; um < 4; ++um)
  for (auto& c : cfgs) {
    if (um < 3 && &c != &cfgs[0]) continue;
    printf(um == 3 ? "[shfl warp idx + broadcast]  " : um == 2 ? "[uniform + shfl-broadcast]   " : um ? "[warp-uniform, elected lane] " : "[single divergent thread]    ");
    rate_kernel<<<grid, 128, 64 * 1024>>>(d, c.N, c.nacc, c.layout, nmma, c.a_bytes_step, c.mn_major, um);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    std::vector<long long> h(grid * 2);
    cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    std::vector<double> issue, total;
    for (int i = 0; i < grid; ++i) { issue.push_back((double)h[2 * i] / nmma); total.push_back((double)h[2 * i + 1] / nmma); }
    std::sort(issue.begin(), issue.end()); std::sort(total.begin(), total.end());
    const double ideal = 128.0 * c.N / 256.0 / 1.0;   // M*N*K / (8192 MAC/clk... ) -> M*N/256 cycles for K=16
    printf("%-48s issue %6.1f  complete median %6.1f max %6.1f cyc/MMA (ideal %5.1f)\n", c.name, issue[grid / 2], total[grid / 2], total[grid - 1], ideal);
  }
  return 0;
}
