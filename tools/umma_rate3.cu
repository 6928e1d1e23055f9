// Micro-benchmark 3: why do tcgen05.mma batches run slower inside the conv kernels than in umma_rate2?
// One issuing warp per CTA replays the row-scatter batch of conv3x3_c32_s1_tc_kernel (7 MMAs per input row
// into an 8-slot accumulator ring) under different conditions: uniform vs mixed N, commits per batch,
// bystander warps polling mbarriers (with and without back-off).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate3 tools/umma_rate3.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <algorithm>
#include <vector>
#include "../driving-dirty_b200/csrc/umma.cuh"

constexpr int PS = 2176, WN = 1536;

// MODE bits: 32 / 64 / 128 = tcgen05.fence::after_thread_sync / fence.proxy.async / wait on a completed mbarrier before each batch;
//  1 = mixed N (64 + 32 + 5 x 96) instead of 7 x 96; 2 = two commits per batch; 4 = 12 bystander warps poll a
// barrier (plain try_wait loop); 8 = bystanders poll with __nanosleep(200); 16 = bystanders poll with try_wait suspend hint
template <int MODE>
__global__ void __launch_bounds__(416) rate_kernel(long long* out, int rows) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar, never, ring[32], done0;
  __shared__ uint32_t tbase;
  __shared__ volatile int stop;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    umma::mbar_init(&bar, 1); umma::mbar_init(&never, 1); umma::mbar_init(&done0, 1); umma::mbar_arrive(&done0);
    for (int i = 0; i < 32; ++i) umma::mbar_init(&ring[i], 1);
    umma::fence_mbar_init();
    stop = 0;
  }
  if (threadIdx.x < 32) umma::tmem_alloc(&tbase, 256);
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) {
    constexpr uint32_t idesc32 = umma::make_idesc_bf16(128, 32, false, false);
    constexpr uint32_t NSTEP = (32u >> 3) << 17;
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    const uint32_t sa = __shfl_sync(0xffffffffu, umma::smem_u32(smem), 0);
    const uint32_t a_lo0 = umma::desc_lo(sa + 20 * 1024, PS), b_lo0 = umma::desc_lo(sa, WN);
    constexpr uint32_t hi = umma::desc_hi(128);
    const long long t0 = clock64();
    for (int s = 0; s < rows; ++s) {
      if (MODE & 128) umma::mbar_wait(&done0, 0);
      if (MODE & 64) umma::fence_proxy_async_smem();
      if (MODE & 32) umma::tc_fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t slab_lo = a_lo0 + (s % 8) * (4 * PS >> 4);
        const uint32_t sl = (s % 6);                 // rows s..s+2 of the ring without wrap
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a = slab_lo + ((kw * 16 + 2 * ks * PS) >> 4), b = b_lo0 + (((kw * 4 + 2 * ks) * WN) >> 4);
            if ((MODE & 1) && kw == 0 && ks == 0) {
              umma::mma_bf16_lohi(tb + sl * 32, a, hi, b, hi, idesc32 + NSTEP, 1u);
              umma::mma_bf16_lohi(tb + sl * 32 + 64, a, hi, b + 64, hi, idesc32, 0u);
            } else {
              umma::mma_bf16_lohi(tb + sl * 32, a, hi, b, hi, idesc32 + 2 * NSTEP, 1u);
              if (!(MODE & 1) && kw == 0 && ks == 0) umma::mma_bf16_lohi(tb + sl * 32, a, hi, b, hi, idesc32 + 2 * NSTEP, 1u);
            }
          }
        }
        if (MODE & 2) { umma::mma_commit(&ring[s % 16]); umma::mma_commit(&ring[16 + s % 8]); }
      }
      __syncwarp();
    }
    if (umma::elect_one()) umma::mma_commit(&bar);
    __syncwarp();
    const long long t1 = clock64();
    umma::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
    stop = 1;
  } else if (MODE & (4 | 8 | 16)) {
    while (!stop) {
      if (MODE & 16) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(umma::smem_u32(&never)), "r"(0u), "r"(2000u) : "memory");
      } else {
        umma::mbar_try_wait(&never, 0);
        if (MODE & 8) __nanosleep(200);
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tbase, 256);
}

template <int MODE>
void run(const char* name, long long* d) {
  const int rows = 2048, grid = 148;
  cudaFuncSetAttribute(rate_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024 + 1024);
  rate_kernel<MODE><<<grid, 416, 96 * 1024 + 1024>>>(d, rows);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-72s CUDA error %s\n", name, cudaGetErrorString(e)); return; }
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  std::vector<double> issue, total;
  for (int i = 0; i < grid; ++i) { issue.push_back((double)h[2 * i] / rows); total.push_back((double)h[2 * i + 1] / rows); }
  std::sort(issue.begin(), issue.end()); std::sort(total.begin(), total.end());
  printf("%-72s issue %6.1f  complete median %6.1f max %6.1f cycles per 7-MMA batch (expected 368-392)\n", name, issue[grid / 2],
         total[grid / 2], total[grid - 1]);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  run<0>("7 x N=96, no commits", d);
  run<1>("N=64 + N=32 + 5 x N=96, no commits", d);
  run<2>("7 x N=96, 2 commits per batch", d);
  run<3>("mixed N, 2 commits per batch", d);
  run<3 | 32>("mixed N, commits, tcgen05.fence::after_thread_sync per batch", d);
  run<3 | 64>("mixed N, commits, fence.proxy.async per batch", d);
  run<3 | 128>("mixed N, commits, wait on a completed mbarrier per batch", d);
  run<3 | 32 | 64 | 128>("mixed N, commits, all three", d);
  run<3 | 4>("mixed N, commits, 12 warps polling try_wait", d);
  run<3 | 8>("mixed N, commits, 12 warps polling try_wait + nanosleep(200)", d);
  run<3 | 16>("mixed N, commits, 12 warps polling try_wait with suspend hint", d);
  return 0;
}
