// Micro-benchmark 2: tcgen05.mma (kind::f16, M=128, K=16, SS) throughput with descriptors that are pure
// compile-time offsets from uniform bases (no per-instruction R2UR traffic): 16 MMAs straight-line per loop
// trip.  Reports cycles per MMA for N and for the number of independent accumulators (dependent chains).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate2 tools/umma_rate2.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <algorithm>
#include <vector>
#include "../driving-dirty_b200/csrc/umma.cuh"

template <int N, int NACC, int ASTEP, int TS, int MN = 0>
__global__ void __launch_bounds__(128) rate_kernel(long long* out, int trips) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
  if (threadIdx.x < 32) umma::tmem_alloc(&tbase, 512);
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp == 0) {
    constexpr uint32_t idesc = umma::make_idesc_bf16(128, N, MN != 0, MN != 0);
    const uint32_t tb = __shfl_sync(0xffffffffu, tbase, 0);
    const uint32_t sa = __shfl_sync(0xffffffffu, umma::smem_u32(smem), 0);
    // MN-major (both operands, as in the weight-gradient kernels): LBO = 128 between the 8-row K groups, SBO = plane stride
    const uint32_t a_lo = umma::desc_lo(sa, MN ? 128 : 2176), b_lo = umma::desc_lo(sa + 40 * 1024, MN ? 128 : 512);
    constexpr uint32_t hi = umma::desc_hi(MN ? 2176 : 128);
    const long long t0 = clock64();
    if (umma::elect_one()) {
      for (int t = 0; t < trips; ++t) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (TS) {
            // A from TMEM: columns 256.. hold the operand (contents irrelevant for timing)
            asm volatile(
                "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
                "setp.ne.b32 p, %5, 0;\n\t"
                "mov.b64 db, {%2, %3};\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
                ::"r"(tb + (j % NACC) * N), "r"(tb + 256 + (j % 8) * 8), "r"(b_lo + (j % 9) * 64), "r"(hi), "r"(idesc), "r"(1u)
                : "memory");
          } else {
            umma::mma_bf16_lohi(tb + (j % NACC) * N, a_lo + (((j % 8) * ASTEP) >> 4), hi, b_lo + (j % 9) * 64, hi, idesc, 1u);
          }
        }
      }
      umma::mma_commit(&bar);
    }
    __syncwarp();
    const long long t1 = clock64();
    umma::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tbase, 512);
}

template <int N, int NACC, int ASTEP, int TS, int MN = 0>
void run(const char* name, long long* d) {
  const int trips = 256, grid = 148, nmma = trips * 16;
  cudaFuncSetAttribute(rate_kernel<N, NACC, ASTEP, TS, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024 + 1024);
  rate_kernel<N, NACC, ASTEP, TS, MN><<<grid, 128, 64 * 1024 + 1024>>>(d, trips);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-56s CUDA error %s\n", name, cudaGetErrorString(e)); return; }
  std::vector<long long> h(grid * 2);
  cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  std::vector<double> issue, total;
  for (int i = 0; i < grid; ++i) { issue.push_back((double)h[2 * i] / nmma); total.push_back((double)h[2 * i + 1] / nmma); }
  std::sort(issue.begin(), issue.end()); std::sort(total.begin(), total.end());
  printf("%-56s issue %6.1f  complete median %6.1f max %6.1f cyc/MMA (tensor floor %5.1f, smem A+B %5.1f)\n", name, issue[grid / 2],
         total[grid / 2], total[grid - 1], 128.0 * N / 256.0, (TS ? N * 32.0 : (128 + N) * 32.0) / 128.0);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 2 * sizeof(long long));
  run<32, 1, 16, 0>("SS N=32  1 accumulator  (A shifted 16 B per MMA)", d);
  run<32, 2, 16, 0>("SS N=32  2 accumulators", d);
  run<32, 4, 16, 0>("SS N=32  4 accumulators", d);
  run<32, 8, 16, 0>("SS N=32  8 accumulators", d);
  run<32, 4, 0, 0>("SS N=32  4 accumulators, same A", d);
  run<64, 1, 16, 0>("SS N=64  1 accumulator", d);
  run<64, 4, 16, 0>("SS N=64  4 accumulators", d);
  run<128, 1, 16, 0>("SS N=128 1 accumulator", d);
  run<128, 2, 16, 0>("SS N=128 2 accumulators", d);
  run<256, 1, 16, 0>("SS N=256 1 accumulator", d);
  run<48, 1, 16, 0>("SS N=48  1 accumulator", d);
  run<80, 1, 16, 0>("SS N=80  1 accumulator", d);
  run<96, 1, 16, 0>("SS N=96  1 accumulator", d);
  run<96, 2, 16, 0>("SS N=96  2 accumulators", d);
  run<112, 1, 16, 0>("SS N=112 1 accumulator", d);
  run<160, 1, 16, 0>("SS N=160 1 accumulator", d);
  run<192, 1, 16, 0>("SS N=192 1 accumulator", d);
  run<224, 1, 16, 0>("SS N=224 1 accumulator", d);
  run<32, 1, 256, 0, 1>("SS MN-major A and B, N=32", d);
  run<64, 1, 256, 0, 1>("SS MN-major A and B, N=64 (c2 weight gradient)", d);
  run<32, 1, 16, 1>("TS (A in TMEM) N=32 1 accumulator", d);
  run<32, 4, 16, 1>("TS (A in TMEM) N=32 4 accumulators", d);
  run<64, 4, 16, 1>("TS (A in TMEM) N=64 4 accumulators", d);
  return 0;
}
