// HBM bandwidth by direction (settles what roof a read-only kernel such as a weight gradient can be held to):
// the copy figure in MEASURED_PEAKS.json counts read + write bytes of one torch copy; this probe times, over a
// buffer much larger than the 126 MB L2, (a) a read-only reduction, (b) a write-only fill, (c) a copy, each as a
// grid-stride kernel of 16-byte accesses with 8 independent accesses in flight per thread, and (d) a read-only
// pass through TMA (cp.async.bulk, 16 KB per request, 8 requests in flight per CTA).
//   make -C tools hbm_rw_probe && gpurun -- ./tools/hbm_rw_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../driving-dirty_b200/csrc/umma.cuh"

constexpr int UN = 8;

__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ src, size_t n16, uint32_t* __restrict__ sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * UN) {
    uint4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const size_t j = i + u * stride;
      v[u] = make_uint4(0, 0, 0, 0);
      if (j < n16) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + j));
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;       // never true for the fill pattern; keeps the loads alive
}

__global__ void __launch_bounds__(256) write_kernel(uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = v;
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * UN) {
    uint4 v[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const size_t j = i + u * stride;
      if (j < n16) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src + j));
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const size_t j = i + u * stride;
      if (j < n16) dst[j] = v[u];
    }
  }
}

// read-only through the TMA unit: one thread per CTA keeps STAGES bulk copies of CHUNK bytes in flight
constexpr int CHUNK = 16384, STAGES = 8;
__global__ void __launch_bounds__(32) tma_read_kernel(const uint8_t* __restrict__ src, size_t nchunks) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t full[STAGES];
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) umma::mbar_init(&full[s], 1);
    umma::fence_mbar_init();
    size_t n = 0;
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++n) {
      const int s = (int)(n % STAGES);
      if (n >= STAGES) umma::mbar_wait(&full[s], ((n / STAGES) - 1) & 1);
      umma::mbar_expect_tx(&full[s], CHUNK);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(umma::smem_u32(smem + s * CHUNK)), "l"(src + c * CHUNK), "r"(CHUNK), "r"(umma::smem_u32(&full[s])) : "memory");
    }
    for (size_t k = (n > STAGES ? n - STAGES : 0); k < n; ++k) umma::mbar_wait(&full[k % STAGES], (k / STAGES) & 1);
  }
}

template <class F>
static float time_ms(F f, int iters) {
  f();
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) f();
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms / iters;
}

int main() {
  const size_t bytes = (size_t)4 << 30;        // 4 GiB per buffer, 32x the L2
  uint4 *a, *b;
  uint32_t* sink;
  if (cudaMalloc(&a, bytes) != cudaSuccess || cudaMalloc(&b, bytes) != cudaSuccess) { printf("cudaMalloc failed\n"); return 1; }
  cudaMalloc(&sink, 4);
  cudaMemset(a, 1, bytes); cudaMemset(b, 2, bytes);
  const size_t n16 = bytes / 16;
  cudaFuncSetAttribute(tma_read_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CHUNK * STAGES);
  for (int per_sm : {4, 8, 16}) {
    const int grid = 148 * per_sm;
    const float tr = time_ms([&] { read_kernel<<<grid, 256>>>(a, n16, sink); }, 5);
    const float tw = time_ms([&] { write_kernel<<<grid, 256>>>(b, n16); }, 5);
    const float tc = time_ms([&] { copy_kernel<<<grid, 256>>>(a, b, n16); }, 5);
    printf("grid %4d x 256: read-only %7.1f GB/s   write-only %7.1f GB/s   copy %7.1f GB/s (read + write bytes)\n", grid,
           bytes / tr / 1e6, bytes / tw / 1e6, 2.0 * bytes / tc / 1e6);
  }
  for (int per_sm : {1, 2}) {
    const int grid = 148 * per_sm;
    const float tt = time_ms([&] { tma_read_kernel<<<grid, 32, CHUNK * STAGES>>>((const uint8_t*)a, bytes / CHUNK); }, 5);
    printf("TMA bulk read, grid %3d, %d x %d KB in flight per CTA: %7.1f GB/s\n", grid, STAGES, CHUNK / 1024, bytes / tt / 1e6);
  }
  const float tm = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 5);
  printf("cudaMemcpy D2D: %7.1f GB/s (read + write bytes)\n", 2.0 * bytes / tm / 1e6);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
