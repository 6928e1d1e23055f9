// Probe for the next round's producers (self-checking; prints PASS / FAIL per variant):
//
//  1. K-major A operand in the 64-byte-swizzle layout, written by ONE TMA box of whole pixels ([136 px][32 ch] = 64 B
//     rows, CU_TENSOR_MAP_SWIZZLE_64B), read by tcgen05.mma with the start address shifted by s pixels (s * 64 B) --
//     the "tap = shifted start address" trick of the conv kernels, but with full-sector TMA rows instead of four
//     16-byte planes.  Open question: does the swizzle follow the absolute shared-memory address (then any shift
//     works with base_offset 0), or does the descriptor's base_offset field have to carry the phase?  All eight
//     base_offset values are tried for s = 0..3.
//  2. TMA element strides: a box that takes every second pixel ([8 ch] x 129 of 257 pixels), as the stride-2 kernels'
//     even / odd pixel planes need, including a start coordinate of -1 (zero fill).
//
//   make -C tools tma_layout_probe && gpurun -- ./tools/tma_layout_probe
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../driving-dirty_b200/csrc/tma_host.h"
#include "../driving-dirty_b200/csrc/umma.cuh"

constexpr int NPX = 136;          // pixels per box (rows of 64 B)
constexpr int A_BYTES = NPX * 64;

__device__ __forceinline__ uint32_t desc_hi_sw64(uint32_t sbo_bytes, uint32_t base_offset) {
  // high word of the matrix descriptor: SBO (bits 32-45), base_offset (49-51), version 1 (46), layout_type 4 = SWIZZLE_64B (61-63)
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | ((base_offset & 7u) << 17) | (4u << 29);
}

// D[128 px][32 n] = sum_c X[p0 + shift + m][c] * Wt[n][c]
__global__ void __launch_bounds__(128) probe_sw64_kernel(const __grid_constant__ CUtensorMap map_x, const __nv_bfloat16* __restrict__ wt,
                                                         float* __restrict__ d_out, int p0, int shift, uint32_t base_offset,
                                                         uint32_t lbo_units) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_a = smem;                       // 1024-aligned, [136 px][64 B] swizzled by the TMA unit
  uint8_t* s_b = smem + 9216;                // weights, SWIZZLE_NONE K-major: [cg][n][8 k]
  __shared__ uint64_t full, done;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 32 * 32; i += 128) {
    const int k = i & 31, n = i >> 5;
    *reinterpret_cast<__nv_bfloat16*>(s_b + (k >> 3) * 512 + n * 16 + (k & 7) * 2) = wt[n * 32 + k];
  }
  if (tid == 0) { umma::mbar_init(&full, 1); umma::mbar_init(&done, 1); umma::fence_mbar_init(); }
  if (warp == 0) umma::tmem_alloc(&tbase, 32);
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  if (tid == 0) {
    umma::mbar_expect_tx(&full, A_BYTES);
    umma::tma_load_2d(umma::smem_u32(s_a), &map_x, 0, p0, &full);
  }
  umma::mbar_wait(&full, 0);
  if (warp == 0) {
    if (umma::elect_one()) {
      constexpr uint32_t idesc = umma::make_idesc_bf16(128, 32, false, false);
      const uint32_t a_addr = umma::smem_u32(s_a) + shift * 64;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint32_t a_lo = (((a_addr + ks * 32) >> 4) & 0x3FFF) | ((lbo_units & 0x3FFF) << 16);
        const uint32_t b_lo = umma::desc_lo(umma::smem_u32(s_b) + ks * 2 * 512, 512);
        umma::mma_bf16_lohi(tbase, a_lo, desc_hi_sw64(512, base_offset), b_lo, umma::desc_hi(128), idesc, ks ? 1u : 0u);
      }
      umma::mma_commit(&done);
    }
    __syncwarp();
  }
  umma::mbar_wait(&done, 0);
  umma::tc_fence_after_sync();
  uint32_t r[32];
  umma::tmem_ld_32x32(tbase + ((uint32_t)(warp * 32) << 16), r);
  umma::tmem_ld_wait();
  for (int n = 0; n < 32; ++n) d_out[(warp * 32 + lane) * 32 + n] = __uint_as_float(r[n]);
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 32);
}

// strided box: copies what the TMA unit wrote (129 x 16 B) back to global memory
__global__ void __launch_bounds__(128) probe_stride2_kernel(const __grid_constant__ CUtensorMap map, __nv_bfloat16* __restrict__ out,
                                                            int cg, int w0, int h, int b, int bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full;
  if (threadIdx.x == 0) { umma::mbar_init(&full, 1); umma::fence_mbar_init(); }
  for (int i = threadIdx.x; i < 4096 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x7fc07fc0u;   // NaN pattern
  umma::fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    umma::mbar_expect_tx(&full, bytes);
    umma::tma_load_4d(umma::smem_u32(smem), &map, cg * 8, w0, h, b, &full);
  }
  umma::mbar_wait(&full, 0);
  for (int i = threadIdx.x; i < 4096 / 2; i += 128) out[i] = reinterpret_cast<__nv_bfloat16*>(smem)[i];
}

static int encode_stride2(CUtensorMap* map, const void* base, uint64_t B, uint64_t H, uint64_t W, uint32_t box_w) {
  dd::EncodeTiledFn enc = dd::tma_encoder();
  if (!enc) return -1;
  cudaFree(nullptr);
  const cuuint64_t dims[4] = {32, W, H, B};
  const cuuint64_t strides[3] = {64, W * 64, H * W * 64};
  const cuuint32_t box[4] = {8, box_w, 1, 1};
  const cuuint32_t estr[4] = {1, 2, 1, 1};
  return (int)enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

static int encode_sw64(CUtensorMap* map, const void* base, uint64_t npix) {
  dd::EncodeTiledFn enc = dd::tma_encoder();
  if (!enc) return -1;
  cudaFree(nullptr);
  const cuuint64_t dims[2] = {32, npix};
  const cuuint64_t strides[1] = {64};
  const cuuint32_t box[2] = {32, NPX};
  const cuuint32_t estr[2] = {1, 1};
  return (int)enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
  // ---------------------------------------------------------------- data: small integers, exact in bf16
  const int P = 1024;
  std::vector<__nv_bfloat16> hx(P * 32), hw(32 * 32);
  std::vector<float> fx(P * 32), fw(32 * 32);
  for (int p = 0; p < P; ++p)
    for (int c = 0; c < 32; ++c) { fx[p * 32 + c] = (float)(((p * 7 + c * 3) % 13) - 6); hx[p * 32 + c] = __float2bfloat16(fx[p * 32 + c]); }
  for (int n = 0; n < 32; ++n)
    for (int k = 0; k < 32; ++k) { fw[n * 32 + k] = (float)(((n * 5 + k * 11) % 7) - 3); hw[n * 32 + k] = __float2bfloat16(fw[n * 32 + k]); }
  __nv_bfloat16 *dx, *dw, *dplane;
  float* dd_out;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dw, hw.size() * 2); cudaMalloc(&dd_out, 128 * 32 * 4); cudaMalloc(&dplane, 4096);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);

  // ---------------------------------------------------------------- 1. swizzle-64B K-major A with pixel shifts
  CUtensorMap mx;
  int rc = encode_sw64(&mx, dx, P);
  printf("1. K-major SWIZZLE_64B A from one TMA box, start address shifted by s pixels (encode rc %d)\n", rc);
  if (rc == 0) {
    cudaFuncSetAttribute(probe_sw64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    const int p0 = 40;
    for (uint32_t lbo = 0; lbo < 2; ++lbo)
      for (int shift = 0; shift < 4; ++shift) {
        printf("   LBO %u  shift %d px: base_offset ->", lbo, shift);
        for (uint32_t bo = 0; bo < 8; ++bo) {
          probe_sw64_kernel<<<1, 128, 16384>>>(mx, dw, dd_out, p0, shift, bo, lbo);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf(" [%u: CUDA error %s]", bo, cudaGetErrorString(e)); return 1; }
          std::vector<float> hd(128 * 32);
          cudaMemcpy(hd.data(), dd_out, hd.size() * 4, cudaMemcpyDeviceToHost);
          double worst = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 32; ++n) {
              double ref = 0;
              for (int c = 0; c < 32; ++c) ref += (double)fx[(p0 + shift + m) * 32 + c] * fw[n * 32 + c];
              worst = fmax(worst, fabs(ref - hd[m * 32 + n]));
            }
          printf(" %u:%s", bo, worst == 0 ? "PASS" : "fail");
        }
        printf("\n");
      }
  }

  // ---------------------------------------------------------------- 2. element stride 2 along the pixel axis
  const int B = 1, H = 4, W = 256;          // the first 1024 pixels of hx seen as [1][4][256][32]
  CUtensorMap ms;
  rc = encode_stride2(&ms, dx, B, H, W, 255);        // box extents are capped at 256 SOURCE pixels: 255 -> 128 loaded
  printf("2. TMA box {8 ch, 255 px, element stride 2} -> expect 128 px x 16 B (encode rc %d)\n", rc);
  if (rc == 0) {
    cudaFuncSetAttribute(probe_stride2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    for (int w0 = -1; w0 <= 1; ++w0) {
      const int cg = 2, h = 1;
      probe_stride2_kernel<<<1, 128, 8192>>>(ms, dplane, cg, w0, h, 0, 128 * 16);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("   w0 %d: CUDA error %s (transaction bytes probably differ from 129 x 16)\n", w0, cudaGetErrorString(e)); return 1; }
      std::vector<__nv_bfloat16> hp(2048);
      cudaMemcpy(hp.data(), dplane, 4096, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int i = 0; i < 128; ++i)
        for (int c = 0; c < 8; ++c) {
          const int w = w0 + 2 * i;
          const float ref = (w >= 0 && w < W) ? fx[((h * W) + w) * 32 + cg * 8 + c] : 0.f;
          if (__bfloat162float(hp[i * 8 + c]) != ref) ++bad;
        }
      printf("   start w0 = %2d: %s (%d of %d elements differ)\n", w0, bad == 0 ? "PASS" : "FAIL", bad, 128 * 8);
    }
  }
  return 0;
}
