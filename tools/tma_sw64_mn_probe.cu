// Probe for the weight-gradient kernels (self-checking; prints PASS / fail per variant):
//
//  MN-major A and B operands in the 64-byte-swizzle layout, each x / dy row written by ONE TMA box of whole pixels
//  ([NPX px][32 ch] = 64-byte rows, CU_TENSOR_MAP_SWIZZLE_64B) -- the contraction of a weight gradient runs over
//  pixels (K), channels are the M / N index, so "[pixel][32 ch]" NHWC rows ARE the canonical MN-major SW64 atom
//  (8 K-rows x 64 B).  Questions:
//    1. LBO = stride between 32-channel blocks (the next x row's tile), SBO = 512 B (next 8-pixel group)?
//    2. does a start address shifted by s pixels (s * 64 B = the kw tap) read the right pixels with base_offset 0,
//       or must base_offset carry the phase?  (all eight values tried for s = 0..3)
//    3. K step of 16 pixels = start address + 1024 B.
//    4. the same with a TMA element stride of 2 along the pixel axis (even / odd pixel planes of the stride-2 conv).
//
//   make -C tools tma_sw64_mn_probe && gpurun -- ./tools/tma_sw64_mn_probe
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include "../driving-dirty_b200/csrc/tma_host.h"
#include "../driving-dirty_b200/csrc/umma.cuh"

constexpr int NPX = 136;                 // pixels per x box
constexpr int XROW = NPX * 64;           // 8704 B = 17 x 512: x row tile stride (LBO of A)
constexpr int DPX = 128;                 // pixels per dy box
constexpr int DROW = DPX * 64;           // 8192 B
constexpr int KPX = 32;                  // pixels contracted (two K = 16 steps)

__device__ __forceinline__ uint32_t desc_hi_sw64(uint32_t sbo_bytes, uint32_t base_offset) {
  return ((sbo_bytes >> 4) & 0x3FFF) | (1u << 14) | ((base_offset & 7u) << 17) | (4u << 29);
}

// D[r*32 + ci][q*32 + co] = sum_{p < KPX} X_r[shift + p][ci] * Y_q[p][co]     (r < 4 x rows, q < 2 dy rows)
__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                                                    float* __restrict__ d_out, int shift, uint32_t base_offset, int swap_lbo_sbo,
                                                    int xcoord_mul, int xbox_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_x = smem;                       // 4 x row tiles
  uint8_t* s_y = smem + 4 * XROW + 1024 - ((4 * XROW) & 1023);   // 1024-aligned dy tiles (2 rows)
  __shared__ uint64_t full, done;
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { umma::mbar_init(&full, 1); umma::mbar_init(&done, 1); umma::fence_mbar_init(); }
  if (warp == 0) umma::tmem_alloc(&tbase, 64);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  if (tid == 0) {
    umma::mbar_expect_tx(&full, 4 * xbox_bytes + 2 * DROW);
    for (int r = 0; r < 4; ++r) umma::tma_load_2d(umma::smem_u32(s_x) + r * XROW, &map_x, 0, (r * 200) * xcoord_mul, &full);
    for (int q = 0; q < 2; ++q) umma::tma_load_2d(umma::smem_u32(s_y) + q * DROW, &map_y, 0, q * 300, &full);
  }
  umma::mbar_wait(&full, 0);
  if (warp == 0) {
    if (umma::elect_one()) {
      constexpr uint32_t idesc = umma::make_idesc_bf16(128, 64, true, true);
      const uint32_t lbo_a = swap_lbo_sbo ? 512 : XROW, sbo_a = swap_lbo_sbo ? XROW : 512;
      const uint32_t lbo_b = swap_lbo_sbo ? 512 : DROW, sbo_b = swap_lbo_sbo ? DROW : 512;
#pragma unroll
      for (int ks = 0; ks < KPX / 16; ++ks) {
        const uint32_t a_lo = umma::desc_lo(umma::smem_u32(s_x) + shift * 64 + ks * 1024, lbo_a);
        const uint32_t b_lo = umma::desc_lo(umma::smem_u32(s_y) + ks * 1024, lbo_b);
        umma::mma_bf16_lohi(tbase, a_lo, desc_hi_sw64(sbo_a, base_offset), b_lo, desc_hi_sw64(sbo_b, 0), idesc, ks ? 1u : 0u);
      }
      umma::mma_commit(&done);
    }
    __syncwarp();
  }
  umma::mbar_wait(&done, 0);
  umma::tc_fence_after_sync();
  for (int half = 0; half < 2; ++half) {
    uint32_t r[32];
    umma::tmem_ld_32x32(tbase + ((uint32_t)(warp * 32) << 16) + half * 32, r);
    umma::tmem_ld_wait();
    for (int n = 0; n < 32; ++n) d_out[(warp * 32 + lane) * 64 + half * 32 + n] = __uint_as_float(r[n]);
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tbase, 64);
}

static int encode_sw64(CUtensorMap* map, const void* base, uint64_t npix, uint32_t box_px, uint32_t estride) {
  dd::EncodeTiledFn enc = dd::tma_encoder();
  if (!enc) return -1;
  cudaFree(nullptr);
  const cuuint64_t dims[2] = {32, npix};
  const cuuint64_t strides[1] = {64};
  const cuuint32_t box[2] = {32, box_px};
  const cuuint32_t estr[2] = {1, estride};
  return (int)enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
  const int P = 4096;
  std::vector<__nv_bfloat16> hx(P * 32), hy(P * 32);
  std::vector<float> fx(P * 32), fy(P * 32);
  for (int p = 0; p < P; ++p)
    for (int c = 0; c < 32; ++c) {
      fx[p * 32 + c] = (float)(((p * 7 + c * 3 + (p >> 3)) % 13) - 6); hx[p * 32 + c] = __float2bfloat16(fx[p * 32 + c]);
      fy[p * 32 + c] = (float)(((p * 5 + c * 11 + (p >> 2)) % 7) - 3); hy[p * 32 + c] = __float2bfloat16(fy[p * 32 + c]);
    }
  __nv_bfloat16 *dx, *dy;
  float* dd_out;
  cudaMalloc(&dx, hx.size() * 2); cudaMalloc(&dy, hy.size() * 2); cudaMalloc(&dd_out, 128 * 64 * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dy, hy.data(), hy.size() * 2, cudaMemcpyHostToDevice);
  const int SMEM = 4 * XROW + 2 * DROW + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);

  for (int estride = 1; estride <= 2; ++estride) {
    CUtensorMap mx, my;
    // with an element stride of 2 the box extent counts SOURCE pixels: 2*NPX - 1 source pixels -> NPX loaded
    const int xpx = estride == 1 ? NPX : 120;           // box extents are capped at 256 source pixels
    const int rx = encode_sw64(&mx, dx, P, estride == 1 ? NPX : 2 * xpx - 1, estride);
    const int ry = encode_sw64(&my, dy, P, DPX, 1);
    printf("MN-major SWIZZLE_64B operands from whole-pixel TMA boxes, x element stride %d (encode rc %d %d)\n", estride, rx, ry);
    if (rx || ry) continue;
    for (int swap = 0; swap < 2; ++swap)
      for (int shift = 0; shift < 4; ++shift) {
        printf("   %s  shift %d px: base_offset ->", swap ? "LBO=512 SBO=row" : "LBO=row SBO=512", shift);
        for (uint32_t bo = 0; bo < 8; ++bo) {
          probe_kernel<<<1, 128, SMEM>>>(mx, my, dd_out, shift, bo, swap, 1, xpx * 64);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf(" [%u: CUDA error %s]\n", bo, cudaGetErrorString(e)); return 1; }
          std::vector<float> hd(128 * 64);
          cudaMemcpy(hd.data(), dd_out, hd.size() * 4, cudaMemcpyDeviceToHost);
          int bad = 0;
          for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 64; ++n) {
              const int r = m >> 5, ci = m & 31, q = n >> 5, co = n & 31;
              double ref = 0;
              for (int p = 0; p < KPX; ++p)
                ref += (double)fx[(r * 200 + (shift + p) * estride) * 32 + ci] * fy[(q * 300 + p) * 32 + co];
              if (ref != hd[m * 64 + n]) ++bad;
            }
          printf(" %u:%s", bo, bad == 0 ? "PASS" : "fail");
        }
        printf("\n");
      }
  }
  return 0;
}
