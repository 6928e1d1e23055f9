// 32->32 3x3 convolutions of the encoder as implicit GEMMs on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM), bf16 NHWC activations, fp32 accumulate.
// Reference ops: components.py:20-21,42-43 (c2, c3) and their input gradients.
//
// Formulation ("row marching"): a CTA owns a strip of 128 output pixels of one image and walks
// down ROWS output rows.  GEMM per output row:  D[128 pix x 32 co] = sum over 9 taps of
// A_tap[128 pix x 32 ci] * W_tap[32 ci x 32 co]  -> 18 tcgen05.mma (M=128, N=32, K=16) into one
// 32-column TMEM accumulator.
//
// Shared-memory layout of an input row ("slab"): four channel-group planes, each [pixel][8 ch] =
// 16 B per pixel (core matrices of the SWIZZLE_NONE K-major canonical layout: 8 pixels x 16 B).
// Because the layout is unswizzled, the A operand of tap kw is the SAME slab addressed kw*16 B
// further on: every input row is loaded ONCE per strip and reused by 9 taps / 3 output rows
// (verified on hardware by tools/umma_probe.cu, profiles/r1_umma_descriptor_probe.txt).
// Stride 2 keeps even and odd input pixels in separate planes so that consecutive output pixels
// still read consecutive 16-byte rows.
//
// Warp roles (192 threads): warp 0 = producer (cp.async 16 B, zero-fill = conv padding),
// warp 1 = MMA issuer (one lane) + TMEM owner, warps 2..5 = epilogue (tcgen05.ld -> bias/ReLU or
// ReLU-mask -> bf16 -> global).  mbarrier rings: slab full/empty (8 deep), accumulator
// full/empty (4 x 32 TMEM columns), so loads, MMAs and epilogues of different rows overlap.
#include "dd_common.cuh"
#include "umma.cuh"

namespace {

constexpr int C = 32;
constexpr int TILE_M = 128;
constexpr int ROWS = 32;            // output rows per work item
constexpr int RING = 8;             // slab ring depth
constexpr int INFLIGHT = 4;         // cp.async groups in flight before the oldest is published
constexpr int NACC = 4;             // TMEM accumulator buffers (32 columns each)
constexpr int NPAD = 136;           // pixels per plane (>= 130, and 129 per parity plane for stride 2)
constexpr int PS = NPAD * 16;       // plane stride in bytes
constexpr int W_BYTES = 9 * 4 * 512;  // bf16 weights [tap][cg][co][8 ci]
constexpr int NTHREADS = 192;

template <int STRIDE>
struct Geo {
  static constexpr int NPIX = (TILE_M - 1) * STRIDE + 3;      // input pixels per slab: 130 / 257
  static constexpr int PLANES = 4 * STRIDE;                   // stride 2: even + odd pixel planes
  static constexpr int SLAB_BYTES = PLANES * PS;
  static constexpr int SMEM = W_BYTES + RING * SLAB_BYTES + 1024;
};

struct Bars {
  uint64_t full[RING], empty[RING], acc_full[NACC], acc_empty[NACC];
  uint32_t tmem_base;
};

// MODE 0: forward (bias + ReLU).  MODE 1: stride-1 input gradient (flipped/transposed filter,
// epilogue multiplies by mask > 0).
template <int STRIDE, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) conv3x3_c32_tc_kernel(const __nv_bfloat16* __restrict__ in,
                                                                      const float* __restrict__ w_oihw,
                                                                      const float* __restrict__ bias,
                                                                      const __nv_bfloat16* __restrict__ mask,
                                                                      __nv_bfloat16* __restrict__ out, int B, int H,
                                                                      int W, int Ho, int Wo) {
  using G = Geo<STRIDE>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_slab = smem + W_BYTES;
  Bars* bars = reinterpret_cast<Bars*>(smem + W_BYTES + RING * G::SLAB_BYTES);
  __shared__ float s_bias[C];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wtiles = (Wo + TILE_M - 1) / TILE_M;
  const int hsegs = (Ho + ROWS - 1) / ROWS;
  const int items = B * wtiles * hsegs;

  // ---- one-time setup: weights -> bf16 UMMA layout, barriers, TMEM ---------------------------
  for (int i = tid; i < 9 * C * C; i += NTHREADS) {
    const int ci = i & 31, co = (i >> 5) & 31, tap = i >> 10;     // (in-channel role, out-channel role)
    float v;
    if (MODE == 0) v = w_oihw[(co * C + ci) * 9 + tap];
    else v = w_oihw[(ci * C + co) * 9 + (8 - tap)];               // dgrad: W[co=in][ci=out][flipped tap]
    *reinterpret_cast<__nv_bfloat16*>(s_w + (tap * 4 + (ci >> 3)) * 512 + co * 16 + (ci & 7) * 2) = __float2bfloat16_rn(v);
  }
  if (tid < C) s_bias[tid] = (MODE == 0) ? bias[tid] : 0.f;
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) { umma::mbar_init(&bars->full[i], 1); umma::mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < NACC; ++i) { umma::mbar_init(&bars->acc_full[i], 1); umma::mbar_init(&bars->acc_empty[i], 4); }
    umma::fence_mbar_init();
  }
  if (warp == 1) umma::tmem_alloc(&bars->tmem_base, NACC * 32);
  umma::fence_proxy_async_smem();      // weights were written with generic stores
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // =========================== producer ======================================================
    uint32_t g = 0;                      // global slab counter
    uint32_t published = 0;              // slabs whose full barrier has been signalled
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int wt = it % wtiles, hs = (it / wtiles) % hsegs, b = it / (wtiles * hsegs);
      const int ho0 = hs * ROWS;
      const int rows = min(ROWS, Ho - ho0);
      const int nslabs = (rows - 1) * STRIDE + 3;
      const int r0 = ho0 * STRIDE - 1;                 // first input row
      const int c0 = wt * TILE_M * STRIDE - 1;         // first input column
      const __nv_bfloat16* img = in + (size_t)b * H * W * C;
      for (int s = 0; s < nslabs; ++s, ++g) {
        const uint32_t slot = g % RING;
        umma::mbar_wait(&bars->empty[slot], ((g / RING) & 1) ^ 1);
        const int r = r0 + s;
        const bool row_ok = (r >= 0) && (r < H);
        const uint32_t dst0 = umma::smem_u32(s_slab + slot * G::SLAB_BYTES);
        const __nv_bfloat16* rowp = img + (size_t)(row_ok ? r : 0) * W * C;
        for (int c = lane; c < G::NPIX * 4; c += 32) {
          const int li = c >> 2, cg = c & 3;
          const int col = c0 + li;
          const bool ok = row_ok && (col >= 0) && (col < W);
          uint32_t dst;
          if (STRIDE == 1) dst = dst0 + cg * PS + li * 16;
          else dst = dst0 + ((li & 1) * 4 + cg) * PS + (li >> 1) * 16;
          umma::cp_async16(dst, rowp + (size_t)(ok ? col : 0) * C + cg * 8, ok ? 16u : 0u);
        }
        umma::cp_async_commit();
        if (g + 1 - published >= INFLIGHT) {           // publish the oldest in-flight slab
          umma::cp_async_wait<INFLIGHT - 1>();
          umma::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(&bars->full[published % RING]);
          ++published;
        }
      }
    }
    umma::cp_async_wait<0>();
    umma::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0)
      for (; published < g; ++published) umma::mbar_arrive(&bars->full[published % RING]);
  } else if (warp == 1) {
    // =========================== MMA issuer ====================================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma::make_idesc_bf16(TILE_M, C, false, false);
      const uint32_t wbase = umma::smem_u32(s_w);
      const uint32_t sbase = umma::smem_u32(s_slab);
      uint32_t g0 = 0;            // slab counter at the start of the current item
      uint32_t waited = 0;        // slabs [0, waited) are known to be full
      uint32_t row_ctr = 0;       // global output-row counter (accumulator ring)
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int hs = (it / wtiles) % hsegs;
        const int rows = min(ROWS, Ho - hs * ROWS);
        const int nslabs = (rows - 1) * STRIDE + 3;
        for (int j = 0; j < rows; ++j, ++row_ctr) {
          const uint32_t need = g0 + j * STRIDE + 3;
          for (; waited < need; ++waited) umma::mbar_wait(&bars->full[waited % RING], (waited / RING) & 1);
          const uint32_t buf = row_ctr % NACC;
          umma::mbar_wait(&bars->acc_empty[buf], ((row_ctr / NACC) & 1) ^ 1);
          umma::tc_fence_after_sync();
          const uint32_t d_tmem = tmem + buf * 32;
          uint32_t first = 1;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh) {
            const uint32_t slab = sbase + ((g0 + j * STRIDE + kh) % RING) * G::SLAB_BYTES;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              uint32_t a_off;
              if (STRIDE == 1) a_off = kw * 16;
              else a_off = (kw == 1 ? 4 * PS : 0) + (kw == 2 ? 16 : 0);    // kw: 0 -> even[i], 1 -> odd[i], 2 -> even[i+1]
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint64_t da = umma::make_desc(slab + a_off + (2 * ks) * PS, PS, 128);
                const uint64_t db = umma::make_desc(wbase + ((kh * 3 + kw) * 4 + 2 * ks) * 512, 512, 128);
                umma::mma_bf16(d_tmem, da, db, idesc, first ? 0u : 1u);
                first = 0;
              }
            }
          }
          umma::mma_commit(&bars->acc_full[buf]);
          // input rows that no later output row of this item reads
          const int nrel = (j == rows - 1) ? 3 : STRIDE;
          for (int q = 0; q < nrel; ++q) umma::mma_commit(&bars->empty[(g0 + j * STRIDE + q) % RING]);
        }
        g0 += nslabs;
      }
    }
  } else {
    // =========================== epilogue (warps 2..5) =========================================
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) belong to this warp
    uint32_t row_ctr = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int wt = it % wtiles, hs = (it / wtiles) % hsegs, b = it / (wtiles * hsegs);
      const int ho0 = hs * ROWS;
      const int rows = min(ROWS, Ho - ho0);
      const int wo = wt * TILE_M + quarter * 32 + lane;
      for (int j = 0; j < rows; ++j, ++row_ctr) {
        const uint32_t buf = row_ctr % NACC;
        umma::mbar_wait(&bars->acc_full[buf], (row_ctr / NACC) & 1);
        umma::tc_fence_after_sync();
        uint32_t r[32];
        umma::tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + buf * 32, r);
        umma::tmem_ld_wait();
        umma::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(&bars->acc_empty[buf]);
        if (wo < Wo) {
          const size_t off = (((size_t)b * Ho + ho0 + j) * Wo + wo) * C;
          float v[C];
#pragma unroll
          for (int k = 0; k < C; ++k) v[k] = __uint_as_float(r[k]);
          if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < C; ++k) v[k] = fmaxf(v[k] + s_bias[k], 0.f);
          } else if (mask) {
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              float m[8];
              dd::ld8<__nv_bfloat16>(mask + off + gq * 8, m);
#pragma unroll
              for (int k = 0; k < 8; ++k) v[gq * 8 + k] = m[k] > 0.f ? v[gq * 8 + k] : 0.f;
            }
          }
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            float t8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t8[k] = v[gq * 8 + k];
            dd::st8<__nv_bfloat16>(out + off + gq * 8, t8);
          }
        }
      }
    }
  }
  // ---- teardown ---------------------------------------------------------------------------------
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, NACC * 32);
}

template <int STRIDE, int MODE>
int launch(const void* in, const float* w, const float* bias, const void* mask, void* out, int B, int H, int W,
           cudaStream_t st) {
  using G = Geo<STRIDE>;
  const int Ho = (H - 1) / STRIDE + 1, Wo = (W - 1) / STRIDE + 1;
  const int items = B * ((Wo + TILE_M - 1) / TILE_M) * ((Ho + ROWS - 1) / ROWS);
  auto k = conv3x3_c32_tc_kernel<STRIDE, MODE>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
  if (e != cudaSuccess) return dd::fail((int)e, "conv_tc: cudaFuncSetAttribute(%d): %s", G::SMEM, cudaGetErrorString(e));
  const int grid = items < dd::kSMs ? items : dd::kSMs;
  k<<<grid, NTHREADS, G::SMEM, st>>>((const __nv_bfloat16*)in, w, bias, (const __nv_bfloat16*)mask,
                                     (__nv_bfloat16*)out, B, H, W, Ho, Wo);
  return dd::check_launch("conv3x3_c32_tc");
}


// ================================================================================================
// Input gradient of the stride-2 conv (c3) = stride-2 transposed conv, on the tensor cores.
//   dx[h,w,ci] = sum over taps with (h+1-kh), (w+1-kw) even of dy[(h+1-kh)/2,(w+1-kw)/2,:] . W[:,ci,kh,kw]
// A CTA marches down dy rows m for a strip of 128 column PAIRS i (dx columns 2i, 2i+1).  Each m
// yields dx rows 2m and 2m+1; even and odd columns are separate GEMMs over the same dy slabs
// (plane layout, tap = 16-byte start shift), 9 taps = 18 MMAs per m into four 32-column TMEM
// accumulators {row 2m, 2m+1} x {even, odd}; the epilogue interleaves even/odd so that each thread
// stores two adjacent pixels (128 contiguous bytes) after the ReLU mask of the layer input.
// ================================================================================================
constexpr int DG_MROWS = 16;      // dy rows per work item (32 dx rows)
constexpr int DG_SMEM = W_BYTES + RING * 4 * PS + 1024;

struct DgBars {
  uint64_t full[RING], empty[RING], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(NTHREADS, 1) conv3x3_c32_dgrad_s2_tc_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                               const float* __restrict__ w_oihw,
                                                                               const __nv_bfloat16* __restrict__ mask,
                                                                               __nv_bfloat16* __restrict__ dx, int B,
                                                                               int H, int W, int Ho, int Wo) {
  constexpr int SLAB = 4 * PS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* s_w = smem;
  uint8_t* s_slab = smem + W_BYTES;
  DgBars* bars = reinterpret_cast<DgBars*>(smem + W_BYTES + RING * SLAB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Hp = (H + 1) / 2, Wp = (W + 1) / 2;
  const int wtiles = (Wp + TILE_M - 1) / TILE_M;
  const int hsegs = (Hp + DG_MROWS - 1) / DG_MROWS;
  const int items = B * wtiles * hsegs;

  // B operand of tap t: [n = ci][k = co] = W[co][ci][t]
  for (int i = tid; i < 9 * C * C; i += NTHREADS) {
    const int co = i & 31, ci = (i >> 5) & 31, tap = i >> 10;
    *reinterpret_cast<__nv_bfloat16*>(s_w + (tap * 4 + (co >> 3)) * 512 + ci * 16 + (co & 7) * 2) =
        __float2bfloat16_rn(w_oihw[(co * C + ci) * 9 + tap]);
  }
  if (tid == 0) {
    for (int i = 0; i < RING; ++i) { umma::mbar_init(&bars->full[i], 1); umma::mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&bars->acc_full[i], 1); umma::mbar_init(&bars->acc_empty[i], 4); }
    umma::fence_mbar_init();
  }
  if (warp == 1) umma::tmem_alloc(&bars->tmem_base, 256);
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    // =========================== producer: dy rows m0 .. m0+rows ================================
    uint32_t g = 0, published = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int wt = it % wtiles, hs = (it / wtiles) % hsegs, b = it / (wtiles * hsegs);
      const int m0 = hs * DG_MROWS;
      const int rows = min(DG_MROWS, Hp - m0);
      const int i0 = wt * TILE_M;
      const __nv_bfloat16* img = dy + (size_t)b * Ho * Wo * C;
      for (int s = 0; s <= rows; ++s, ++g) {
        const uint32_t slot = g % RING;
        umma::mbar_wait(&bars->empty[slot], ((g / RING) & 1) ^ 1);
        const int r = m0 + s;
        const bool row_ok = r < Ho;
        const uint32_t dst0 = umma::smem_u32(s_slab + slot * SLAB);
        const __nv_bfloat16* rowp = img + (size_t)(row_ok ? r : 0) * Wo * C;
        for (int c = lane; c < 129 * 4; c += 32) {
          const int li = c >> 2, cg = c & 3;
          const int col = i0 + li;
          const bool ok = row_ok && col < Wo;
          umma::cp_async16(dst0 + cg * PS + li * 16, rowp + (size_t)(ok ? col : 0) * C + cg * 8, ok ? 16u : 0u);
        }
        umma::cp_async_commit();
        if (g + 1 - published >= INFLIGHT) {
          umma::cp_async_wait<INFLIGHT - 1>();
          umma::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(&bars->full[published % RING]);
          ++published;
        }
      }
    }
    umma::cp_async_wait<0>();
    umma::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0)
      for (; published < g; ++published) umma::mbar_arrive(&bars->full[published % RING]);
  } else if (warp == 1) {
    if (lane == 0) {
      // =========================== MMA issuer ==================================================
      constexpr uint32_t idesc = umma::make_idesc_bf16(TILE_M, C, false, false);
      const uint32_t wbase = umma::smem_u32(s_w);
      const uint32_t sbase = umma::smem_u32(s_slab);
      // {accumulator (0: row 2m even, 1: row 2m odd, 2: row 2m+1 even, 3: row 2m+1 odd), kh, kw, slab (0: m, 1: m+1), shift}
      constexpr int T[9][5] = {{0, 1, 1, 0, 0}, {1, 1, 0, 0, 1}, {1, 1, 2, 0, 0}, {2, 0, 1, 1, 0}, {2, 2, 1, 0, 0},
                               {3, 0, 0, 1, 1}, {3, 0, 2, 1, 0}, {3, 2, 0, 0, 1}, {3, 2, 2, 0, 0}};
      uint32_t g0 = 0, waited = 0, mctr = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int hs = (it / wtiles) % hsegs;
        const int rows = min(DG_MROWS, Hp - hs * DG_MROWS);
        for (int j = 0; j < rows; ++j, ++mctr) {
          const uint32_t need = g0 + j + 2;
          for (; waited < need; ++waited) umma::mbar_wait(&bars->full[waited % RING], (waited / RING) & 1);
          const uint32_t set = mctr & 1;
          umma::mbar_wait(&bars->acc_empty[set], ((mctr >> 1) & 1) ^ 1);
          umma::tc_fence_after_sync();
          uint32_t used = 0;      // bit a set once accumulator a has been written for this m
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const int a = T[t][0], tap = T[t][1] * 3 + T[t][2];
            const uint32_t slab = sbase + ((g0 + j + T[t][3]) % RING) * SLAB + T[t][4] * 16;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t da = umma::make_desc(slab + (2 * ks) * PS, PS, 128);
              const uint64_t db = umma::make_desc(wbase + (tap * 4 + 2 * ks) * 512, 512, 128);
              umma::mma_bf16(tmem + set * 128 + a * 32, da, db, idesc, (used >> a) & 1u);
              used |= 1u << a;
            }
          }
          umma::mma_commit(&bars->acc_full[set]);
          umma::mma_commit(&bars->empty[(g0 + j) % RING]);
          if (j == rows - 1) umma::mma_commit(&bars->empty[(g0 + j + 1) % RING]);
        }
        g0 += rows + 1;
      }
    }
  } else {
    // =========================== epilogue (warps 2..5) =========================================
    const int quarter = warp & 3;
    uint32_t mctr = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int wt = it % wtiles, hs = (it / wtiles) % hsegs, b = it / (wtiles * hsegs);
      const int m0 = hs * DG_MROWS;
      const int rows = min(DG_MROWS, Hp - m0);
      const int i = wt * TILE_M + quarter * 32 + lane;
      for (int j = 0; j < rows; ++j, ++mctr) {
        const uint32_t set = mctr & 1;
        umma::mbar_wait(&bars->acc_full[set], (mctr >> 1) & 1);
        umma::tc_fence_after_sync();
#pragma unroll 1
        for (int a = 0; a < 4; ++a) {
          uint32_t r[32];
          umma::tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + set * 128 + a * 32, r);
          umma::tmem_ld_wait();
          if (a == 3) {            // all four accumulators of this set are in registers / stored
            umma::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) umma::mbar_arrive(&bars->acc_empty[set]);
          }
          const int h = 2 * (m0 + j) + (a >> 1), w = 2 * i + (a & 1);
          if (h < H && w < W) {
            const size_t off = (((size_t)b * H + h) * W + w) * C;
            float v[C];
#pragma unroll
            for (int k = 0; k < C; ++k) v[k] = __uint_as_float(r[k]);
            if (mask) {
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                float mk[8];
                dd::ld8<__nv_bfloat16>(mask + off + gq * 8, mk);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[gq * 8 + k] = mk[k] > 0.f ? v[gq * 8 + k] : 0.f;
              }
            }
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
              float t8[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) t8[k] = v[gq * 8 + k];
              dd::st8<__nv_bfloat16>(dx + off + gq * 8, t8);
            }
          }
        }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 256);
}

// ================================================================================================
// Weight gradient on the tensor cores (stride 1):
//   dW[co][ci][kh][kw] = sum_{b,h,w} x[b,h+kh-1,w+kw-1,ci] * dy[b,h,w,co]
// The contraction runs over PIXELS, so both operands are MN-major: A = x planes (M = 4 input rows x
// 32 ci = 128), B = dy planes (N = 2 dy rows x 32 co = 64), K = 16 pixels per tcgen05.mma.  For the
// dy-row pair (h, h+1) and x rows h-1..h+2, block (r, q) of D holds tap kh = r - q; 6 of the 8
// blocks are useful.  kw is again a 16-byte shift of the A start address, one 64-column TMEM
// accumulator per kw, accumulated over every work item of the CTA and written out ONCE at the end
// as [cta][q][tap][ci][co] partials that an ordered reduction kernel folds (deterministic).
// Stage = 4 dy rows + 6 x rows of a 128-pixel column strip, double buffered; warps 0,1,3 are
// cp.async producers, warp 2 issues the MMAs, all four warps run the final epilogue.
// ================================================================================================
// Stride 2 (c3): one dy row per stage, x rows 2ho-1..2ho+1 (+ one junk row that only feeds the unused
// D block r = 3) in [parity][row][cg] planes so that the (row, cg) chunks keep one uniform stride;
// N = 32, every useful block has q = 0.
constexpr int PSD = TILE_M * 16;                 // dy plane stride (128 pixels)
template <int STRIDE>
struct WgGeo {
  static constexpr int ROWS = STRIDE == 1 ? 4 : 1;            // dy rows per stage
  static constexpr int XROWS = STRIDE == 1 ? 6 : 3;           // x rows loaded per stage
  static constexpr int XROWS_ALLOC = STRIDE == 1 ? 6 : 4;
  static constexpr int XPIX = STRIDE == 1 ? 130 : 257;        // x pixels per row
  static constexpr int X_BYTES = STRIDE * XROWS_ALLOC * 4 * PS;
  static constexpr int DY_BYTES = ROWS * 4 * PSD;
  static constexpr int STAGE_BYTES = X_BYTES + DY_BYTES;
  static constexpr int SMEM = 2 * STAGE_BYTES + 1024;
  static constexpr int NQ = STRIDE == 1 ? 2 : 1;              // q slots in the per-CTA partials
  static constexpr int PARTIAL = NQ * 9 * C * C;              // floats per CTA
  static constexpr int N = 32 * NQ;
};

struct WgBars {
  uint64_t full[2], empty[2], done;
  uint32_t tmem_base;
};

template <int STRIDE>
__global__ void __launch_bounds__(128, 1) conv3x3_c32_wgrad_tc_kernel(const __nv_bfloat16* __restrict__ x,
                                                                       const __nv_bfloat16* __restrict__ dy,
                                                                       float* __restrict__ partial, int B, int H,
                                                                       int W, int Ho, int Wo) {
  using G = WgGeo<STRIDE>;
  constexpr int WG_ROWS = G::ROWS, WG_XROWS = G::XROWS, WG_X_BYTES = G::X_BYTES, WG_STAGE_BYTES = G::STAGE_BYTES;
  constexpr int WG_PARTIAL = G::PARTIAL;
  extern __shared__ __align__(1024) uint8_t smem[];
  WgBars* bars = reinterpret_cast<WgBars*>(smem + 2 * WG_STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wtiles = (Wo + TILE_M - 1) / TILE_M;
  const int hsegs = (Ho + WG_ROWS - 1) / WG_ROWS;
  const int items = B * wtiles * hsegs;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&bars->full[i], 3); umma::mbar_init(&bars->empty[i], 1); }
    umma::mbar_init(&bars->done, 1);
    umma::fence_mbar_init();
  }
  if (warp == 2) umma::tmem_alloc(&bars->tmem_base, 256);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = bars->tmem_base;

  if (warp != 2) {
    // =========================== producers (warps 0, 1, 3) =====================================
    const int pl = (warp == 3 ? 2 : warp) * 32 + lane;      // 0..95
    uint32_t n = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int wt = it % wtiles, hs = (it / wtiles) % hsegs, b = it / (wtiles * hsegs);
      const int h0 = hs * WG_ROWS, w0 = wt * TILE_M;
      const uint32_t stage = n & 1;
      umma::mbar_wait(&bars->empty[stage], ((n >> 1) & 1) ^ 1);
      const uint32_t xs = umma::smem_u32(smem + stage * WG_STAGE_BYTES);
      const uint32_t ds = xs + WG_X_BYTES;
      const __nv_bfloat16* ximg = x + (size_t)b * H * W * C;
      const __nv_bfloat16* dimg = dy + (size_t)b * Ho * Wo * C;
      constexpr int XCH = WG_XROWS * G::XPIX * 4;
      for (int c = pl; c < XCH; c += 96) {
        const int cg = c & 3, t = c >> 2;
        const int r = t / G::XPIX, li = t - r * G::XPIX;
        const int row = h0 * STRIDE - 1 + r, col = w0 * STRIDE - 1 + li;
        const bool ok = row >= 0 && row < H && col >= 0 && col < W;
        uint32_t dst;
        if (STRIDE == 1) dst = xs + (r * 4 + cg) * PS + li * 16;
        else dst = xs + ((li & 1) * (G::XROWS_ALLOC * 4) + r * 4 + cg) * PS + (li >> 1) * 16;
        umma::cp_async16(dst, ximg + ((size_t)(ok ? row : 0) * W + (ok ? col : 0)) * C + cg * 8, ok ? 16u : 0u);
      }
      constexpr int DCH = WG_ROWS * TILE_M * 4;
      for (int c = pl; c < DCH; c += 96) {
        const int cg = c & 3, t = c >> 2;
        const int r = t >> 7, li = t & 127;
        const int row = h0 + r, col = w0 + li;
        const bool ok = row < Ho && col < Wo;
        umma::cp_async16(ds + (r * 4 + cg) * PSD + li * 16,
                         dimg + ((size_t)(ok ? row : 0) * Wo + (ok ? col : 0)) * C + cg * 8, ok ? 16u : 0u);
      }
      umma::cp_async_commit();
      umma::cp_async_wait<0>();
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&bars->full[stage]);
    }
  } else if (lane == 0) {
    // =========================== MMA issuer ====================================================
    constexpr uint32_t idesc = umma::make_idesc_bf16(128, G::N, true, true);
    uint32_t n = 0;
    uint32_t fresh = 1;     // accumulators not yet written
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const uint32_t stage = n & 1;
      umma::mbar_wait(&bars->full[stage], (n >> 1) & 1);
      umma::tc_fence_after_sync();
      const uint32_t xs = umma::smem_u32(smem + stage * WG_STAGE_BYTES);
      const uint32_t ds = xs + WG_X_BYTES;
#pragma unroll
      for (int p = 0; p < (STRIDE == 1 ? WG_ROWS / 2 : 1); ++p) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          // stride 1: x column w+kw-1 = slab index (w-w0)+kw.  stride 2: kw 0 -> even[i], 1 -> odd[i], 2 -> even[i+1]
          const uint32_t a_off = STRIDE == 1 ? (2 * p * 4) * PS + kw * 16
                                             : (kw == 1 ? G::XROWS_ALLOC * 4 * PS : 0) + (kw == 2 ? 16 : 0);
#pragma unroll
          for (int ks = 0; ks < TILE_M / 16; ++ks) {
            const uint64_t da = umma::make_desc(xs + a_off + ks * 256, 128, PS);
            const uint64_t db = umma::make_desc(ds + (2 * p * 4) * PSD + ks * 256, 128, PSD);
            umma::mma_bf16(tmem + kw * 64, da, db, idesc, (fresh && p == 0 && ks == 0) ? 0u : 1u);
          }
        }
      }
      fresh = 0;
      umma::mma_commit(&bars->empty[stage]);
    }
    umma::mma_commit(&bars->done);
  }
  // =========================== epilogue: TMEM -> per-CTA partials ================================
  __syncwarp();
  umma::mbar_wait(&bars->done, 0);
  umma::tc_fence_after_sync();
  {
    const int r = warp;                      // TMEM lane quarter = x row offset r; lane = ci
    float* out = partial + (size_t)blockIdx.x * WG_PARTIAL;
#pragma unroll 1
    for (int kw = 0; kw < 3; ++kw) {
#pragma unroll 1
      for (int q = 0; q < G::NQ; ++q) {
        uint32_t v[32];
        umma::tmem_ld_32x32(tmem + ((uint32_t)(r * 32) << 16) + kw * 64 + q * 32, v);
        umma::tmem_ld_wait();
        const int kh = r - q;
        if (kh >= 0 && kh <= 2) {
          float4* dst = reinterpret_cast<float4*>(out + (size_t)q * 9 * C * C + ((kh * 3 + kw) * C + lane) * C);
#pragma unroll
          for (int g4 = 0; g4 < 8; ++g4)
            dst[g4] = make_float4(__uint_as_float(v[4 * g4]), __uint_as_float(v[4 * g4 + 1]),
                                  __uint_as_float(v[4 * g4 + 2]), __uint_as_float(v[4 * g4 + 3]));
        }
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 2) umma::tmem_dealloc(tmem, 256);
}

// per-channel sums of an NHWC bf16 tensor (bias gradient): partial[blk][32]
__global__ void __launch_bounds__(256) colsum_nhwc_bf16_kernel(const __nv_bfloat16* __restrict__ dy, long long npix,
                                                               float* __restrict__ partial) {
  __shared__ float red[64][33];
  const int cg = threadIdx.x & 3, pl = threadIdx.x >> 2;     // 64 pixel lanes x 4 channel groups
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long p = (long long)blockIdx.x * 64 + pl; p < npix; p += (long long)gridDim.x * 64) {
    float v[8];
    dd::ld8<__nv_bfloat16>(dy + p * C + cg * 8, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += v[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[pl][cg * 8 + k] = s[k];
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = 0.f;
    for (int i = 0; i < 64; ++i) t += red[i][threadIdx.x];
    partial[(size_t)blockIdx.x * C + threadIdx.x] = t;
  }
}

// dw[co][ci][tap] = sum over CTAs and their q slots (nslots = CTAs x NQ); db[co] = sum over colsum CTAs
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ partial, int nslots, const float* __restrict__ dbp,
                                       int ndb, float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 9 * C * C) {
    float s = 0.f;
    for (int blk = 0; blk < nslots; ++blk) s += partial[(size_t)blk * (9 * C * C) + i];
    const int co = i & 31, ci = (i >> 5) & 31, tap = i >> 10;
    dw[(co * C + ci) * 9 + tap] = s;
  } else if (i < 9 * C * C + C) {
    const int co = i - 9 * C * C;
    float s = 0.f;
    for (int blk = 0; blk < ndb; ++blk) s += dbp[(size_t)blk * C + co];
    db[co] = s;
  }
}

constexpr int kDbBlocks = dd::kSMs * 4;

}  // namespace

namespace dd {

// mode: 0 forward, 1 input gradient, 2 weight gradient
bool conv_tc_supported(int H, int W, int stride, int mode) {
  if (H < 1 || W < 1) return false;
  if (mode == 0) return stride == 1 || stride == 2;
  return mode >= 0 && mode <= 2 && (stride == 1 || stride == 2);
}

int conv3x3_c32_fwd_tc(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int stride,
                       int mode, const void* mask, cudaStream_t st) {
  if (mode == 0 && stride == 1) return launch<1, 0>(in, w, bias, nullptr, out, B, H, W, st);
  if (mode == 0 && stride == 2) return launch<2, 0>(in, w, bias, nullptr, out, B, H, W, st);
  if (mode == 1 && stride == 1) return launch<1, 1>(in, w, nullptr, mask, out, B, H, W, st);
  if (mode == 1 && stride == 2) {
    // here `in` = dy [B,Ho,Wo,32], `out` = dx [B,H,W,32]
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int items = B * (((W + 1) / 2 + TILE_M - 1) / TILE_M) * (((H + 1) / 2 + DG_MROWS - 1) / DG_MROWS);
    cudaError_t e = cudaFuncSetAttribute(conv3x3_c32_dgrad_s2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DG_SMEM);
    if (e != cudaSuccess) return fail((int)e, "dgrad_s2_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    conv3x3_c32_dgrad_s2_tc_kernel<<<items < kSMs ? items : kSMs, NTHREADS, DG_SMEM, st>>>(
        (const __nv_bfloat16*)in, w, (const __nv_bfloat16*)mask, (__nv_bfloat16*)out, B, H, W, Ho, Wo);
    return check_launch("conv3x3_c32_dgrad_s2_tc");
  }
  return fail(DD_ERR_UNSUPPORTED, "conv_tc: mode %d stride %d", mode, stride);
}

template <int STRIDE>
static int wgrad_tc_launch(const void* x, const void* dy, float* dw, float* db, void* ws, size_t ws_bytes, int B, int H,
                           int W, cudaStream_t st) {
  using G = WgGeo<STRIDE>;
  const int Ho = (H - 1) / STRIDE + 1, Wo = (W - 1) / STRIDE + 1;
  const int items = B * ((Wo + TILE_M - 1) / TILE_M) * ((Ho + G::ROWS - 1) / G::ROWS);
  const int grid = items < kSMs ? items : kSMs;
  const size_t need = ((size_t)grid * G::PARTIAL + (size_t)kDbBlocks * C) * sizeof(float);
  if (ws_bytes < need) return fail(DD_ERR_WORKSPACE, "tcgen05 wgrad: workspace %zu < %zu", ws_bytes, need);
  float* partial = (float*)ws;
  float* dbp = partial + (size_t)grid * G::PARTIAL;
  auto k = conv3x3_c32_wgrad_tc_kernel<STRIDE>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
  if (e != cudaSuccess) return fail((int)e, "wgrad_tc: cudaFuncSetAttribute(%d): %s", G::SMEM, cudaGetErrorString(e));
  k<<<grid, 128, G::SMEM, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, partial, B, H, W, Ho, Wo);
  if (int err = check_launch("conv3x3_c32_wgrad_tc")) return err;
  const long long npix = (long long)B * Ho * Wo;
  const int dbg = (int)((npix + 63) / 64 < kDbBlocks ? (npix + 63) / 64 : kDbBlocks);
  colsum_nhwc_bf16_kernel<<<dbg, 256, 0, st>>>((const __nv_bfloat16*)dy, npix, dbp);
  if (int err = check_launch("colsum_nhwc_bf16")) return err;
  wgrad_tc_reduce_kernel<<<(9 * C * C + C + 255) / 256, 256, 0, st>>>(partial, grid * G::NQ, dbp, dbg, dw, db);
  return check_launch("wgrad_tc_reduce");
}

int conv3x3_c32_wgrad_tc(const void* x, const void* dy, float* dw, float* db, void* ws, size_t ws_bytes, int B, int H,
                         int W, int stride, cudaStream_t st) {
  if (stride == 1) return wgrad_tc_launch<1>(x, dy, dw, db, ws, ws_bytes, B, H, W, st);
  if (stride == 2) return wgrad_tc_launch<2>(x, dy, dw, db, ws, ws_bytes, B, H, W, st);
  return fail(DD_ERR_UNSUPPORTED, "tcgen05 wgrad: stride %d", stride);
}

}  // namespace dd
