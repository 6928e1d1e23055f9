// Skinny, weight-streaming linear layers on the 5th-gen tensor cores: fp32 weights go HBM -> TMA
// (128-byte-swizzled boxes) -> shared memory -> tcgen05.mma kind::tf32 -> TMEM, once per pass.
//   Encoder.fc1.fc1  Linear(940,032 -> hidden)   components.py:26,105
//   roadmap head     Linear(latent -> 640,000)   roadmap_bce_v2.py:50,75
//   Decoder.fc2.fc1  Linear(hidden -> 1,253,376) components.py:69
// All three passes are HBM-bound (the weight, or its gradient, crosses HBM exactly once per pass over
// the batch); the batch is the narrow MMA dimension: 32 rows per pass in the training step (forward,
// dgrad, wgrad), up to 256 rows per pass in the forward (inference batches; SCfg below).
//
//   fwd   y[b][n]  = sum_k x[b][k] W[n][k]     D[128 n x 32 b]:  A = W tile, K-major (SWIZZLE_128B);
//                                              B = x tile, K-major
//   dgrad dx[b][k] = sum_n dy[b][n] W[n][k]    D[128 k x 32 b]:  A = W^T tile, MN-major -- for tf32 that
//                                              is SWIZZLE_128B_BASE32B (TMA ..._ATOM_32B);  B = dy, K-major
//   wgrad dW[n][k] = sum_b dy[b][n] x[b][k]    D[128 n x 256 k]: A = dy^T, B = x, both MN-major
// Operand descriptors were verified on hardware with tools/umma_tf32_probe.cu
// (profiles/r1_umma_tf32_tma_probe.txt).
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue (TMEM -> registers -> global).  Persistent CTAs walk (tile, split) work items;
// split contractions go to ordered partial buffers folded by a second kernel (deterministic).
#include "dd_common.cuh"
#include "tma_host.h"
#include "umma.cuh"

namespace {

constexpr int NB = 32;                    // batch rows per launch (MMA N, or K for wgrad)
constexpr int kLinearTcMaxRows = 256;     // forward only: rows per pass (SCfg below)
constexpr int THREADS = 192;
constexpr int BOX = 32 * 128;             // a 32-row x 128-byte TMA box
constexpr int kChunk = 32;                // contraction elements per stage (one 128-byte swizzle span)

__device__ __forceinline__ uint8_t* align1024(uint8_t* p) {
  return reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(p) + 1023) & ~static_cast<uintptr_t>(1023));
}
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!umma::mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins > (1u << 22)) __trap();
  }
}

// ------------------------------------------------------------------------------------------------
// fwd (MODE 0) and dgrad (MODE 1): out[b * LD + r] = sum over the contraction, r = row of the 128-row tile
// ------------------------------------------------------------------------------------------------
// NBT = batch rows per pass (the MMA N): 32 for the training step; the inference front end sends up to 256 scenes per
// call, and with NBT = 256 the weights still cross HBM once (they were re-read for every 32 rows: 8 passes over the
// 0.96 GB of Encoder.fc1 at batch 256)
template <int NBT> struct SCfg {
  static constexpr int STAGES = NBT <= 32 ? 8 : NBT <= 64 ? 8 : NBT <= 128 ? 6 : 4;
  static constexpr int STAGE_BYTES = 4 * BOX + NBT * 128;      // A: 128 x 128 B (fwd) or 4 boxes (dgrad); B: NBT rows x 128 B
  static constexpr int SMEM = STAGES * STAGE_BYTES + 2048;
};
constexpr int S_MAX_STAGES = 8;

struct StreamBars {
  uint64_t full[S_MAX_STAGES], empty[S_MAX_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct StreamGeo {
  int B, LD;                // batch rows; output leading dimension (N for fwd, K for dgrad)
  int tiles, splits;        // 128-row output tiles; contraction splits
  int chunks, chunks_per_split;
};

template <int MODE, typename TOUT, int NBT = NB>
__global__ void __launch_bounds__(THREADS, 1) linear_stream_tc_kernel(const __grid_constant__ CUtensorMap map_w,
                                                                       const __grid_constant__ CUtensorMap map_v,
                                                                       const float* __restrict__ bias,
                                                                       TOUT* __restrict__ out, float* __restrict__ partial,
                                                                       StreamGeo g) {
  constexpr int S_STAGES = SCfg<NBT>::STAGES, S_STAGE_BYTES = SCfg<NBT>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  StreamBars* bars = reinterpret_cast<StreamBars*>(smem + S_STAGES * S_STAGE_BYTES);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int items = g.tiles * g.splits;

  if (tid == 0) {
    for (int i = 0; i < S_STAGES; ++i) { umma::mbar_init(&bars->full[i], 1); umma::mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&bars->acc_full[i], 1); umma::mbar_init(&bars->acc_empty[i], 4); }
    umma::fence_mbar_init();
    umma::tma_prefetch_desc(&map_w);
    umma::tma_prefetch_desc(&map_v);
  }
  if (warp == 1) umma::tmem_alloc(&bars->tmem_base, 2 * NBT);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, bars->tmem_base, 0);

  if (warp == 0) {
    // =========================== TMA producer ===================================================
    if (umma::elect_one()) {
      uint32_t c = 0;
      for (int it = blockIdx.x; it < items; it += gridDim.x) {
        const int ti = it % g.tiles, si = it / g.tiles;
        const int c0 = si * g.chunks_per_split;
        const int c1 = min(c0 + g.chunks_per_split, g.chunks);
        for (int ch = c0; ch < c1; ++ch, ++c) {
          const uint32_t s = c % S_STAGES;
          umma::mbar_wait(&bars->empty[s], ((c / S_STAGES) & 1) ^ 1);
          const uint32_t dst = umma::smem_u32(smem + s * S_STAGE_BYTES);
          umma::mbar_expect_tx(&bars->full[s], S_STAGE_BYTES);
          if (MODE == 0) {
            umma::tma_load_2d(dst, &map_w, ch * kChunk, ti * 128, &bars->full[s]);             // W[n tile][k chunk]
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)                                                         // W[n chunk][k tile], 4 x 32 k
              umma::tma_load_2d(dst + i * BOX, &map_w, ti * 128 + 32 * i, ch * kChunk, &bars->full[s]);
          }
          umma::tma_load_2d(dst + 4 * BOX, &map_v, ch * kChunk, 0, &bars->full[s]);             // x / dy [b][chunk]
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer =======================================================
    constexpr uint32_t idesc = umma::make_idesc_tf32(128, NBT, MODE == 1, false);
    uint32_t c = 0, n_item = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n_item) {
      const int si = it / g.tiles;
      const int c0 = si * g.chunks_per_split;
      const int c1 = min(c0 + g.chunks_per_split, g.chunks);
      const uint32_t a = n_item & 1;
      umma::mbar_wait(&bars->acc_empty[a], ((n_item >> 1) & 1) ^ 1);
      umma::tc_fence_after_sync();
      for (int ch = c0; ch < c1; ++ch, ++c) {
        const uint32_t s = c % S_STAGES;
        umma::mbar_wait(&bars->full[s], (c / S_STAGES) & 1);
        umma::tc_fence_after_sync();
        const uint32_t base = umma::smem_u32(smem + s * S_STAGE_BYTES);
        if (umma::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            uint32_t a_lo, a_hi;
            if (MODE == 0) { a_lo = umma::desc_lo(base + k * 32, 16); a_hi = umma::desc_hi_sw128(1024); }
            else { a_lo = umma::desc_lo(base + k * 1024, BOX); a_hi = umma::desc_hi_sw128_base32(512); }
            umma::mma_tf32_lohi(tmem + a * NBT, a_lo, a_hi, umma::desc_lo(base + 4 * BOX + k * 32, 16), umma::desc_hi_sw128(1024),
                                idesc, (ch > c0 || k > 0) ? 1u : 0u);
          }
          umma::mma_commit(&bars->empty[s]);
          if (ch == c1 - 1) umma::mma_commit(&bars->acc_full[a]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue (warps 2..5) =============================================
    const int quarter = warp & 3;
    uint32_t n_item = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n_item) {
      const int ti = it % g.tiles, si = it / g.tiles;
      const uint32_t a = n_item & 1;
      mbar_wait_relaxed(&bars->acc_full[a], (n_item >> 1) & 1);
      umma::tc_fence_after_sync();
      const int r = ti * 128 + quarter * 32 + lane;
      const float bv = (g.splits == 1 && bias && r < g.LD) ? __ldg(bias + r) : 0.f;
#pragma unroll 1
      for (int nb = 0; nb < NBT; nb += 32) {
        uint32_t v[32];
        umma::tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + a * NBT + nb, v);
        umma::tmem_ld_wait();
        const bool last = nb + 32 >= NBT || nb + 32 >= g.B;
        if (last) {                                      // last group read: hand the accumulator back before the stores
          umma::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) umma::mbar_arrive(&bars->acc_empty[a]);
        }
        if (r < g.LD) {
          if (g.splits == 1) {
#pragma unroll
            for (int b = 0; b < 32; ++b)
              if (nb + b < g.B) dd::st<TOUT>(out + (size_t)(nb + b) * g.LD + r, __uint_as_float(v[b]) + bv);
          } else {
            float* p = partial + (size_t)si * g.B * g.LD + (size_t)nb * g.LD + r;
#pragma unroll
            for (int b = 0; b < 32; ++b)
              if (nb + b < g.B) p[(size_t)b * g.LD] = __uint_as_float(v[b]);
          }
        }
        if (last) break;
      }
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 2 * NBT);
}

// out[i] = sum over splits (in order) of partial[s][i] (+ bias[i % LD])
template <typename TOUT>
__global__ void split_fold_kernel(const float* __restrict__ partial, const float* __restrict__ bias, TOUT* __restrict__ out,
                                  int splits, long long n, int LD) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int k = 0;
  for (; k + 4 <= splits; k += 4) {       // four independent chains: the loads overlap, the order stays fixed
    s0 += partial[(size_t)k * n + i];
    s1 += partial[(size_t)(k + 1) * n + i];
    s2 += partial[(size_t)(k + 2) * n + i];
    s3 += partial[(size_t)(k + 3) * n + i];
  }
  for (; k < splits; ++k) s0 += partial[(size_t)k * n + i];
  dd::st<TOUT>(out + i, ((s0 + s1) + (s2 + s3)) + (bias ? __ldg(bias + (i % LD)) : 0.f));
}

// ------------------------------------------------------------------------------------------------
// wgrad: dW[n][k] (+)= sum_b dy[b][n] x[b][k]; one 128 x KT output tile per stage (no accumulation across
// stages), written straight from TMEM to HBM.
// ------------------------------------------------------------------------------------------------
constexpr int W_STAGES = 3;
constexpr int W_STAGE_BYTES = 4 * BOX + 8 * BOX;     // dy^T: 4 boxes of 32 n; x: up to 8 boxes of 32 k
constexpr int W_TPAD = 36;                           // floats per row of a warp's 32 x 32 transpose block
// epilogue warps: 4 for the plain weight gradient (192 threads x ~118 registers: co-resides with the data-parallel update CTA
// of csrc/adam.cu), 8 for the form with Adam folded in (two per TMEM lane quarter, alternate 32-column chunks: it needs
// the bytes in flight)
template <bool ADAM> struct WCfg { static constexpr int EPI = ADAM ? 8 : 4; static constexpr int THREADS = 32 * (2 + EPI); };
constexpr int W_SMEM = W_STAGES * W_STAGE_BYTES + 8 * 32 * W_TPAD * 4 + 2048;

// Adam folded into the weight-gradient epilogue (world size 1): the gradient tile never goes to HBM -- the epilogue reads
// the weight and its two moments where it would have written dW, and writes them back updated: 24 bytes per parameter
// instead of 4 (dW write) + 28 (a separate Adam pass).  Same arithmetic as csrc/adam.cu (torch.optim.Adam, amsgrad off).
struct AdamFuse {
  float* p; float* m; float* v;          // weight [N][K], exp_avg, exp_avg_sq (same layout); p == nullptr: plain weight gradient
  float beta1, beta2, eps, step_size, inv_sqrt_bc2, weight_decay;
};
__device__ __forceinline__ float adam_fused_one(float p, float g, float& m, float& v, const AdamFuse& h) {
  if (h.weight_decay != 0.f) g = fmaf(h.weight_decay, p, g);
  m = m + (1.f - h.beta1) * (g - m);
  v = h.beta2 * v + (1.f - h.beta2) * g * g;
  const float denom = sqrtf(v) * h.inv_sqrt_bc2 + h.eps;
  return p - h.step_size * (m / denom);
}

struct WgradBars {
  uint64_t full[W_STAGES], empty[W_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

struct WgradGeo {
  int B, N;
  long long K;
  int n_tiles, k_tiles, KT;      // KT = 128 or 256 columns per tile
  int accumulate;
  AdamFuse adam;
};

template <bool ADAM>
__global__ void __launch_bounds__(WCfg<ADAM>::THREADS, 1) linear_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                                                 const __grid_constant__ CUtensorMap map_x,
                                                                                 float* __restrict__ dw, const WgradGeo g) {
  constexpr int W_EPI = WCfg<ADAM>::EPI;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  float* stage_out = reinterpret_cast<float*>(smem + W_STAGES * W_STAGE_BYTES);
  WgradBars* bars = reinterpret_cast<WgradBars*>(smem + W_STAGES * W_STAGE_BYTES + 8 * 32 * W_TPAD * 4);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const long long items = (long long)g.n_tiles * g.k_tiles;
  const int xboxes = g.KT / 32;
  const uint32_t stage_tx = (4 + xboxes) * BOX;

  if (tid == 0) {
    for (int i = 0; i < W_STAGES; ++i) { umma::mbar_init(&bars->full[i], 1); umma::mbar_init(&bars->empty[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&bars->acc_full[i], 1); umma::mbar_init(&bars->acc_empty[i], W_EPI); }
    umma::fence_mbar_init();
    umma::tma_prefetch_desc(&map_dy);
    umma::tma_prefetch_desc(&map_x);
  }
  if (warp == 1) umma::tmem_alloc(&bars->tmem_base, 512);
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tmem = __shfl_sync(0xffffffffu, bars->tmem_base, 0);

  // item -> (n tile, k tile): consecutive items of a CTA share the n tile (its dy^T boxes stay in L2)
  if (warp == 0) {
    if (umma::elect_one()) {
      uint32_t c = 0;
      for (long long it = blockIdx.x; it < items; it += gridDim.x, ++c) {
        const int ni = (int)(it / g.k_tiles), ki = (int)(it % g.k_tiles);
        const uint32_t s = c % W_STAGES;
        umma::mbar_wait(&bars->empty[s], ((c / W_STAGES) & 1) ^ 1);
        const uint32_t dst = umma::smem_u32(smem + s * W_STAGE_BYTES);
        umma::mbar_expect_tx(&bars->full[s], stage_tx);
#pragma unroll
        for (int i = 0; i < 4; ++i) umma::tma_load_2d(dst + i * BOX, &map_dy, ni * 128 + 32 * i, 0, &bars->full[s]);
        for (int j = 0; j < xboxes; ++j) umma::tma_load_2d(dst + (4 + j) * BOX, &map_x, ki * g.KT + 32 * j, 0, &bars->full[s]);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma::make_idesc_tf32(128, g.KT, true, true);
    uint32_t c = 0;
    for (long long it = blockIdx.x; it < items; it += gridDim.x, ++c) {
      const uint32_t s = c % W_STAGES, a = c & 1;
      umma::mbar_wait(&bars->acc_empty[a], ((c >> 1) & 1) ^ 1);
      umma::mbar_wait(&bars->full[s], (c / W_STAGES) & 1);
      umma::tc_fence_after_sync();
      const uint32_t base = umma::smem_u32(smem + s * W_STAGE_BYTES);
      if (umma::elect_one()) {
#pragma unroll
        for (int k = 0; k < NB / 8; ++k)
          umma::mma_tf32_lohi(tmem + a * 256, umma::desc_lo(base + k * 1024, BOX), umma::desc_hi_sw128_base32(512),
                              umma::desc_lo(base + 4 * BOX + k * 1024, BOX), umma::desc_hi_sw128_base32(512), idesc, k > 0 ? 1u : 0u);
        umma::mma_commit(&bars->empty[s]);
        umma::mma_commit(&bars->acc_full[a]);
      }
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const int sub = (warp - 2) >> 2;               // ADAM: which of the quarter's two warps (takes the 32-column chunks of its parity)
    uint32_t c = 0;
    for (long long it = blockIdx.x; it < items; it += gridDim.x, ++c) {
      const int ni = (int)(it / g.k_tiles), ki = (int)(it % g.k_tiles);
      const uint32_t a = c & 1;
      mbar_wait_relaxed(&bars->acc_full[a], (c >> 1) & 1);
      umma::tc_fence_after_sync();
      const int n0 = ni * 128 + quarter * 32;
      const long long k0 = (long long)ki * g.KT;
      float* blk = stage_out + (warp - 2) * (32 * W_TPAD);
      for (int cb = 32 * sub; cb < g.KT; cb += 32 * (W_EPI / 4)) {
        uint32_t v[32];
        umma::tmem_ld_32x32(tmem + ((uint32_t)(quarter * 32) << 16) + a * 256 + cb, v);
        umma::tmem_ld_wait();
        // lane = row n0 + lane holds 32 consecutive k: transpose through shared memory so that each store
        // instruction below writes four whole 128-byte row segments instead of 32 scattered 16-byte pieces
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(blk + lane * W_TPAD + 4 * q) =
              make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                          __uint_as_float(v[4 * q + 3]));
        __syncwarp();
        const int cq = lane & 7;
        if (ADAM) {
          // fused Adam: all eight rows of the chunk at once -> 24 independent 16-byte loads per thread in flight (with
          // four rows per batch the kernel sat at 4.1 TB/s: ~49 KB in flight per SM do not cover the loaded-DRAM latency)
          {
            float4 pp[8], mm[8], vv[8];
            bool ok[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int r = (lane >> 3) + 4 * q;
              const int n = n0 + r;
              const long long k = k0 + cb + 4 * cq;
              ok[q] = n < g.N && k < g.K;
              if (ok[q]) {
                const size_t e = (size_t)n * g.K + k;
                pp[q] = *reinterpret_cast<const float4*>(g.adam.p + e);
                mm[q] = __ldcs(reinterpret_cast<const float4*>(g.adam.m + e));
                vv[q] = __ldcs(reinterpret_cast<const float4*>(g.adam.v + e));
              }
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (!ok[q]) continue;
              const int r = (lane >> 3) + 4 * q;
              const size_t e = (size_t)(n0 + r) * g.K + (k0 + cb + 4 * cq);
              const float4 gr = *reinterpret_cast<const float4*>(blk + r * W_TPAD + 4 * cq);
              pp[q].x = adam_fused_one(pp[q].x, gr.x, mm[q].x, vv[q].x, g.adam);
              pp[q].y = adam_fused_one(pp[q].y, gr.y, mm[q].y, vv[q].y, g.adam);
              pp[q].z = adam_fused_one(pp[q].z, gr.z, mm[q].z, vv[q].z, g.adam);
              pp[q].w = adam_fused_one(pp[q].w, gr.w, mm[q].w, vv[q].w, g.adam);
              *reinterpret_cast<float4*>(g.adam.p + e) = pp[q];
              __stcs(reinterpret_cast<float4*>(g.adam.m + e), mm[q]);
              __stcs(reinterpret_cast<float4*>(g.adam.v + e), vv[q]);
            }
          }
        }
#pragma unroll
        for (int itr = 0; itr < (ADAM ? 0 : 8); ++itr) {
          const int r = (lane >> 3) + 4 * itr;
          const int n = n0 + r;
          const long long k = k0 + cb + 4 * cq;
          if (n < g.N && k < g.K) {                 // K % 4 == 0
            float4 o = *reinterpret_cast<const float4*>(blk + r * W_TPAD + 4 * cq);
            float4* dst = reinterpret_cast<float4*>(dw + (size_t)n * g.K + k);
            if (g.accumulate) { const float4 old = *dst; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
            __stcs(dst, o);
          }
        }
      }
      umma::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&bars->acc_empty[a]);
    }
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

int ceil_div_ll(long long a, long long b) { return (int)((a + b - 1) / b); }

template <typename K>
int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  dd::prefer_max_smem(kernel);
  if (e != cudaSuccess) return dd::fail((int)e, "linear_tc: cudaFuncSetAttribute(%d): %s", bytes, cudaGetErrorString(e));
  return 0;
}

// split the contraction so that ~kSMs work items exist; rows = 128-row output tiles
void plan_stream(int tiles, int chunks, int& splits, int& cps) {
  splits = 1;
  if (tiles < dd::kSMs) {
    splits = dd::kSMs / tiles;
    const int maxs = (chunks + 7) / 8;            // at least 8 chunks (256 contraction elements) per split
    if (splits > maxs) splits = maxs;
    if (splits < 1) splits = 1;
  }
  cps = (chunks + splits - 1) / splits;
  splits = (chunks + cps - 1) / cps;
}

}  // namespace

namespace dd {

bool linear_tc_supported(int B, int N, long long K) {
  // TMA needs 16-byte row pitches; the streaming design pays off only for wide layers
  return B >= 1 && N % 4 == 0 && K % 4 == 0 && (long long)N * K >= (1ll << 22) && tma_encoder() != nullptr;
}

size_t linear_tc_workspace_bytes(int B, int N, long long K) {
  const int Bf = B < kLinearTcMaxRows ? B : kLinearTcMaxRows;     // forward: up to 256 rows per pass
  const int Bd = B < NB ? B : NB;                                 // input gradient: 32
  int s1, c1, s2, c2;
  plan_stream(ceil_div_ll(N, 128), ceil_div_ll(K, kChunk), s1, c1);      // fwd: tiles over N, contraction K
  plan_stream(ceil_div_ll(K, 128), ceil_div_ll(N, kChunk), s2, c2);      // dgrad: tiles over K, contraction N
  const size_t a = s1 > 1 ? (size_t)s1 * Bf * N * sizeof(float) : 0;
  const size_t b = s2 > 1 ? (size_t)s2 * Bd * K * sizeof(float) : 0;
  return (a > b ? a : b) + 256;
}

template <int NBT>
int launch_fwd(const CUtensorMap& mw, const CUtensorMap& mx, const float* bias, float* y, float* ws, const StreamGeo& g,
               cudaStream_t st) {
  auto k = linear_stream_tc_kernel<0, float, NBT>;
  if (int e = set_smem(k, SCfg<NBT>::SMEM)) return e;
  const int items = g.tiles * g.splits;
  k<<<items < kSMs ? items : kSMs, THREADS, SCfg<NBT>::SMEM, st>>>(mw, mx, bias, y, ws, g);
  return check_launch("linear_fwd_tc");
}

// x fp32 [B][K] (B <= 256: one pass over the weights), W fp32 [N][K], y fp32 [B][N]
int linear_fwd_tc(const float* x, const float* w, const float* bias, float* y, void* ws, size_t ws_bytes, int B, int N,
                  long long K, cudaStream_t st) {
  if (B > kLinearTcMaxRows) return fail(DD_ERR_UNSUPPORTED, "linear_fwd_tc: %d rows per pass (max %d)", B, kLinearTcMaxRows);
  StreamGeo g;
  g.B = B; g.LD = N;
  g.tiles = ceil_div_ll(N, 128);
  g.chunks = ceil_div_ll(K, kChunk);
  plan_stream(g.tiles, g.chunks, g.splits, g.chunks_per_split);
  if (g.splits > 1 && ws_bytes < (size_t)g.splits * B * N * sizeof(float))
    return fail(DD_ERR_WORKSPACE, "linear_fwd_tc: workspace %zu too small for %d splits", ws_bytes, g.splits);
  const int nbt = B <= 32 ? 32 : B <= 64 ? 64 : B <= 128 ? 128 : 256;
  CUtensorMap mw, mx;
  char msg[256];
  if (tma_map_2d_checked(&mw, w, 4, N, K, K, 32, 128, false, msg, sizeof msg) ||
      tma_map_2d_checked(&mx, x, 4, B, K, K, 32, nbt, false, msg, sizeof msg))
    return fail(DD_ERR_UNSUPPORTED, "linear_fwd_tc: %s", msg);
  int e = nbt == 32 ? launch_fwd<32>(mw, mx, bias, y, (float*)ws, g, st)
        : nbt == 64 ? launch_fwd<64>(mw, mx, bias, y, (float*)ws, g, st)
        : nbt == 128 ? launch_fwd<128>(mw, mx, bias, y, (float*)ws, g, st)
                     : launch_fwd<256>(mw, mx, bias, y, (float*)ws, g, st);
  if (e) return e;
  if (g.splits > 1) {
    const long long n = (long long)B * N;
    split_fold_kernel<float><<<ceil_div_ll(n, 256), 256, 0, st>>>((const float*)ws, bias, y, g.splits, n, N);
    return check_launch("linear_fwd_fold");
  }
  return 0;
}

// dy fp32 [B][N], W fp32 [N][K], dx [B][K] fp32 or bf16
int linear_dgrad_tc(const float* dy, const float* w, void* dx, int dx_dtype, void* ws, size_t ws_bytes, int B, int N,
                    long long K, cudaStream_t st) {
  StreamGeo g;
  g.B = B; g.LD = (int)K;
  g.tiles = ceil_div_ll(K, 128);
  g.chunks = ceil_div_ll(N, kChunk);
  plan_stream(g.tiles, g.chunks, g.splits, g.chunks_per_split);
  if (g.splits > 1 && ws_bytes < (size_t)g.splits * B * K * sizeof(float))
    return fail(DD_ERR_WORKSPACE, "linear_dgrad_tc: workspace %zu too small for %d splits", ws_bytes, g.splits);
  CUtensorMap mw, mdy;
  char msg[256];
  if (tma_map_2d_checked(&mw, w, 4, N, K, K, 32, 32, /*atom32=*/true, msg, sizeof msg) ||
      tma_map_2d_checked(&mdy, dy, 4, B, N, N, 32, 32, false, msg, sizeof msg))
    return fail(DD_ERR_UNSUPPORTED, "linear_dgrad_tc: %s", msg);
  const int items = g.tiles * g.splits;
  const int grid = items < kSMs ? items : kSMs;
  const long long n = (long long)B * K;
  if (dx_dtype == DD_F32) {
    auto k = linear_stream_tc_kernel<1, float>;
    if (int e = set_smem(k, SCfg<NB>::SMEM)) return e;
    k<<<grid, THREADS, SCfg<NB>::SMEM, st>>>(mw, mdy, nullptr, (float*)dx, (float*)ws, g);
    if (int e = check_launch("linear_dgrad_tc")) return e;
    if (g.splits > 1) split_fold_kernel<float><<<ceil_div_ll(n, 256), 256, 0, st>>>((const float*)ws, nullptr, (float*)dx, g.splits, n, (int)K);
  } else {
    auto k = linear_stream_tc_kernel<1, __nv_bfloat16>;
    if (int e = set_smem(k, SCfg<NB>::SMEM)) return e;
    k<<<grid, THREADS, SCfg<NB>::SMEM, st>>>(mw, mdy, nullptr, (__nv_bfloat16*)dx, (float*)ws, g);
    if (int e = check_launch("linear_dgrad_tc")) return e;
    if (g.splits > 1)
      split_fold_kernel<__nv_bfloat16><<<ceil_div_ll(n, 256), 256, 0, st>>>((const float*)ws, nullptr, (__nv_bfloat16*)dx, g.splits, n, (int)K);
  }
  return g.splits > 1 ? check_launch("linear_dgrad_fold") : 0;
}

// dy fp32 [B][N], x fp32 [B][K] -> dW fp32 [N][K] (overwritten, or accumulated into); with adam_pmv != nullptr the epilogue
// applies the Adam update to {weight, exp_avg, exp_avg_sq} = adam_pmv[0..2] instead of writing dW (hyper: beta1, beta2,
// eps, step_size = lr / (1 - beta1^t), inv_sqrt_bc2 = 1 / sqrt(1 - beta2^t), weight_decay)
int linear_wgrad_tc(const float* dy, const float* x, float* dw, int B, int N, long long K, int accumulate, cudaStream_t st,
                    float* const* adam_pmv = nullptr, const float* hyper = nullptr) {
  WgradGeo g = {};
  g.B = B; g.N = N; g.K = K;
  if (adam_pmv) {
    g.adam.p = adam_pmv[0]; g.adam.m = adam_pmv[1]; g.adam.v = adam_pmv[2];
    g.adam.beta1 = hyper[0]; g.adam.beta2 = hyper[1]; g.adam.eps = hyper[2]; g.adam.step_size = hyper[3];
    g.adam.inv_sqrt_bc2 = hyper[4]; g.adam.weight_decay = hyper[5];
  }
  g.KT = K >= 256 ? 256 : 128;
  g.n_tiles = ceil_div_ll(N, 128);
  g.k_tiles = ceil_div_ll(K, g.KT);
  g.accumulate = accumulate;
  CUtensorMap mdy, mx;
  char msg[256];
  if (tma_map_2d_checked(&mdy, dy, 4, B, N, N, 32, 32, true, msg, sizeof msg) ||
      tma_map_2d_checked(&mx, x, 4, B, K, K, 32, 32, true, msg, sizeof msg))
    return fail(DD_ERR_UNSUPPORTED, "linear_wgrad_tc: %s", msg);
  const long long items = (long long)g.n_tiles * g.k_tiles;
  const int grid = (int)(items < kSMs ? items : kSMs);
  if (adam_pmv) {
    if (int e = set_smem(linear_wgrad_tc_kernel<true>, W_SMEM)) return e;
    linear_wgrad_tc_kernel<true><<<grid, WCfg<true>::THREADS, W_SMEM, st>>>(mdy, mx, dw, g);
  } else {
    if (int e = set_smem(linear_wgrad_tc_kernel<false>, W_SMEM)) return e;
    linear_wgrad_tc_kernel<false><<<grid, WCfg<false>::THREADS, W_SMEM, st>>>(mdy, mx, dw, g);
  }
  return check_launch("linear_wgrad_tc");
}

}  // namespace dd
