// Hardware probe for the shared-memory matrix-descriptor semantics conv_tc.cu relies on
// (SWIZZLE_NONE core-matrix layouts, start-address shifts, MN-major operands).  The host builds
// byte images of the operands, the kernel copies them to smem verbatim, issues tcgen05.mma with the
// host-given descriptor fields and returns the accumulator.  Prints max |D - ref| per experiment.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../driving-dirty_b200/csrc/umma.cuh"

struct Params {
  uint32_t a_bytes, b_bytes;      // image sizes
  uint32_t a_start, b_start;      // descriptor start offsets inside the images
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_kstep, b_kstep;      // start-address advance per MMA
  uint32_t nk, idesc, N;
};

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* a_img, const uint8_t* b_img, float* d, Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((p.a_bytes + 1023) / 1024) * 1024;
  for (uint32_t i = threadIdx.x; i < p.a_bytes; i += 128) sa[i] = a_img[i];
  for (uint32_t i = threadIdx.x; i < p.b_bytes; i += 128) sb[i] = b_img[i];
  if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
  if (threadIdx.x < 32) umma::tmem_alloc(&tmem_base, 128);
  umma::fence_proxy_async_smem();
  umma::tc_fence_before_sync();
  __syncthreads();
  umma::tc_fence_after_sync();
  const uint32_t tb = tmem_base;
  if (threadIdx.x == 0) {
    for (uint32_t k = 0; k < p.nk; ++k) {
      const uint64_t da = umma::make_desc(umma::smem_u32(sa) + p.a_start + k * p.a_kstep, p.a_lbo, p.a_sbo);
      const uint64_t db = umma::make_desc(umma::smem_u32(sb) + p.b_start + k * p.b_kstep, p.b_lbo, p.b_sbo);
      umma::mma_bf16(tb, da, db, p.idesc, k > 0);
    }
    umma::mma_commit(&bar);
  }
  umma::mbar_wait(&bar, 0);
  umma::tc_fence_after_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t c = 0; c < p.N; c += 32) {
    uint32_t r[32];
    umma::tmem_ld_32x32(tb + ((uint32_t)(warp * 32) << 16) + c, r);
    umma::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) d[(size_t)(warp * 32 + lane) * p.N + c + j] = __uint_as_float(r[j]);
  }
  umma::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) umma::tmem_dealloc(tb, 128);
}

static uint16_t bf16(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }   // exact for small ints
static float frand(int i, int j, int salt) { return (float)(((i * 7 + j * 13 + salt * 5) % 9) - 4); }

struct Exp {
  const char* name;
  std::vector<uint8_t> a, b;
  Params p;
  std::vector<float> ref;   // [128][N]
};

static void put(std::vector<uint8_t>& img, size_t off, float v) {
  if (off + 2 > img.size()) img.resize(off + 2, 0);
  uint16_t h = bf16(v);
  memcpy(&img[off], &h, 2);
}

// K-major, no swizzle: element (r, k) at (r/8)*sbo + (k/8)*lbo + (r%8)*16 + (k%8)*2
static Exp make_kmajor(const char* name, int N, int K, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo,
                       int row_shift, bool swap_fields) {
  Exp e; e.name = name;
  const int M = 128, rows = M + row_shift;
  for (int r = 0; r < rows; ++r)
    for (int k = 0; k < K; ++k) put(e.a, (size_t)(r / 8) * a_sbo + (size_t)(k / 8) * a_lbo + (r % 8) * 16 + (k % 8) * 2, frand(r, k, 1));
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) put(e.b, (size_t)(n / 8) * b_sbo + (size_t)(k / 8) * b_lbo + (n % 8) * 16 + (k % 8) * 2, frand(n, k, 2));
  e.a.resize(e.a.size() + 64, 0); e.b.resize(e.b.size() + 64, 0);
  Params& p = e.p; memset(&p, 0, sizeof(p));
  p.a_bytes = (uint32_t)e.a.size(); p.b_bytes = (uint32_t)e.b.size();
  p.a_start = row_shift * 16; p.b_start = 0;
  p.a_lbo = swap_fields ? a_sbo : a_lbo; p.a_sbo = swap_fields ? a_lbo : a_sbo;
  p.b_lbo = swap_fields ? b_sbo : b_lbo; p.b_sbo = swap_fields ? b_lbo : b_sbo;
  p.a_kstep = 2 * a_lbo; p.b_kstep = 2 * b_lbo; p.nk = K / 16; p.N = N;
  p.idesc = umma::make_idesc_bf16(128, N, false, false);
  e.ref.assign((size_t)M * N, 0.f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += frand(m + row_shift, k, 1) * frand(n, k, 2);
      e.ref[(size_t)m * N + n] = s;
    }
  return e;
}

// MN-major, no swizzle: element (m, k) at (m/8)*sbo + (k/8)*lbo + (k%8)*16 + (m%8)*2   (K = 16 per MMA)
static Exp make_mnmajor(const char* name, int N, int K, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo,
                        int k_shift, bool swap_fields) {
  Exp e; e.name = name;
  const int M = 128, krows = K + k_shift;
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < krows; ++k) put(e.a, (size_t)(m / 8) * a_sbo + (size_t)(k / 8) * a_lbo + (k % 8) * 16 + (m % 8) * 2, frand(m, k, 3));
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) put(e.b, (size_t)(n / 8) * b_sbo + (size_t)(k / 8) * b_lbo + (k % 8) * 16 + (n % 8) * 2, frand(n, k, 4));
  e.a.resize(e.a.size() + 64, 0); e.b.resize(e.b.size() + 64, 0);
  Params& p = e.p; memset(&p, 0, sizeof(p));
  p.a_bytes = (uint32_t)e.a.size(); p.b_bytes = (uint32_t)e.b.size();
  p.a_start = k_shift * 16; p.b_start = 0;     // shifting the K (pixel) axis by one row of 16 B
  p.a_lbo = swap_fields ? a_sbo : a_lbo; p.a_sbo = swap_fields ? a_lbo : a_sbo;
  p.b_lbo = swap_fields ? b_sbo : b_lbo; p.b_sbo = swap_fields ? b_lbo : b_sbo;
  p.a_kstep = 2 * a_lbo; p.b_kstep = 2 * b_lbo; p.nk = K / 16; p.N = N;
  p.idesc = umma::make_idesc_bf16(128, N, true, true);
  e.ref.assign((size_t)M * N, 0.f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < K; ++k) s += frand(m, k + k_shift, 3) * frand(n, k, 4);
      e.ref[(size_t)m * N + n] = s;
    }
  return e;
}

int main() {
  std::vector<Exp> exps;
  // conv fwd operand geometry: A = [pixels x 16 ch] in 16-byte-pixel planes (plane stride 2176), B = weights
  exps.push_back(make_kmajor("K-major  N=32 K=16  lbo=plane sbo=128", 32, 16, 2176, 128, 512, 128, 0, false));
  exps.push_back(make_kmajor("K-major  start shifted by 1 row (16 B) ", 32, 16, 2176, 128, 512, 128, 1, false));
  exps.push_back(make_kmajor("K-major  start shifted by 2 rows       ", 32, 16, 2176, 128, 512, 128, 2, false));
  exps.push_back(make_kmajor("K-major  K=32 (2 MMAs, accumulate)     ", 32, 32, 2176, 128, 512, 128, 1, false));
  exps.push_back(make_kmajor("K-major  N=96 K=32                     ", 96, 32, 2176, 128, 1536, 128, 1, false));
  // (swapping the LBO/SBO fields of experiment 1 faults with an illegal address: orientation confirmed)
  // wgrad operand geometry: A = x planes [chunk of 8 ch][pixel] (MN-major), B = dy planes, K = pixels
  exps.push_back(make_mnmajor("MN-major N=64 K=16 sbo=plane lbo=128   ", 64, 16, 128, 2176, 128, 2176, 0, false));
  exps.push_back(make_mnmajor("MN-major K shifted by 1 pixel row      ", 64, 16, 128, 2176, 128, 2176, 1, false));
  exps.push_back(make_mnmajor("MN-major K=64 (4 MMAs)                 ", 64, 64, 128, 2176, 128, 2176, 1, false));
  float* d_d; cudaMalloc(&d_d, 128 * 256 * sizeof(float));
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (auto& e : exps) {
    uint8_t *da, *db;
    cudaMalloc(&da, e.a.size()); cudaMalloc(&db, e.b.size());
    cudaMemcpy(da, e.a.data(), e.a.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, e.b.data(), e.b.size(), cudaMemcpyHostToDevice);
    cudaMemset(d_d, 0xff, 128 * 256 * sizeof(float));
    size_t smem = ((e.a.size() + 1023) / 1024) * 1024 + e.b.size() + 1024;
    probe_kernel<<<1, 128, smem>>>(da, db, d_d, e.p);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("%-44s CUDA error: %s\n", e.name, cudaGetErrorString(err)); return 1; }
    std::vector<float> out((size_t)128 * e.p.N);
    cudaMemcpy(out.data(), d_d, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (size_t i = 0; i < out.size(); ++i) { double er = fabs((double)out[i] - e.ref[i]); if (!(er <= 1e-3)) ++bad; if (er > maxerr || er != er) maxerr = er; }
    printf("%-44s smem %6zu B  max|err| %-10.3g mismatches %d / %zu  %s\n", e.name, smem, maxerr, bad, out.size(), bad ? "FAIL" : "ok");
    cudaFree(da); cudaFree(db);
  }
  return 0;
}
