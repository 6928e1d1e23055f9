// K5: flatten + max_pool1d(kernel_size=4) over the NCHW-flat index (components.py:46-47), on
// NHWC activations; its backward with the ReLU mask of the pooled activation folded in; and the
// NHWC <-> NCHW layout changes used at the API edge.  HBM-bound byte movers.
//
// flat index f = c*H*W + pix (what x.view(B,-1) of an NCHW tensor enumerates); window j covers
// f = 4j..4j+3.  When H*W % 4 == 0 a window never straddles channels and the fast kernels work
// on [128 pixel][32 channel] tiles staged in shared memory so both global sides are coalesced.
#include "dd_common.cuh"

namespace {
constexpr int C = 32, TP = 128, PAD = 33;

template <typename T>
__device__ __forceinline__ void load_tile(const T* __restrict__ img, long long p0, int npix, float* s, int tid) {
  for (int i = tid; i < TP * 4; i += 256) {
    const int cg = i & 3, p = i >> 2;
    float v[8];
    if (p < npix) dd::ld8<T>(img + (p0 + p) * C + cg * 8, v);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s[p * PAD + cg * 8 + k] = v[k];
  }
}

template <typename T, typename TO>
__global__ void __launch_bounds__(256) pool4_fwd_tiled(const T* __restrict__ a3, TO* __restrict__ pooled, int HW) {
  __shared__ float s[TP * PAD];
  __shared__ float so[C * PAD];
  const int tid = threadIdx.x, b = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * TP;
  const int npix = (int)min((long long)TP, HW - p0);
  load_tile<T>(a3 + (size_t)b * HW * C, p0, npix, s, tid);
  __syncthreads();
  const int c = tid & 31;
  for (int ql = tid >> 5; ql < TP / 4; ql += 8) {
    const float* q = s + (4 * ql) * PAD + c;
    so[c * PAD + ql] = fmaxf(fmaxf(q[0], q[PAD]), fmaxf(q[2 * PAD], q[3 * PAD]));
  }
  __syncthreads();
  const int Q = HW / 4, nq = npix / 4;
  TO* out = pooled + (size_t)b * C * Q + p0 / 4;
  for (int i = tid; i < C * (TP / 4); i += 256) {
    const int ql = i & 31, cc = i >> 5;
    if (ql < nq) dd::st<TO>(out + (size_t)cc * Q + ql, so[cc * PAD + ql]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) pool4_bwd_tiled(const T* __restrict__ a3, const T* __restrict__ dpooled,
                                                       T* __restrict__ da3, int HW) {
  __shared__ float s[TP * PAD];
  __shared__ float sg[C * PAD];
  const int tid = threadIdx.x, b = blockIdx.y;
  const long long p0 = (long long)blockIdx.x * TP;
  const int npix = (int)min((long long)TP, HW - p0);
  const int Q = HW / 4, nq = npix / 4;
  load_tile<T>(a3 + (size_t)b * HW * C, p0, npix, s, tid);
  const T* gin = dpooled + (size_t)b * C * Q + p0 / 4;
  for (int i = tid; i < C * (TP / 4); i += 256) {
    const int ql = i & 31, cc = i >> 5;
    sg[cc * PAD + ql] = ql < nq ? dd::ld<T>(gin + (size_t)cc * Q + ql) : 0.f;
  }
  __syncthreads();
  const int c = tid & 31;
  for (int ql = tid >> 5; ql < TP / 4; ql += 8) {
    float* q = s + (4 * ql) * PAD + c;
    const float v0 = q[0], v1 = q[PAD], v2 = q[2 * PAD], v3 = q[3 * PAD];
    int m = 0;
    float mx = v0;
    if (v1 > mx) { mx = v1; m = 1; }
    if (v2 > mx) { mx = v2; m = 2; }
    if (v3 > mx) { mx = v3; m = 3; }
    const float g = mx > 0.f ? sg[c * PAD + ql] : 0.f;   // ReLU'(a3) at the selected element
    q[0] = m == 0 ? g : 0.f;
    q[PAD] = m == 1 ? g : 0.f;
    q[2 * PAD] = m == 2 ? g : 0.f;
    q[3 * PAD] = m == 3 ? g : 0.f;
  }
  __syncthreads();
  T* out = da3 + ((size_t)b * HW + p0) * C;
  for (int i = tid; i < TP * 4; i += 256) {
    const int cg = i & 3, p = i >> 2;
    if (p < npix) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = s[p * PAD + cg * 8 + k];
      dd::st8<T>(out + (size_t)p * C + cg * 8, v);
    }
  }
}

// any H*W: one thread per window
template <typename T, bool BWD, typename TO = T>
__global__ void pool4_generic(const T* __restrict__ a3, const T* __restrict__ dpooled, TO* __restrict__ out, int HW,
                              long long windows_per_img, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / windows_per_img, j = idx - b * windows_per_img;
    const T* img = a3 + (size_t)b * HW * C;
    size_t off[4];
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long f = 4 * j + e;
      const int c = (int)(f / HW), pix = (int)(f - (long long)c * HW);
      off[e] = (size_t)pix * C + c;
      v[e] = dd::ld<T>(img + off[e]);
    }
    int m = 0;
    float mx = v[0];
#pragma unroll
    for (int e = 1; e < 4; ++e)
      if (v[e] > mx) { mx = v[e]; m = e; }
    if (!BWD) {
      dd::st<TO>(out + idx, mx);
    } else {
      const float g = mx > 0.f ? dd::ld<T>(dpooled + idx) : 0.f;
      TO* o = out + (size_t)b * HW * C;
#pragma unroll
      for (int e = 0; e < 4; ++e) dd::st<TO>(o + off[e], e == m ? g : 0.f);
    }
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int Cn, int HW, long long total) {
  // out index (b, c, pix) -> in (b, pix, c); coalesced writes
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int pix = (int)(idx % HW);
    const long long bc = idx / HW;
    const int c = (int)(bc % Cn);
    const long long b = bc / Cn;
    out[idx] = dd::ld<T>(in + ((size_t)b * HW + pix) * Cn + c);
  }
}
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, T* __restrict__ out, int Cn, int HW, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % Cn);
    const long long bp = idx / Cn;
    const int pix = (int)(bp % HW);
    const long long b = bp / HW;
    dd::st<T>(out + idx, in[((size_t)b * Cn + c) * HW + pix]);
  }
}

int grid1d(long long total) {
  long long g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g < dd::kSMs * 16 ? g : dd::kSMs * 16));
}

template <typename T>
int pool_dispatch(const T* a3, const T* dpooled, T* out, int B, int H, int W, bool bwd, cudaStream_t st) {
  const int HW = H * W;
  if (HW % 4 == 0) {
    dim3 grid((HW + TP - 1) / TP, B);
    if (!bwd) pool4_fwd_tiled<T, T><<<grid, 256, 0, st>>>(a3, out, HW);
    else pool4_bwd_tiled<T><<<grid, 256, 0, st>>>(a3, dpooled, out, HW);
  } else {
    const long long wpi = (long long)C * HW / 4, total = wpi * B;
    if (!bwd) pool4_generic<T, false><<<grid1d(total), 256, 0, st>>>(a3, nullptr, out, HW, wpi, total);
    else pool4_generic<T, true><<<grid1d(total), 256, 0, st>>>(a3, dpooled, out, HW, wpi, total);
  }
  return dd::check_launch(bwd ? "pool4_bwd" : "pool4_fwd");
}
// forward with fp32 features out of a bf16 activation: the max of bf16 values is a bf16 value, so the fp32 features hold
// exactly what the bf16 ones would, and the tf32 linear that follows reads them without a conversion pass
int pool_fwd_bf16_to_f32(const __nv_bfloat16* a3, float* out, int B, int H, int W, cudaStream_t st) {
  const int HW = H * W;
  if (HW % 4 == 0) {
    pool4_fwd_tiled<__nv_bfloat16, float><<<dim3((HW + TP - 1) / TP, B), 256, 0, st>>>(a3, out, HW);
  } else {
    const long long wpi = (long long)C * HW / 4, total = wpi * B;
    pool4_generic<__nv_bfloat16, false, float><<<grid1d(total), 256, 0, st>>>(a3, nullptr, out, HW, wpi, total);
  }
  return dd::check_launch("pool4_fwd");
}
}  // namespace

extern "C" int dd_pool4_fwd_f32(const void* a3, float* pooled, int dtype, int B, int H, int W, void* stream) {
  DD_REQUIRE(a3 && pooled, DD_ERR_BAD_ARG, "dd_pool4_fwd_f32: null pointer");
  DD_REQUIRE(B >= 0 && H > 0 && W > 0, DD_ERR_BAD_ARG, "dd_pool4_fwd_f32: bad shape");
  if (B == 0) return 0;
  cudaStream_t st = dd::as_stream(stream);
  if (dtype == DD_F32) return pool_dispatch<float>((const float*)a3, nullptr, pooled, B, H, W, false, st);
  if (dtype == DD_BF16) return pool_fwd_bf16_to_f32((const __nv_bfloat16*)a3, pooled, B, H, W, st);
  return dd::fail(DD_ERR_UNSUPPORTED, "dd_pool4_fwd_f32: dtype %d", dtype);
}

extern "C" int dd_pool4_fwd(const void* a3, void* pooled, int dtype, int B, int H, int W, void* stream) {
  DD_REQUIRE(a3 && pooled, DD_ERR_BAD_ARG, "dd_pool4_fwd: null pointer");
  DD_REQUIRE(B >= 0 && H > 0 && W > 0, DD_ERR_BAD_ARG, "dd_pool4_fwd: bad shape");
  if (B == 0) return 0;
  cudaStream_t st = dd::as_stream(stream);
  if (dtype == DD_F32) return pool_dispatch<float>((const float*)a3, nullptr, (float*)pooled, B, H, W, false, st);
  if (dtype == DD_BF16)
    return pool_dispatch<__nv_bfloat16>((const __nv_bfloat16*)a3, nullptr, (__nv_bfloat16*)pooled, B, H, W, false, st);
  return dd::fail(DD_ERR_UNSUPPORTED, "dd_pool4_fwd: dtype %d", dtype);
}

extern "C" int dd_pool4_bwd(const void* a3, const void* dpooled, void* da3, int dtype, int B, int H, int W,
                            void* stream) {
  DD_REQUIRE(a3 && dpooled && da3, DD_ERR_BAD_ARG, "dd_pool4_bwd: null pointer");
  DD_REQUIRE(B >= 0 && H > 0 && W > 0, DD_ERR_BAD_ARG, "dd_pool4_bwd: bad shape");
  if (B == 0) return 0;
  cudaStream_t st = dd::as_stream(stream);
  if (dtype == DD_F32) return pool_dispatch<float>((const float*)a3, (const float*)dpooled, (float*)da3, B, H, W, true, st);
  if (dtype == DD_BF16)
    return pool_dispatch<__nv_bfloat16>((const __nv_bfloat16*)a3, (const __nv_bfloat16*)dpooled, (__nv_bfloat16*)da3, B, H, W, true, st);
  return dd::fail(DD_ERR_UNSUPPORTED, "dd_pool4_bwd: dtype %d", dtype);
}

extern "C" int dd_nhwc_to_nchw_f32(const void* in, int in_dtype, float* out, int B, int Cn, int H, int W,
                                   void* stream) {
  DD_REQUIRE(in && out, DD_ERR_BAD_ARG, "dd_nhwc_to_nchw_f32: null pointer");
  DD_REQUIRE(B >= 0 && Cn > 0 && H > 0 && W > 0, DD_ERR_BAD_ARG, "dd_nhwc_to_nchw_f32: bad shape");
  if (B == 0) return 0;
  const long long total = (long long)B * Cn * H * W;
  cudaStream_t st = dd::as_stream(stream);
  if (in_dtype == DD_F32) nhwc_to_nchw_kernel<float><<<grid1d(total), 256, 0, st>>>((const float*)in, out, Cn, H * W, total);
  else if (in_dtype == DD_BF16)
    nhwc_to_nchw_kernel<__nv_bfloat16><<<grid1d(total), 256, 0, st>>>((const __nv_bfloat16*)in, out, Cn, H * W, total);
  else return dd::fail(DD_ERR_UNSUPPORTED, "dd_nhwc_to_nchw_f32: dtype %d", in_dtype);
  return dd::check_launch("nhwc_to_nchw");
}

extern "C" int dd_nchw_f32_to_nhwc(const float* in, void* out, int out_dtype, int B, int Cn, int H, int W,
                                   void* stream) {
  DD_REQUIRE(in && out, DD_ERR_BAD_ARG, "dd_nchw_f32_to_nhwc: null pointer");
  DD_REQUIRE(B >= 0 && Cn > 0 && H > 0 && W > 0, DD_ERR_BAD_ARG, "dd_nchw_f32_to_nhwc: bad shape");
  if (B == 0) return 0;
  const long long total = (long long)B * Cn * H * W;
  cudaStream_t st = dd::as_stream(stream);
  if (out_dtype == DD_F32) nchw_to_nhwc_kernel<float><<<grid1d(total), 256, 0, st>>>(in, (float*)out, Cn, H * W, total);
  else if (out_dtype == DD_BF16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid1d(total), 256, 0, st>>>(in, (__nv_bfloat16*)out, Cn, H * W, total);
  else return dd::fail(DD_ERR_UNSUPPORTED, "dd_nchw_f32_to_nhwc: dtype %d", out_dtype);
  return dd::check_launch("nchw_to_nhwc");
}
