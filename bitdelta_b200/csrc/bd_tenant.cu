// Per-tenant dense leaves of the multi-tenant decode step (reference demo/demo_backend.py:62-79, DataParallelModule):
// row t of the batch goes through tenant t's OWN full-precision weight.  The reference loops over the tenants on the
// host (swap weight.data, call the module on x[t,None]) and pads ragged outputs with finfo.min through a nested tensor;
// here each leaf kind is one launch over all tenants:
//
//   tenant_linear   lm_head:      y[t,i,v] = x[t,i,:] . W_t[v,:]   (v < V_t; columns V_t..ldy-1 get the lowest finite value)
//   tenant_rmsnorm  RMSNorm:      y[t,i,:] = W_t * round16(x[t,i,:] * rsqrt(mean(x^2) + eps))   (HF LlamaRMSNorm arithmetic)
//   tenant_embed    embed_tokens: y[t,i,:] = E_t[ids[t,i], :]
//
// tenant_linear is a pure HBM stream (decode: m <= 4 rows per tenant, T * V_t * K * 2 bytes of weights read once), so it
// is CUDA-core code organised for bytes in flight: a persistent grid of 2 CTAs per SM, 16 warps per CTA, every lane keeps
// 8 independent 16-byte weight loads outstanding (2 weight rows x 4 K-chunks), activations come from shared memory.
#include "bd_common.cuh"

namespace bd {
namespace {

constexpr int kMaxTenants = 32;   // per launch; the host loops over groups of 32
constexpr int kLinThreads = 512;
constexpr int kLinWarps = kLinThreads / 32;
constexpr int kRowsPerWarp = 2;
constexpr int kRowsPerUnit = kLinWarps * kRowsPerWarp;  // 32 weight rows per work unit
constexpr int kUnroll = 4;                              // K chunks (of 32 lanes x 8 elements) in flight per row

struct TenantTable {
  const void* w[kMaxTenants];
  int n_out[kMaxTenants];
  int unit0[kMaxTenants + 1];  // prefix sum of ceil(n_out / kRowsPerUnit)
};

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

template <typename T16>
__device__ __forceinline__ T16 lowest_value();
template <>
__device__ __forceinline__ __nv_bfloat16 lowest_value<__nv_bfloat16>() { return __ushort_as_bfloat16((unsigned short)0xFF7Fu); }
template <>
__device__ __forceinline__ __half lowest_value<__half>() { return __ushort_as_half((unsigned short)0xFBFFu); }

template <typename T16, int M>
__global__ void __launch_bounds__(kLinThreads, 2)
tenant_linear_kernel(const T16* __restrict__ x, const __grid_constant__ TenantTable tab, const T16* __restrict__ bias,
                     T16* __restrict__ y, int T, int K, int64_t ldy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T16* xs = reinterpret_cast<T16*>(smem_raw);  // [M][K] of the current tenant
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_units = tab.unit0[T];
  int cur_t = -1, t = 0;
  for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    while (unit >= tab.unit0[t + 1]) ++t;  // units are tenant-major and this CTA's sequence is increasing
    if (t != cur_t) {
      __syncthreads();  // everyone is done with the previous tenant's activations
      const uint4* src = reinterpret_cast<const uint4*>(x + (int64_t)t * M * K);
      for (int i = threadIdx.x; i < M * K / 8; i += kLinThreads) reinterpret_cast<uint4*>(xs)[i] = src[i];
      __syncthreads();
      cur_t = t;
    }
    const int n_out = tab.n_out[t];
    const int row0 = (unit - tab.unit0[t]) * kRowsPerUnit + warp * kRowsPerWarp;
    const T16* wt = reinterpret_cast<const T16*>(tab.w[t]);
    float acc[kRowsPerWarp][M];
#pragma unroll
    for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
      for (int i = 0; i < M; ++i) acc[r][i] = 0.f;

    if (row0 < n_out) {
      // rows past the end re-read the last valid row (their result is dropped): no divergence inside the K loop
      const T16* wrow[kRowsPerWarp];
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r) wrow[r] = wt + (int64_t)min(row0 + r, n_out - 1) * K;
      for (int k0 = lane * 8; k0 < K; k0 += 256 * kUnroll) {
        uint4 wv[kRowsPerWarp][kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int k = k0 + u * 256;
#pragma unroll
          for (int r = 0; r < kRowsPerWarp; ++r) wv[r][u] = k < K ? ldg_stream(wrow[r] + k) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
          const int k = k0 + u * 256;
          if (k < K) {
#pragma unroll
            for (int i = 0; i < M; ++i) {
              const uint4 xv = *reinterpret_cast<const uint4*>(xs + i * K + k);
              const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
              for (int r = 0; r < kRowsPerWarp; ++r) {
                const uint32_t ww[4] = {wv[r][u].x, wv[r][u].y, wv[r][u].z, wv[r][u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  acc[r][i] = fmaf(F16<T16>::lo(ww[j]), F16<T16>::lo(xw[j]), acc[r][i]);
                  acc[r][i] = fmaf(F16<T16>::hi(ww[j]), F16<T16>::hi(xw[j]), acc[r][i]);
                }
              }
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < kRowsPerWarp; ++r)
#pragma unroll
        for (int i = 0; i < M; ++i) {
          float v = acc[r][i];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          acc[r][i] = v;
        }
      if (lane < kRowsPerWarp * M) {
        const int r = lane / M, i = lane % M;
        float v = 0.f;
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; ++rr)
#pragma unroll
          for (int ii = 0; ii < M; ++ii)
            if (rr == r && ii == i) v = acc[rr][ii];
        const int row = row0 + r;
        if (row < n_out) {
          if (bias) v += F16<T16>::to_f32(bias[row]);
          y[((int64_t)t * M + i) * ldy + row] = F16<T16>::from_f32(v);
        }
      }
    }
    // the CTA that owns a tenant's last row block also writes that tenant's padding columns (demo_backend.py:78-79)
    if (unit == tab.unit0[t + 1] - 1 && n_out < ldy) {
      const int pad = (int)(ldy - n_out);
      for (int idx = threadIdx.x; idx < M * pad; idx += kLinThreads)
        y[((int64_t)t * M + idx / pad) * ldy + n_out + idx % pad] = lowest_value<T16>();
    }
  }
}

// ---- RMSNorm with a per-tenant weight (HF LlamaRMSNorm.forward: fp32 statistics, two roundings to the model dtype) ----
template <typename T16>
__global__ void __launch_bounds__(256)
tenant_rmsnorm_kernel(const T16* __restrict__ x, const __grid_constant__ TenantTable tab, T16* __restrict__ y, int m, int H,
                      float eps) {
  __shared__ float red[8];
  __shared__ float s_scale;
  const int row = blockIdx.x, t = row / m;
  const T16* xr = x + (int64_t)row * H;
  const T16* w = reinterpret_cast<const T16*>(tab.w[t]);
  float ss = 0.f;
  for (int h = threadIdx.x; h < H; h += 256) {
    const float v = F16<T16>::to_f32(xr[h]);
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    s_scale = rsqrtf(tot / (float)H + eps);
  }
  __syncthreads();
  const float scale = s_scale;
  for (int h = threadIdx.x; h < H; h += 256) {
    const T16 n = F16<T16>::from_f32(F16<T16>::to_f32(xr[h]) * scale);
    y[(int64_t)row * H + h] = F16<T16>::from_f32(F16<T16>::to_f32(w[h]) * F16<T16>::to_f32(n));
  }
}

// ---- embedding gather with a per-tenant table ----
template <typename T16>
__global__ void __launch_bounds__(128)
tenant_embed_kernel(const int64_t* __restrict__ ids, const __grid_constant__ TenantTable tab, T16* __restrict__ y, int m, int H) {
  const int row = blockIdx.x, t = row / m;
  int64_t id = ids[row];
  const int64_t n = tab.n_out[t];
  id = id < 0 ? 0 : (id >= n ? n - 1 : id);  // memory safety only; ids must be in range like for nn.Embedding
  const T16* src = reinterpret_cast<const T16*>(tab.w[t]) + id * H;
  T16* dst = y + (int64_t)row * H;
  if (H % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    for (int h = threadIdx.x; h < H / 8; h += 128) reinterpret_cast<uint4*>(dst)[h] = reinterpret_cast<const uint4*>(src)[h];
  } else {
    for (int h = threadIdx.x; h < H; h += 128) dst[h] = src[h];
  }
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T16, int M>
int launch_linear(const T16* x, const TenantTable& tab, const T16* bias, T16* y, int T, int K, int64_t ldy, cudaStream_t s) {
  const size_t smem = (size_t)M * K * sizeof(T16);
  auto kern = tenant_linear_kernel<T16, M>;
  if (smem > 48 * 1024) BD_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int units = tab.unit0[T];
  const int grid = units < 2 * sm_count() ? units : 2 * sm_count();
  kern<<<grid, kLinThreads, smem, s>>>(x, tab, bias, y, T, K, ldy);
  count_launch();
  return check_launch("tenant_linear_kernel");
}

template <typename T16>
int launch_linear_m(int m, const T16* x, const TenantTable& tab, const T16* bias, T16* y, int T, int K, int64_t ldy,
                    cudaStream_t s) {
  switch (m) {
    case 1: return launch_linear<T16, 1>(x, tab, bias, y, T, K, ldy, s);
    case 2: return launch_linear<T16, 2>(x, tab, bias, y, T, K, ldy, s);
    case 3: return launch_linear<T16, 3>(x, tab, bias, y, T, K, ldy, s);
    default: return launch_linear<T16, 4>(x, tab, bias, y, T, K, ldy, s);
  }
}

}  // namespace
}  // namespace bd

using namespace bd;

extern "C" BD_API int bd_tenant_linear(const void* x, const void* const* w, const int64_t* n_out, const void* bias, void* y,
                                       int dtype, int64_t T, int64_t m, int64_t K, int64_t ldy, void* stream) {
  BD_REQUIRE(dtype == BD_BF16 || dtype == BD_FP16, "tenant_linear: dtype must be BD_BF16 or BD_FP16");
  BD_REQUIRE(x && w && n_out && y, "tenant_linear: null pointer");
  BD_REQUIRE(T > 0 && m >= 1 && m <= BD_TENANT_LINEAR_MAX_ROWS, "tenant_linear: m must be 1..%d rows per tenant (got %lld)",
             BD_TENANT_LINEAR_MAX_ROWS, (long long)m);
  BD_REQUIRE(K > 0 && K % 8 == 0, "tenant_linear: K must be a positive multiple of 8");
  BD_REQUIRE((size_t)m * K * 2 <= 200 * 1024, "tenant_linear: m*K too large for the shared-memory activation tile");
  BD_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "tenant_linear: x must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  for (int64_t t0 = 0; t0 < T; t0 += kMaxTenants) {
    const int nt = (int)((T - t0 < kMaxTenants) ? T - t0 : kMaxTenants);
    TenantTable tab;
    tab.unit0[0] = 0;
    for (int i = 0; i < nt; ++i) {
      const int64_t n = n_out[t0 + i];
      BD_REQUIRE(w[t0 + i] && (reinterpret_cast<uintptr_t>(w[t0 + i]) & 15) == 0, "tenant_linear: weight %lld null or not 16-byte aligned",
                 (long long)(t0 + i));
      BD_REQUIRE(n > 0 && n <= ldy && n < (1ll << 31), "tenant_linear: n_out[%lld] = %lld out of range (ldy %lld)", (long long)(t0 + i),
                 (long long)n, (long long)ldy);
      BD_REQUIRE(!bias || n == ldy, "tenant_linear: a shared bias needs equal output widths");
      tab.w[i] = w[t0 + i];
      tab.n_out[i] = (int)n;
      tab.unit0[i + 1] = tab.unit0[i] + (int)((n + kRowsPerUnit - 1) / kRowsPerUnit);
    }
    const size_t xoff = (size_t)t0 * m * K, yoff = (size_t)t0 * m * ldy;
    int rc;
    if (dtype == BD_BF16)
      rc = launch_linear_m<__nv_bfloat16>((int)m, (const __nv_bfloat16*)x + xoff, tab, (const __nv_bfloat16*)bias,
                                          (__nv_bfloat16*)y + yoff, nt, (int)K, ldy, s);
    else
      rc = launch_linear_m<__half>((int)m, (const __half*)x + xoff, tab, (const __half*)bias, (__half*)y + yoff, nt, (int)K, ldy, s);
    if (rc) return rc;
  }
  return BD_OK;
}

extern "C" BD_API int bd_tenant_rmsnorm(const void* x, const void* const* w, void* y, int dtype, int64_t T, int64_t m, int64_t H,
                                        float eps, void* stream) {
  BD_REQUIRE(dtype == BD_BF16 || dtype == BD_FP16, "tenant_rmsnorm: dtype must be BD_BF16 or BD_FP16");
  BD_REQUIRE(x && w && y, "tenant_rmsnorm: null pointer");
  BD_REQUIRE(T > 0 && m > 0 && H > 0 && T * m < (1ll << 31) && H < (1ll << 31), "tenant_rmsnorm: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  for (int64_t t0 = 0; t0 < T; t0 += kMaxTenants) {
    const int nt = (int)((T - t0 < kMaxTenants) ? T - t0 : kMaxTenants);
    TenantTable tab;
    for (int i = 0; i < nt; ++i) {
      BD_REQUIRE(w[t0 + i], "tenant_rmsnorm: weight %lld is null", (long long)(t0 + i));
      tab.w[i] = w[t0 + i];
      tab.n_out[i] = (int)H;
    }
    const size_t off = (size_t)t0 * m * H;
    if (dtype == BD_BF16)
      tenant_rmsnorm_kernel<__nv_bfloat16><<<(unsigned)(nt * m), 256, 0, s>>>((const __nv_bfloat16*)x + off, tab, (__nv_bfloat16*)y + off,
                                                                            (int)m, (int)H, eps);
    else
      tenant_rmsnorm_kernel<__half><<<(unsigned)(nt * m), 256, 0, s>>>((const __half*)x + off, tab, (__half*)y + off, (int)m, (int)H, eps);
    count_launch();
    int rc = check_launch("tenant_rmsnorm_kernel");
    if (rc) return rc;
  }
  return BD_OK;
}

extern "C" BD_API int bd_tenant_embed(const int64_t* ids, const void* const* w, const int64_t* n_rows, void* y, int dtype, int64_t T,
                                      int64_t m, int64_t H, void* stream) {
  BD_REQUIRE(dtype == BD_BF16 || dtype == BD_FP16, "tenant_embed: dtype must be BD_BF16 or BD_FP16");
  BD_REQUIRE(ids && w && n_rows && y, "tenant_embed: null pointer");
  BD_REQUIRE(T > 0 && m > 0 && H > 0 && T * m < (1ll << 31) && H < (1ll << 31), "tenant_embed: bad sizes");
  cudaStream_t s = (cudaStream_t)stream;
  for (int64_t t0 = 0; t0 < T; t0 += kMaxTenants) {
    const int nt = (int)((T - t0 < kMaxTenants) ? T - t0 : kMaxTenants);
    TenantTable tab;
    for (int i = 0; i < nt; ++i) {
      BD_REQUIRE(w[t0 + i] && n_rows[t0 + i] > 0 && n_rows[t0 + i] < (1ll << 31), "tenant_embed: table %lld null or empty", (long long)(t0 + i));
      tab.w[i] = w[t0 + i];
      tab.n_out[i] = (int)n_rows[t0 + i];
    }
    const size_t off = (size_t)t0 * m;
    if (dtype == BD_BF16)
      tenant_embed_kernel<__nv_bfloat16><<<(unsigned)(nt * m), 128, 0, s>>>(ids + off, tab, (__nv_bfloat16*)y + off * H, (int)m, (int)H);
    else
      tenant_embed_kernel<__half><<<(unsigned)(nt * m), 128, 0, s>>>(ids + off, tab, (__half*)y + off * H, (int)m, (int)H);
    count_launch();
    int rc = check_launch("tenant_embed_kernel");
    if (rc) return rc;
  }
  return BD_OK;
}
