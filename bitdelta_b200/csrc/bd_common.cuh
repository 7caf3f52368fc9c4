// Shared helpers for libbitdelta_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/bitdelta_b200.h"

namespace bd {

// ---- error plumbing (thread-local message, integer status; nothing throws across the ABI) ----
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define BD_CUDA_OK(expr)                                                                                   \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess) return ::bd::fail(BD_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define BD_REQUIRE(cond, ...) \
  do {                        \
    if (!(cond)) return ::bd::fail(BD_ERR_INVALID, __VA_ARGS__); \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    return fail(BD_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  }
  return BD_OK;
}

// ---- 16-bit float traits ----
template <typename T>
struct F16;
template <>
struct F16<__nv_bfloat16> {
  using vec2 = __nv_bfloat162;
  static __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
  // low / high element of a packed pair as fp32 (bf16 -> fp32 is a 16-bit shift)
  static __device__ __forceinline__ float lo(uint32_t p) { return __uint_as_float(p << 16); }
  static __device__ __forceinline__ float hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }
  static __device__ __forceinline__ __nv_bfloat16 sub(__nv_bfloat16 a, __nv_bfloat16 b) { return __hsub(a, b); }
};
template <>
struct F16<__half> {
  using vec2 = __half2;
  static __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ float lo(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p & 0xFFFFu))); }
  static __device__ __forceinline__ float hi(uint32_t p) { return __half2float(__ushort_as_half((unsigned short)(p >> 16))); }
  static __device__ __forceinline__ __half sub(__half a, __half b) { return __hsub(a, b); }
};

__device__ __forceinline__ float load_coeff(const void* p, int dtype, int64_t i) {
  if (dtype == BD_FP32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == BD_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}

// ---- problem descriptor shared by the forward kernels ----
struct FwdProblem {
  const void* x;         // [rows, K]   rows = T*m, row r belongs to tenant r / m
  const void* w;         // [N, K] or nullptr (delta-only: binary_bmm)
  const int32_t* masks;  // tenant t at masks + t*mask_tenant_stride, [K/32, N]
  const void* coeff;     // T values or nullptr (delta-only: coefficient 1)
  int coeff_dtype;
  void* y;               // [rows, N]
  int dtype;
  int64_t T, m, K, N;
  int64_t mask_tenant_stride;
  void* workspace;
  size_t workspace_bytes;
  cudaStream_t stream;
  bool static_operands = false;  // BD_FLAG_STATIC_OPERANDS: w / sign words may be prefetched ahead of the stream's preceding kernel
  bool fp32_out = false;         // BD_FLAG_FP32_OUT: y is fp32 (tensor-parallel partial sums)
  // Grouped launch (optional): `nseg` > 1 matrices that share x, K, T, m -- e.g. q/k/v or gate/up.  Segment 0 is the
  // main w/masks/coeff/y/N above; segments 1.. are listed here.  Every N but the last must be a multiple of 128.
  int nseg = 1;
  const void* seg_w[2] = {nullptr, nullptr};
  const int32_t* seg_masks[2] = {nullptr, nullptr};
  const void* seg_coeff[2] = {nullptr, nullptr};
  void* seg_y[2] = {nullptr, nullptr};
  int64_t seg_N[2] = {0, 0};
  int64_t seg_mask_tenant_stride[2] = {0, 0};
};

// Workspace layout shared by the forward kernels: two zero-initialised counter regions, then scratch.
constexpr size_t kWsCounterBytes = 4096;        // per kernel family
constexpr size_t kWsUmmaCounterOffset = 4096;   // tcgen05 kernel's counters
constexpr size_t kWsScratchOffset = 8192;       // split-K slots start here

int launch_fwd_simt(const FwdProblem& p);
int launch_fwd_umma(const FwdProblem& p);
bool umma_supports(const FwdProblem& p, const char** why);
size_t simt_workspace_bytes(int64_t rows, int64_t N);
size_t umma_workspace_bytes(int64_t rows, int64_t N);
#ifdef BD_BRINGUP
void umma_set_trace(long long* buf);
void umma_set_debug(int flags, int load_group);
#endif

}  // namespace bd
