// C ABI entry points of the forward path + error plumbing.  Declarations: include/bitdelta_b200.h.
#include "bd_common.cuh"

namespace bd {

std::string& last_error_ref() {
  static thread_local std::string msg;
  return msg;
}

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

std::atomic<uint64_t> g_launches{0};

static int validate(const FwdProblem& p, const char* who) {
  if (p.dtype != BD_BF16 && p.dtype != BD_FP16) return fail(BD_ERR_UNSUPPORTED, "%s: dtype must be BD_BF16 or BD_FP16", who);
  if (p.T <= 0 || p.m <= 0 || p.K <= 0 || p.N <= 0) return fail(BD_ERR_INVALID, "%s: empty problem (T=%lld m=%lld K=%lld N=%lld)", who,
                                                               (long long)p.T, (long long)p.m, (long long)p.K, (long long)p.N);
  if (p.K % 32 != 0) return fail(BD_ERR_INVALID, "%s: K must be divisible by n_bits (K=%lld)", who, (long long)p.K);
  if (!p.x || !p.masks || !p.y) return fail(BD_ERR_INVALID, "%s: null pointer", who);
  if (p.w && !p.coeff) return fail(BD_ERR_INVALID, "%s: coeff is required with a base weight", who);
  if (p.w && p.coeff_dtype != BD_FP32 && p.coeff_dtype != BD_BF16 && p.coeff_dtype != BD_FP16)
    return fail(BD_ERR_INVALID, "%s: bad coeff dtype %d", who, p.coeff_dtype);
  if (((uintptr_t)p.x | (uintptr_t)p.w | (uintptr_t)p.y) % 16 != 0) return fail(BD_ERR_INVALID, "%s: x, w and y must be 16-byte aligned", who);
  if ((uintptr_t)p.masks % 4 != 0) return fail(BD_ERR_INVALID, "%s: masks must be 4-byte aligned", who);
  if (p.mask_tenant_stride != 0 && p.mask_tenant_stride < (p.K / 32) * p.N) return fail(BD_ERR_INVALID, "%s: mask tenant stride too small", who);
  return BD_OK;
}

// `kernel` argument of the entry points = bd_kernel selector in the low byte + bd_launch_flag bits
static int apply_flags(FwdProblem& p, int kernel, const char* who) {
  if (kernel & ~(0xFF | BD_FLAG_STATIC_OPERANDS | BD_FLAG_FP32_OUT)) return fail(BD_ERR_INVALID, "%s: unknown launch flags 0x%x", who, kernel & ~0xFF);
  p.static_operands = (kernel & BD_FLAG_STATIC_OPERANDS) != 0;
  p.fp32_out = (kernel & BD_FLAG_FP32_OUT) != 0;
  return BD_OK;
}

static int dispatch(FwdProblem& p, int kernel, const char* who) {
  int rc = apply_flags(p, kernel, who);
  if (rc) return rc;
  kernel &= 0xFF;
  rc = validate(p, who);
  if (rc) return rc;
  const char* why = "";
  switch (kernel) {
    case BD_KERNEL_SIMT: return launch_fwd_simt(p);
    case BD_KERNEL_UMMA:
      if (!umma_supports(p, &why)) return fail(BD_ERR_UNSUPPORTED, "%s: tcgen05 kernel does not support this problem: %s", who, why);
      return launch_fwd_umma(p);
    case BD_KERNEL_AUTO: return umma_supports(p, &why) ? launch_fwd_umma(p) : launch_fwd_simt(p);
    default: return fail(BD_ERR_INVALID, "%s: unknown kernel selector %d", who, kernel);
  }
}

}  // namespace bd

using namespace bd;

extern "C" BD_API int bd_abi_version(void) { return BD_ABI_VERSION; }
extern "C" BD_API const char* bd_last_error(void) { return last_error_ref().c_str(); }
extern "C" BD_API uint64_t bd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

#ifdef BD_BRINGUP
// Bring-up library only (libbitdelta_b200_bringup.so): device buffer of 64 x 16 int64 clock64 stamps written by CTA 0 of
// the tcgen05 kernel (nullptr disables it) and the A/B knobs of dbg_flags() in bd_umma.cu.
extern "C" BD_API void bd_debug_set_trace(void* device_buffer) { umma_set_trace(reinterpret_cast<long long*>(device_buffer)); }
extern "C" BD_API void bd_debug_set_flags(int flags, int load_group) { umma_set_debug(flags, load_group); }
#endif

extern "C" BD_API size_t bd_workspace_bytes(int64_t max_rows, int64_t max_n) {
  if (max_rows <= 0 || max_n <= 0) return 0;
  size_t a = simt_workspace_bytes(max_rows, max_n), b = umma_workspace_bytes(max_rows, max_n);
  return a > b ? a : b;
}

extern "C" BD_API int bd_select_kernel(int dtype, int64_t T, int64_t m, int64_t K, int64_t N, int has_base) {
  FwdProblem p{};
  p.dtype = dtype; p.T = T; p.m = m; p.K = K; p.N = N;
  p.mask_tenant_stride = (K / 32) * N;  // contiguous [T, K/32, N] sign words
  p.w = has_base ? reinterpret_cast<const void*>(16) : nullptr;  // only null-ness and alignment are inspected
  p.x = p.y = reinterpret_cast<void*>(16);
  const char* why = "";
  return umma_supports(p, &why) ? BD_KERNEL_UMMA : BD_KERNEL_SIMT;
}

extern "C" BD_API int bd_binary_bmm(const void* a, const int32_t* words, void* c, int dtype, int64_t B, int64_t M, int64_t K, int64_t N,
                             int64_t b_batch_stride, void* workspace, size_t workspace_bytes, int kernel, void* stream) {
  FwdProblem p{};
  p.x = a; p.w = nullptr; p.masks = words; p.coeff = nullptr; p.coeff_dtype = BD_FP32; p.y = c; p.dtype = dtype;
  p.T = B; p.m = M; p.K = K; p.N = N; p.mask_tenant_stride = b_batch_stride;
  p.workspace = workspace; p.workspace_bytes = workspace_bytes; p.stream = (cudaStream_t)stream;
  return dispatch(p, kernel, "bd_binary_bmm");
}

extern "C" BD_API int bd_binarydiff_fwd_grouped(const void* x, int nseg, const void* const* w, const int32_t* const* masks, const void* const* coeff,
                                                 int coeff_dtype, void* const* y, const int64_t* N, const int64_t* mask_tenant_stride, int dtype,
                                                 int64_t T, int64_t m, int64_t K, void* workspace, size_t workspace_bytes, int kernel,
                                                 void* stream) {
  if (nseg < 1 || nseg > 3) return fail(BD_ERR_INVALID, "bd_binarydiff_fwd_grouped: nseg must be 1..3 (got %d)", nseg);
  if (!w || !masks || !coeff || !y || !N || !mask_tenant_stride) return fail(BD_ERR_INVALID, "bd_binarydiff_fwd_grouped: null pointer array");
  FwdProblem segs[3];
  for (int sg = 0; sg < nseg; ++sg) {
    FwdProblem& p = segs[sg];
    p = FwdProblem{};
    if (!w[sg]) return fail(BD_ERR_INVALID, "bd_binarydiff_fwd_grouped: w[%d] is required", sg);
    p.x = x; p.w = w[sg]; p.masks = masks[sg]; p.coeff = coeff[sg]; p.coeff_dtype = coeff_dtype; p.y = y[sg]; p.dtype = dtype;
    p.T = T; p.m = m; p.K = K; p.N = N[sg]; p.mask_tenant_stride = mask_tenant_stride[sg];
    p.workspace = workspace; p.workspace_bytes = workspace_bytes; p.stream = (cudaStream_t)stream;
    int rc = apply_flags(p, kernel, "bd_binarydiff_fwd_grouped");
    if (rc) return rc;
    rc = validate(p, "bd_binarydiff_fwd_grouped");
    if (rc) return rc;
  }
  kernel &= 0xFF;
  const char* why = "";
  bool umma_ok = kernel != BD_KERNEL_SIMT;
  for (int sg = 0; sg < nseg && umma_ok; ++sg) umma_ok = umma_supports(segs[sg], &why);
  if (kernel == BD_KERNEL_UMMA && !umma_ok) return fail(BD_ERR_UNSUPPORTED, "bd_binarydiff_fwd_grouped: tcgen05 kernel does not support this problem: %s", why);
  if (umma_ok) {
    FwdProblem g = segs[0];
    g.nseg = nseg;
    for (int sg = 1; sg < nseg; ++sg) {
      g.seg_w[sg - 1] = segs[sg].w; g.seg_masks[sg - 1] = segs[sg].masks; g.seg_coeff[sg - 1] = segs[sg].coeff; g.seg_y[sg - 1] = segs[sg].y;
      g.seg_N[sg - 1] = segs[sg].N; g.seg_mask_tenant_stride[sg - 1] = segs[sg].mask_tenant_stride;
    }
    return launch_fwd_umma(g);
  }
  for (int sg = 0; sg < nseg; ++sg) {  // general kernel: one launch per matrix
    int rc = launch_fwd_simt(segs[sg]);
    if (rc) return rc;
  }
  return BD_OK;
}

extern "C" BD_API int bd_binarydiff_fwd_batched(const void* x, const void* w, const int32_t* masks, const void* coeff, int coeff_dtype, void* y,
                                         int dtype, int64_t T, int64_t m, int64_t K, int64_t N, int64_t mask_tenant_stride,
                                         void* workspace, size_t workspace_bytes, int kernel, void* stream) {
  if (!w) return fail(BD_ERR_INVALID, "bd_binarydiff_fwd_batched: w is required (use bd_binary_bmm for the delta-only product)");
  FwdProblem p{};
  p.x = x; p.w = w; p.masks = masks; p.coeff = coeff; p.coeff_dtype = coeff_dtype; p.y = y; p.dtype = dtype;
  p.T = T; p.m = m; p.K = K; p.N = N; p.mask_tenant_stride = mask_tenant_stride;
  p.workspace = workspace; p.workspace_bytes = workspace_bytes; p.stream = (cudaStream_t)stream;
  return dispatch(p, kernel, "bd_binarydiff_fwd_batched");
}
