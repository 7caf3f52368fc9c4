/*
 * bitdelta_b200 -- C ABI of the B200-native BitDelta hot path (libbitdelta_b200.so).
 *
 * The reference (FasterDecoding/BitDelta) has no FFI: its boundary is the Python module
 * bitdelta/binary_gemm_kernel.py plus the nn.Module classes that replace nn.Linear leaves.
 * Every entry point below replaces the body of one of those Python symbols; the Python package
 * bitdelta_b200/ keeps the reference's names and signatures and calls these through ctypes
 * (see INTEGRATION.md for the binding a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers + sizes, no C++/torch types; all device pointers belong to the caller.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every device entry point is
 *     asynchronous on that stream, performs no allocation and no host sync, and is CUDA-graph capturable.
 *   - return value: 0 = BD_OK, negative = error; bd_last_error() returns a thread-local message.
 *     No C++ exception crosses the ABI.
 *   - 16-bit float tensors are selected by `dtype` (BD_BF16 / BD_FP16); sign words are int32, LSB-first
 *     along K exactly as reference pack() lays them out (binary_gemm_kernel.py:6-32).
 */
#ifndef BITDELTA_B200_H
#define BITDELTA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BD_ABI_VERSION 1

#if defined(__GNUC__)
#define BD_API __attribute__((visibility("default")))
#else
#define BD_API
#endif

enum bd_status {
  BD_OK = 0,
  BD_ERR_INVALID = -1,     /* bad argument (shape, alignment, null pointer) */
  BD_ERR_UNSUPPORTED = -2, /* dtype / shape outside what the kernels implement */
  BD_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed; message carries cudaGetErrorString */
  BD_ERR_WORKSPACE = -4    /* workspace missing or smaller than bd_workspace_bytes() */
};

enum bd_dtype { BD_BF16 = 0, BD_FP16 = 1, BD_FP32 = 2 };

/* kernel selection for the fused forward (bd_binarydiff_fwd_batched / bd_binary_bmm) */
enum bd_kernel {
  BD_KERNEL_AUTO = 0, /* tcgen05 kernel when the shape qualifies, otherwise the general SIMT kernel */
  BD_KERNEL_SIMT = 1, /* general CUDA-core kernel: any N, K % 32 == 0 */
  BD_KERNEL_UMMA = 2  /* tcgen05/TMEM/TMA kernel: fails with BD_ERR_UNSUPPORTED if the shape does not qualify */
};

/* Launch flags, OR-ed into the `kernel` argument of the forward entry points (bits 8 and up; the low byte is bd_kernel).
 *
 * BD_FLAG_STATIC_OPERANDS: the caller guarantees that `w` and the sign words were NOT written by work that precedes
 *   this call on `stream` since the last synchronisation point (long-lived weights: BinaryDiff / DiffCompressModule
 *   buffers).  The tcgen05 kernel is launched with programmatic dependent launch; with this flag it requests its first
 *   weight / sign tiles while the preceding kernel of the stream is still draining.  WITHOUT the flag every load is issued
 *   after griddepcontrol.wait, so operands produced by the immediately preceding kernel -- binary_bmm(a, pack(b)) from the
 *   reference's notebook, a freshly made .contiguous() copy -- are always seen.  Activations `x` are never prefetched.
 * BD_FLAG_FP32_OUT: y is written as fp32 [T,m,N] (no rounding to `dtype`): the partial sums of a row-parallel
 *   tensor-parallel shard, reduced across ranks in fp32 and rounded once (bitdelta_b200/parallel.py).
 */
enum bd_launch_flag {
  BD_FLAG_STATIC_OPERANDS = 1 << 8,
  BD_FLAG_FP32_OUT = 1 << 9
};

BD_API int bd_abi_version(void);
BD_API const char* bd_last_error(void);

/* Number of CUDA kernels this library has launched in the calling process (bench.py's `gpu_launches`). */
BD_API uint64_t bd_launch_count(void);

/* ---- a1 / a2: the bit codec ------------------------------------------------------------------------------
 * pack  replaces bitdelta/binary_gemm_kernel.py:6-32   bool (*,K,N)      -> int (*,K/n_bits,N)
 * unpack replaces bitdelta/binary_gemm_kernel.py:34-46 int (*,K/n_bits,N)-> bool (*,K,N)
 * `bits` is a byte-per-bool array (torch.bool storage); `words` has n_bits/8 bytes per element
 * (n_bits in {8,16,32,64}; uint8/int16/int32/int64 as in the reference).  `batch` = product of leading dims.
 * The *_host variants run on host memory (diff.pt tensors live on the CPU in save_diff/load_diff).
 */
BD_API int bd_pack(const uint8_t* bits, void* words, int n_bits, int64_t batch, int64_t K, int64_t N, void* stream);
BD_API int bd_unpack(const void* words, uint8_t* bits, int n_bits, int64_t batch, int64_t J, int64_t N, void* stream);
BD_API int bd_pack_host(const uint8_t* bits, void* words, int n_bits, int64_t batch, int64_t K, int64_t N);
BD_API int bd_unpack_host(const void* words, uint8_t* bits, int n_bits, int64_t batch, int64_t J, int64_t N);

/* ---- a5: BinaryDiff.__init__ (bitdelta/diff.py:9-31) on the device -----------------------------------------
 * base, finetune: [N,K] row-major, `dtype`.  Writes mask int32 [K/32,N] (bit = !(finetune-base < 0), the
 * subtraction rounded to `dtype` first as the reference does) and *coeff = mean(|finetune-base|) as fp32
 * (device pointer).  `scratch` is a device double the kernel accumulates into; it must be zero on entry
 * and is re-zeroed on exit.
 */
BD_API int bd_compress(const void* base, const void* finetune, int dtype, int64_t N, int64_t K, int32_t* mask, float* coeff,
                double* scratch, void* stream);

/* ---- diff.py:93-95 fold: w[n,k] += round_dtype(coeff * (2*bit(k,n)-1)), in `dtype` arithmetic like load_diff -- */
BD_API int bd_fold(void* w, const int32_t* mask, const float* coeff, int dtype, int64_t N, int64_t K, void* stream);

/* ---- a3 / a4: binary_matmul / binary_bmm (binary_gemm_kernel.py:153-184, :297-335) -------------------------
 * c[b,M,N] = a[b,M,K] . (2*unpack(words[b])-1), fp32 accumulate, one rounding to `dtype`.
 * `b_batch_stride` is the distance in int32 words between consecutive batch entries of `words`
 * (K/32*N for a contiguous [B,K/32,N]; 0 broadcasts one sign matrix, which replaces the mask.repeat of diff.py:38).
 */
BD_API int bd_binary_bmm(const void* a, const int32_t* words, void* c, int dtype, int64_t B, int64_t M, int64_t K, int64_t N,
                  int64_t b_batch_stride, void* workspace, size_t workspace_bytes, int kernel, void* stream);

/* ---- a6 / a7: the fused BinaryDiff / DiffCompressModule forward --------------------------------------------
 * y[t,i,:] = x[t,i,:] . w^T + coeff[t] * ( x[t,i,:] . (2*unpack(masks[t])-1) )       t < T tenants, i < m rows
 *   replaces  bitdelta/diff.py:33-39            (T = 1: x is [B*seq, K], one mask, one fp32 coeff)
 *   and       demo/demo_backend.py:93-98        (T tenants share w; masks [T,K/32,N]; coeff [T])
 * x: [T,m,K] `dtype` contiguous; w: [N,K] row-major `dtype` (nn.Linear.weight; BinaryDiff.base is its .T view);
 * masks: int32, tenant stride `mask_tenant_stride` words; coeff: T values of `coeff_dtype` (BD_FP32, or BD_BF16/BD_FP16
 * as the demo stores them) on the device; y: [T,m,N] `dtype`.  Both products accumulate in fp32 and the sum is
 * rounded once.  K % 32 == 0.  `workspace` must hold bd_workspace_bytes() bytes, zero-filled once at allocation;
 * the kernels leave it zeroed for the next call (calls sharing a workspace must be stream-ordered).
 */
BD_API int bd_binarydiff_fwd_batched(const void* x, const void* w, const int32_t* masks, const void* coeff, int coeff_dtype, void* y,
                              int dtype, int64_t T, int64_t m, int64_t K, int64_t N, int64_t mask_tenant_stride,
                              void* workspace, size_t workspace_bytes, int kernel, void* stream);

/* Grouped form: `nseg` (1..3) BinaryDiff linears that consume the SAME activations in one launch -- q/k/v of an attention
 * block or gate/up of an MLP (the callers of a6/a7 in SURVEY.md 3.1/3.2 issue them back to back on one input).  Arrays
 * have `nseg` entries: w[s] [N[s], K], masks[s] (tenant stride mask_tenant_stride[s] words), coeff[s] (T values),
 * y[s] [T, m, N[s]].  Semantics per matrix are exactly bd_binarydiff_fwd_batched's; one launch streams all of them, so the
 * per-launch fixed cost is paid once.  Every N but the last must be a multiple of 128 for the single-launch path
 * (otherwise, or on the SIMT kernel, the matrices are processed one after the other). */
BD_API int bd_binarydiff_fwd_grouped(const void* x, int nseg, const void* const* w, const int32_t* const* masks, const void* const* coeff,
                              int coeff_dtype, void* const* y, const int64_t* N, const int64_t* mask_tenant_stride, int dtype,
                              int64_t T, int64_t m, int64_t K, void* workspace, size_t workspace_bytes, int kernel, void* stream);

/* ---- a8: per-tenant dense leaves, DataParallelModule.forward (demo/demo_backend.py:69-79) --------------------
 * Row t of the batch goes through tenant t's own full-precision weight; one launch covers all tenants instead of
 * the reference's host loop (weight.data swap + module call per tenant, :72-74) and nested-tensor padding (:77-79).
 * `w` is a HOST array of T device pointers (tenant weights are separate tensors in the checkpoints); it is copied
 * into the kernel parameters, so the calls are asynchronous and graph-capturable.
 *
 * bd_tenant_linear  (lm_head):  y[t,i,v] = x[t,i,:] . w[t][v,:]  for v < n_out[t], fp32 accumulate, one rounding;
 *   columns n_out[t]..ldy-1 are filled with the lowest finite value of `dtype` (torch.finfo(dtype).min, :78).
 *   x: [T,m,K], w[t]: [n_out[t],K] row-major, y: [T,m,ldy]; m <= BD_TENANT_LINEAR_MAX_ROWS (decode; larger m is a
 *   plain per-tenant library GEMM and stays with the caller), K % 8 == 0, 16-byte aligned pointers.  `bias` (may be
 *   NULL) is the wrapped module's shared bias [ldy]; it requires equal widths.
 * bd_tenant_rmsnorm (HF Llama/Mistral RMSNorm, the class the reference's model instantiates):
 *   y[t,i,:] = w[t] * round(x[t,i,:] * rsqrt(mean(x[t,i,:]^2) + eps)), statistics in fp32, both roundings to `dtype`.
 * bd_tenant_embed   (embed_tokens): y[t,i,:] = w[t][ids[t,i], :]; ids int64 [T,m] on the device, n_rows[t] table rows.
 */
#define BD_TENANT_LINEAR_MAX_ROWS 4
BD_API int bd_tenant_linear(const void* x, const void* const* w, const int64_t* n_out, const void* bias, void* y, int dtype,
                     int64_t T, int64_t m, int64_t K, int64_t ldy, void* stream);
BD_API int bd_tenant_rmsnorm(const void* x, const void* const* w, void* y, int dtype, int64_t T, int64_t m, int64_t H, float eps,
                      void* stream);
BD_API int bd_tenant_embed(const int64_t* ids, const void* const* w, const int64_t* n_rows, void* y, int dtype, int64_t T, int64_t m,
                    int64_t H, void* stream);

/* ---- e: tensor-parallel exchange for row-parallel BinaryDiff linears (BASELINE config 5) ------------------------
 * The reference has no multi-GPU data path (its only multi-GPU artefact is accelerate layer placement,
 * bitdelta/utils.py:82-101); a Megatron split of BinaryDiff.forward (bitdelta/diff.py:33-39) needs ONE sum over the ranks
 * after each row-parallel linear (o_proj, down_proj).  bd_tp_allreduce does that sum in one kernel over NVLink peer memory
 * instead of a library collective: every rank pushes its fp32 partial sums (the output of a forward launched with
 * BD_FLAG_FP32_OUT) into all ranks' exchange buffers, signals, waits for the other ranks' signals and adds the slots in
 * rank order -- deterministic, identical on every rank, one rounding to `dtype`.
 *
 * One exchange buffer per rank, created with bd_tp_buffer_create (cudaMalloc + zero fill; `ipc_handle64` receives the
 * 64-byte cudaIpcMemHandle_t to send to the other ranks' processes) and mapped by them with bd_tp_buffer_open.
 * bd_tp_buffer_bytes(max_elems, world) sizes it for sums of up to `max_elems` fp32 values.  bd_tp_allreduce takes the HOST
 * array `bufs` of `world` device pointers (bufs[r] = rank r's buffer as mapped in this process, bufs[rank] the local one),
 * all ranks must call it the same number of times with the same `n` (a multiple of 4); the call is asynchronous on
 * `stream` and CUDA-graph capturable (its epoch counter lives in the buffer). */
BD_API size_t bd_tp_buffer_bytes(int64_t max_elems, int world);
BD_API int bd_tp_buffer_create(size_t bytes, void** dev_ptr, void* ipc_handle64);
BD_API int bd_tp_buffer_open(const void* ipc_handle64, void** dev_ptr);
BD_API int bd_tp_buffer_close(void* dev_ptr, int opened_from_handle);
BD_API int bd_tp_allreduce(void* const* bufs, size_t buffer_bytes, int rank, int world, const float* partial, int64_t n, void* y,
                           int dtype, void* stream);

/* Upper bound of the workspace any forward of at most `max_rows` = T*m rows and `max_n` outputs needs on the
 * current device. */
BD_API size_t bd_workspace_bytes(int64_t max_rows, int64_t max_n);

/* Which kernel BD_KERNEL_AUTO would pick for this problem (BD_KERNEL_SIMT or BD_KERNEL_UMMA). */
BD_API int bd_select_kernel(int dtype, int64_t T, int64_t m, int64_t K, int64_t N, int has_base);

/* Bring-up instrumentation is NOT part of this library: the per-unit clock64 trace and the A/B knobs of the tcgen05 kernel
 * exist only in libbitdelta_b200_bringup.so (python bitdelta_b200/build.py --bringup, compiled with -DBD_BRINGUP), which
 * tools/umma_trace.py and tools/kernel_bench.py load explicitly.  The release library has no mutable global state besides
 * mutex-guarded caches. */
#ifdef BD_BRINGUP
#define BD_BRINGUP_API BD_API
BD_BRINGUP_API void bd_debug_set_trace(void* device_buffer); /* 64 x 16 int64 stamps of CTA 0; NULL disables */
BD_BRINGUP_API void bd_debug_set_flags(int flags, int reserved);
#endif

#ifdef __cplusplus
}
#endif
#endif /* BITDELTA_B200_H */
