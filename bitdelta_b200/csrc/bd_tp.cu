// Tensor-parallel exchange for row-parallel BinaryDiff linears (BASELINE config 5: Llama-2-70B, W_base and the sign words
// split over the GPUs of one NVSwitch box, one sum per row-parallel layer).
//
// The reference has no multi-GPU data path at all; what this replaces is the NCCL all-reduce a Megatron
// split would call after o_proj / down_proj.  At decode the payload is tiny ([T tokens, hidden] fp32 = 256 KB for
// 8 x 8192), so the cost of a collective is launch + protocol latency, not bandwidth.  One kernel does the whole
// exchange over peer memory:
//
//   1. PUSH   every CTA copies its chunk of this rank's fp32 partial sums into slot [parity][rank] of EVERY rank's
//             exchange buffer (plain stores to peer memory through NVLink / NVSwitch -- fire and forget);
//   2. SIGNAL the last CTA to finish pushing writes this rank's epoch into flag[rank] of every peer (release, system scope);
//   3. WAIT   every CTA spins on its OWN buffer's flags until all ranks have signalled this epoch (local loads only);
//   4. SUM    the `world` slots are added in rank order in fp32 (the same order on every rank: replicas are bit-identical
//             and the result does not depend on timing) and rounded ONCE to the activation dtype.
//
// No trailing barrier: slots are double-buffered by the parity of a device-resident epoch counter, and a rank can only
// start pushing epoch e+2 (same parity as e) after it passed the barrier of e+1, which every peer enters only after its
// kernel of epoch e has completed (stream order).  The epoch lives in device memory, so a captured CUDA graph replays.
#include "bd_common.cuh"

namespace bd {
namespace {

constexpr int kTpMaxWorld = 16;
constexpr size_t kTpHeaderBytes = 4096;  // [0] epoch, [4] arrive counter, [8] done counter, [128 + 4r] flag of rank r
constexpr int kTpThreads = 256;

struct TpArgs {
  char* bufs[kTpMaxWorld];  // rank r's exchange buffer as mapped into this process (bufs[rank] is the local one)
  const float* partial;     // this rank's fp32 partial sums [n]
  void* y;                  // result [n] in `dtype`
  size_t slot_bytes;        // capacity of one slot
  int64_t n;                // fp32 elements (multiple of 4)
  int rank, world, dtype;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_volatile_f4(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kTpThreads) tp_allreduce_kernel(const TpArgs a) {
  // Programmatic dependent launch on both sides: this grid is scheduled while the row-parallel forward that produces
  // `partial` is still running (it needs no shared memory and few registers, so it co-resides) and blocks here until that
  // forward has completed and flushed; the NEXT forward of the stream may start its prologue and prefetch its static
  // weight / sign tiles while the exchange is in flight (its activation loads wait for this grid to complete).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  char* mine = a.bufs[a.rank];
  unsigned* hdr = reinterpret_cast<unsigned*>(mine);
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(hdr) + 1u;  // bumped by the last CTA of the previous call
  __syncthreads();
  const unsigned e = s_epoch;
  const size_t par_off = kTpHeaderBytes + (size_t)(e & 1u) * a.world * a.slot_bytes;
  const int64_t n4 = a.n >> 2;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;

  // 1. push this rank's partial sums into slot [parity][rank] of every rank's buffer
  const float4* src = reinterpret_cast<const float4*>(a.partial);
  for (int64_t i = i0; i < n4; i += stride) {
    const float4 v = src[i];
    for (int r = 0; r < a.world; ++r)
      reinterpret_cast<float4*>(a.bufs[r] + par_off + (size_t)a.rank * a.slot_bytes)[i] = v;
  }
  // 2. the last CTA to finish pushing publishes the epoch to every rank.  Every CTA fences its pushes at system scope
  //    BEFORE it counts itself in, so when the last one arrives all of this rank's data is already performed at the peers:
  //    the flags can then go out as plain stores, one per thread, in parallel (a release store per peer from one thread
  //    serialised eight NVLink round trips: 26 us per exchange at 8 ranks).
  __shared__ unsigned s_last;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(hdr + 1, 1u);
    s_last = (t == gridDim.x - 1) ? 1u : 0u;
    if (s_last) hdr[1] = 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x < a.world)
    *reinterpret_cast<volatile unsigned*>(reinterpret_cast<unsigned*>(a.bufs[threadIdx.x] + 128) + a.rank) = e;
  // 3. wait until every rank has published this epoch (flags in the LOCAL buffer; bounded spin -> trap, never a hang)
  if (threadIdx.x < a.world) {
    const unsigned* f = reinterpret_cast<const unsigned*>(mine + 128) + threadIdx.x;
    unsigned spins = 0;
    while ((int)(ld_acquire_sys(f) - e) < 0) {
      if (++spins > (1u << 28)) __trap();
    }
  }
  __syncthreads();
  // 4. sum the slots in rank order, round once
  for (int64_t i = i0; i < n4; i += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < a.world; ++r) {
      const float4 v = ld_volatile_f4(reinterpret_cast<const float4*>(mine + par_off + (size_t)r * a.slot_bytes) + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (a.dtype == BD_FP32) {
      reinterpret_cast<float4*>(a.y)[i] = acc;
    } else if (a.dtype == BD_BF16) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
      reinterpret_cast<uint2*>(a.y)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    } else {
      const __half2 lo = __floats2half2_rn(acc.x, acc.y), hi = __floats2half2_rn(acc.z, acc.w);
      reinterpret_cast<uint2*>(a.y)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
    }
  }
  // the last CTA to finish advances the epoch for the next call
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned d = atomicAdd(hdr + 2, 1u);
    if (d == gridDim.x - 1) {
      hdr[2] = 0u;
      *reinterpret_cast<volatile unsigned*>(hdr) = e;
    }
  }
}

}  // namespace
}  // namespace bd

using namespace bd;

extern "C" BD_API size_t bd_tp_buffer_bytes(int64_t max_elems, int world) {
  if (max_elems <= 0 || world <= 0 || world > kTpMaxWorld) return 0;
  const size_t slot = ((size_t)max_elems * sizeof(float) + 1023) / 1024 * 1024;
  return kTpHeaderBytes + 2 * (size_t)world * slot;
}

extern "C" BD_API int bd_tp_buffer_create(size_t bytes, void** dev_ptr, void* ipc_handle64) {
  BD_REQUIRE(dev_ptr && ipc_handle64 && bytes >= kTpHeaderBytes, "bd_tp_buffer_create: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  BD_CUDA_OK(cudaMalloc(&p, bytes));
  BD_CUDA_OK(cudaMemset(p, 0, bytes));
  BD_CUDA_OK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  BD_CUDA_OK(cudaIpcGetMemHandle(&h, p));
  memcpy(ipc_handle64, &h, sizeof(h));
  *dev_ptr = p;
  return BD_OK;
}

extern "C" BD_API int bd_tp_buffer_open(const void* ipc_handle64, void** dev_ptr) {
  BD_REQUIRE(dev_ptr && ipc_handle64, "bd_tp_buffer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle64, sizeof(h));
  void* p = nullptr;
  BD_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *dev_ptr = p;
  return BD_OK;
}

extern "C" BD_API int bd_tp_buffer_close(void* dev_ptr, int opened_from_handle) {
  if (!dev_ptr) return BD_OK;
  if (opened_from_handle) BD_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
  else BD_CUDA_OK(cudaFree(dev_ptr));
  return BD_OK;
}

extern "C" BD_API int bd_tp_allreduce(void* const* bufs, size_t buffer_bytes, int rank, int world, const float* partial, int64_t n, void* y,
                                      int dtype, void* stream) {
  BD_REQUIRE(bufs && partial && y, "bd_tp_allreduce: null pointer");
  BD_REQUIRE(world >= 1 && world <= kTpMaxWorld && rank >= 0 && rank < world, "bd_tp_allreduce: bad rank %d / world %d", rank, world);
  BD_REQUIRE(n > 0 && n % 4 == 0, "bd_tp_allreduce: n must be a positive multiple of 4 (got %lld)", (long long)n);
  BD_REQUIRE(dtype == BD_BF16 || dtype == BD_FP16 || dtype == BD_FP32, "bd_tp_allreduce: bad dtype");
  BD_REQUIRE(((uintptr_t)partial | (uintptr_t)y) % 16 == 0, "bd_tp_allreduce: partial and y must be 16-byte aligned");
  BD_REQUIRE(buffer_bytes > kTpHeaderBytes, "bd_tp_allreduce: exchange buffer too small");
  TpArgs a{};
  a.slot_bytes = (buffer_bytes - kTpHeaderBytes) / (2 * (size_t)world) / 1024 * 1024;
  if ((size_t)n * sizeof(float) > a.slot_bytes)
    return fail(BD_ERR_WORKSPACE, "bd_tp_allreduce: %lld elements do not fit a %zu-byte slot", (long long)n, a.slot_bytes);
  for (int r = 0; r < world; ++r) {
    BD_REQUIRE(bufs[r] != nullptr, "bd_tp_allreduce: buffer of rank %d is null", r);
    a.bufs[r] = reinterpret_cast<char*>(bufs[r]);
  }
  a.partial = partial; a.y = y; a.n = n; a.rank = rank; a.world = world; a.dtype = dtype;
  // one float4 per thread up to 64 CTAs (all co-resident: the CTAs of a rank wait for flags that only peers can set)
  int64_t ctas = (n / 4 + kTpThreads - 1) / kTpThreads;
  if (ctas > 64) ctas = 64;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(kTpThreads);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // see the kernel's griddepcontrol use
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  BD_CUDA_OK(cudaLaunchKernelEx(&cfg, tp_allreduce_kernel, a));
  count_launch();
  return check_launch("tp_allreduce_kernel");
}
