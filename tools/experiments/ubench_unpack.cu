// Stand-alone microbenchmark of the decode kernel's sign-unpack loop (bd_umma.cu, unpack_unit, 8-bit path): per "unit" a warp
// reads 12 sign words from shared memory, turns each into eight e4m3x4 registers (IMAD.SHL + LOP3 per register) and writes
// them to tensor memory with tcgen05.st.32x32b.x8, then tcgen05.wait::st.  How long does that take per unit with one or two
// warps per TMEM lane quadrant, and with other warps of the CTA spinning on shared memory the way the kernel's waiting
// roles do?   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_unpack tools/experiments/ubench_unpack.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}

// MODE 0: full loop; 1: no tcgen05.st (registers consumed by an empty asm); 2: stores only (no ALU: constant registers)
template <int MODE>
__global__ void __launch_bounds__(512) k(const uint32_t* in, int units, int unpack_warps, int spinners, long long* out) {
  __shared__ uint32_t words[12 * 128 * 2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int flag;
  for (int i = threadIdx.x; i < 12 * 128 * 2; i += blockDim.x) words[i] = in[i];
  if (threadIdx.x == 0) flag = 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  uint32_t sign_mask = 0x80808080u;
  asm volatile("" : "+r"(sign_mask));
  const uint32_t kOne = 0x38383838u;
  long long t0 = 0, t1 = 0;
  if (warp < unpack_warps) {
    const int quad = warp & 3, row = quad * 32 + lane;
    const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (warp >> 2) * 192;
    t0 = clock64();
    for (int u = 0; u < units; ++u) {
      const uint32_t* mw = words + (u & 1) * 12 * 128 + row;
#pragma unroll 1
      for (int t0i = 0; t0i < 6; t0i += 3) {
        uint32_t wv[3][2];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) wv[q][jj] = mw[((t0i + q) * 2 + jj) * 128];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int jj = 0; jj < 2; ++jj) {
            uint32_t r[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              if (MODE == 2) r[c] = wv[q][jj];
              else asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[c]) : "r"(wv[q][jj] << (7 - c)), "r"(sign_mask), "r"(kOne));
            }
            if (MODE == 1) asm volatile("" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
            else tmem_st8(ta + (t0i + q) * 16 + jj * 8, r);
          }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    t1 = clock64();
    __syncwarp();
    if (threadIdx.x == 0) flag = 1;
  } else if (warp < unpack_warps + spinners) {
    // waiting roles: poll a shared-memory word (ld.acquire) like wait_released() does
    int v;
    do {
      asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&flag)) : "memory");
    } while (v == 0);
  }
  __syncthreads();
  if (lane == 0 && warp < unpack_warps && blockIdx.x == 0) out[warp] = t1 - t0;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

template <int MODE>
void run(const char* name, const uint32_t* in, long long* out, int unpack_warps, int spinners) {
  const int units = 400;
  k<MODE><<<148, 512>>>(in, 10, unpack_warps, spinners, out);
  k<MODE><<<148, 512>>>(in, units, unpack_warps, spinners, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[16];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int w = 0; w < unpack_warps; ++w) mx = h[w] > mx ? h[w] : mx;
  printf("%-28s unpack warps %2d, spinning warps %d: %7.1f cycles per unit per warp (12 stores of 1 KB) -> %6.1f cycles per SM-unit of 48 KB\n", name,
         unpack_warps, spinners, mx / units, mx / units * 4.0 / unpack_warps);
}

int main() {
  uint32_t* in; long long* out;
  cudaMalloc(&in, 12 * 128 * 2 * 4); cudaMemset(in, 0x5a, 12 * 128 * 2 * 4);
  cudaMalloc(&out, 16 * 8);
  for (int uw : {4, 8}) for (int sp : {0, 6}) {
    run<0>("LDS + ALU + tcgen05.st", in, out, uw, sp);
    run<1>("LDS + ALU only", in, out, uw, sp);
    run<2>("LDS + tcgen05.st only", in, out, uw, sp);
  }
  return 0;
}
