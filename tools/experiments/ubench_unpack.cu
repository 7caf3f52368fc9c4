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

// ---- optional concurrent MMA stream (the kernel's per-unit mix: 4 bf16 SS MMAs N=16 + 12 e4m3 x e5m2 TS MMAs N=8, K=32) ----
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kDesc8LoLbo = (128u >> 4) << 16;
constexpr uint32_t kDesc8Hi = (512u >> 4) | (1u << 14);
__device__ __forceinline__ void mma_ss(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts8(uint32_t d, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 db, {%2, %3};\n\t"
               "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], db, %4, p;\n\t}" ::"r"(d), "r"(a_tmem), "r"(b_lo | kDesc8LoLbo), "r"(kDesc8Hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// MODE 0: full loop; 1: no tcgen05.st (registers consumed by an empty asm); 2: stores only (no ALU: constant registers)
template <int MODE>
__global__ void __launch_bounds__(512) k(const uint32_t* in, int units, int unpack_warps, int spinners, long long* out, int with_mma) {
  __shared__ __align__(1024) uint8_t tiles[16384 + 4096];  // a W tile (SWIZZLE_128B) and B operand tiles; contents are irrelevant
  __shared__ uint64_t bar_mma_arr[3];
  uint64_t& bar_mma = bar_mma_arr[0];
  __shared__ uint32_t words[12 * 128 * 2];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int flag;
  for (int i = threadIdx.x; i < 12 * 128 * 2; i += blockDim.x) words[i] = in[i];
  if (threadIdx.x == 0) {
    flag = 0;
    for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar_mma_arr[i])), "r"(1u) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < (16384 + 4096) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tiles)[i] = 0x3c003c00u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
  } else if (with_mma && warp == 15) {
    // MMA issuer: per unit 4 SS + 12 TS MMAs and a commit; at most 2 units in flight (waits on the commit of unit u-2)
    const bool leader = elect_one();
    const uint32_t w_lo = (smem_u32(tiles) & 0x3FFFFu) >> 4, x_lo = (smem_u32(tiles + 16384) & 0x3FFFFu) >> 4;
    const uint32_t idesc_b = (1u << 4) | (1u << 7) | (1u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc_d = (1u << 4) | (0u << 7) | (1u << 10) | ((8u >> 3) << 17) | ((128u >> 4) << 24);
    uint32_t phase = 0;
    int v = 0, ucount = 0;
    long long t_issue = 0, t_commit = 0, t_wait = 0;
    const long long m0 = clock64();
    for (int u = 0; v == 0; ++u, ++ucount) {
      const long long ta = clock64();
      long long tb = ta, tc = ta;
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          mma_ss(tmem_base + 400, w_lo + ks * 2, x_lo + ks * 2, idesc_b, 1u);
          if ((ks & 1) == 0)
            for (int t = 0; t < 6; ++t) mma_ts8(tmem_base + 416 + t * 8, tmem_base + (u & 1) * 192 + t * 16 + (ks >> 1) * 8, x_lo + 16 + t * 32 + (ks >> 1) * 16, idesc_d, 1u);
        }
        tb = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
        tc = clock64();
      }
      __syncwarp();
      while (!mbar_try_wait(&bar_mma, phase)) {}
      phase ^= 1u;
      if (leader) { t_issue += tb - ta; t_commit += tc - tb; t_wait += clock64() - tc; }
      asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&flag)) : "memory");
    }
    if (leader && blockIdx.x == 0) { out[13] = clock64() - m0; out[14] = ucount; out[10] = t_issue; out[11] = t_commit; out[12] = t_wait; }
  } else if (with_mma == 4 && warp < unpack_warps + spinners) {
    // waiting roles that spin on an mbarrier (the kernel's sync warp and TMA producer): try_wait on a phase that never completes
    int v = 0;
    do {
      for (int rep = 0; rep < 4; ++rep) (void)mbar_try_wait(&bar_mma + 1 + (warp & 1), 0);
      asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&flag)) : "memory");
    } while (v == 0);
  } else if (with_mma >= 2 && warp < unpack_warps + spinners) {
    // instruction-cache probe: the other warps run a dependent ALU chain at the same issue rate either from a 16-instruction
    // loop (with_mma == 2) or from a 1024-instruction straight-line body = 16 KB of code (with_mma == 3)
    uint32_t x = lane, v = 0;
    if (with_mma == 2) {
      do {
#pragma unroll 1
        for (int rep = 0; rep < 64; ++rep) {
#pragma unroll
          for (int i = 0; i < 16; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(i * 77u + 1u), "r"(0x9e3779b9u));
        }
        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&flag)) : "memory");
      } while (v == 0);
    } else {
      do {
#pragma unroll
        for (int i = 0; i < 1024; ++i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(i * 77u + 1u), "r"(0x9e3779b9u));
        asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32((const void*)&flag)) : "memory");
      } while (v == 0);
    }
    if (x == 0x12345u) out[15] = x;
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
void run(const char* name, const uint32_t* in, long long* out, int unpack_warps, int spinners, int with_mma = 0) {
  const int units = 400;
  k<MODE><<<148, 512>>>(in, 10, unpack_warps, spinners, out, with_mma);
  k<MODE><<<148, 512>>>(in, units, unpack_warps, spinners, out, with_mma);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[16];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int w = 0; w < unpack_warps; ++w) mx = h[w] > mx ? h[w] : mx;
  if (with_mma == 1) printf("   MMA warp: %lld units of 4 SS + 12 TS MMAs + commit + wait in %lld cycles = %.1f cycles per unit (issue of the 16 MMAs %.1f, commit %.1f, completion wait %.1f)\n", h[14], h[13], (double)h[13] / (double)h[14], (double)h[10] / h[14], (double)h[11] / h[14], (double)h[12] / h[14]);
  printf("%-28s %s unpack warps %2d, spinning warps %d: %7.1f cycles per unit per warp (12 stores of 1 KB) -> %6.1f cycles per SM-unit of 48 KB\n", name,
         with_mma == 1 ? "+ concurrent MMA stream," : with_mma == 2 ? "+ ALU chain, 16-instr loop," : with_mma == 3 ? "+ ALU chain, 16 KB body," : with_mma == 4 ? "+ mbarrier try_wait spinners," : "", unpack_warps, spinners, mx / units, mx / units * 4.0 / unpack_warps);
}

int main() {
  uint32_t* in; long long* out;
  cudaMalloc(&in, 12 * 128 * 2 * 4); cudaMemset(in, 0x5a, 12 * 128 * 2 * 4);
  cudaMalloc(&out, 16 * 8);
  for (int uw : {4, 8}) for (int sp : {0, 6}) {
    run<0>("LDS + ALU + tcgen05.st", in, out, uw, sp);
    run<1>("LDS + ALU only", in, out, uw, sp);
    run<2>("LDS + tcgen05.st only", in, out, uw, sp);
    run<0>("LDS + ALU + tcgen05.st", in, out, uw, sp, 1);
    run<2>("LDS + tcgen05.st only", in, out, uw, sp, 1);
    if (sp) {
      run<0>("LDS + ALU + tcgen05.st", in, out, uw, sp, 2);
      run<0>("LDS + ALU + tcgen05.st", in, out, uw, sp, 3);
      run<0>("LDS + ALU + tcgen05.st", in, out, uw, sp, 4);
      run<2>("LDS + tcgen05.st only", in, out, uw, sp, 4);
    }
  }
  return 0;
}
