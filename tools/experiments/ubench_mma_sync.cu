// Stand-alone microbenchmark: issue rate of the warp-level mma.sync forms a register-operand decode kernel would use on
// sm_100a (sign operand built in registers, no tensor-memory store): m16n8k32 e4m3 x e5m2 and m16n8k16 bf16, alone and
// with the sign-unpack ALU work (one shift + one LOP3 per operand register) in the same loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_mma_sync tools/experiments/ubench_mma_sync.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_fp8(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e5m2.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ void mma_s8(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MODE 4: u8 x s8 IMMAs only; 5: IMMAs + bit-extract ALU; 6: decode unit with IMMA delta (4 bf16 HMMA + 12 IMMA, fresh
// integer accumulators per unit, converted and scaled into fp32 totals)
template <int MODE, int CHAINS>
__global__ void __launch_bounds__(1024) ki(const uint32_t* __restrict__ in, float* out, int iters, long long* cycles) {
  __shared__ uint32_t words[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) words[i] = in[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int d[CHAINS][4];
  float tot[6][4], base[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0;
#pragma unroll
  for (int t = 0; t < 6; ++t) tot[t][0] = tot[t][1] = tot[t][2] = tot[t][3] = 0.f;
  uint32_t a[4] = {in[lane], in[lane + 32], in[lane + 64], in[lane + 96]};
  uint32_t b[2] = {in[lane + 128], in[lane + 160]};
  const uint32_t kOnes = 0x01010101u;
  const int g = lane >> 2, cc = lane & 3;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if constexpr (MODE == 4) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) mma_s8(d[c], a, b);
    } else if constexpr (MODE == 5) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        const uint32_t w0 = words[(it * 8 + c * 16 + g) & 2047], w1 = words[(it * 8 + c * 16 + g + 8) & 2047];
        const uint32_t r[4] = {(w0 >> cc) & kOnes, (w1 >> cc) & kOnes, (w0 >> (cc + 4)) & kOnes, (w1 >> (cc + 4)) & kOnes};
        mma_s8(d[c], r, b);
      }
    } else {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t wa[4];
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&words[((it * 4 + ks) * 128 + lane * 4) & 2047]);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(wa[0]), "=r"(wa[1]), "=r"(wa[2]), "=r"(wa[3]) : "r"(addr));
        const uint32_t xb[2] = {words[(ks * 64 + lane) & 2047], words[(ks * 64 + 32 + lane) & 2047]};
        mma_bf16(base, wa, xb);
      }
#pragma unroll
      for (int t = 0; t < 6; ++t) {
        int acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t w0 = words[(it * 8 + (t * 2 + j) * 128 + warp * 16 + g) & 2047];
          const uint32_t w1 = words[(it * 8 + (t * 2 + j) * 128 + warp * 16 + g + 8) & 2047];
          const uint2 xb2 = *reinterpret_cast<const uint2*>(&words[((t * 2 + j) * 64 + lane * 2) & 2047]);
          const uint32_t xb[2] = {xb2.x, xb2.y};
          const uint32_t r[4] = {(w0 >> cc) & kOnes, (w1 >> cc) & kOnes, (w0 >> (cc + 4)) & kOnes, (w1 >> (cc + 4)) & kOnes};
          mma_s8(acc, r, xb);
        }
        const float sc = __uint_as_float(words[(it + t) & 2047] | 0x30000000u);
#pragma unroll
        for (int i = 0; i < 4; ++i) tot[t][i] = fmaf(__int_as_float(0x4B400000 + acc[i]) - 12582912.0f, sc, tot[t][i]);
      }
    }
  }
  const long long t1 = clock64();
  float s = base[0] + base[1] + base[2] + base[3];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += (float)(d[c][0] + d[c][1] + d[c][2] + d[c][3]);
#pragma unroll
  for (int t = 0; t < 6; ++t) s += tot[t][0] + tot[t][1] + tot[t][2] + tot[t][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// MODE 0: fp8 MMAs only; 1: bf16 MMAs only; 2: fp8 MMAs + unpack ALU (8 ops per MMA); 3: decode-unit mix
// (per "unit": 4 bf16 + 12 fp8 MMAs with unpack, sign words from shared memory)
template <int MODE, int CHAINS>
__global__ void __launch_bounds__(1024) k(const uint32_t* __restrict__ in, float* out, int iters, long long* cycles) {
  __shared__ uint32_t words[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) words[i] = in[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float d[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
  uint32_t a[4] = {in[lane], in[lane + 32], in[lane + 64], in[lane + 96]};
  uint32_t b[2] = {in[lane + 128], in[lane + 160]};
  const uint32_t kMask = 0x80808080u, kOne = 0x38383838u;
  const int g = lane >> 2, cc = lane & 3;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if constexpr (MODE == 0) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) mma_fp8(d[c], a, b);
    } else if constexpr (MODE == 1) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) mma_bf16(d[c], a, b);
    } else if constexpr (MODE == 2) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) {
        const uint32_t w0 = words[(it * 8 + c * 16 + g) & 2047], w1 = words[(it * 8 + c * 16 + g + 8) & 2047];
        uint32_t r[4];
        asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[0]) : "r"(w0 << (7 - cc)), "r"(kMask), "r"(kOne));
        asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[1]) : "r"(w1 << (7 - cc)), "r"(kMask), "r"(kOne));
        asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[2]) : "r"(w0 << (3 - cc)), "r"(kMask), "r"(kOne));
        asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[3]) : "r"(w1 << (3 - cc)), "r"(kMask), "r"(kOne));
        mma_fp8(d[c], r, b);
      }
    } else {
      // one decode unit of a warp: 16 weight rows x 64 k, 6 tenants
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t wa[4];
        const uint32_t addr = (uint32_t)__cvta_generic_to_shared(&words[((it * 4 + ks) * 128 + lane * 4) & 2047]);
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(wa[0]), "=r"(wa[1]), "=r"(wa[2]), "=r"(wa[3]) : "r"(addr));
        const uint32_t xb[2] = {words[(ks * 64 + lane) & 2047], words[(ks * 64 + 32 + lane) & 2047]};
        mma_bf16(d[0], wa, xb);
      }
#pragma unroll
      for (int t = 0; t < 6; ++t) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t w0 = words[(it * 8 + (t * 2 + j) * 128 + warp * 16 + g) & 2047];
          const uint32_t w1 = words[(it * 8 + (t * 2 + j) * 128 + warp * 16 + g + 8) & 2047];
          const uint2 xb2 = *reinterpret_cast<const uint2*>(&words[((t * 2 + j) * 64 + lane * 2) & 2047]);
          const uint32_t xb[2] = {xb2.x, xb2.y};
          uint32_t r[4];
          asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[0]) : "r"(w0 << (7 - cc)), "r"(kMask), "r"(kOne));
          asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[1]) : "r"(w1 << (7 - cc)), "r"(kMask), "r"(kOne));
          asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[2]) : "r"(w0 << (3 - cc)), "r"(kMask), "r"(kOne));
          asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[3]) : "r"(w1 << (3 - cc)), "r"(kMask), "r"(kOne));
          mma_fp8(d[1 + (t % (CHAINS - 1))], r, xb);
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE, int CHAINS>
void runi(const char* name, int warps, const uint32_t* in, float* out, long long* cyc, int per_iter) {
  const int iters = 2000;
  ki<MODE, CHAINS><<<148, warps * 32>>>(in, out, 10, cyc);
  ki<MODE, CHAINS><<<148, warps * 32>>>(in, out, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per = (double)h / iters;
  printf("%-34s warps/CTA %2d (%d per SMSP): %8.1f cyc/iter, %6.2f cyc per MMA per warp, %6.2f cyc per MMA per SMSP\n", name, warps,
         warps / 4, per, per / per_iter, per / per_iter / (warps / 4.0));
}

template <int MODE, int CHAINS>
void run(const char* name, int warps, const uint32_t* in, float* out, long long* cyc, int per_iter) {
  const int iters = 2000;
  k<MODE, CHAINS><<<148, warps * 32>>>(in, out, 10, cyc);
  k<MODE, CHAINS><<<148, warps * 32>>>(in, out, iters, cyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h;
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per = (double)h / iters;
  printf("%-34s warps/CTA %2d (%d per SMSP): %8.1f cyc/iter, %6.2f cyc per MMA per warp, %6.2f cyc per MMA per SMSP\n", name, warps,
         warps / 4, per, per / per_iter, per / per_iter / (warps / 4.0));
}

int main() {
  uint32_t* in;
  float* out;
  long long* cyc;
  cudaMalloc(&in, 2048 * 4);
  cudaMemset(in, 0x3c, 2048 * 4);
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  for (int warps : {4, 8, 12, 16}) {
    run<0, 1>("fp8 m16n8k32, 1 chain", warps, in, out, cyc, 1);
    run<0, 4>("fp8 m16n8k32, 4 chains", warps, in, out, cyc, 4);
    run<0, 8>("fp8 m16n8k32, 8 chains", warps, in, out, cyc, 8);
    run<1, 1>("bf16 m16n8k16, 1 chain", warps, in, out, cyc, 1);
    run<1, 4>("bf16 m16n8k16, 4 chains", warps, in, out, cyc, 4);
    run<2, 4>("fp8 + unpack ALU, 4 chains", warps, in, out, cyc, 4);
    run<2, 8>("fp8 + unpack ALU, 8 chains", warps, in, out, cyc, 8);
    run<3, 7>("decode unit (4 bf16 + 12 fp8)", warps, in, out, cyc, 16);
    runi<4, 1>("u8xs8 m16n8k32, 1 chain", warps, in, out, cyc, 1);
    runi<4, 4>("u8xs8 m16n8k32, 4 chains", warps, in, out, cyc, 4);
    runi<5, 4>("u8xs8 + bit-extract ALU, 4 chains", warps, in, out, cyc, 4);
    runi<6, 1>("decode unit (4 bf16 + 12 s8)", warps, in, out, cyc, 16);
  }
  return 0;
}
