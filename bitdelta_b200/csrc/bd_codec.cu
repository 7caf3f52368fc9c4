// Bit codec, BinaryDiff compression and delta fold kernels.
//   pack / unpack : bitdelta/binary_gemm_kernel.py:6-46   (bit i of word [j,n] <-> K index n_bits*j+i, LSB first)
//   compress      : bitdelta/diff.py:9-31                 (diff in weight dtype, coeff = mean|diff|, bit = !(diff<0))
//   fold          : bitdelta/diff.py:93-95                (W += ((2b-1)*coeff).T cast to the weight dtype)
// All of these are HBM-bound byte/bit shuffles: coalesced along the contiguous axis, no tensor cores.
#include "bd_common.cuh"

namespace bd {

// ------------------------------------------------------------------ pack / unpack (device)
// One thread per output word (b, j, n); consecutive threads walk n, so both the 32 strided byte reads and the
// word write are coalesced across the warp.
template <typename WordU, int NBITS>
__global__ void pack_kernel(const uint8_t* __restrict__ bits, WordU* __restrict__ words, int64_t total, int64_t J, int64_t N) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int64_t n = idx % N;
  int64_t bj = idx / N;  // b*J + j
  const uint8_t* src = bits + (bj * NBITS) * N + n;
  WordU w = 0;
#pragma unroll
  for (int i = 0; i < NBITS; ++i) w |= (WordU)(src[(int64_t)i * N] != 0 ? 1 : 0) << i;
  words[idx] = w;
}

template <typename WordU, int NBITS>
__global__ void unpack_kernel(const WordU* __restrict__ words, uint8_t* __restrict__ bits, int64_t total, int64_t J, int64_t N) {
  int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int64_t n = idx % N;
  int64_t bj = idx / N;
  WordU w = words[idx];
  uint8_t* dst = bits + (bj * NBITS) * N + n;
#pragma unroll
  for (int i = 0; i < NBITS; ++i) dst[(int64_t)i * N] = (uint8_t)((w >> i) & 1);
}

template <typename WordU, int NBITS>
static int pack_launch(const uint8_t* bits, void* words, int64_t batch, int64_t K, int64_t N, cudaStream_t s) {
  int64_t J = K / NBITS, total = batch * J * N;
  if (total == 0) return BD_OK;
  int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  pack_kernel<WordU, NBITS><<<(unsigned)blocks, threads, 0, s>>>(bits, (WordU*)words, total, J, N);
  count_launch();
  return check_launch("pack_kernel");
}
template <typename WordU, int NBITS>
static int unpack_launch(const void* words, uint8_t* bits, int64_t batch, int64_t J, int64_t N, cudaStream_t s) {
  int64_t total = batch * J * N;
  if (total == 0) return BD_OK;
  int threads = 256;
  int64_t blocks = (total + threads - 1) / threads;
  unpack_kernel<WordU, NBITS><<<(unsigned)blocks, threads, 0, s>>>((const WordU*)words, bits, total, J, N);
  count_launch();
  return check_launch("unpack_kernel");
}

// ------------------------------------------------------------------ pack / unpack (host)
template <typename WordU, int NBITS>
static void pack_host_t(const uint8_t* bits, WordU* words, int64_t batch, int64_t K, int64_t N) {
  int64_t J = K / NBITS;
  for (int64_t b = 0; b < batch; ++b)
    for (int64_t j = 0; j < J; ++j) {
      WordU* out = words + (b * J + j) * N;
      for (int64_t n = 0; n < N; ++n) out[n] = 0;
      for (int i = 0; i < NBITS; ++i) {
        const uint8_t* row = bits + ((b * K) + j * NBITS + i) * N;
        for (int64_t n = 0; n < N; ++n) out[n] |= (WordU)(row[n] != 0 ? 1 : 0) << i;
      }
    }
}
template <typename WordU, int NBITS>
static void unpack_host_t(const WordU* words, uint8_t* bits, int64_t batch, int64_t J, int64_t N) {
  for (int64_t b = 0; b < batch; ++b)
    for (int64_t j = 0; j < J; ++j) {
      const WordU* in = words + (b * J + j) * N;
      for (int i = 0; i < NBITS; ++i) {
        uint8_t* row = bits + ((b * J + j) * NBITS + i) * N;
        for (int64_t n = 0; n < N; ++n) row[n] = (uint8_t)((in[n] >> i) & 1);
      }
    }
}

// ------------------------------------------------------------------ compress (BinaryDiff.__init__)
// A warp owns one weight row n and sweeps K in 256-element steps: each lane loads 8 consecutive elements (16 B),
// turns them into one byte of sign bits, and 4 neighbouring lanes merge their bytes into the int32 word of their
// 32-element K group.  |diff| is summed per thread in fp32 over <= K/32 values, then reduced in fp64.
template <typename T>
__global__ void __launch_bounds__(256) compress_kernel(const T* __restrict__ base, const T* __restrict__ fine,
                                                       int32_t* __restrict__ mask, double* __restrict__ abs_sum, int64_t N, int64_t K) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  int64_t n = blockIdx.x * (int64_t)warps_per_block + (threadIdx.x >> 5);
  double local = 0.0;
  if (n < N) {
    const T* brow = base + n * K;
    const T* frow = fine + n * K;
    for (int64_t k0 = 0; k0 < K; k0 += 256) {
      int64_t k = k0 + lane * 8;
      uint32_t byte = 0;
      float part = 0.f;
      if (k < K) {  // K % 32 == 0 and k % 8 == 0 -> the whole 8-group is in range
        uint4 bv = *reinterpret_cast<const uint4*>(brow + k);
        uint4 fv = *reinterpret_cast<const uint4*>(frow + k);
        const T* bp = reinterpret_cast<const T*>(&bv);
        const T* fp = reinterpret_cast<const T*>(&fv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float d = F16<T>::to_f32(F16<T>::sub(fp[i], bp[i]));  // rounded to the weight dtype (diff.py:11)
          part += fabsf(d);
          byte |= (d < 0.f ? 0u : 1u) << i;  // diff == 0 (and NaN) keep bit 1 (diff.py:14-15)
        }
      }
      local += (double)part;
      uint32_t v = byte << (8 * (lane & 3));
      v |= __shfl_xor_sync(0xffffffffu, v, 1);
      v |= __shfl_xor_sync(0xffffffffu, v, 2);
      if ((lane & 3) == 0 && k < K) mask[(k >> 5) * N + n] = (int32_t)v;
    }
  }
  // block reduction of |diff|
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  __shared__ double red[8];
  if (lane == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < warps_per_block; ++i) s += red[i];
    atomicAdd(abs_sum, s);
  }
}

__global__ void compress_finalize_kernel(double* abs_sum, float* coeff, double inv_count) {
  *coeff = (float)(*abs_sum * inv_count);
  *abs_sum = 0.0;
}

// ------------------------------------------------------------------ fold (load_diff)
// w[n, k] = round(w[n,k] + round(+-coeff)): a warp owns row n; lane l handles k = 32j + l, so the word for
// (j, n) is one broadcast load per 32 elements.  The mask read is strided by N (4 B per 64 B of weights).
template <typename T>
__global__ void __launch_bounds__(256) fold_kernel(T* __restrict__ w, const int32_t* __restrict__ mask, const float* __restrict__ coeff,
                                                   int64_t N, int64_t K) {
  const int lane = threadIdx.x & 31;
  int64_t n = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= N) return;
  const float c = *coeff;
  const float pos = F16<T>::to_f32(F16<T>::from_f32(c)), neg = F16<T>::to_f32(F16<T>::from_f32(-c));
  T* row = w + n * K;
  for (int64_t j = 0; j < K / 32; ++j) {
    uint32_t word = (uint32_t)__ldg(mask + j * N + n);
    float d = ((word >> lane) & 1u) ? pos : neg;
    int64_t k = j * 32 + lane;
    row[k] = F16<T>::from_f32(F16<T>::to_f32(row[k]) + d);
  }
}

}  // namespace bd

// =================================================================== C ABI
using namespace bd;

extern "C" BD_API int bd_pack(const uint8_t* bits, void* words, int n_bits, int64_t batch, int64_t K, int64_t N, void* stream) {
  BD_REQUIRE(n_bits == 8 || n_bits == 16 || n_bits == 32 || n_bits == 64, "pack: n_bits must be 8, 16, 32 or 64 (got %d)", n_bits);
  BD_REQUIRE(batch >= 0 && K >= 0 && N >= 0, "pack: negative dimension");
  BD_REQUIRE(K % n_bits == 0, "K must be divisible by n_bits");
  BD_REQUIRE(batch * K * N == 0 || (bits && words), "pack: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  switch (n_bits) {
    case 8: return pack_launch<uint8_t, 8>(bits, words, batch, K, N, s);
    case 16: return pack_launch<uint16_t, 16>(bits, words, batch, K, N, s);
    case 32: return pack_launch<uint32_t, 32>(bits, words, batch, K, N, s);
    default: return pack_launch<unsigned long long, 64>(bits, words, batch, K, N, s);
  }
}

extern "C" BD_API int bd_unpack(const void* words, uint8_t* bits, int n_bits, int64_t batch, int64_t J, int64_t N, void* stream) {
  BD_REQUIRE(n_bits == 8 || n_bits == 16 || n_bits == 32 || n_bits == 64, "unpack: n_bits must be 8, 16, 32 or 64 (got %d)", n_bits);
  BD_REQUIRE(batch >= 0 && J >= 0 && N >= 0, "unpack: negative dimension");
  BD_REQUIRE(batch * J * N == 0 || (bits && words), "unpack: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  switch (n_bits) {
    case 8: return unpack_launch<uint8_t, 8>(words, bits, batch, J, N, s);
    case 16: return unpack_launch<uint16_t, 16>(words, bits, batch, J, N, s);
    case 32: return unpack_launch<uint32_t, 32>(words, bits, batch, J, N, s);
    default: return unpack_launch<unsigned long long, 64>(words, bits, batch, J, N, s);
  }
}

extern "C" BD_API int bd_pack_host(const uint8_t* bits, void* words, int n_bits, int64_t batch, int64_t K, int64_t N) {
  BD_REQUIRE(n_bits == 8 || n_bits == 16 || n_bits == 32 || n_bits == 64, "pack: n_bits must be 8, 16, 32 or 64 (got %d)", n_bits);
  BD_REQUIRE(batch >= 0 && K >= 0 && N >= 0, "pack: negative dimension");
  BD_REQUIRE(K % n_bits == 0, "K must be divisible by n_bits");
  BD_REQUIRE(batch * K * N == 0 || (bits && words), "pack: null pointer");
  switch (n_bits) {
    case 8: pack_host_t<uint8_t, 8>(bits, (uint8_t*)words, batch, K, N); break;
    case 16: pack_host_t<uint16_t, 16>(bits, (uint16_t*)words, batch, K, N); break;
    case 32: pack_host_t<uint32_t, 32>(bits, (uint32_t*)words, batch, K, N); break;
    default: pack_host_t<unsigned long long, 64>(bits, (unsigned long long*)words, batch, K, N); break;
  }
  return BD_OK;
}

extern "C" BD_API int bd_unpack_host(const void* words, uint8_t* bits, int n_bits, int64_t batch, int64_t J, int64_t N) {
  BD_REQUIRE(n_bits == 8 || n_bits == 16 || n_bits == 32 || n_bits == 64, "unpack: n_bits must be 8, 16, 32 or 64 (got %d)", n_bits);
  BD_REQUIRE(batch >= 0 && J >= 0 && N >= 0, "unpack: negative dimension");
  BD_REQUIRE(batch * J * N == 0 || (bits && words), "unpack: null pointer");
  switch (n_bits) {
    case 8: unpack_host_t<uint8_t, 8>((const uint8_t*)words, bits, batch, J, N); break;
    case 16: unpack_host_t<uint16_t, 16>((const uint16_t*)words, bits, batch, J, N); break;
    case 32: unpack_host_t<uint32_t, 32>((const uint32_t*)words, bits, batch, J, N); break;
    default: unpack_host_t<unsigned long long, 64>((const unsigned long long*)words, bits, batch, J, N); break;
  }
  return BD_OK;
}

extern "C" BD_API int bd_compress(const void* base, const void* finetune, int dtype, int64_t N, int64_t K, int32_t* mask, float* coeff,
                           double* scratch, void* stream) {
  BD_REQUIRE(dtype == BD_BF16 || dtype == BD_FP16, "compress: dtype must be BD_BF16 or BD_FP16");
  BD_REQUIRE(N > 0 && K > 0, "compress: empty weight");
  BD_REQUIRE(K % 32 == 0, "K must be divisible by n_bits");
  BD_REQUIRE(base && finetune && mask && coeff && scratch, "compress: null pointer");
  BD_REQUIRE(((uintptr_t)base % 16 == 0) && ((uintptr_t)finetune % 16 == 0), "compress: weights must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int threads = 256, rows_per_block = threads / 32;
  unsigned blocks = (unsigned)((N + rows_per_block - 1) / rows_per_block);
  if (dtype == BD_BF16)
    compress_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>((const __nv_bfloat16*)base, (const __nv_bfloat16*)finetune, mask, scratch, N, K);
  else
    compress_kernel<__half><<<blocks, threads, 0, s>>>((const __half*)base, (const __half*)finetune, mask, scratch, N, K);
  int rc = check_launch("compress_kernel");
  if (rc) return rc;
  compress_finalize_kernel<<<1, 1, 0, s>>>(scratch, coeff, 1.0 / ((double)N * (double)K));
  count_launch(2);
  return check_launch("compress_finalize_kernel");
}

extern "C" BD_API int bd_fold(void* w, const int32_t* mask, const float* coeff, int dtype, int64_t N, int64_t K, void* stream) {
  BD_REQUIRE(dtype == BD_BF16 || dtype == BD_FP16, "fold: dtype must be BD_BF16 or BD_FP16");
  BD_REQUIRE(N > 0 && K > 0 && K % 32 == 0, "fold: K must be a positive multiple of 32");
  BD_REQUIRE(w && mask && coeff, "fold: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const int threads = 256, rows_per_block = threads / 32;
  unsigned blocks = (unsigned)((N + rows_per_block - 1) / rows_per_block);
  if (dtype == BD_BF16)
    fold_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>((__nv_bfloat16*)w, mask, coeff, N, K);
  else
    fold_kernel<__half><<<blocks, threads, 0, s>>>((__half*)w, mask, coeff, N, K);
  count_launch();
  return check_launch("fold_kernel");
}
