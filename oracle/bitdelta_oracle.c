/* CPU restatement (plain C + OpenMP) of the BitDelta hot path -- TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Same arithmetic as oracle/bitdelta_oracle.py (which is pinned to the reference's golden vectors; this file is
 * checked against it in tests/test_oracle_c.py).  Used by bench.py's `cpu_baseline` leg and `--impl reference` arm
 * as the "port" of the reference's CPU unpack-matmul path
 *     y = x @ base + coeff * (x @ (unpack(mask)*2-1))        bitdelta/diff.py:39 + :93, binary_gemm_kernel.py:34-46
 * and its multi-tenant form (demo/demo_backend.py:93-98).  Nothing under bitdelta_b200/ links or loads this.
 *
 * Layouts: x [T,m,K] bf16 bits; w [N,K] bf16 bits (nn.Linear.weight); masks [T,K/32,N] int32, bit i of word [j,n] is
 * K index 32j+i (binary_gemm_kernel.py:6-32); coeff [T] float; y [T,m,N] float32 (fp32 accumulate, no output rounding).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float bf16_to_f32(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

/* unpack (binary_gemm_kernel.py:34-46): words [J,N] -> bits [32J,N] bytes */
void oracle_unpack_i32(const int32_t* words, uint8_t* bits, int64_t J, int64_t N) {
  for (int64_t j = 0; j < J; ++j)
    for (int i = 0; i < 32; ++i)
      for (int64_t n = 0; n < N; ++n) bits[(j * 32 + i) * N + n] = (uint8_t)((words[j * N + n] >> i) & 1);
}

/* pack (binary_gemm_kernel.py:6-32) */
void oracle_pack_i32(const uint8_t* bits, int32_t* words, int64_t K, int64_t N) {
  for (int64_t j = 0; j < K / 32; ++j)
    for (int64_t n = 0; n < N; ++n) {
      uint32_t w = 0;
      for (int i = 0; i < 32; ++i) w |= (uint32_t)(bits[(j * 32 + i) * N + n] != 0) << i;
      words[j * N + n] = (int32_t)w;
    }
}

/* fused forward, all tenants; threads <= 0 means "all the host threads OpenMP gives us" */
void oracle_fwd_batched_bf16(const uint16_t* x, const uint16_t* w, const int32_t* masks, const float* coeff, float* y,
                             int64_t T, int64_t m, int64_t K, int64_t N, int threads) {
  const int64_t rows = T * m;
  float* xf = (float*)malloc(sizeof(float) * rows * K);
  float* yd = (float*)malloc(sizeof(float) * rows * N); /* delta product, combined with the base product at the end */
  for (int64_t i = 0; i < rows * K; ++i) xf[i] = bf16_to_f32(x[i]);
  const int64_t NB = 32; /* columns per work item: one 128-byte run of sign words per (tenant, j) */
  const int64_t nblocks = (N + NB - 1) / NB;
#ifdef _OPENMP
  omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#endif
  /* work items: (column block) x (part), part 0 = base product for every row, part t+1 = tenant t's delta product */
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
  for (int64_t nb = 0; nb < nblocks; ++nb) {
    for (int64_t part = 0; part <= T; ++part) {
      const int64_t n0 = nb * NB, n1 = (n0 + NB < N) ? n0 + NB : N;
      if (part == 0) {
        /* base product: x . w^T (w may be NULL for the delta-only binary_bmm) */
        for (int64_t r = 0; r < rows; ++r)
          for (int64_t n = n0; n < n1; ++n) {
            float acc = 0.f;
            if (w) {
              const uint16_t* wr = w + n * K;
              const float* xr = xf + r * K;
#pragma omp simd reduction(+ : acc)
              for (int64_t k = 0; k < K; ++k) acc += xr[k] * bf16_to_f32(wr[k]);
            }
            y[r * N + n] = acc;
          }
      } else {
        /* delta product: x . (2*bits-1) for tenant t */
        const int64_t t = part - 1;
        const int32_t* mt = masks + t * (K / 32) * N;
        for (int64_t i = 0; i < m; ++i) {
          const float* xr = xf + (t * m + i) * K;
          float acc[32];
          for (int64_t c = 0; c < NB; ++c) acc[c] = 0.f;
          for (int64_t j = 0; j < K / 32; ++j) {
            const int32_t* wj = mt + j * N + n0;
            const float* xj = xr + j * 32;
            for (int b = 0; b < 32; ++b) {
              const float xv = xj[b];
              for (int64_t c = 0; c < n1 - n0; ++c) acc[c] += ((wj[c] >> b) & 1) ? xv : -xv;
            }
          }
          for (int64_t c = 0; c < n1 - n0; ++c) yd[(t * m + i) * N + n0 + c] = acc[c];
        }
      }
    }
  }
  for (int64_t t = 0; t < T; ++t) {
    const float cf = coeff ? coeff[t] : 1.0f;
    for (int64_t i = 0; i < m * N; ++i) {
      const int64_t o = t * m * N + i;
      y[o] = w ? y[o] + cf * yd[o] : yd[o];
    }
  }
  free(yd);
  free(xf);
}
