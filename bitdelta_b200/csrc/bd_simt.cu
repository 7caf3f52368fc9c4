// General CUDA-core forward kernel: any N, any K % 32 == 0, any number of rows.
//
//   y[r, n] = sum_k x[r,k] w[n,k]  +  coeff[t(r)] * sum_k x[r,k] * s_{t(r)}[k,n],      s = +1 if bit else -1
//
// It is the shape-agnostic member of the family (bitdelta/diff.py:39, demo/demo_backend.py:95-98,
// binary_gemm_kernel.py:297-335) and the first kernel that was brought up; the tcgen05 kernel in bd_umma.cu takes
// over whenever the shape qualifies.  HBM-bound design: a CTA owns 64 output columns and one K split, streams the
// 64xK_split slab of w with 16-byte loads (a warp per row, lanes along K -> fully coalesced) and the matching
// [K_split/32, 64] slab of sign words (thread per column -> coalesced along N), keeps both accumulators in fp32 and
// combines K splits deterministically through per-split slots in the workspace (the last CTA to arrive sums them in
// split order, so results do not depend on scheduling).
#include "bd_common.cuh"

namespace bd {

namespace {
constexpr int kBN = 64;        // output columns per CTA
constexpr int kKS = 256;       // K elements staged per step
constexpr int kThreads = 256;  // 8 warps
constexpr int kTM = 8;         // rows (tokens) per CTA pass
constexpr int kMaxSplits = 32;

struct SimtArgs {
  const void* x;
  const void* w;
  const int32_t* masks;
  const void* coeff;
  int coeff_dtype;
  void* y;
  float* partial;      // [splits][rows][N]
  unsigned* counters;  // [row_chunks * n_tiles]
  int64_t rows, m, K, N, mask_tenant_stride;
  int splits, slabs_per_split;
  int fp32_out;  // BD_FLAG_FP32_OUT: y is fp32 (no rounding)
};

template <typename T, bool HAS_BASE>
__global__ void __launch_bounds__(kThreads) fwd_simt_kernel(SimtArgs a) {
  __shared__ __align__(16) float xs[kTM][kKS];
  __shared__ float red_base[kTM][kBN];
  __shared__ float red_delta[4][kTM][kBN];
  __shared__ unsigned s_last;

  const T* __restrict__ x = reinterpret_cast<const T*>(a.x);
  const T* __restrict__ w = reinterpret_cast<const T*>(a.w);
  T* __restrict__ y = reinterpret_cast<T*>(a.y);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t n0 = (int64_t)blockIdx.x * kBN;
  const int split = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.z * kTM;
  const int nrows = (int)min((int64_t)kTM, a.rows - r0);
  const int64_t K = a.K, N = a.N;

  float accb[8][kTM];
  float accd[kTM];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int t = 0; t < kTM; ++t) accb[i][t] = 0.f;
#pragma unroll
  for (int t = 0; t < kTM; ++t) accd[t] = 0.f;

  const int dcol = tid & (kBN - 1);  // delta phase: column within the tile
  const int dq = tid >> 6;           // delta phase: which quarter of the slab's words
  const int64_t dn = n0 + dcol;

  const int64_t slab_begin = (int64_t)split * a.slabs_per_split;
  const int64_t total_slabs = (K + kKS - 1) / kKS;
  const int64_t slab_end = min(total_slabs, slab_begin + a.slabs_per_split);

  for (int64_t slab = slab_begin; slab < slab_end; ++slab) {
    const int64_t k0 = slab * kKS;
    const int kw = (int)min((int64_t)kKS, K - k0);  // multiple of 32
    // ---- stage the activation slab as fp32 (zero padded) ----
    for (int idx = tid; idx < kTM * kKS; idx += kThreads) {
      int t = idx / kKS, kk = idx % kKS;
      float v = 0.f;
      if (t < nrows && kk < kw) v = F16<T>::to_f32(x[(r0 + t) * K + k0 + kk]);
      xs[t][kk] = v;
    }
    __syncthreads();

    if (HAS_BASE) {
      const int kk = lane * 8;
      if (kk < kw) {
        float xr[kTM][8];
#pragma unroll
        for (int t = 0; t < kTM; ++t) {
          float4 v0 = *reinterpret_cast<const float4*>(&xs[t][kk]);
          float4 v1 = *reinterpret_cast<const float4*>(&xs[t][kk + 4]);
          xr[t][0] = v0.x; xr[t][1] = v0.y; xr[t][2] = v0.z; xr[t][3] = v0.w;
          xr[t][4] = v1.x; xr[t][5] = v1.y; xr[t][6] = v1.z; xr[t][7] = v1.w;
        }
        uint4 wv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {  // issue all 8 row loads before using them
          int64_t n = n0 + warp * 8 + i;
          wv[i] = make_uint4(0, 0, 0, 0);
          if (n < N) wv[i] = __ldg(reinterpret_cast<const uint4*>(w + n * K + k0 + kk));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t p[4] = {wv[i].x, wv[i].y, wv[i].z, wv[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float lo = F16<T>::lo(p[e]), hi = F16<T>::hi(p[e]);
#pragma unroll
            for (int t = 0; t < kTM; ++t) {
              accb[i][t] = fmaf(lo, xr[t][2 * e], accb[i][t]);
              accb[i][t] = fmaf(hi, xr[t][2 * e + 1], accb[i][t]);
            }
          }
        }
      }
    }

    // ---- delta: thread <-> column, words jj = dq, dq+4, ... of this slab ----
    if (dn < N) {
      for (int jj = dq; jj < kw / 32; jj += 4) {
        const int64_t j = (k0 >> 5) + jj;
#pragma unroll
        for (int t = 0; t < kTM; ++t) {
          if (t < nrows) {
            const int64_t tenant = (r0 + t) / a.m;
            const uint32_t word = (uint32_t)__ldg(a.masks + tenant * a.mask_tenant_stride + j * N + dn);
            const float* xv = &xs[t][jj * 32];
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float v = xv[i];  // warp-uniform address -> shared-memory broadcast
              s += ((word >> i) & 1u) ? v : -v;
            }
            accd[t] += s;
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- CTA reduction ----
  if (HAS_BASE) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int t = 0; t < kTM; ++t) {
        float v = accb[i][t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red_base[t][warp * 8 + i] = v;
      }
  }
#pragma unroll
  for (int t = 0; t < kTM; ++t) red_delta[dq][t][dcol] = accd[t];
  __syncthreads();

  // thread -> (row t, column c): 8 x 64 = 512 outputs, two per thread
  for (int idx = tid; idx < kTM * kBN; idx += kThreads) {
    const int t = idx / kBN, c = idx % kBN;
    const int64_t n = n0 + c;
    if (t >= nrows || n >= N) continue;
    const int64_t r = r0 + t;
    float d = red_delta[0][t][c] + red_delta[1][t][c] + red_delta[2][t][c] + red_delta[3][t][c];
    float v = d;
    if (HAS_BASE) v = red_base[t][c] + load_coeff(a.coeff, a.coeff_dtype, r / a.m) * d;
    if (a.splits == 1) {
      if (a.fp32_out) reinterpret_cast<float*>(a.y)[r * N + n] = v;
      else y[r * N + n] = F16<T>::from_f32(v);
    } else
      a.partial[((int64_t)split * a.rows + r) * N + n] = v;
  }
  if (a.splits == 1) return;

  // ---- deterministic split-K combine: last CTA of this (tile, row chunk) sums the slots in order ----
  __threadfence();
  __syncthreads();
  const unsigned tile_id = blockIdx.z * gridDim.x + blockIdx.x;
  if (tid == 0) {
    unsigned old = atomicAdd(&a.counters[tile_id], 1u);
    s_last = (old == (unsigned)a.splits - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int idx = tid; idx < kTM * kBN; idx += kThreads) {
    const int t = idx / kBN, c = idx % kBN;
    const int64_t n = n0 + c;
    if (t >= nrows || n >= N) continue;
    const int64_t r = r0 + t;
    float v = 0.f;
    for (int s = 0; s < a.splits; ++s) v += __ldcg(&a.partial[((int64_t)s * a.rows + r) * N + n]);
    if (a.fp32_out) reinterpret_cast<float*>(a.y)[r * N + n] = v;
    else y[r * N + n] = F16<T>::from_f32(v);
  }
  if (tid == 0) a.counters[tile_id] = 0u;  // leave the workspace clean for the next launch
}

struct SimtPlan {
  int splits, slabs_per_split;
  int64_t n_tiles, row_chunks;
  size_t partial_bytes, counter_bytes;
};

SimtPlan make_plan(int64_t rows, int64_t K, int64_t N) {
  SimtPlan p;
  p.n_tiles = (N + kBN - 1) / kBN;
  p.row_chunks = (rows + kTM - 1) / kTM;
  int64_t slabs = (K + kKS - 1) / kKS;
  int sms = 148;
  int64_t ctas = p.n_tiles * p.row_chunks;
  // aim for ~4 CTAs per SM worth of parallelism, but keep >= 2 slabs per split
  int64_t want = ctas >= 4 * (int64_t)sms ? 1 : (4 * (int64_t)sms + ctas - 1) / ctas;
  int64_t splits = want < 1 ? 1 : want;
  if (splits > slabs / 2) splits = slabs / 2;
  if (splits < 1) splits = 1;
  if (splits > kMaxSplits) splits = kMaxSplits;
  p.slabs_per_split = (int)((slabs + splits - 1) / splits);
  p.splits = (int)((slabs + p.slabs_per_split - 1) / p.slabs_per_split);
  p.partial_bytes = p.splits > 1 ? (size_t)p.splits * rows * N * sizeof(float) : 0;
  p.counter_bytes = (size_t)(p.n_tiles * p.row_chunks) * sizeof(unsigned);
  return p;
}
}  // namespace

size_t simt_workspace_bytes(int64_t rows, int64_t N) {
  // Workspace layout (bd_common.cuh): [0, 4 KiB) SIMT tile counters, [4 KiB, 8 KiB) tcgen05 tile counters -- both must
  // stay zero between launches -- then scratch for the split slots.  K is split only while the grid has fewer than
  // 4*148 CTAs (so at most 4*148 counters), and then
  // splits * rows * N <= (4*148/ctas + 1) * ctas * kTM * kBN <= 2 * 4*148 * kTM * kBN floats.
  (void)rows; (void)N;
  return kWsScratchOffset + (size_t)2 * 4 * 148 * kTM * kBN * sizeof(float);
}

int launch_fwd_simt(const FwdProblem& p) {
  const int64_t rows = p.T * p.m;
  SimtPlan plan = make_plan(rows, p.K, p.N);
  const size_t counters = kWsScratchOffset;
  if (plan.splits > 1) {
    if (!p.workspace || p.workspace_bytes < counters + plan.partial_bytes)
      return fail(BD_ERR_WORKSPACE, "simt forward needs %zu workspace bytes, got %zu", counters + plan.partial_bytes, p.workspace_bytes);
    if (plan.counter_bytes > kWsCounterBytes) return fail(BD_ERR_UNSUPPORTED, "simt forward: internal error, %zu counter bytes", plan.counter_bytes);
  }
  if (plan.row_chunks > 65535) return fail(BD_ERR_UNSUPPORTED, "simt forward: too many rows (%lld)", (long long)rows);
  SimtArgs a;
  a.x = p.x; a.w = p.w; a.masks = p.masks; a.coeff = p.coeff; a.coeff_dtype = p.coeff_dtype; a.y = p.y;
  a.counters = reinterpret_cast<unsigned*>(p.workspace);
  a.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(p.workspace) + counters);
  a.rows = rows; a.m = p.m; a.K = p.K; a.N = p.N; a.mask_tenant_stride = p.mask_tenant_stride;
  a.splits = plan.splits; a.slabs_per_split = plan.slabs_per_split;
  a.fp32_out = p.fp32_out ? 1 : 0;
  dim3 grid((unsigned)plan.n_tiles, (unsigned)plan.splits, (unsigned)plan.row_chunks);
  const bool has_base = p.w != nullptr;
  if (p.dtype == BD_BF16) {
    if (has_base) fwd_simt_kernel<__nv_bfloat16, true><<<grid, kThreads, 0, p.stream>>>(a);
    else fwd_simt_kernel<__nv_bfloat16, false><<<grid, kThreads, 0, p.stream>>>(a);
  } else {
    if (has_base) fwd_simt_kernel<__half, true><<<grid, kThreads, 0, p.stream>>>(a);
    else fwd_simt_kernel<__half, false><<<grid, kThreads, 0, p.stream>>>(a);
  }
  count_launch();
  return check_launch("fwd_simt_kernel");
}

}  // namespace bd
