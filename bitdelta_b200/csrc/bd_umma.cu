// tcgen05 / TMEM / TMA fused forward kernel for sm_100a.
//
//   y[t,i,n] = sum_k x[t,i,k] w[n,k]  +  coeff[t] * sum_k x[t,i,k] s_t[k,n]          s_t = 2*unpack(masks[t]) - 1
//
// replaces BinaryDiff.forward (bitdelta/diff.py:33-39), DiffCompressModule.forward (demo/demo_backend.py:93-98) and,
// without the base term, binary_bmm (bitdelta/binary_gemm_kernel.py:297-335).
//
// Design (swap-AB, weights on MMA-M = 128, tokens on MMA-N; DESIGN.md 3.1 has the measurements behind each choice):
//   * work unit = (128-row tile of N, 64-wide block of K); units are dealt to a persistent grid of one CTA per SM in
//     contiguous runs (stream-K), so every SM streams the same number of weight bytes whatever the layer shape.
//   * warps 0-7 (unpack / read-out): thread <-> weight row (= TMEM lane).  A thread reads its row's sign words from shared
//     memory (conflict-free), turns them into +-1.0 operand registers with one shift + one LOP3 per register and writes
//     them straight into TENSOR MEMORY with tcgen05.st: the unpacked sign tile is the A operand of a tcgen05.mma read from
//     TMEM and never touches shared memory.  Decode (one row per tenant): e4m3 signs, kind::f8f6f4, the two warps of a lane
//     quadrant take alternate units; otherwise +-1.0 in the activation dtype, kind::f16, the two warps split the tenants.
//   * warps 8-10 (activation permute): the K order inside a 32-group of the TMEM operand is fixed by the bit -> register
//     mapping, so these warps write the matching K-permuted copy of each tenant's activation rows (decode: scaled by the
//     row's power of two and split into three / four e5m2 pieces) as the B operand of the delta MMAs.
//   * warp 11 (TMA producer): per unit three loads into one pipeline stage: the bf16 W tile [128 x 64] (SWIZZLE_128B,
//     K-major, exactly how nn.Linear.weight sits in memory), the sign words [T x 2 x 128] int32 (the natural [K/32, N]
//     pack layout: one 512-byte run per (tenant, 32-K group)) and the activation block [rows x 64].
//   * warp 12 (sync): the only one of the unpack group that waits on mbarriers; publishes released units in a shared-memory
//     counter.  One barrier per stage carries both "stage landed" and "A buffer free".
//   * warp 13 (MMA issuer, one thread): per unit 4 MMAs  D_base += W_tile . X^T  (A and B from shared memory) and 2 (8-bit)
//     or 4 (16-bit) per tenant  D_delta[:, cols(t)] += S_t . X_t^T  (A from TMEM).  Two fp32 accumulators in TMEM so that
//     the per-tenant coefficient is applied exactly, in fp32, in the read-out (the reference's "TODO: Fuse coeff").
//   * read-out (warps 0-7 after the last unit of a tile run; on the decode path one unit late, see the unit loop):
//     tcgen05.ld, y = base + coeff * delta, one rounding.  A run that covers only part of K writes fp32 partials to a
//     per-CTA slot; the last CTA to finish a tile sums the slots in K order (deterministic) and stores y.
//
// HBM-bound by construction at decode sizes (algorithmic bytes per unit: 16 KiB of W + T KiB of signs); in steady state
// the kernel streams at the rate of a pure TMA copy, what is left is per launch (prologue, first operands, tail).
#include <cuda.h>
#include <cuda_fp8.h>

#include <mutex>
#include <type_traits>

#include "bd_common.cuh"

namespace bd {
namespace {

constexpr int kTileN = 128;   // weight rows per tile (MMA M)
constexpr int kBlockK = 64;   // K per unit (one 128-byte swizzle atom of bf16)
constexpr int kMaxABuf = 8;   // TMEM A-operand buffers (as many as fit: they hide the MMA completion latency)
constexpr int kUnpackWarps = 8;
constexpr int kXpermWarps = 3;  // one per SM sub-partition 0..2 (the issue slots of a sub-partition are the scarce resource)
constexpr int kThreads = 32 * (3 + kXpermWarps + kUnpackWarps);  // 8 unpack/epilogue, 3 activation-permute, sync, TMA producer, MMA issuer
// Warp roles.  The unpack warps come first so that warp % 4 is their TMEM lane quadrant.
constexpr int kScanWarps = kUnpackWarps + kXpermWarps;         // the warps that take part in the row-scale pass
// Helper roles on warps 8..13 = sub-partitions 0, 1, 2, 3, 0, 1.  Measured placements (T = 6 / T = 1 gate_proj, us):
// producer next to two unpack warps AND a permute warp (map 0) 33.3 / 27.3; producer alone with two unpack warps (map 2)
// 32.7 / 25.9 -- the slowest of a group's four unpack warps (one per sub-partition) paces the whole chain.
#ifndef BD_ROLE_MAP
#define BD_ROLE_MAP 2
#endif
#if BD_ROLE_MAP == 0
constexpr int kWarpXperm0 = 8, kWarpXperm1 = 9, kWarpXperm2 = 10;
constexpr int kWarpSync = 11, kWarpProducer = 12, kWarpMma = 13;
#elif BD_ROLE_MAP == 1
constexpr int kWarpXperm0 = 8, kWarpXperm1 = 9, kWarpXperm2 = 10;
constexpr int kWarpMma = 11, kWarpProducer = 12, kWarpSync = 13;
#elif BD_ROLE_MAP == 2
constexpr int kWarpXperm0 = 8, kWarpXperm1 = 9, kWarpXperm2 = 10;
constexpr int kWarpProducer = 11, kWarpSync = 12, kWarpMma = 13;
#elif BD_ROLE_MAP == 3  // MMA issuer and producer each alone with two unpack warps
constexpr int kWarpXperm0 = 8, kWarpXperm1 = 9, kWarpXperm2 = 13;
constexpr int kWarpMma = 10, kWarpProducer = 11, kWarpSync = 12;
#else                   // permute warps paired on sub-partition 0
constexpr int kWarpXperm0 = 8, kWarpXperm1 = 10, kWarpXperm2 = 12;
constexpr int kWarpMma = 9, kWarpProducer = 11, kWarpSync = 13;
#endif
// index of a permute warp among the permute warps, -1 for the other roles
__device__ __forceinline__ int xperm_index(int warp) {
  static_assert(kXpermWarps == 3, "three permute warps");
  return warp == kWarpXperm0 ? 0 : warp == kWarpXperm1 ? 1 : warp == kWarpXperm2 ? 2 : -1;
}
constexpr int kAFullThreads = (kUnpackWarps + kXpermWarps + 1) * 32;  // named barrier kBarAFull0+b: unpack + permute warps + MMA warp
constexpr int kMaxStages = 8;
constexpr int kMaxRows = 128;  // rows (tokens) per launch
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSpinLimit = 1u << 24;
// 8-bit delta path: every activation row is scaled by a power of two taken from the largest exponent of the row inside this
// CTA's K range, so that its three e5m2 pieces sit in e5m2's exponent window (see the row-scale pass and xperm_job)
constexpr bool kD8WideStore = false; // 8-bit path: tcgen05.st.x16 per tenant and unit (A/B against two .x8: see DESIGN.md 5b)
constexpr int kD8MaxTenants = 16;   // rows of s_rowexp (the TMEM budget allows 10)
constexpr int kBarRowScale = 12;    // named barrier of the row-scale pass (unpack + permute warps)
constexpr int kD8TopExp = 14;       // the row maximum is scaled to [2^14, 2^15): the first piece stays below e5m2's 57344

// ---------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU.  The loop must NOT be unrolled: the
// kernel has five roles' worth of code and the instruction cache is small (an unrolled copy per call site cost ~40 KB).
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
#pragma unroll 1
  while (!mbar_try_wait(bar, parity)) {
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    if (++spins > kSpinLimit) __trap();
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_commit_addr(uint32_t bar_smem) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_smem) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
  return __uint_as_float(r);
}

// kind::f16 instruction descriptor: fp32 accumulate, K-major A and B, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(int fmt /*0 = f16, 1 = bf16*/, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------- kernel
constexpr int kMaxSeg = 3;  // matrices sharing one activation block in one launch (q/k/v, gate/up)

struct UmmaMaps {
  CUtensorMap w[kMaxSeg], m[kMaxSeg], x;
};

struct UmmaArgs {
  const void* x;       // activations [rows, K] (the 8-bit path scans them for the row scales; everything else goes through TMA)
  const void* coeff;   // segment 0 (kept for the single-matrix case)
  int coeff_dtype;
  void* y;
  // grouped launch: segment s owns tiles [seg_tile0[s], seg_tile0[s+1]) and writes y_seg[s] ([rows, n_seg[s]])
  int nseg;
  int seg_tile0[kMaxSeg + 1];
  int n_seg[kMaxSeg];
  void* y_seg[kMaxSeg];
  const void* coeff_seg[kMaxSeg];
  float* partial;     // [grid][2][rows][128]
  unsigned* counters; // [n_tiles]
  int T, m, rows;     // tenants, rows per tenant, T*m
  uint32_t inv_m;     // ceil(65536 / m): row / m == (row * inv_m) >> 16 for row < 128 (no integer division in the kernel)
  int mp;             // rows per tenant padded to 16 (delta accumulator columns per tenant)
  int ntb;            // T*m padded to 16 (base accumulator columns)
  int K, N;
  int kblocks;        // ceil(K / 64)
  int n_tiles;
  int total_units, units_per_cta, units_rem;  // CTA c owns units quantum * [c*U + min(c,R), ...)
  int unit_quantum;   // 1 = stream-K over single units; kblocks = whole (tile, row chunk) runs per CTA (no split-K fix-up)
  int m_chunks;       // row chunks of a.m rows each (prefill-size launches of one tenant); tile id = n_tile * m_chunks + chunk
  int m_total;        // rows of the tenant over all chunks
  int tile_stride;    // 0: a CTA's tiles are consecutive; else (whole-tile launches of a multi-tenant prefill) CTA c works on
                      // tiles c, c + stride, c + 2*stride, ...: all CTAs sweep the tile list together, one tenant at a time
  int tpt;            // N tiles per tenant: tile column nt belongs to tenant nt / tpt (multi-tenant prefill: all tenants in one
                      // launch, tenant outermost so that one tenant's activations stay in L2 while its tiles are swept)
  int stages;
  int n_abuf;         // TMEM A-operand buffers in use (2..kMaxABuf)
  int a_cols_tenant;  // TMEM columns of one tenant's sign tile per unit: 32 (16-bit signs) or 16 (e4m3 signs)
  uint32_t stage_bytes, off_masks, off_x;     // stage layout: [W tile][masks][X tile]
  uint32_t off_xp, xp_buf_bytes;              // permuted activation tiles, one per A buffer
  uint32_t tx_bytes;
  int static_ops;    // BD_FLAG_STATIC_OPERANDS: weight / sign tiles may be requested before griddepcontrol.wait
  int fp32_out;      // BD_FLAG_FP32_OUT: y is fp32 (no rounding)
  int dbg;           // bring-up builds only (-DBD_BRINGUP): see dbg_flags() below; always 0 in the release library
  long long* trace;  // optional [64 units][16 slots] clock64 timestamps of CTA 0 (bring-up instrumentation)
};

// Bring-up knobs exist only in the -DBD_BRINGUP build (libbitdelta_b200_bringup.so, used by tools/): the release library
// compiles them to the constant 0 and has no trace instantiation, no mutable globals and no bd_debug_* entry points.
// bit 0 = stream the operands but skip unpack / MMA / epilogue, bit 2 = producer waits for the TMEM rendezvous,
// bit 3 = unpack warps split tenants (not units), bit 4 = unpack without tcgen05.st, bit 5 = no MMAs (commits only),
// bit 6 = no activation permute / split, bit 7 = no row-scale scan (rows assumed to peak in [1, 2)),
// bit 10 = sign stores without the conversion ALU work (wrong results; isolates the tcgen05.st path),
// bit 8 = half the sign work (only the first 32-K group of every unit is unpacked, stored and multiplied: wrong results, the
// cost profile of a 4-bit sign operand).
#if defined(BD_BRINGUP) && !defined(BD_NO_KNOBS)  // (-DBD_NO_KNOBS: the trace without the knobs = the release kernel, traced)
__device__ __forceinline__ int dbg_flags(const UmmaArgs& a) { return a.dbg; }
#else
__device__ __forceinline__ constexpr int dbg_flags(const UmmaArgs&) { return 0; }
#endif

template <bool TRACE>
__device__ __forceinline__ void trace_mark(const UmmaArgs& a, int it, int slot) {
  if (TRACE && a.trace != nullptr && blockIdx.x == 0 && it < 64) a.trace[it * 16 + slot] = clock64();
}
// Output store: the activation dtype (one rounding), or fp32 partial sums for tensor-parallel shards (BD_FLAG_FP32_OUT).
template <typename T16>
__device__ __forceinline__ void store_y(T16* y, int64_t idx, float v, int fp32) {
  if (fp32) reinterpret_cast<float*>(y)[idx] = v;
  else y[idx] = F16<T16>::from_f32(v);
}
__device__ __forceinline__ int cta_unit_begin(const UmmaArgs& a, int c) { return a.unit_quantum * (c * a.units_per_cta + min(c, a.units_rem)); }

// Hardware named barriers (ids 0..15): far cheaper than mbarriers.  id 0 = __syncthreads, 1 = split-K fix-up, 2 = accumulators
// read out, kBarAFull0 + b = "A buffer b is written" (unpack warps + permute warp arrive, the MMA warp syncs).
constexpr int kBarAFull0 = 4;
constexpr int kBarDEmpty = 2;     // unpack warps arrive after reading a run's accumulators, the MMA warp syncs before the next run
constexpr int kBarTmemReady = 3;  // every warp but the TMA producer: tensor memory allocated, activation tiles zeroed
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__device__ __forceinline__ void st_release_shared(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_shared(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
// Wait until the sync warp has released unit `it` (bounded spin on a shared-memory counter: ~30 cycles when it is
// already released, which is the common case -- no rendezvous of the whole group per unit).
#ifndef BD_STATIC_PREFETCH
#define BD_STATIC_PREFETCH 4
#endif
constexpr int kStaticPrefetch = BD_STATIC_PREFETCH;  // stages whose static operands are requested ahead of griddepcontrol.wait
#ifndef BD_XPERM_SLEEP
#define BD_XPERM_SLEEP 0
#endif
#ifndef BD_PROD_SLEEP
#define BD_PROD_SLEEP 0
#endif
#ifndef BD_UNPACK_SLEEP
#define BD_UNPACK_SLEEP 0
#endif
template <int SLEEP_NS = 0>
__device__ __forceinline__ void wait_released(const int* counter, int it) {
  uint32_t spins = 0;
#pragma unroll 1
  while (ld_acquire_shared(counter) <= it) {
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    if (++spins > kSpinLimit) __trap();
  }
}

// Ring-buffer cursor: index + phase bit, advanced without integer division.
struct Ring {
  int idx = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++idx == n) { idx = 0; phase ^= 1u; }
  }
};

// Position in the unit space: K block kb of row chunk mc of N tile nt (tile id = nt * m_chunks + mc).
struct UnitCursor {
  int nt, mc, kb;
  int tt;  // tenant of tile column nt (0 unless it is a multi-tenant prefill launch): tracked without a division per unit
  __device__ __forceinline__ void next(int kblocks, int m_chunks, int tpt, int tile_stride) {
    if (++kb == kblocks) {
      kb = 0;
      if (tile_stride) {  // strided whole-tile schedule: jump to tile id + stride (one division per tile)
        const int tile = nt * m_chunks + mc + tile_stride;
        nt = tile / m_chunks;
        mc = tile - nt * m_chunks;
        tt = nt / tpt;
      } else if (++mc == m_chunks) {
        mc = 0;
        if (++nt - tt * tpt == tpt) ++tt;
      }
    }
  }
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// The 64-bit shared-memory descriptors differ only in their low word (start address >> 4); the high word is constant:
// SBO = 1024 B (>> 4 = 64) at bits [32,46), version 1 at bit 46, SWIZZLE_128B (2) at bits [61,64).
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ void mma_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(kDescHi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 8-bit delta path: A = e4m3 signs in TMEM, B = e5m2 activation pieces in shared memory, no swizzle (K-major
// "interleave" layout: 8-row x 16-byte core matrices, LBO = 128 B between K-adjacent cores, SBO = 512 B between 8-row
// groups of a 64-byte-wide tile).  LBO sits in the low descriptor word.
constexpr uint32_t kDesc8LoLbo = (128u >> 4) << 16;
constexpr uint32_t kDesc8Hi = (512u >> 4) | (1u << 14);
__device__ __forceinline__ void mma_ts8_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo | kDesc8LoLbo), "r"(kDesc8Hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the LBO field already OR-ed into b_lo (per-tenant offsets are added to one pre-built low word: the address field
// is 14 bits, offsets never carry into the LBO field at bit 16).
__device__ __forceinline__ void mma_ts8_raw(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo_lbo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], db, %4, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo_lbo), "r"(kDesc8Hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The 4 base + 2 * TT delta MMAs of one unit on the 8-bit path, for a COMPILE-TIME number of tenants: every operand is the
// unit's base value plus a constant, so the one issuing thread -- the most loaded single role of the pipeline -- runs a
// branch-free sequence (the generic loop over a run-time tenant count costs ~35 cycles per MMA, this one ~17).
template <int TT, bool HAS_BASE>
__device__ __forceinline__ void issue_unit_d8(uint32_t d_base, uint32_t d_delta, uint32_t w_lo, uint32_t x_lo, uint32_t a_tmem0,
                                              uint32_t bl0, uint32_t idesc_base, uint32_t idesc_delta, uint32_t acc0) {
#pragma unroll
  for (int ks = 0; ks < kBlockK / 16; ++ks) {
    if (HAS_BASE) mma_ss_lo(d_base, w_lo + ks * 2, x_lo + ks * 2, idesc_base, ks == 0 ? acc0 : 1u);
    if ((ks & 1) == 0) {
      const int k8 = ks >> 1;
#pragma unroll
      for (int t = 0; t < TT; ++t)
        mma_ts8_raw(d_delta + t * 8, a_tmem0 + k8 * 8 + t * 16, bl0 + k8 * (256u >> 4) + t * (512u >> 4), idesc_delta, k8 == 0 ? acc0 : 1u);
    }
  }
}
// kind::f8f6f4 instruction descriptor: A = e4m3 (0), B = e5m2 (1), fp32 accumulate, K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc8(int n) {
  return (1u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileN >> 4) << 24);
}

// NATK (one tenant on the 16-bit path): the sign tile is unpacked in NATURAL K order (register i = elements 2i, 2i+1), so the
// delta MMAs read the very same TMA-loaded activation tile as the base MMAs and no K-permuted copy is built at all.  It
// costs 4 ALU ops per register instead of 2, which is nothing for a single tenant -- and prefill is tensor-bound.
template <typename T16, bool HAS_BASE, bool DELTA8, bool NATK, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
fwd_umma_kernel(const __grid_constant__ UmmaMaps maps, const UmmaArgs a) {
  const CUtensorMap& tmap_x = maps.x;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles and their UMMA descriptors need 1024-byte alignment: align by hand (the host adds 1 KiB of slack)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar_full[kMaxStages], bar_empty[kMaxStages], bar_dfull;
  __shared__ uint32_t tmem_base_slot;
  __shared__ unsigned s_is_last;
  __shared__ int s_released;  // number of units the sync warp has released to the unpack group (release/acquire)
  __shared__ int s_rowexp[kD8MaxTenants];  // 8-bit path: largest biased bf16 exponent of tenant t's row over this CTA's K range

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int xi = xperm_index(warp);  // 0..kXpermWarps-1 for the activation-permute warps, else -1
  const int cta = blockIdx.x;
  if (TRACE && a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { a.trace[63 * 16 + 0] = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a.trace[63 * 16 + 8])); }
  if (TRACE && a.trace != nullptr && threadIdx.x == 0) {  // per-CTA lifetime: [1024 + 4*cta] = entry ns, exit ns, SM id
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a.trace[1024 + 4 * blockIdx.x]));
    a.trace[1024 + 4 * blockIdx.x + 2] = smid;
  }
  const int u_begin = cta_unit_begin(a, cta), u_end = cta_unit_begin(a, cta + 1);
  // first tile: where the CTA's contiguous unit range starts -- or, with the strided whole-tile schedule, tile `cta`
  const int tile0 = a.tile_stride ? cta : u_begin / a.kblocks;
  const int kb0 = a.tile_stride ? 0 : u_begin - tile0 * a.kblocks;
  const int nt0 = tile0 / a.m_chunks, mc0 = tile0 - nt0 * a.m_chunks;       // n tile and row chunk of the first tile
  // The activation permutation is done by the dedicated warp alone when it is small (decode), otherwise shared with
  // the unpack warps.
  const int xjobs = DELTA8 ? a.rows * 16 : a.rows * 8;
  const bool xperm_shared = xjobs > kXpermWarps * 32 * 2;
  // Decode (8-bit delta path, small row counts): the two unpack warps of a TMEM lane quadrant take ALTERNATE units (all
  // tenants of a unit each) instead of splitting the tenants of every unit.  tcgen05.st drains at ~16 B/clk per quadrant
  // (measured: 24 KB per quadrant and round in ~1640 cycles), which at 6 tenants is most of the per-unit HBM time; with
  // alternating units one warp's stores drain while the other warp is in its per-unit synchronisation (release check,
  // fences, wait::st, barrier arrive, loop), instead of both warps paying that with the store port idle.
  const bool alt_units = DELTA8 && !xperm_shared && !(dbg_flags(a) & 8);
  const int afull_threads = alt_units ? (kUnpackWarps / 2 + kXpermWarps + 1) * 32 : kAFullThreads;

  // Programmatic dependent launch: let the next kernel of the stream be scheduled as soon as SMs free up.  Its CTAs run
  // their prologue and prefetch their first weight / sign tiles (static data) while this grid drains; only its
  // activation loads wait for this grid to complete (griddepcontrol.wait in the producer).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // ---- one-time setup ----
  if (warp == kWarpProducer) {
    // the producer warp's lanes initialise the mbarriers in parallel (one thread doing all ~21 of them cost ~0.3 us of every
    // launch's prologue): lanes 0..7 the stage barriers, 8..15 the A-buffer barriers, 16 the accumulator barrier; lane 31
    // prefetches the tensor maps
    if (lane < kMaxStages) {
      if (lane < a.stages) {
        // "unit u may be unpacked" = its stage has landed AND the A buffer it will use is free.  Both events arrive on the
        // SAME barrier: the producer's arrive.expect_tx (+ the TMA bytes) and the MMA commit of unit u - n_abuf, which
        // targets the stage of unit u (n_abuf <= stages, so that stage's previous phase is long complete).  One mbarrier
        // wait per unit instead of two: mbarrier operations cost the sync warp ~300 cycles each and it paces the kernel.
        mbar_init(&bar_full[lane], 2);
        // A stage is released by the MMA commit alone: the MMAs of a unit are issued only after every unpack warp and
        // the permute warp have arrived on bar_afull, i.e. after they are done reading the stage.
        mbar_init(&bar_empty[lane], 1);
      }
    } else if (lane == kMaxStages + kMaxABuf) {
      mbar_init(&bar_dfull, 1);
    } else if (lane == 31) {
      for (int sg = 0; sg < a.nseg; ++sg) {
        if (HAS_BASE) prefetch_tmap(&maps.w[sg]);
        prefetch_tmap(&maps.m[sg]);
      }
      prefetch_tmap(&tmap_x);
    }
    fence_barrier_init();
    __syncwarp();  // lane 0 (the TMA producer) uses barriers its sibling lanes initialised
    if (lane < a.n_abuf) mbar_arrive(&bar_full[lane]);  // the first n_abuf units find their A buffer free (n_abuf <= stages)
  }
  if (threadIdx.x == 0) s_released = 0;
  if (DELTA8 && threadIdx.x < kD8MaxTenants) s_rowexp[threadIdx.x] = 0;
  // ONE rendezvous for the whole prologue: the producer warp (which initialised the mbarriers above) only ARRIVES and starts
  // requesting the first weight / sign stages right away; the other warps meanwhile allocate tensor memory and zero the
  // permuted-activation tiles, then wait for each other and for the producer's arrival (= mbarriers initialised).
  if (warp == kWarpProducer && !(dbg_flags(a) & 4)) {
    named_bar_arrive(kBarTmemReady, kThreads);
  } else {
    if (warp == kWarpMma) tmem_alloc(&tmem_base_slot, kTmemCols);
    // zero the permuted-activation tiles once: rows >= m of every tenant tile stay zero for the whole kernel
    const bool all = (dbg_flags(a) & 4) != 0;
    const uint32_t part = all ? threadIdx.x : threadIdx.x - (warp > kWarpProducer ? 32u : 0u);
    const uint32_t nparts = all ? kThreads : kThreads - 32;
    for (uint32_t i = part * 16; i < a.n_abuf * a.xp_buf_bytes; i += nparts * 16)
      *reinterpret_cast<uint4*>(smem + a.off_xp + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
    tc_fence_before();
    named_bar_sync(kBarTmemReady, kThreads);
    tc_fence_after();
  }
  if (TRACE && a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.trace[63 * 16 + 1] = clock64();
  const uint32_t tmem_base = (warp != kWarpProducer) ? tmem_base_slot : 0u;
  // TMEM columns: [0, ntb) base accumulator, [ntb, ntb + T*mp) delta accumulator, then the A-operand buffers
  const uint32_t col_dbase = 0, col_ddelta = a.ntb;
  const uint32_t a_cols_per_buf = (uint32_t)a.T * a.a_cols_tenant;
  const uint32_t col_abuf0 = kTmemCols - a.n_abuf * a_cols_per_buf;

  // ---- 8-bit path: row scales ----
  // e5m2 has fp16's 5-bit exponent, bf16 has 8: before an activation is split into e5m2 pieces it is multiplied by
  // 2^(kD8TopExp - emax_t), emax_t = the largest exponent of tenant t's row over the K blocks THIS CTA consumes (its unit run;
  // the scale only has to be the same for everything one CTA accumulates -- partial sums leave the CTA descaled, in fp32).
  // The row maximum then sits in [2^14, 2^15); three pieces (3 + 3 + 2 significant bits) hold every element down to
  // 2^-23 of the row maximum exactly, and smaller ones to an absolute error below 2^-31 of the row maximum -- the order of
  // the rounding error of the fp32 accumulation itself, whatever the magnitude or dynamic range of the row.  inf / NaN
  // elements become NaN pieces (the row's outputs are NaN, like the reference's).
  // The unpack and permute warps scan the rows (from L2: T x <= K values per CTA) while the first stages are in flight.
  if constexpr (DELTA8) {
    if (dbg_flags(a) & 128) {  // bring-up A/B: no scan, rows assumed to peak in [1, 2)
      if (threadIdx.x < kD8MaxTenants) s_rowexp[threadIdx.x] = 127;
      if (warp < kUnpackWarps || xi >= 0) named_bar_sync(kBarRowScale, kScanWarps * 32);  // (orders the stores above)
    } else
    if (warp < kUnpackWarps || xi >= 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");  // the activations are produced by the previous kernel of the stream
      constexpr int kScanThreads = kScanWarps * 32;
      constexpr bool kIsBf16 = std::is_same<T16, __nv_bfloat16>::value;
      constexpr uint32_t kExpMask = kIsBf16 ? 0x7F807F80u : 0x7C007C00u;  // exponent fields of a packed pair
      constexpr int kExpShift = kIsBf16 ? 7 : 10;
      constexpr int kExpRebias = kIsBf16 ? 0 : 127 - 15;                  // s_rowexp holds fp32-biased exponents
      const int nkb = min(u_end - u_begin, a.kblocks);    // K blocks kb0, kb0 + 1, ... (mod kblocks) of this CTA's run
      const int chunks = nkb * (kBlockK / 8);             // 16-byte chunks (8 values) per row
      const T16* xg = reinterpret_cast<const T16*>(a.x);
      uint32_t mx[10];
#pragma unroll
      for (int t = 0; t < 10; ++t) mx[t] = 0u;
      for (int ci = (xi >= 0 ? (kUnpackWarps + xi) * 32 + lane : (int)threadIdx.x); ci < chunks; ci += kScanThreads) {
        int kb = kb0 + (ci >> 3);
        if (kb >= a.kblocks) kb -= a.kblocks;
        const int k = kb * kBlockK + (ci & 7) * 8;
        if (k < a.K) {  // K % 8 == 0: a chunk is inside the row or outside it
          // all tenants' loads first (independent L2 requests in flight), then the comparisons
          // (unconditional: rows past the last tenant re-read the last tenant's chunk, an L2 hit, instead of predicating
          // the loads -- predicated loads ended up serialised through one destination register)
          uint4 v[10];
#pragma unroll
          for (int t = 0; t < 10; ++t) v[t] = __ldcg(reinterpret_cast<const uint4*>(xg + (size_t)min(t, a.T - 1) * a.K + k));
#pragma unroll
          for (int t = 0; t < 10; ++t) {
            // exponent fields of the two bf16 halves of a word, compared as packed unsigned 16-bit values
            mx[t] = __vmaxu2(__vmaxu2(mx[t], v[t].x & kExpMask), __vmaxu2(v[t].y & kExpMask, __vmaxu2(v[t].z & kExpMask, v[t].w & kExpMask)));
          }
        }
      }
      // warp maxima, tenant t's into lane t, then ONE shared-memory atomic per warp (one lane per tenant)
      uint32_t mine = 0u;
#pragma unroll
      for (int t = 0; t < 10; ++t) {
        if (t < a.T) {
          const uint32_t e = max(mx[t] & 0xFFFFu, mx[t] >> 16) >> kExpShift;
          const uint32_t wmax = __reduce_max_sync(0xffffffffu, e);
          if (lane == t) mine = wmax;
        }
      }
      // (an fp16 subnormal row maximum, field 0, is below 2^-14: field 1's exponent is a valid upper bound; a bf16
      // subnormal, field 0, takes the clamped largest scale, which cannot overflow it)
      if (lane < a.T) atomicMax(&s_rowexp[lane], (int)(kIsBf16 ? mine : max(mine, 1u)) + kExpRebias);
      // Only the permute warps need the scales right away; the unpack warps read them in the epilogue (ordered behind this
      // barrier through the MMA completion they wait for), so they arrive without waiting and go on to their first unit.
      if (xi >= 0) named_bar_sync(kBarRowScale, kScanThreads);
      else named_bar_arrive(kBarRowScale, kScanThreads);
    }
  }
  // scale applied to tenant t's activations before the split, as the exponent field of a float: 2^(kD8TopExp - (emax - 127)),
  // clamped to a normal number whose reciprocal is normal too
  auto d8_scale_field = [&](int t) -> uint32_t { return (uint32_t)min(max(254 + kD8TopExp - s_rowexp[t], 1), 253); };

  // K-permuted copy of the activation rows for A buffer `b` (see the header comment): job = (row r, 16-byte output
  // chunk c of the 64-K block); out chunk c of a 32-group = x[4c..4c+3] interleaved with x[4c+16..4c+19]; source and
  // destination tiles use the 128-byte swizzle (chunk index XOR row % 8).
  // pieces per activation on the 8-bit path: 3 + 3 + 2 significant bits cover bf16's 8, 3 + 3 + 3 + 2 cover fp16's 11
  constexpr int kD8Pieces = std::is_same<T16, __nv_bfloat16>::value ? 3 : 4;
  auto xperm_job = [&](const uint8_t* xsrc, uint8_t* xp, int job, uint32_t scale_field) {
    if constexpr (DELTA8) {
      // 8-bit delta path (one row per tenant).  job = (row r, 32-group g, c): one 32-bit output word per B-operand row,
      // holding K slots 4c..4c+3 of the group = activations k = c + 8q, q = 0..3 (the order the e4m3 sign registers are
      // built in).  Jobs are kept this small on purpose: a warp executes a job's instructions once however many lanes
      // are active, and the issue slots of the sub-partition this warp shares with two unpack warps are the scarce resource.
      //
      // The activation, scaled by the row's power of two (see the row-scale pass), is split into three e5m2 pieces
      // p1 + p2 + p3 by a round-to-nearest residual chain (3 + 3 + 2 significant bits cover bf16's 8); each piece is its own
      // B-operand row (3 of the tenant's 8 accumulator columns) and the epilogue adds the three partial sums and undoes
      // the scale in fp32.
      const int r = job >> 4, g = (job >> 3) & 1, c = job & 7;
      const uint8_t* src = xsrc + r * 128 + 2 * (c & 7);
      const float scale = __uint_as_float(scale_field << 23);
      float f[4];
      uint32_t bad = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // element 32g + c + 8q lives in 16-byte chunk 4g + q (swizzled by the row)
        const uint32_t bits = *reinterpret_cast<const unsigned short*>(src + (((4 * g + q) ^ (r & 7)) << 4));
        if constexpr (std::is_same<T16, __nv_bfloat16>::value) {
          bad |= ((bits & 0x7F80u) == 0x7F80u ? 0xFFu : 0u) << (8 * q);  // inf / NaN
          f[q] = __uint_as_float(bits << 16) * scale;
        } else {
          bad |= ((bits & 0x7C00u) == 0x7C00u ? 0xFFu : 0u) << (8 * q);
          f[q] = __half2float(__ushort_as_half((unsigned short)bits)) * scale;
        }
      }
      uint32_t pw[kD8Pieces];
#pragma unroll
      for (int piece = 0; piece < kD8Pieces; ++piece) {
        const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(f[0], f[1]), __NV_SATFINITE, __NV_E5M2);
        const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(f[2], f[3]), __NV_SATFINITE, __NV_E5M2);
        pw[piece] = lo | (hi << 16);
        if (piece < kD8Pieces - 1) {  // residuals: e5m2 is the top byte of fp16
          const uint32_t h01 = __byte_perm(lo, 0, 0x1404), h23 = __byte_perm(hi, 0, 0x1404);
          const float2 b01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
          const float2 b23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
          f[0] -= b01.x; f[1] -= b01.y; f[2] -= b23.x; f[3] -= b23.y;
        }
      }
      pw[0] = (pw[0] & ~bad) | (0x7E7E7E7Eu & bad);  // non-finite activation: e5m2 NaN
      // tenant tile: [8 rows x 64 B] as four 8-row x 16-byte core matrices; word c of the group sits at byte 32g + 4c
      uint8_t* tile = xp + r * 512 + (2 * g + (c >> 2)) * 128 + (c & 3) * 4;
#pragma unroll
      for (int piece = 0; piece < kD8Pieces; ++piece) *reinterpret_cast<uint32_t*>(tile + piece * 16) = pw[piece];
      return;
    }
    const int r = job >> 3, c = job & 7;
    const int g = c >> 2, cc = c & 3;
    const int ca = 4 * g + (cc >> 1), cb = ca + 2;
    const uint8_t* src = xsrc + r * 128;
    const uint2 va = *reinterpret_cast<const uint2*>(src + ((ca ^ (r & 7)) << 4) + ((cc & 1) << 3));
    const uint2 vb = *reinterpret_cast<const uint2*>(src + ((cb ^ (r & 7)) << 4) + ((cc & 1) << 3));
    uint4 o;
    o.x = __byte_perm(va.x, vb.x, 0x5410);
    o.y = __byte_perm(va.x, vb.x, 0x7632);
    o.z = __byte_perm(va.y, vb.y, 0x5410);
    o.w = __byte_perm(va.y, vb.y, 0x7632);
    const int t = (int)(((uint32_t)r * a.inv_m) >> 16), i = r - t * a.m;
    *reinterpret_cast<uint4*>(xp + (t * a.mp + i) * 128 + ((c ^ (i & 7)) << 4)) = o;
  };

  if (warp == kWarpProducer) {
    // ===================================================== TMA producer
    if (lane == 0) {
      Ring st;
      UnitCursor cur{nt0, mc0, kb0, nt0 / a.tpt};
      // Prologue under PDL: when the caller declares the weight and sign tiles static (BD_FLAG_STATIC_OPERANDS: long-lived
      // module buffers) those of the first `stages` units are requested right away; the activation tiles are produced by
      // the previous kernel of the stream, so those loads are issued only after griddepcontrol.wait.  Each stage's barrier
      // expects all three loads.
      {
        // The TMA unit works through its requests in order, and the first MMA needs the (tiny) activation tile of stage 0:
        // queued behind the weight and sign tiles of all eight stages it arrived 2.2 us after its request.  So only the
        // first kStaticPrefetch stages are requested ahead of griddepcontrol.wait, and after it every stage's activation
        // tile goes first.
        const int npre = min(a.stages, u_end - u_begin);
        const int nstat = a.static_ops ? min(npre, kStaticPrefetch) : 0;
        UnitCursor c2 = cur;
        auto load_w_masks = [&](int i) {
          uint8_t* sp = smem + (size_t)i * a.stage_bytes;
          const int sg = (c2.nt >= a.seg_tile0[1]) + (c2.nt >= a.seg_tile0[2]);
          const int lt = c2.nt - a.seg_tile0[sg] - c2.tt * a.tpt;
          tma_load_3d(sp + a.off_masks, &maps.m[sg], &bar_full[i], lt * kTileN, c2.kb * (kBlockK / 32), c2.tt, kEvictFirst);
          if (HAS_BASE) tma_load_2d(sp, &maps.w[sg], &bar_full[i], c2.kb * kBlockK, lt * kTileN, kEvictFirst);
          c2.next(a.kblocks, a.m_chunks, a.tpt, a.tile_stride);
        };
        for (int i = 0; i < nstat; ++i) {
          mbar_arrive_expect_tx(&bar_full[i], a.tx_bytes);
          load_w_masks(i);
        }
        // Operands the caller did not declare static may have been written by the preceding kernel: everything waits.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        for (int i = 0; i < npre; ++i) {
          trace_mark<TRACE>(a, i, 8);
          if (i >= nstat) mbar_arrive_expect_tx(&bar_full[i], a.tx_bytes);
          tma_load_2d(smem + (size_t)i * a.stage_bytes + a.off_x, &tmap_x, &bar_full[i], cur.kb * kBlockK, cur.tt * a.m_total + cur.mc * a.m, kEvictLast);
          if (i >= nstat) load_w_masks(i);
          st.advance(a.stages);
          cur.next(a.kblocks, a.m_chunks, a.tpt, a.tile_stride);
        }
      }
#pragma unroll 1
      for (int u = u_begin + min(a.stages, u_end - u_begin); u < u_end; ++u) {
        mbar_wait<BD_PROD_SLEEP>(&bar_empty[st.idx], st.phase ^ 1u);
        trace_mark<TRACE>(a, u - u_begin, 8);
        uint8_t* sp = smem + (size_t)st.idx * a.stage_bytes;
        mbar_arrive_expect_tx(&bar_full[st.idx], a.tx_bytes);
        const int sg = (cur.nt >= a.seg_tile0[1]) + (cur.nt >= a.seg_tile0[2]);
        const int lt = cur.nt - a.seg_tile0[sg] - cur.tt * a.tpt;
        // with row chunks the same W / sign tile is re-read for every chunk of the tile: keep it in L2 (evict-last)
        const uint64_t whint = a.m_chunks > 1 ? kEvictLast : kEvictFirst;
        if (HAS_BASE) tma_load_2d(sp, &maps.w[sg], &bar_full[st.idx], cur.kb * kBlockK, lt * kTileN, whint);
        tma_load_3d(sp + a.off_masks, &maps.m[sg], &bar_full[st.idx], lt * kTileN, cur.kb * (kBlockK / 32), cur.tt, whint);
        tma_load_2d(sp + a.off_x, &tmap_x, &bar_full[st.idx], cur.kb * kBlockK, cur.tt * a.m_total + cur.mc * a.m, kEvictLast);
        st.advance(a.stages);
        cur.next(a.kblocks, a.m_chunks, a.tpt, a.tile_stride);
      }
    }
  } else if (warp == kWarpMma) {
    // ===================================================== MMA issuer
    constexpr int fmt = std::is_same<T16, __nv_bfloat16>::value ? 1 : 0;
    const uint32_t idesc_base = make_idesc(fmt, a.ntb);
    const uint32_t idesc_delta = DELTA8 ? make_idesc8(a.mp) : make_idesc(fmt, a.mp);
    const bool leader = elect_one();
    // descriptor low words (address >> 4) of stage 0 / A buffer 0 and their strides
    const uint32_t w_lo0 = (smem_u32(smem) & 0x3FFFFu) >> 4, stage_lo = a.stage_bytes >> 4;
    const uint32_t x_off_lo = a.off_x >> 4;
    const uint32_t xp_lo0 = (smem_u32(smem + a.off_xp) & 0x3FFFFu) >> 4, xp_buf_lo = a.xp_buf_bytes >> 4, xp_t_lo = DELTA8 ? (512u >> 4) : (uint32_t)a.mp * 8;
    // 8-bit path, common tenant counts: the whole loop is specialised on the tenant count and keeps everything a unit needs
    // (descriptor bases, barrier addresses, ring positions) in registers, stepped incrementally.  With the unpack warps out
    // of the way this ONE thread paces the kernel -- per unit its MMAs occupy the tensor pipe for ~37 cycles each and
    // everything else it executes between two units (barrier, commits, loop) adds on top, because the pipe's issue queue
    // is only a few MMAs deep.
    bool fast_done = false;
    if constexpr (DELTA8) {
      auto fast_loop = [&](auto ttc) {
        constexpr int TT = decltype(ttc)::value;
        uint32_t n_stages = a.stages, n_abuf = a.n_abuf, kblocks = a.kblocks, aft = afull_threads;
        asm volatile("" : "+r"(n_stages), "+r"(n_abuf), "+r"(kblocks), "+r"(aft));  // loop invariants: registers, not constant-bank reloads
        const uint32_t empty0 = smem_u32(&bar_empty[0]), full0 = smem_u32(&bar_full[0]), dfull_addr = smem_u32(&bar_dfull);
        const uint32_t d_base = tmem_base + col_dbase, d_delta = tmem_base + col_ddelta, a_tmem00 = tmem_base + col_abuf0;
        const uint32_t bl00 = xp_lo0 | kDesc8LoLbo;
        uint32_t sti = 0, abi = 0, kb = kb0;
        uint32_t w_lo = w_lo0, bl0 = bl00, a_tmem0 = a_tmem00;
        uint32_t ati = n_abuf == n_stages ? 0u : n_abuf;  // stage of unit u + n_abuf, the next user of this unit's A buffer
        uint32_t bar_e = empty0, bar_a = full0 + 8 * ati;
        int left = u_end - u_begin;
        bool first = true;
#pragma unroll 1
        for (; left > 0; --left) {
          const bool seg_first = first || kb == 0;
          const bool seg_last = left == 1 || kb + 1 == kblocks;
          if (lane == 0) trace_mark<TRACE>(a, (u_end - u_begin) - left, 5);
          named_bar_sync(kBarAFull0 + abi, aft);
          if (lane == 0) trace_mark<TRACE>(a, (u_end - u_begin) - left, 2);
          if (seg_first && !first) named_bar_sync(kBarDEmpty, kUnpackWarps * 32 + 32);  // previous run's accumulators read out
          tc_fence_after();
          if (leader) {
            trace_mark<TRACE>(a, (u_end - u_begin) - left, 6);
            issue_unit_d8<TT, HAS_BASE>(d_base, d_delta, w_lo, w_lo + x_off_lo, a_tmem0, bl0, idesc_base, idesc_delta, seg_first ? 0u : 1u);
            trace_mark<TRACE>(a, (u_end - u_begin) - left, 9);
            tc_commit_addr(bar_e);   // stage (W tile, masks, X tile) may be overwritten once these MMAs retire
            tc_commit_addr(bar_a);   // so may the TMEM A buffer and its permuted-X tiles: second arrival of unit u + n_abuf's barrier
            if (seg_last) tc_commit_addr(dfull_addr);
            trace_mark<TRACE>(a, (u_end - u_begin) - left, 7);
          }
          __syncwarp();
          if (lane == 0) trace_mark<TRACE>(a, (u_end - u_begin) - left, 12);
          first = false;
          if (++sti == n_stages) { sti = 0; w_lo = w_lo0; bar_e = empty0; } else { w_lo += stage_lo; bar_e += 8; }
          if (++abi == n_abuf) { abi = 0; bl0 = bl00; a_tmem0 = a_tmem00; } else { bl0 += xp_buf_lo; a_tmem0 += a_cols_per_buf; }
          if (++ati == n_stages) { ati = 0; bar_a = full0; } else { bar_a += 8; }
          if (++kb == kblocks) kb = 0;
        }
      };
      if (dbg_flags(a) == 0) {
        fast_done = true;
        switch (a.T) {
          case 1: fast_loop(std::integral_constant<int, 1>{}); break;
          case 2: fast_loop(std::integral_constant<int, 2>{}); break;
          case 3: fast_loop(std::integral_constant<int, 3>{}); break;
          case 4: fast_loop(std::integral_constant<int, 4>{}); break;
          case 6: fast_loop(std::integral_constant<int, 6>{}); break;
          case 8: fast_loop(std::integral_constant<int, 8>{}); break;
          default: fast_done = false; break;  // other tenant counts: the generic loop below
        }
      }
    }
    Ring st, ab;
    int kb = kb0;
#pragma unroll 1
    for (int u = fast_done ? u_end : u_begin; u < u_end; ++u) {
      const bool seg_first = (u == u_begin) || (kb == 0);
      const bool seg_last = (u + 1 == u_end) || (kb + 1 == a.kblocks);
      if (lane == 0) trace_mark<TRACE>(a, u - u_begin, 5);
      // "A buffer written" is a hardware named barrier (8 unpack warps + permute warp arrive, this warp syncs).  It also
      // implies that the stage has landed: those warps only get there after the stage's mbarrier completed.  mbarrier
      // operations are slow and serialised per SM, so every role touches as few of them as it can.
      named_bar_sync(kBarAFull0 + ab.idx, afull_threads);
      if (seg_first && u != u_begin) named_bar_sync(kBarDEmpty, kUnpackWarps * 32 + 32);  // previous run's accumulators read out
      tc_fence_after();
      Ring st_next = st;
      st_next.advance(a.stages);
      if (leader) {
        trace_mark<TRACE>(a, u - u_begin, 6);
        if (!(dbg_flags(a) & (1 | 32))) {
        const uint32_t w_lo = w_lo0 + st.idx * stage_lo;
        const uint32_t x_lo = w_lo + x_off_lo;
        const uint32_t xp_lo = xp_lo0 + ab.idx * xp_buf_lo;
        const uint32_t a_tmem0 = tmem_base + col_abuf0 + ab.idx * a_cols_per_buf;
        const uint32_t d_base = tmem_base + col_dbase, d_delta = tmem_base + col_ddelta;
        // k-step outermost: consecutive MMAs go to different accumulators (base, tenant 0, tenant 1, ...)
#pragma unroll
        for (int ks = 0; ks < kBlockK / 16; ++ks) {
          const uint32_t acc = (seg_first && ks == 0) ? 0u : 1u;
          if (HAS_BASE) mma_ss_lo(d_base, w_lo + ks * 2, x_lo + ks * 2, idesc_base, acc);
          if constexpr (DELTA8) {
            if ((ks & 1) == 0 && !((dbg_flags(a) & 256) && ks != 0)) {  // K = 32 per 8-bit MMA: two per tenant per unit
              const int k8 = ks >> 1;
              const uint32_t acc8 = (seg_first && k8 == 0) ? 0u : 1u;
              uint32_t d = d_delta, at = a_tmem0 + k8 * 8, bl = xp_lo + k8 * (256u >> 4);
#pragma unroll 2
              for (int t = 0; t < a.T; ++t) {
                mma_ts8_lo(d, at, bl, idesc_delta, acc8);
                d += a.mp; at += 16; bl += xp_t_lo;
              }
            }
          } else {
            uint32_t d = d_delta, at = a_tmem0 + ks * 8, bl = (NATK ? x_lo : xp_lo) + ks * 2;
#pragma unroll 2
            for (int t = 0; t < a.T; ++t) {
              mma_ts_lo(d, at, bl, idesc_delta, acc);
              d += a.mp; at += kBlockK / 2; bl += xp_t_lo;
            }
          }
        }
        }
        tc_commit(&bar_empty[st.idx]);   // stage (W tile, masks, X tile) may be overwritten once these MMAs retire
        {  // so may the TMEM A buffer and its permuted-X tiles: second arrival on the barrier of unit u + n_abuf (see the set-up)
          int nxt = st.idx + a.n_abuf;
          if (nxt >= a.stages) nxt -= a.stages;
          tc_commit(&bar_full[nxt]);
        }
        if (seg_last) tc_commit(&bar_dfull);
        trace_mark<TRACE>(a, u - u_begin, 7);
      }
      __syncwarp();
      st = st_next;
      ab.advance(a.n_abuf);
      if (++kb == a.kblocks) kb = 0;
    }
  } else if (xi >= 0) {
    // ===================================================== activation-permute warps
    // Released per unit by the sync warp (shared-memory counter); permutes / splits the activations while the unpack warps
    // convert the signs; arrives with them on the A buffer's named barrier.
    Ring st, ab;
    // 8-bit path: this lane's (at most two) jobs always belong to the same tenants: their scales are loop invariants
    uint32_t sf0 = 0, sf1 = 0;
    if constexpr (DELTA8) {
      const int j0 = xi * 32 + lane, j1 = j0 + kXpermWarps * 32;
      if (j0 < xjobs) sf0 = d8_scale_field(j0 >> 4);
      if (j1 < xjobs) sf1 = d8_scale_field(j1 >> 4);
    }
#pragma unroll 1
    for (int u = u_begin; u < u_end; ++u) {
      wait_released<BD_XPERM_SLEEP>(&s_released, u - u_begin);
      if (!NATK && !(dbg_flags(a) & (1 | 64))) {
        const uint8_t* xsrc = smem + (size_t)st.idx * a.stage_bytes + a.off_x;
        uint8_t* xp = smem + a.off_xp + (size_t)ab.idx * a.xp_buf_bytes;
        if constexpr (DELTA8) {
          // at most T <= 10 tenants x 16 jobs: one or two jobs per lane
#pragma unroll 1
          for (int ji = 0; ji < 2; ++ji) {  // not unrolled: one copy of the job's code in the instruction cache
            const int job = xi * 32 + lane + ji * kXpermWarps * 32;
            if (job < xjobs) xperm_job(xsrc, xp, job, ji ? sf1 : sf0);
          }
        } else {
          // small row counts: these three warps do all of it; prefill-size row counts: shared with the 8 unpack warps
          const int j0 = xperm_shared ? kUnpackWarps * 32 : 0, jstride = xperm_shared ? (kUnpackWarps + kXpermWarps) * 32 : kXpermWarps * 32;
          for (int job = j0 + xi * 32 + lane; job < xjobs; job += jstride) xperm_job(xsrc, xp, job, 0u);
        }
        fence_proxy_async();
      }
      if (xi == 0 && lane == 0) trace_mark<TRACE>(a, u - u_begin, 10);
      named_bar_arrive(kBarAFull0 + ab.idx, afull_threads);
      st.advance(a.stages);
      ab.advance(a.n_abuf);
    }
  } else if (warp == kWarpSync) {
    // ===================================================== sync warp
    // The only warp of the unpack group that talks to the mbarriers (they are slow and serialised per SM): waits until
    // the stage has landed and an A buffer is free, then publishes the unit in a shared-memory counter (release store;
    // the unpack and permute warps acquire-load it).  It runs ahead of them, so they normally never wait.
    Ring st;
#pragma unroll 1
    for (int u = u_begin; u < u_end; ++u) {
      mbar_wait(&bar_full[st.idx], st.phase);  // stage landed + A buffer free (both arrive on this barrier)
      if (lane == 0) trace_mark<TRACE>(a, u - u_begin, 15);
      __syncwarp();
      if (lane == 0) st_release_shared(&s_released, u - u_begin + 1);
      st.advance(a.stages);
    }
  } else {
    // ===================================================== unpack + epilogue warps
    const int uw = warp;                // 0..7
    const int quad = warp & 3;          // TMEM lane quadrant this warp may touch
    const int grp = uw >> 2;            // two warps per quadrant: they split the tenants
    const int row = quad * 32 + lane;   // weight row inside the tile == TMEM lane
    const int ut = threadIdx.x;         // 0..255
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    constexpr uint32_t kOne = DELTA8 ? 0x38383838u : (std::is_same<T16, __nv_bfloat16>::value ? 0x3F803F80u : 0x3C003C00u);
    uint32_t sign_mask = DELTA8 ? 0x80808080u : 0x80008000u;
    asm volatile("" : "+r"(sign_mask));  // keep the mask in a register so mask + constant fit one LOP3
    Ring st, ab;
    int tile = tile0, nt = nt0, mc = mc0, kb = kb0, seg_kb0 = kb0;  // tile id = nt * m_chunks + mc
    auto advance_tile = [&]() {
      if (a.tile_stride) {
        tile += a.tile_stride;
        nt = tile / a.m_chunks;
        mc = tile - nt * a.m_chunks;
      } else {
        ++tile;
        if (++mc == a.m_chunks) { mc = 0; ++nt; }
      }
    };
    bool seg_is_first = true;  // the current (tile, K run) is the first one of this CTA
    bool pend_valid = false, pend_full = false, pend_first = false;  // run whose read-out is deferred by one round
    int pend_tile = 0, pend_nt = 0, pend_mc = 0;
    uint32_t dphase = 0;
    // Sign words of one unit -> +-1.0 operand registers -> TMEM A buffer `abi` (this warp's tenants, this thread's row).
    auto unpack_unit = [&](const uint8_t* sp, int abi, int tfirst, int tstep) {
      const uint32_t* mw = reinterpret_cast<const uint32_t*>(sp + a.off_masks) + row;
      const uint32_t ta = tmem_base + lane_addr + col_abuf0 + abi * a_cols_per_buf;
#pragma unroll 1
      for (int t0 = tfirst; t0 < a.T; t0 += 3 * tstep) {  // up to three tenants (t0, t0+tstep, t0+2*tstep) per pass
        uint32_t wv[3][kBlockK / 32];
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
          for (int jj = 0; jj < kBlockK / 32; ++jj)
            wv[q][jj] = (t0 + tstep * q < a.T) ? mw[((t0 + tstep * q) * (kBlockK / 32) + jj) * kTileN] : 0u;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int t = t0 + tstep * q;
          if (t >= a.T) break;
          if (DELTA8 && kD8WideStore) {  // one 16-column store per tenant (both 32-K groups of the unit) instead of two 8-column stores
            uint32_t r[16];
#pragma unroll
            for (int c = 0; c < 16; ++c)
              asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[c]) : "r"(wv[q][c >> 3] << (7 - (c & 7))), "r"(sign_mask), "r"(kOne));
            tmem_st16(ta + t * 16, r);
            continue;
          }
#pragma unroll
          for (int jj = 0; jj < kBlockK / 32; ++jj) {
            const uint32_t w = wv[q][jj];
            if (DELTA8 && (dbg_flags(a) & 256) && jj != 0) continue;  // bring-up: half the sign work (what a 4-bit operand would leave)
            if constexpr (DELTA8) {
              uint32_t r[8];
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint32_t sh = w << (7 - c);                // bits c, c+8, c+16, c+24 -> 7, 15, 23, 31
                // r = (~sh & 0x80808080) | 0x38383838 : four e4m3 values, +1.0 = 0x38, -1.0 = 0xB8 (one LOP3, LUT 0xAE)
                asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[c]) : "r"(sh), "r"(sign_mask), "r"(kOne));
              }
              if (dbg_flags(a) & 1024) {  // bring-up: stores without the conversion (timing of the store path alone)
#pragma unroll
                for (int c = 0; c < 8; ++c) r[c] = w;
              }
              if (dbg_flags(a) & 16) {
                asm volatile("" ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]));
              } else {
                tmem_st8(ta + t * 16 + jj * 8, r);
              }
            } else {
              uint32_t r[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                uint32_t sh;
                if constexpr (NATK) {
                  // natural K order: bits 2i, 2i+1 -> 15, 31:  ((w >> 2i) & 3) * 0x40008000 has them at 15/30 and 16/31
                  sh = ((w >> (2 * i)) & 3u) * 0x40008000u;
                } else {
                  sh = w << (15 - i);                            // bit i -> 15, bit i+16 -> 31
                }
                // r = (~sh & 0x80008000) | one : sign = ~bit, bit 1 -> +1.0, bit 0 -> -1.0.  One LOP3 (LUT 0xAE).
                asm("lop3.b32 %0, %1, %2, %3, 0xAE;" : "=r"(r[i]) : "r"(sh), "r"(sign_mask), "r"(kOne));
              }
              tmem_st16(ta + t * (kBlockK / 2) + jj * 16, r);
            }
          }
        }
      }
    };
    // Read-out of one (tile, K run): accumulators -> y (whole K) or -> this CTA's split-K slot + fix-up.  `e_next`: another run
    // follows, the MMA warp waits for these warps to have read the accumulators before it overwrites them (kBarDEmpty).
    auto read_out = [&](int e_tile, int e_nt, int e_mc, bool e_full, bool e_first, bool e_next) {
      if (TRACE && a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.trace[63 * 16 + 2] = clock64();
      // ===================================================== read-out of one (tile, K run)
      mbar_wait(&bar_dfull, dphase);
      dphase ^= 1u;
      tc_fence_after();
      // which matrix of a grouped launch this tile belongs to
      const int sg = (e_nt >= a.seg_tile0[1]) + (e_nt >= a.seg_tile0[2]);
      // row chunk of a prefill-size launch: rows [mc*m, mc*m + m_here) of tenant tt (tt = 0 unless it is a multi-tenant prefill)
      const int tt = e_nt / a.tpt;
      const int ltile = e_nt - a.seg_tile0[sg] - tt * a.tpt;
      const int r_off = tt * a.m_total + e_mc * a.m;
      const int m_here = min(a.m, a.m_total - e_mc * a.m);
      const int64_t seg_n = a.n_seg[sg];
      T16* __restrict__ y = reinterpret_cast<T16*>(a.y_seg[sg]);
      const void* seg_coeff = a.coeff_seg[sg];
      const int64_t n = (int64_t)ltile * kTileN + row;
      // this CTA's partial slot: 0 if the run starts the CTA's unit range, else 1 (only the first and the last run
      // of a CTA can be partial; the runs in between cover whole tiles)
      const int slot = e_first ? 0 : 1;
      float* part = a.partial + ((size_t)(cta * 2 + slot) * a.rows) * kTileN;

      if constexpr (DELTA8) {
        // delta accumulator: 8 columns per tenant (one row per tenant), columns 0..2 = the three pieces of the scaled
        // activation; the two warps of a quadrant take alternate tenants
        for (int t = grp; t < a.T; t += 2) {
          const float cf = HAS_BASE ? load_coeff(seg_coeff, a.coeff_dtype, t) : 1.0f;
          float d0[8], bv = 0.f;
          tmem_ld8(tmem_base + lane_addr + col_ddelta + t * 8, d0);
          if (HAS_BASE) bv = tmem_ld1(tmem_base + lane_addr + col_dbase + t);
          tc_wait_ld();
          const float dsum = (d0[0] + d0[1] + d0[2] + (kD8Pieces > 3 ? d0[3] : 0.f)) * __uint_as_float((254u - d8_scale_field(t)) << 23);  // undo the row scale
          const float v = HAS_BASE ? fmaf(cf, dsum, bv) : dsum;
          if (e_full) {
            if (n < seg_n) store_y(y, (int64_t)t * seg_n + n, v, a.fp32_out);
          } else {
            part[(size_t)t * kTileN + row] = v;
          }
        }
      } else
      for (int t = 0; t < a.T; ++t) {
        const float cf = HAS_BASE ? load_coeff(seg_coeff, a.coeff_dtype, t + tt) : 1.0f;
        for (int c8 = 0; c8 < a.mp / 8; ++c8) {
          if (((t * (a.mp / 8) + c8) & 1) != grp) continue;   // the two warps of a quadrant split the column chunks
          if (c8 * 8 >= m_here) continue;
          float dv[8], bv[8];
          tmem_ld8(tmem_base + lane_addr + col_ddelta + t * a.mp + c8 * 8, dv);
          if (HAS_BASE && a.T == 1) {
            tmem_ld8(tmem_base + lane_addr + col_dbase + c8 * 8, bv);  // one tenant: base columns line up with the chunk
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              bv[i] = 0.f;
              if (HAS_BASE && c8 * 8 + i < m_here) bv[i] = tmem_ld1(tmem_base + lane_addr + col_dbase + t * a.m + c8 * 8 + i);
            }
          }
          tc_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int ii = c8 * 8 + i;
            if (ii >= m_here) continue;
            const int r = t * a.m + ii;
            const float v = HAS_BASE ? fmaf(cf, dv[i], bv[i]) : dv[i];
            if (e_full) {
              if (n < seg_n) store_y(y, (int64_t)(r_off + r) * seg_n + n, v, a.fp32_out);
            } else {
              part[(size_t)r * kTileN + row] = v;
            }
          }
        }
      }
      tc_fence_before();

      // every unpack warp has read its part of the accumulators: the MMA warp may start the next run (it waits for all
      // eight warps here -- with alternating units a unit's hand-over involves only half of them)
      if (e_next) named_bar_arrive(kBarDEmpty, kUnpackWarps * 32 + 32);

      if (!e_full) {
        // ---- split-K fix-up: the last CTA to arrive sums every contributor's slot in K order ----
        const int first_unit = e_tile * a.kblocks, last_unit = first_unit + a.kblocks - 1;
        const int big = a.units_rem * (a.units_per_cta + 1);  // units owned by the CTAs that got one extra unit
        const int c_first = first_unit < big ? first_unit / (a.units_per_cta + 1) : a.units_rem + (first_unit - big) / a.units_per_cta;
        const int c_last = last_unit < big ? last_unit / (a.units_per_cta + 1) : a.units_rem + (last_unit - big) / a.units_per_cta;
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(kUnpackWarps * 32) : "memory");
        if (ut == 0) {
          const unsigned old = atomicAdd(&a.counters[e_tile], 1u);
          s_is_last = (old == (unsigned)(c_last - c_first)) ? 1u : 0u;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kUnpackWarps * 32) : "memory");
        if (s_is_last) {
          __threadfence();
          // Deterministic reduction with memory-level parallelism: work item = (output row r, 4 consecutive weight
          // rows); the slots of all contributors are fetched with independent 16-byte L2 loads, 8 in flight per
          // thread, and added in K order.  (A serial loop over ~18 contributors cost a 10 us tail per launch.)
          const int items = (a.m_chunks > 1 ? m_here : a.rows) * (kTileN / 4);
          for (int item = ut; item < items; item += kUnpackWarps * 32) {
            const int r = item / (kTileN / 4), q4 = item - r * (kTileN / 4);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c0 = c_first; c0 <= c_last; c0 += 8) {
              float4 v[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = c0 + j;
                v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c <= c_last) {
                  const int cs = (cta_unit_begin(a, c) >= first_unit) ? 0 : 1;
                  v[j] = __ldcg(reinterpret_cast<const float4*>(a.partial + ((size_t)(c * 2 + cs) * a.rows + r) * kTileN) + q4);
                }
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
            }
            const int64_t n4 = (int64_t)ltile * kTileN + q4 * 4;
            const int64_t o4 = (int64_t)(r_off + r) * seg_n + n4;
            if (n4 + 3 < seg_n) {  // N % 4 == 0 and rows of y are 8-byte (fp32: 16-byte) aligned for these 4 elements
              if (a.fp32_out) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + o4) = acc;
              } else {
                const T16 o[4] = {F16<T16>::from_f32(acc.x), F16<T16>::from_f32(acc.y), F16<T16>::from_f32(acc.z), F16<T16>::from_f32(acc.w)};
                *reinterpret_cast<uint2*>(y + o4) = *reinterpret_cast<const uint2*>(o);
              }
            }
          }
          if (ut == 0) a.counters[e_tile] = 0u;  // leave the workspace clean for the next launch
        }
      }
      if (TRACE && a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) a.trace[63 * 16 + 3] = clock64();
    };
    if (alt_units) {
      // ---- decode path: the two warp groups take alternate units.  One warp's whole per-unit loop is: release check,
      // unpack, store drain, barrier arrive, two ring steps -- the run bookkeeping (tile cursor, read-out) sits outside it:
      // as straight-line code of one loop it cost these warps ~500 cycles of branches and constant loads per unit.
      // A run that is followed by another one is read out one unit LATE: these warps first hand over their first unit of
      // the next run, so that the MMA warp has work ready the moment the accumulators are free, and the run's MMAs have
      // long completed when these warps come to wait for them (read out right away, every run boundary emptied the
      // pipeline: CTAs with two runs finished ~3 us after those with one).
      const int n_units = u_end - u_begin;
      int it = grp;                       // this warp's next unit (CTA-relative): grp, grp + 2, ...
      int sti = grp, abi = grp;           // its stage / A buffer (stages, n_abuf >= 2)
      if (sti >= a.stages) sti -= a.stages;
      if (abi >= a.n_abuf) abi -= a.n_abuf;
      auto own_unit = [&]() {
        if (uw == 0 && lane == 0) trace_mark<TRACE>(a, it, 0);
        wait_released<BD_UNPACK_SLEEP>(&s_released, it);
        tc_fence_after();
        if (uw == 0 && lane == 0) trace_mark<TRACE>(a, it, 1);
        if (!(dbg_flags(a) & 1)) {
          unpack_unit(smem + (size_t)sti * a.stage_bytes, abi, 0, 1);
          if (uw == 0 && lane == 0) trace_mark<TRACE>(a, it, 3);
          tc_wait_st();
        }
        tc_fence_before();
        if (uw == 0 && lane == 0) trace_mark<TRACE>(a, it, 4);
        named_bar_arrive(kBarAFull0 + abi, afull_threads);  // this warp's rows of A buffer abi are written
        if (uw == 0 && lane == 0) trace_mark<TRACE>(a, it, 11);
        sti += 2; if (sti >= a.stages) sti -= a.stages;
        abi += 2; if (abi >= a.n_abuf) abi -= a.n_abuf;
        if (uw == 0 && lane == 0) trace_mark<TRACE>(a, it, 13);
        it += 2;
      };
      int rb = 0;                         // CTA-relative first unit of the current run
      bool more = n_units > 0, pend_next = false;
      int re = 0;
      while (true) {
        if (more) {
          re = min(n_units, rb + a.kblocks - kb);
          if (it < re) own_unit();        // first own unit of the run: ahead of the previous run's read-out
        }
        if (pend_valid) {
          if (TRACE && a.trace != nullptr && threadIdx.x == 0 && !pend_next) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a.trace[1024 + 4 * blockIdx.x + 3]));  // last unit handed over
          if (dbg_flags(a) & 1) {  // bring-up: stream only
            mbar_wait(&bar_dfull, dphase);
            dphase ^= 1u;
            if (pend_next) named_bar_arrive(kBarDEmpty, kUnpackWarps * 32 + 32);
          } else {
            read_out(pend_tile, pend_nt, pend_mc, pend_full, pend_first, pend_next);
          }
          pend_valid = false;
        }
        if (!more) break;
#pragma unroll 1
        while (it < re) own_unit();
        // the run [rb, re) is handed over: remember it for the read-out, step to the next tile
        pend_valid = true; pend_tile = tile; pend_nt = nt; pend_mc = mc;
        pend_full = (kb == 0) && (kb + (re - rb) == a.kblocks);
        pend_first = (rb == 0);
        pend_next = (re < n_units);
        kb += re - rb;
        if (kb == a.kblocks) { kb = 0; advance_tile(); }
        rb = re;
        more = rb < n_units;
      }
    } else {
    // ---- prefill-size row counts and the 16-bit operand path: the two warps of a quadrant split every unit's tenants.
    // Units are processed in rounds of up to two (never across the end of a (tile, K run)): the per-round costs of this
    // in-order warp -- release check, tcgen05 fences, tcgen05.wait::st, loop bookkeeping, ~800 cycles -- are paid once per
    // round instead of once per unit.  Two rounds fit the >= 4 A buffers, so unpacking still overlaps with the MMAs.
    const int round_units = a.n_abuf >= 4 ? 2 : 1;
#pragma unroll 1
    for (int u = u_begin; u < u_end;) {
      const int it = u - u_begin;
      const int seg_left = min(a.kblocks - kb, u_end - u);
      const int g = min(round_units, seg_left);
      const bool seg_last = (g == seg_left);
      const bool tr = (uw == 0 && lane == 0);
      // The sync warp waits on the mbarriers for the whole group and publishes released units in s_released.
      if (tr) trace_mark<TRACE>(a, it, 0);
      wait_released(&s_released, it + g - 1);
      tc_fence_after();
      if (tr) trace_mark<TRACE>(a, it, 1);
      Ring st_i = st, ab_i = ab;
      if (!(dbg_flags(a) & 1)) {
        for (int i = 0; i < g; ++i) {
          const uint8_t* sp = smem + (size_t)st_i.idx * a.stage_bytes;
          if (!NATK && xperm_shared) {  // large row counts: the unpack warps share the activation permutation
            uint8_t* xp = smem + a.off_xp + (size_t)ab_i.idx * a.xp_buf_bytes;
            for (int job = ut; job < xjobs; job += (kUnpackWarps + kXpermWarps) * 32) xperm_job(sp + a.off_x, xp, job, 0u);
          }
          unpack_unit(sp, ab_i.idx, grp, 2);
          st_i.advance(a.stages);
          ab_i.advance(a.n_abuf);
        }
        if (tr) trace_mark<TRACE>(a, it, 3);
        tc_wait_st();
        if (!NATK && xperm_shared) fence_proxy_async();
      }
      tc_fence_before();
      if (tr) trace_mark<TRACE>(a, it, 4);
      if (uw == kUnpackWarps - 1 && lane == 0) trace_mark<TRACE>(a, it, 9);
      for (int i = 0; i < g; ++i) {
        named_bar_arrive(kBarAFull0 + ab.idx, afull_threads);  // this warp's part of A buffer ab.idx is written
        st.advance(a.stages);
        ab.advance(a.n_abuf);
      }
      if (tr) trace_mark<TRACE>(a, it, 11);
      if (tr) trace_mark<TRACE>(a, it, 12);
      u += g;
      kb += g - 1;  // kb = K block of the last unit of the round (the epilogue below looks at it)
      if ((dbg_flags(a) & 1) && seg_last) {
        mbar_wait(&bar_dfull, dphase);
        dphase ^= 1u;
        if (u < u_end) named_bar_arrive(kBarDEmpty, kUnpackWarps * 32 + 32);
        if (++kb == a.kblocks) { kb = 0; advance_tile(); }
        seg_kb0 = kb;
        seg_is_first = false;
        continue;
      }

      if (seg_last) {
        if (TRACE && a.trace != nullptr && threadIdx.x == 0 && u == u_end) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a.trace[1024 + 4 * blockIdx.x + 3]));  // last unit handed over
        read_out(tile, nt, mc, (seg_kb0 == 0) && (kb + 1 == a.kblocks), seg_is_first, u < u_end);
        seg_is_first = false;
      }
      if (++kb == a.kblocks) { kb = 0; advance_tile(); }
      if (seg_last) seg_kb0 = kb;  // the next run starts at the next unit (kb == 0 unless the CTA's range ended)
      if (tr) trace_mark<TRACE>(a, it, 13);
    }
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (TRACE && a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { a.trace[63 * 16 + 4] = clock64(); asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a.trace[63 * 16 + 9])); }
  if (TRACE && a.trace != nullptr && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(a.trace[1024 + 4 * blockIdx.x + 1]));
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

#ifdef BD_BRINGUP
long long* g_trace_buf = nullptr;
int g_dbg_flags = 0;  // device bits: see dbg_flags(); host bit 1 = force the 16-bit delta path
#else
constexpr long long* g_trace_buf = nullptr;
constexpr int g_dbg_flags = 0;
#endif

struct DeviceInfo {
  int sms = 0, smem_optin = 0, cc_major = 0;
};
DeviceInfo device_info() {
  static DeviceInfo info[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return DeviceInfo{};
  std::lock_guard<std::mutex> lock(mu);
  if (info[dev].sms == 0) {
    cudaDeviceGetAttribute(&info[dev].sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&info[dev].smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&info[dev].cc_major, cudaDevAttrComputeCapabilityMajor, dev);
  }
  return info[dev];
}

struct UmmaPlan {
  bool ok = false;
  const char* why = "";
  int mp = 0, ntb = 0, stages = 0, n_abuf = 0, a_cols_tenant = 0;
  bool d8 = false, natk = false;
  uint32_t stage_bytes = 0, off_masks = 0, off_x = 0, off_xp = 0, xp_buf_bytes = 0, smem_bytes = 0, tx_bytes = 0;
};

UmmaPlan plan_umma(int64_t T, int64_t m, int64_t K, int64_t N, bool has_base, bool d8) {
  UmmaPlan p;
  p.d8 = d8;
  if (d8 && m != 1) { p.why = "the 8-bit delta path takes one row per tenant"; return p; }
  const int64_t rows = T * m;
  if (rows > kMaxRows) { p.why = "more than 128 rows per launch"; return p; }
  if (T > 1 && m > 16) { p.why = "multi-tenant launches support at most 16 rows per tenant"; return p; }
  if (N % 4 != 0) { p.why = "N must be a multiple of 4 (TMA row pitch of the sign words)"; return p; }
  if (K % 32 != 0) { p.why = "K must be a multiple of 32"; return p; }
  if ((N + kTileN - 1) / kTileN > (int64_t)(kWsCounterBytes / sizeof(unsigned))) { p.why = "too many N tiles"; return p; }
  p.mp = d8 ? 8 : (int)((m + 15) / 16 * 16);  // delta accumulator columns per tenant (8-bit path: the 3 pieces of the one row)
  p.a_cols_tenant = d8 ? kBlockK / 4 : kBlockK / 2;
  p.ntb = (int)((rows + 15) / 16 * 16);
  const int64_t a_cols_one = T * p.a_cols_tenant;
  int64_t nbuf = ((int64_t)kTmemCols - p.ntb - T * p.mp) / a_cols_one;
  if (nbuf < 2) { p.why = "accumulators + sign operand buffers exceed 512 TMEM columns"; return p; }
  if (nbuf > kMaxABuf) nbuf = kMaxABuf;
  // the K-permuted activation tiles (one set per A buffer) live in shared memory: keep them under 48 KiB
  p.natk = !d8 && T == 1;  // one tenant, 16-bit signs: natural K order, no permuted activation copy
  const int64_t xp_one = p.natk ? 1024 : (d8 ? T * 512 : T * p.mp * 128);
  while (nbuf > 2 && nbuf * xp_one > 48 * 1024) --nbuf;
  p.n_abuf = (int)nbuf;
  const uint32_t w_bytes = has_base ? kTileN * kBlockK * 2 : 0;
  const uint32_t m_bytes = (uint32_t)T * (kBlockK / 32) * kTileN * 4;
  const uint32_t x_bytes = (uint32_t)p.ntb * 128;
  auto up1k = [](uint32_t v) { return (v + 1023u) & ~1023u; };
  p.off_masks = w_bytes;
  p.off_x = up1k(w_bytes + m_bytes);
  p.stage_bytes = up1k(p.off_x + x_bytes);
  p.tx_bytes = w_bytes + m_bytes + x_bytes;
  p.xp_buf_bytes = (uint32_t)xp_one;
  const uint32_t budget = 216u * 1024u;
  const uint32_t fixed = p.n_abuf * p.xp_buf_bytes;
  if (fixed + 3 * p.stage_bytes > budget) { p.why = "tile does not fit in shared memory"; return p; }
  p.stages = (int)((budget - fixed) / p.stage_bytes);
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  // An A buffer is held for a sub-interval of its unit's stage, so more A buffers than stages buy nothing -- and the kernel
  // relies on n_abuf <= stages (the MMA commit of unit u arrives on the stage barrier of unit u + n_abuf).
  if (p.n_abuf > p.stages) p.n_abuf = p.stages;
  p.off_xp = p.stages * p.stage_bytes;
  p.smem_bytes = p.off_xp + fixed + 1024;  // + slack for the manual 1 KiB alignment
  p.ok = true;
  return p;
}

// Picks the 8-bit delta path (e4m3 signs x e5m2 activation pieces: half the TMEM traffic, half the unpack work, half the
// MMAs) whenever it applies -- bf16 activations, one row per tenant (decode) -- else the 16-bit path.
UmmaPlan choose_plan(int dtype, int64_t T, int64_t m, int64_t K, int64_t N, bool has_base) {
  if ((dtype == BD_BF16 || dtype == BD_FP16) && m == 1 && !(g_dbg_flags & 2)) {
    UmmaPlan p8 = plan_umma(T, m, K, N, has_base, true);
    if (p8.ok) return p8;
  }
  return plan_umma(T, m, K, N, has_base, false);
}

int encode_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
               const cuuint32_t* box, CUtensorMapSwizzle swz, const char* what) {
  // The driver entry point needs a current context on the calling thread; a thread that has only ever had its device
  // *selected* (e.g. PyTorch's autograd worker) has none until its first runtime call that touches the device.
  static thread_local int ctx_dev = -1;
  int cur_dev = -1;
  if (cudaGetDevice(&cur_dev) == cudaSuccess && cur_dev != ctx_dev) {
    (void)cudaFree(nullptr);
    ctx_dev = cur_dev;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(BD_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(BD_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return BD_OK;
}

template <typename T16, bool HAS_BASE, bool DELTA8, bool NATK, bool TRACE>
int launch_typed(const FwdProblem& p, const UmmaPlan& plan, const UmmaMaps& maps, const UmmaArgs& args, int grid) {
  auto kern = fwd_umma_kernel<T16, HAS_BASE, DELTA8, NATK, TRACE>;
  static std::once_flag once;  // one per template instantiation
  static cudaError_t attr_err = cudaSuccess;
  std::call_once(once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024); });
  if (attr_err != cudaSuccess) return fail(BD_ERR_CUDA, "cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(attr_err));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = plan.smem_bytes;
  cfg.stream = p.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // PDL: see the kernel's griddepcontrol use
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, maps, args);
  if (le != cudaSuccess) return fail(BD_ERR_CUDA, "cudaLaunchKernelEx(fwd_umma_kernel) failed: %s", cudaGetErrorString(le));
  count_launch();
  return check_launch("fwd_umma_kernel");
}

}  // namespace

// Largest number of tenants (<= T) that one launch can take at m rows per tenant; 0 if even one does not fit.
static int tenants_per_launch(int dtype, int64_t T, int64_t m, int64_t K, int64_t N, bool has_base) {
  for (int64_t g = T < 10 ? T : 10; g >= 1; --g)
    if (choose_plan(dtype, g, m, K, N, has_base).ok) return (int)g;
  return 0;
}

bool umma_supports(const FwdProblem& p, const char** why) {
  int64_t T = p.T, m = p.m;
  if (p.mask_tenant_stride == 0 && T > 1) {
    if (p.w) { *why = "broadcast sign matrix with several tenants"; return false; }
    m = T * m; T = 1;  // binary_bmm with one shared sign matrix == one tenant with T*m rows
  }
  if (p.dtype != BD_BF16 && p.dtype != BD_FP16) { *why = "dtype"; return false; }
  // Problems larger than one launch are decomposed by launch_fwd_umma (tenant groups, then 128-row chunks), so only
  // the per-launch constraints matter here: check the smallest piece.
  UmmaPlan plan = choose_plan(p.dtype, 1, m < kMaxRows ? m : kMaxRows, p.K, p.N, p.w != nullptr);
  if (!plan.ok) { *why = plan.why; return false; }
  *why = "";
  return true;
}

size_t umma_workspace_bytes(int64_t rows, int64_t N) {
  (void)N;
  if (rows > kMaxRows) rows = kMaxRows;
  return kWsScratchOffset + (size_t)160 * 2 * rows * kTileN * sizeof(float);  // up to 160 SMs
}

static int launch_one(const FwdProblem& p) {
  int64_t T = p.T, m = p.m;
  // Prefill-size launches of ONE tenant are processed in row chunks of 128 inside the launch: tile = (N tile, row chunk)
  // ... and so are multi-tenant launches with more than 16 rows per tenant (tile = (N tile, tenant, row chunk); inside the
  // kernel it is a one-tenant problem per tile: the tile's tenant selects the sign words, the coefficient and the rows).
  const int64_t m_total = m;
  const int64_t tenants_real = T;
  int64_t m_chunks = 1;
  if ((T == 1 && m > kMaxRows) || (T > 1 && m > 16)) {
    if (p.nseg > 1) return fail(BD_ERR_UNSUPPORTED, "grouped launch: more than %d rows", kMaxRows);
    m_chunks = (m + kMaxRows - 1) / kMaxRows;
    if (m > kMaxRows) m = kMaxRows;
    T = 1;
  }
  const bool has_base = p.w != nullptr;
  const int nseg = p.nseg < 1 ? 1 : p.nseg;
  if (nseg > kMaxSeg) return fail(BD_ERR_INVALID, "tcgen05 kernel: at most %d matrices per grouped launch", kMaxSeg);
  // per-segment views
  const void* ws[kMaxSeg] = {p.w, p.seg_w[0], p.seg_w[1]};
  const int32_t* ms[kMaxSeg] = {p.masks, p.seg_masks[0], p.seg_masks[1]};
  const void* cs[kMaxSeg] = {p.coeff, p.seg_coeff[0], p.seg_coeff[1]};
  void* ys[kMaxSeg] = {p.y, p.seg_y[0], p.seg_y[1]};
  int64_t Ns[kMaxSeg] = {p.N, p.seg_N[0], p.seg_N[1]};
  int64_t strides[kMaxSeg] = {p.mask_tenant_stride, p.seg_mask_tenant_stride[0], p.seg_mask_tenant_stride[1]};
  int64_t n_max = 0;
  for (int sg = 0; sg < nseg; ++sg) {
    if (strides[sg] == 0) strides[sg] = (p.K / 32) * Ns[sg];
    if (sg + 1 < nseg && Ns[sg] % kTileN != 0) return fail(BD_ERR_UNSUPPORTED, "grouped launch: every N but the last must be a multiple of %d", kTileN);
    if (Ns[sg] % 4 != 0) return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel: N must be a multiple of 4");
    if (Ns[sg] > n_max) n_max = Ns[sg];
  }
  UmmaPlan plan = choose_plan(p.dtype, T, m, p.K, n_max, has_base);
  if (!plan.ok) return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel: %s", plan.why);
  DeviceInfo di = device_info();
  if (di.cc_major != 10) return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel needs an sm_100 device (found compute capability %d.x)", di.cc_major);
  if ((uint32_t)di.smem_optin < plan.smem_bytes) return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel needs %u bytes of shared memory", plan.smem_bytes);
  if (di.sms > 160) return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel: more SMs than the workspace layout assumes");

  const int64_t rows = T * m;
  UmmaArgs a{};
  a.x = p.x; a.coeff = p.coeff; a.coeff_dtype = p.coeff_dtype; a.y = p.y;
  a.T = (int)T; a.m = (int)m; a.rows = (int)rows; a.inv_m = (uint32_t)((65536 + m - 1) / m); a.mp = plan.mp; a.ntb = plan.ntb;
  a.K = (int)p.K; a.N = (int)p.N;
  a.kblocks = (int)((p.K + kBlockK - 1) / kBlockK);
  a.nseg = nseg;
  int tiles = 0;
  for (int sg = 0; sg <= kMaxSeg; ++sg) a.seg_tile0[sg] = 0x7fffffff;  // unused segments never match
  for (int sg = 0; sg < nseg; ++sg) {
    a.seg_tile0[sg] = tiles;
    a.n_seg[sg] = (int)Ns[sg];
    a.y_seg[sg] = ys[sg];
    a.coeff_seg[sg] = cs[sg];
    tiles += (int)((Ns[sg] + kTileN - 1) / kTileN);
  }
  if (tiles > (int)(kWsCounterBytes / sizeof(unsigned))) return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel: too many N tiles");
  a.m_chunks = (int)m_chunks;
  a.m_total = (int)m_total;
  a.tpt = tiles;  // tile columns [tt * tiles, (tt + 1) * tiles) belong to tenant tt
  const int64_t total_tiles = (int64_t)tiles * m_chunks * (T == 1 ? tenants_real : 1);
  if (total_tiles * a.kblocks > 0x7fffffff) return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel: problem too large for one launch");
  a.n_tiles = (int)total_tiles;
  a.total_units = a.n_tiles * a.kblocks;
  const int grid = a.total_units < di.sms ? a.total_units : di.sms;
  // Many tiles per CTA: hand out whole tiles (no split-K fix-up, whose partials grow with the row count); otherwise
  // stream-K over single units so that every SM streams the same number of bytes.
  const bool whole_tiles = total_tiles >= 4 * (int64_t)grid;
  a.unit_quantum = whole_tiles ? a.kblocks : 1;
  // multi-tenant prefill: strided tiles, so that at any time all CTAs work on (about) the same tenant and its activations
  // stay in L2 (with consecutive tiles per CTA every tenant's activations are live at once: 168 MB for 4 x 4096 x 5120)
  a.tile_stride = (whole_tiles && T == 1 && tenants_real > 1) ? grid : 0;
  const int sched_units = whole_tiles ? a.n_tiles : a.total_units;
  a.units_per_cta = sched_units / grid;
  a.units_rem = sched_units % grid;
  if (!whole_tiles && a.n_tiles > (int)(kWsCounterBytes / sizeof(unsigned))) return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel: too many tiles for the split-K counters");
  a.n_abuf = plan.n_abuf;
  a.a_cols_tenant = plan.a_cols_tenant;
  a.stages = plan.stages; a.stage_bytes = plan.stage_bytes; a.off_masks = plan.off_masks; a.off_x = plan.off_x;
  a.off_xp = plan.off_xp; a.xp_buf_bytes = plan.xp_buf_bytes; a.tx_bytes = plan.tx_bytes;
  const size_t need = kWsScratchOffset + (size_t)grid * 2 * rows * kTileN * sizeof(float);
  if (!p.workspace || p.workspace_bytes < need) return fail(BD_ERR_WORKSPACE, "tcgen05 forward needs %zu workspace bytes, got %zu", need, p.workspace_bytes);
  a.counters = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(p.workspace) + kWsUmmaCounterOffset);
  a.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(p.workspace) + kWsScratchOffset);
  a.trace = g_trace_buf;
  a.dbg = g_dbg_flags;
  a.static_ops = p.static_operands ? 1 : 0;
  a.fp32_out = p.fp32_out ? 1 : 0;

  const CUtensorMapDataType dt16 = p.dtype == BD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  alignas(64) UmmaMaps maps{};
  int rc;
  for (int sg = 0; sg < nseg; ++sg) {
    if (has_base) {
      if (!ws[sg]) return fail(BD_ERR_INVALID, "grouped launch: segment %d has no base weight", sg);
      cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)Ns[sg]}, str[1] = {(cuuint64_t)p.K * 2};
      cuuint32_t box[2] = {kBlockK, kTileN};
      if ((rc = encode_map(&maps.w[sg], dt16, 2, ws[sg], dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, "w"))) return rc;
    }
    cuuint64_t dims[3] = {(cuuint64_t)Ns[sg], (cuuint64_t)(p.K / 32), (cuuint64_t)tenants_real};
    cuuint64_t str[2] = {(cuuint64_t)Ns[sg] * 4, (cuuint64_t)strides[sg] * 4};
    cuuint32_t box[3] = {kTileN, kBlockK / 32, (cuuint32_t)T};
    if ((rc = encode_map(&maps.m[sg], CU_TENSOR_MAP_DATA_TYPE_INT32, 3, ms[sg], dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE, "masks"))) return rc;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)(tenants_real * m_total)}, str[1] = {(cuuint64_t)p.K * 2};
    cuuint32_t box[2] = {kBlockK, (cuuint32_t)plan.ntb};
    if ((rc = encode_map(&maps.x, dt16, 2, p.x, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, "x"))) return rc;
  }
  if (p.dtype == BD_BF16) {
    if (plan.d8) {
#ifdef BD_BRINGUP
      if (has_base && a.trace) return launch_typed<__nv_bfloat16, true, true, false, true>(p, plan, maps, a, grid);  // instrumented variant
#endif
      return has_base ? launch_typed<__nv_bfloat16, true, true, false, false>(p, plan, maps, a, grid)
                      : launch_typed<__nv_bfloat16, false, true, false, false>(p, plan, maps, a, grid);
    }
    if (plan.natk)
      return has_base ? launch_typed<__nv_bfloat16, true, false, true, false>(p, plan, maps, a, grid)
                      : launch_typed<__nv_bfloat16, false, false, true, false>(p, plan, maps, a, grid);
    return has_base ? launch_typed<__nv_bfloat16, true, false, false, false>(p, plan, maps, a, grid)
                    : launch_typed<__nv_bfloat16, false, false, false, false>(p, plan, maps, a, grid);
  }
  if (plan.d8)
    return has_base ? launch_typed<__half, true, true, false, false>(p, plan, maps, a, grid)
                    : launch_typed<__half, false, true, false, false>(p, plan, maps, a, grid);
  if (plan.natk)
    return has_base ? launch_typed<__half, true, false, true, false>(p, plan, maps, a, grid)
                    : launch_typed<__half, false, false, true, false>(p, plan, maps, a, grid);
  return has_base ? launch_typed<__half, true, false, false, false>(p, plan, maps, a, grid)
                  : launch_typed<__half, false, false, false, false>(p, plan, maps, a, grid);
}

#ifdef BD_BRINGUP
void umma_set_trace(long long* buf) { g_trace_buf = buf; }
void umma_set_debug(int flags, int) { g_dbg_flags = flags; }
#endif

// Decomposes a problem into launches the kernel takes: tenant groups that fit the TMEM budget, then 128-row chunks of a
// single tenant.  Sub-launches are stream-ordered and share the workspace (each leaves its counters at zero).
int launch_fwd_umma(const FwdProblem& p0) {
  FwdProblem p = p0;
  if (p.mask_tenant_stride == 0 && p.T > 1) { p.m = p.T * p.m; p.T = 1; }
  const size_t esz = 2;  // bf16 / fp16
  const size_t ysz = p.fp32_out ? 4 : esz;
  const size_t csz = p.coeff_dtype == BD_FP32 ? 4 : 2;
  const bool has_base = p.w != nullptr;
  if (p.nseg > 1) {
    int64_t n_max = p.N;
    for (int sg = 1; sg < p.nseg; ++sg) n_max = p.seg_N[sg - 1] > n_max ? p.seg_N[sg - 1] : n_max;
    bool aligned = p.N % kTileN == 0;
    for (int sg = 1; sg + 1 < p.nseg; ++sg) aligned = aligned && (p.seg_N[sg - 1] % kTileN == 0);
    if (aligned && choose_plan(p.dtype, p.T, p.m, p.K, n_max, has_base).ok) return launch_one(p);
    // does not fit one grouped launch: run the matrices one by one
    FwdProblem s = p;
    s.nseg = 1;
    int rc = launch_fwd_umma(s);
    for (int sg = 1; sg < p.nseg && rc == BD_OK; ++sg) {
      s = p;
      s.nseg = 1;
      s.w = p.seg_w[sg - 1]; s.masks = p.seg_masks[sg - 1]; s.coeff = p.seg_coeff[sg - 1]; s.y = p.seg_y[sg - 1];
      s.N = p.seg_N[sg - 1]; s.mask_tenant_stride = p.seg_mask_tenant_stride[sg - 1];
      rc = launch_fwd_umma(s);
    }
    return rc;
  }
  if (choose_plan(p.dtype, p.T, p.m, p.K, p.N, has_base).ok) return launch_one(p);
  // multi-tenant prefill: every tenant's row chunks in ONE launch
  if (p.T > 1 && p.m > 16 && p.nseg <= 1 && p.T * p.m <= (1 << 20) && choose_plan(p.dtype, 1, p.m < kMaxRows ? p.m : kMaxRows, p.K, p.N, has_base).ok)
    return launch_one(p);
  if (p.T > 1) {
    int g = p.m <= 16 ? tenants_per_launch(p.dtype, p.T, p.m, p.K, p.N, has_base) : 1;
    if (g < 1) g = 1;
    for (int64_t t0 = 0; t0 < p.T; t0 += g) {
      FwdProblem s = p;
      s.T = (p.T - t0 < g) ? p.T - t0 : g;
      s.x = reinterpret_cast<const char*>(p.x) + (size_t)t0 * p.m * p.K * esz;
      s.y = reinterpret_cast<char*>(p.y) + (size_t)t0 * p.m * p.N * ysz;
      s.masks = p.masks + t0 * p.mask_tenant_stride;
      if (p.coeff) s.coeff = reinterpret_cast<const char*>(p.coeff) + (size_t)t0 * csz;
      int rc = launch_fwd_umma(s);
      if (rc) return rc;
    }
    return BD_OK;
  }
  if (p.T == 1 && p.m > kMaxRows && p.nseg <= 1 && p.m <= (1 << 20)) return launch_one(p);  // row chunks inside the launch
  if (p.m > kMaxRows) {
    for (int64_t r0 = 0; r0 < p.m; r0 += kMaxRows) {
      FwdProblem s = p;
      s.m = (p.m - r0 < kMaxRows) ? p.m - r0 : kMaxRows;
      s.x = reinterpret_cast<const char*>(p.x) + (size_t)r0 * p.K * esz;
      s.y = reinterpret_cast<char*>(p.y) + (size_t)r0 * p.N * ysz;
      int rc = launch_one(s);
      if (rc) return rc;
    }
    return BD_OK;
  }
  return fail(BD_ERR_UNSUPPORTED, "tcgen05 kernel: %s", choose_plan(p.dtype, p.T, p.m, p.K, p.N, has_base).why);
}

}  // namespace bd
