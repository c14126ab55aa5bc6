// stream_spmv.cuh -- the streaming sparse kernel skeleton every hot kernel is built on.
//
// Layout (built once at pdlp_create, see build_sell_plan in pdlp_b200.cu): SELL-32-sigma.  Rows are sorted by
// length inside windows of SIGMA rows and cut into SLICES of 32 rows; a slice stores its rows column-major
// (element k of lane r at [k * 32 + r]) padded to the slice's longest row, so a warp works one row per lane with
// conflict-free shared-memory reads and no cross-lane reduction.  Slices are cut into PIECES of at most KB steps
// (a long slice spans several pieces; several short slices share one piece); a piece is one contiguous block
// [header: its own Piece, 16 B][rows: nslices x 32 int32, only where rows start][cols: steps x 32 int32]
// [vals: steps x 32 T] in global memory, i.e. ONE 1-D TMA bulk copy brings everything a warp needs for it.
//
// Execution: persistent CTAs of W autonomous warps.  Each warp owns a contiguous, weight-balanced range of pieces
// and a private ring of S shared-memory stages; its lane 0 issues cp.async.bulk for piece p + S the moment piece
// p has been consumed (mbarrier complete_tx signals arrival; no producer warp, no empty barriers, no CTA-wide
// barrier inside the loop).  Per piece a lane issues up to U independent gathers of the dense operand before
// the FMAs, accumulates in registers across the pieces of its row, and runs the fused row epilogue (whose
// operands were requested when the row started).  Rows longer than LONG_ROW are handled first, one row per CTA
// straight from the CSR arrays.  All reductions (lane -> warp -> CTA -> grid) have a fixed order: no atomics
// touch data and results are bitwise reproducible for a given plan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>
#include <utility>

namespace pdlp {

#ifndef PDLP_W
#define PDLP_W 12         // warps per CTA
#endif
#ifndef PDLP_SMAX
#define PDLP_SMAX 8       // upper bound on ring stages per warp (the plan picks S <= SMAX from the shared-memory budget)
#endif
#ifndef PDLP_CTAS
#define PDLP_CTAS 1       // CTAs per SM the shared-memory budget is split over
#endif
#ifndef PDLP_OVERSUB
#define PDLP_OVERSUB 1    // grid = resident CTAs x OVERSUB.  Measured on config 2: 1 -> 66 us, 2 -> 79 us, 4 -> 87 us per
#endif                    // iteration: extra waves cost more in pipeline ramp-up than they gain in balance
#ifndef PDLP_SMEM_BUDGET
#define PDLP_SMEM_BUDGET (128 * 1024)   // shared memory per SM given to the rings.  The rest of the 228 KB stays L1, which
#endif                                  // holds the gathers in flight: 64-128 KB -> 66 us, 160 KB -> 68 us, 200 KB -> 108 us
#ifndef PDLP_KB
#define PDLP_KB 8         // steps (elements per lane) per piece
#endif
#ifndef PDLP_U
#define PDLP_U 8          // gathers in flight per lane
#endif
#ifndef PDLP_MINB
#define PDLP_MINB 1       // CTAs per SM the register budget is sized for
#endif
#ifndef PDLP_PREFETCH
#define PDLP_PREFETCH 0   // pieces a warp prefetches into L2 ahead of its ring (0 = off): the ring's bulk copies then hit
#endif                    // L2, so the bytes in flight towards HBM are no longer bounded by the ring's shared memory
#ifndef PDLP_SIGMA
#define PDLP_SIGMA 128    // smallest sorting window (rows) the planner considers
#endif
#ifndef PDLP_LONG_ROW
#define PDLP_LONG_ROW 8192   // rows above this: one row per CTA, straight from CSR
#endif
#ifndef PDLP_WARP_ROW
#define PDLP_WARP_ROW 256    // rows above this (up to PDLP_LONG_ROW): one row per WARP, 32 equal segments, one per lane
#endif

struct Piece {            // 16 bytes
  uint32_t data_off16;    // byte offset / 16 of the block in the packed array
  int32_t perm_off;       // index into perm[] of the piece's first slice
  uint16_t klen;          // steps per slice inside this piece
  uint16_t nslices;       // slices in this piece (1 for a piece of a long slice)
  uint32_t flags;         // PIECE_FIRST: rows start here; PIECE_LAST: rows end here; PIECE_WARP_ROW: the 32 lanes
};                        // hold 32 segments of ONE row (their sums are combined before the row epilogue)
static_assert(sizeof(Piece) == 16, "Piece must be 16 bytes");
constexpr uint32_t PIECE_FIRST = 1u, PIECE_LAST = 2u, PIECE_WARP_ROW = 4u;

constexpr int MAX_RED = 24;

template <typename T>
struct SellView {
  const unsigned char* packed;   // piece blocks
  const Piece* pieces;
  const int* warp_piece_begin;   // gridDim.x * W + 1 entries
  // long rows (CSR)
  const int* long_rows;
  int n_long;
  const int* rowptr;
  const int* colidx;
  const T* vals;
  double* block_partials;        // gridDim.x * MAX_RED
  unsigned int* ticket;
  int stage_bytes;               // ring stage size: the largest piece of this matrix, rounded up to 128
  int nstages;                   // ring stages per warp (2 .. SMAX)
  int grid;                      // CTAs the plan was built for (warp_piece_begin has grid * W + 1 entries)
  // pdlp_time_kernels_in_loop: {start, end} of this launch in %globaltimer nanoseconds -- start when CTA 0 leaves
  // griddepcontrol.wait (the previous kernel has completed), end after the last CTA's finalize; null otherwise
  unsigned long long* stamp = nullptr;
};

template <typename T>
struct StreamCfg {
  static constexpr int W = PDLP_W;
  static constexpr int NT = W * 32;
  static constexpr int SMAX = PDLP_SMAX;
  static constexpr int KB = PDLP_KB;
  static constexpr int U = PDLP_U;
  static constexpr int MINB = PDLP_MINB;
  static constexpr int MAXSL = 16;                 // slices per piece (short rows)
  static constexpr int MAX_STAGE_BYTES = 16 + MAXSL * 128 + KB * 32 * (4 + (int)sizeof(T));
  static constexpr int MAX_SMEM_BYTES = 227 * 1024;
};

// developer experiments (never defined in product builds): isolate pipeline components when timing
#ifdef PDLP_DBG_NOGATHER
#define PDLP_COL(c) ((c) & 0)
#else
#define PDLP_COL(c) (c)
#endif

// ---------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk async copy (TMA)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// The matrix is read exactly once per pass: stream it through L2 with an evict-first policy so that it does not
// push out the gathered vectors (x pairs / y), which are re-read nnz times and should stay L2-resident.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
#ifdef PDLP_NO_L2_HINT
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
#else
  const uint64_t pol = l2_evict_first_policy();
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
#endif
}

// HBM -> L2 only: no shared memory, no completion to wait for
__device__ __forceinline__ void prefetch_l2(const void* gmem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

template <typename T>
__device__ __forceinline__ T shfl_xor_t(T v, int mask) {
  return __shfl_xor_sync(0xffffffffu, v, mask);
}

// ---------------------------------------------------------------------------
// The kernel.  Op contract:
//   using value_type = T;   using Gather = ...;   using Ctx = ...;
//   static constexpr int NV;      accumulators per row
//   static constexpr int NRED;    grid-wide reductions (<= MAX_RED; 0 = none)
//   static constexpr int UDIV;    divides the gather batch (wide Gather types keep registers in check)
//   static constexpr int PIPE;    1: pre-issue the next piece's first gather batch (costs U Gather registers)
//   __device__ bool begin();                         every thread, first; false -> the whole grid exits
//   __device__ Gather gather(int col) const;         operand[col] through the read-only path
//   __device__ void mac(T (&acc)[NV], T a, const Gather& g) const;     acc += a * g
//   __device__ Ctx row_load(int row) const;          request what the row epilogue needs (no use yet)
//   __device__ void row_finish(int row, const Ctx&, const T (&acc)[NV], double (&red)[max(NRED,1)]);
//   __device__ void finalize(const double (&tot)[max(NRED,1)]);        one thread of the last CTA
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ uint32_t piece_bytes(const Piece& d) {
  return 16u + ((d.flags & PIECE_FIRST) ? (uint32_t)d.nslices * 128u : 0u) +
         (uint32_t)d.klen * (uint32_t)d.nslices * 32u * (4u + (uint32_t)sizeof(T));
}

// Per-warp ring state.  `base` counts the pieces this warp has consumed since the barriers were initialised, so the
// ring position and the mbarrier parity carry over from one phase (matrix) to the next inside a persistent kernel.
struct Ring {
  unsigned char* mem;     // this warp's stages
  uint64_t* bars;         // this warp's "full" barriers
  int S;                  // stages
  int stage_bytes;
  unsigned base;
  __device__ __forceinline__ int stage(unsigned seq) const { return (int)(seq % (unsigned)S); }
  __device__ __forceinline__ uint32_t parity(unsigned seq) const { return (seq / (unsigned)S) & 1u; }
};

// Fill the (empty) ring with the first pieces of a phase.  Touches only the constant plan: may run ahead of a
// grid-wide dependency (PDL prologue, or before the grid barrier of the persistent kernel).
template <typename T>
__device__ __forceinline__ void ring_prologue(const SellView<T>& A, const Ring& R, int pb, int pe, int lane) {
  if (lane == 0) {
    for (int i = 0; i < R.S && pb + i < pe; ++i) {
      const Piece d = A.pieces[pb + i];
      const uint32_t bytes = piece_bytes<T>(d);
      const int s = R.stage(R.base + i);
      mbar_arrive_expect_tx(&R.bars[s], bytes);
      tma_load_1d(R.mem + (size_t)s * R.stage_bytes, A.packed + (size_t)d.data_off16 * 16u, bytes, &R.bars[s]);
    }
    if (PDLP_PREFETCH > 0 && pb + R.S < pe) {     // pieces are contiguous in memory: one prefetch for the next few
      const int last = min(pe, pb + R.S + PDLP_PREFETCH) - 1;
      const Piece d0 = A.pieces[pb + R.S], d1 = A.pieces[last];
      const unsigned char* p0 = A.packed + (size_t)d0.data_off16 * 16u;
      const unsigned char* p1 = A.packed + (size_t)d1.data_off16 * 16u + piece_bytes<T>(d1);
      prefetch_l2(p0, (uint32_t)(p1 - p0));
    }
  }
}

// Rows too long for the streaming layouts: one row per CTA, straight from the CSR arrays, U independent gathers per
// thread and batch; the row epilogue runs on thread 0.  s_red: W x SR doubles of shared scratch.
template <class Op, int SR>
__device__ __forceinline__ void long_rows_phase(const int* long_rows, int n_long, const int* rowptr, const int* colidx,
                                                const typename Op::value_type* vals, Op& op,
                                                double (&red)[Op::NRED > 0 ? Op::NRED : 1], double (*s_red)[SR]) {
  using T = typename Op::value_type;
  using Cfg = StreamCfg<T>;
  constexpr int NV = Op::NV;
  constexpr int U = (Cfg::U / Op::UDIV) > 0 ? (Cfg::U / Op::UDIV) : 1;
  const int tid = threadIdx.x;
  const int nthr = blockDim.x;
  const int nwarps = nthr >> 5;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  for (int li = blockIdx.x; li < n_long; li += gridDim.x) {
    const int row = long_rows[li];
    const int b = rowptr[row], e = rowptr[row + 1];
    T acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = T(0);
    for (int i = b + tid; i < e; i += U * nthr) {       // U independent gathers per thread and batch
      typename Op::Gather g[U];
      T a[U];
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int idx = i + j * nthr;
        if (idx < e) {
          g[j] = op.gather(PDLP_COL(colidx[idx]));
          a[j] = vals[idx];
        }
      }
#pragma unroll
      for (int j = 0; j < U; ++j)
        if (i + j * nthr < e) op.mac(acc, a[j], g[j]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[v] += shfl_xor_t(acc[v], o);
    }
    __syncthreads();      // s_red free
    if (lane == 0) {
#pragma unroll
      for (int v = 0; v < NV; ++v) s_red[warp][v] = (double)acc[v];
    }
    __syncthreads();
    if (tid == 0) {
      T tot[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += s_red[w][v];
        tot[v] = (T)t;
      }
      op.row_finish(row, op.row_load(row), tot, red);
    }
  }
}

// One pass over a matrix: long rows (CTA-cooperative, from CSR), then this warp's pieces.  `red` accumulates the
// op's per-lane reduction terms.  s_red: W x SR doubles of shared scratch for the long rows.
template <class Op, int SR>
__device__ __forceinline__ void stream_phase(const SellView<typename Op::value_type>& A, Op& op,
                                             double (&red)[Op::NRED > 0 ? Op::NRED : 1], Ring& R, int pb, int pe,
                                             double (*s_red)[SR]) {
  using T = typename Op::value_type;
  using Cfg = StreamCfg<T>;
  constexpr int NV = Op::NV;
  constexpr int U = (Cfg::U / Op::UDIV) > 0 ? (Cfg::U / Op::UDIV) : 1;
  const int tid = threadIdx.x;
  const int lane = tid & 31;

  // ---------------- long rows: one row per CTA, straight from CSR ----------------
  if (A.n_long > 0) long_rows_phase<Op, SR>(A.long_rows, A.n_long, A.rowptr, A.colidx, A.vals, op, red, s_red);

  // ---------------- SELL pieces: one row per lane ----------------
  // Software pipeline: while piece p is multiplied, reduced into its row and finished, the first batch of
  // gathers of piece p + 1 is already in flight (g_nxt), so a lane keeps up to 2 U gathers outstanding.
  T acc[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) acc[v] = T(0);
  typename Op::Ctx ctx;
  int row = -1;
  constexpr bool PIPE = Op::PIPE != 0;
  typename Op::Gather g_cur[U], g_nxt[PIPE ? U : 1];
  const unsigned base = R.base;
  auto stage_of = [&](int p) { return R.mem + (size_t)R.stage(base + (unsigned)(p - pb)) * R.stage_bytes; };
  // first batch of a piece: slice 0, steps [0, min(U, klen))
  auto issue_first = [&](int p, typename Op::Gather (&g)[U]) {
    const unsigned seq = base + (unsigned)(p - pb);
    mbar_wait(&R.bars[R.stage(seq)], R.parity(seq));
    const unsigned char* st = stage_of(p);
    const Piece d = *reinterpret_cast<const Piece*>(st);
    const int* cols = reinterpret_cast<const int*>(st + 16) + ((d.flags & PIECE_FIRST) ? (int)d.nslices * 32 : 0);
    const int klen = d.klen;
#pragma unroll
    for (int j = 0; j < U; ++j)
      if (j < klen) g[j] = op.gather(PDLP_COL(cols[j * 32 + lane]));
  };
  if (PIPE && pb < pe) issue_first(pb, g_cur);
  for (int p = pb; p < pe; ++p) {
    if (!PIPE) issue_first(p, g_cur);
    Piece d_refill = Piece{0, 0, 0, 0, 0}, d_pref = Piece{0, 0, 0, 0, 0};
    const bool refill = (p + R.S < pe);
    const bool pref = PDLP_PREFETCH > 0 && (p + R.S + PDLP_PREFETCH < pe);
    if (refill && lane == 0) d_refill = A.pieces[p + R.S];        // used after this piece: latency hidden
    if (pref && lane == 0) d_pref = A.pieces[p + R.S + PDLP_PREFETCH];
    const int s = R.stage(base + (unsigned)(p - pb));
    const unsigned char* st = stage_of(p);
    const Piece d = *reinterpret_cast<const Piece*>(st);          // stage already waited for by issue_first
    const int klen = d.klen;
    const int nsl = d.nslices;
    const bool first = (d.flags & PIECE_FIRST) != 0;
    const int* rows = reinterpret_cast<const int*>(st + 16);
    const int* cols = rows + (first ? nsl * 32 : 0);
    const T* vals = reinterpret_cast<const T*>(cols + klen * nsl * 32);
    if (first) {
      row = rows[lane];
#ifndef PDLP_DBG_NOEPI
      if (row >= 0) ctx = op.row_load(row);
#endif
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = T(0);
    }
    if constexpr (PIPE) {
      if (p + 1 < pe) issue_first(p + 1, g_nxt);                  // next piece's gathers fly during the work below
    }
#ifndef PDLP_DBG_NOWORK
    for (int sl = 0; sl < nsl; ++sl) {
      if (sl > 0) {       // further complete slices packed into this piece
        row = rows[sl * 32 + lane];
#ifndef PDLP_DBG_NOEPI
        if (row >= 0) ctx = op.row_load(row);
#endif
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = T(0);
      }
      const int bse = sl * klen;
      int k0 = 0;
      if (sl == 0) {      // the pre-issued batch
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (j < klen) op.mac(acc, vals[j * 32 + lane], g_cur[j]);
        k0 = U;
      }
      for (; k0 < klen; k0 += U) {
        typename Op::Gather g[U];
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (k0 + j < klen) g[j] = op.gather(PDLP_COL(cols[(bse + k0 + j) * 32 + lane]));
#pragma unroll
        for (int j = 0; j < U; ++j)
          if (k0 + j < klen) op.mac(acc, vals[(bse + k0 + j) * 32 + lane], g[j]);
      }
#ifndef PDLP_DBG_NOEPI
      if (d.flags & PIECE_WARP_ROW) {           // warp-uniform: a medium-long row spread over the 32 lanes
        if (d.flags & PIECE_LAST) {
#pragma unroll
          for (int v = 0; v < NV; ++v) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[v] += shfl_xor_t(acc[v], o);
          }
          if (lane == 0) op.row_finish(row, ctx, acc, red);
        }
      } else if ((d.flags & PIECE_LAST) && row >= 0) op.row_finish(row, ctx, acc, red);
#else
      if ((d.flags & PIECE_LAST) && row >= 0) red[0] += (double)acc[0];
#endif
    }
#endif
    __syncwarp();         // every lane is done reading stage s
    if (refill && lane == 0) {
      const uint32_t bytes = piece_bytes<T>(d_refill);
      mbar_arrive_expect_tx(&R.bars[s], bytes);
      tma_load_1d(R.mem + (size_t)s * R.stage_bytes, A.packed + (size_t)d_refill.data_off16 * 16u, bytes, &R.bars[s]);
    }
    if (pref && lane == 0) prefetch_l2(A.packed + (size_t)d_pref.data_off16 * 16u, piece_bytes<T>(d_pref));
    if constexpr (PIPE) {
#pragma unroll
      for (int j = 0; j < U; ++j) g_cur[j] = g_nxt[j];
    }
  }
  R.base = base + (unsigned)(pe - pb);
}

// lane-level terms -> one value per CTA in s_out[0..NRED) (fixed order: xor tree, then warps in order)
template <int NRED, int SR>
__device__ __forceinline__ void block_reduce(double (&red)[NRED > 0 ? NRED : 1], double (*s_red)[SR], double* cta_out) {
  using namespace pdlp;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int W = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NRED; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) red[i] += shfl_xor_t(red[i], o);
  }
  __syncthreads();   // earlier users of s_red are done
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NRED; ++i) s_red[warp][i] = red[i];
  }
  __syncthreads();
  if (tid == 0) {
    for (int i = 0; i < NRED; ++i) {
      double t = 0.0;
      for (int w = 0; w < W; ++w) t += s_red[w][i];
      cta_out[i] = t;
    }
  }
}

// per-CTA values of every CTA -> grid totals, by one warp, fixed order: lane l adds blocks l, l + 32, l + 64, ... from
// +0.0, then the xor tree.  The loads are issued in batches of up to 8 sums x 5 blocks per lane before the first add --
// ONE L2 round trip per batch instead of one per dependent add (this runs in the tail of a kernel, on the critical
// path to the next one); an absent block contributes +0.0, which leaves a sum that started at +0.0 unchanged, so the
// result equals the plain loop's bit for bit.
// (BATCH = false: the plain loop, for callers that have no registers to spare -- the persistent block kernel)
template <int NRED, bool BATCH = true>
__device__ __forceinline__ void grid_totals(const double* partials, int stride, int nblocks, double (&tot)[NRED > 0 ? NRED : 1]) {
  const int lane = threadIdx.x & 31;
  if constexpr (!BATCH) {
    for (int i = 0; i < NRED; ++i) {
      double t = 0.0;
      for (int b = lane; b < nblocks; b += 32) t += __ldcg(&partials[(size_t)b * stride + i]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += shfl_xor_t(t, o);
      tot[i] = t;
    }
    return;
  }
  constexpr int NB = 5;                    // 5 x 32 = 160 >= one CTA per SM on B200 (148)
  constexpr int CH = NRED < 8 ? (NRED > 0 ? NRED : 1) : 8;
#pragma unroll
  for (int i0 = 0; i0 < NRED; i0 += CH) {
    double t[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) t[i] = 0.0;
    for (int c0 = 0; c0 < nblocks; c0 += 32 * NB) {
      double v[CH][NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        const int b = c0 + lane + 32 * j;
#pragma unroll
        for (int i = 0; i < CH; ++i)
          v[i][j] = (i0 + i < NRED && b < nblocks) ? __ldcg(&partials[(size_t)b * stride + i0 + i]) : 0.0;
      }
#pragma unroll
      for (int i = 0; i < CH; ++i) {
#pragma unroll
        for (int j = 0; j < NB; ++j) t[i] += v[i][j];
      }
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      if (i0 + i < NRED) {
        double r = t[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += shfl_xor_t(r, o);
        tot[i0 + i] = r;
      }
    }
  }
}

// Optional members of an op, detected here:
//   bool defer;                 true: the kernel stops after the per-CTA partial sums (no ticket, no finalize) -- the
//                               next kernel's op adds them up itself (PrimalOp -> DualOp: ||dx||^2 is only needed by the
//                               step rule, so K1's tail loses an atomic round trip, a re-read and a dependent store)
//   void pre_finalize();        run by the whole first warp of the last CTA before lane 0 calls finalize
//   void pre_finalize_issue();  same warp, BEFORE the partial sums are read: loads / prefetches the finalize will need,
//                               so that they share the round trip of the partial sums
template <class Op, class = void> struct op_has_defer : std::false_type {};
template <class Op> struct op_has_defer<Op, std::void_t<decltype(std::declval<const Op&>().defer)>> : std::true_type {};
template <class Op, class = void> struct op_has_pre_finalize : std::false_type {};
template <class Op>
struct op_has_pre_finalize<Op, std::void_t<decltype(std::declval<Op&>().pre_finalize())>> : std::true_type {};
template <class Op, class = void> struct op_has_pre_finalize_issue : std::false_type {};
template <class Op>
struct op_has_pre_finalize_issue<Op, std::void_t<decltype(std::declval<Op&>().pre_finalize_issue())>> : std::true_type {};

template <class Op>
__global__ void __launch_bounds__(StreamCfg<typename Op::value_type>::NT, StreamCfg<typename Op::value_type>::MINB)
stream_kernel(const SellView<typename Op::value_type> A, Op op) {
  using T = typename Op::value_type;
  using Cfg = StreamCfg<T>;
  constexpr int NV = Op::NV;
  constexpr int NRED = Op::NRED;
  constexpr int NR = NRED > 0 ? NRED : 1;
  constexpr int SR = NR > NV ? NR : NV;

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[Cfg::W][Cfg::SMAX];
  __shared__ double s_red[Cfg::W][SR];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  Ring R{smem + (size_t)warp * A.nstages * A.stage_bytes, &full_bar[warp][0], A.nstages, A.stage_bytes, 0u};
  if (lane == 0) {
    for (int s = 0; s < R.S; ++s) mbar_init(&R.bars[s], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncwarp();

  double red[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) red[i] = 0.0;

  const int gw = blockIdx.x * Cfg::W + warp;
  const int pb = A.warp_piece_begin[gw];
  const int pe = A.warp_piece_begin[gw + 1];
  ring_prologue<T>(A, R, pb, pe, lane);

  // Programmatic dependent launch: everything above touched only the (constant) matrix plan, so it overlaps the
  // tail of the previous kernel in the stream / graph.  From here on we read what that kernel wrote.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (A.stamp && blockIdx.x == 0 && tid == 0) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    A.stamp[0] = now;
  }
  if (!op.begin()) {
    // drain the copies already in flight before the CTA may retire
    for (int s = 0; s < R.S && pb + s < pe; ++s) mbar_wait(&R.bars[s], 0);
    return;
  }

  stream_phase<Op, SR>(A, op, red, R, pb, pe, s_red);

  // ---------------- deterministic grid reduction + finalize ----------------
  if constexpr (NRED > 0) {
    block_reduce<NRED, SR>(red, s_red, &A.block_partials[(size_t)blockIdx.x * MAX_RED]);
    if constexpr (op_has_defer<Op>::value) {
      if (op.defer) {
        if (A.stamp && tid == 0) {      // timing run: the latest CTA's clock is the end of the launch
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          atomicMax(A.stamp + 1, now);
        }
        return;
      }
    }
    if (tid == 0) {
      __threadfence();
      const unsigned int ticket = atomicAdd(A.ticket, 1u);
      s_last = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last && warp == 0) {
      __threadfence();
      double tot[NR];
      if constexpr (op_has_pre_finalize_issue<Op>::value) op.pre_finalize_issue();
      grid_totals<NRED>(A.block_partials, MAX_RED, (int)gridDim.x, tot);
      if constexpr (op_has_pre_finalize<Op>::value) op.pre_finalize();
      if (lane == 0) {
        op.finalize(tot);
        *A.ticket = 0u;
        if (A.stamp) {
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          A.stamp[1] = now;
        }
      }
    }
  }
}

}  // namespace pdlp
