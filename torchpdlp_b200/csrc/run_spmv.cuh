// run_spmv.cuh -- the "run" layout and its kernels: the matrix as ONE row-major stream of nonzeros, cut into
// fixed-size chunks, rows in their natural order.
//
// Why a second layout next to SELL-32-sigma (stream_spmv.cuh): on B200 the iteration kernels are bound by the number
// of 32-byte sectors that cross the L2 -> SM fabric, not by DRAM bytes (profiles/r1_ncu_full_iteration_kernels.txt:
// DRAM 27 %, nothing saturated, ~1 sector operation per clock and SM).  SELL pays for its one-row-per-lane
// simplicity with sigma-sorted rows, i.e. a row epilogue whose 5-8 vector accesses per row are scattered (one
// sector each), with slot padding, and with row-id arrays.  Here nothing is sorted and nothing is padded:
//
//   stream    entries of row 0, row 1, ... back to back; an empty row is one zero entry; a row that is too long for a
//             warp (handled one row per CTA from CSR, long_rows_phase) is left out and ends the current chunk, so
//             that the rows that END inside a chunk are always CONSECUTIVE: row0, row0 + 1, ...
//   chunk     32 lanes x E steps = 256 entries; lane r owns E consecutive entries of the stream (positions r E ..
//             r E + E - 1), stored column-major (step k of lane r at [k * 32 + r]) so that shared-memory reads are
//             conflict-free.  Block = [vals 256 x T][cols 256 x CT][32 lane words][16-byte trailer]: fixed size, ONE
//             1-D TMA bulk copy.  The top bit of a column index marks the LAST entry of its row; what the walk needs
//             to know about the END bits (per-lane mask, staging slot, rows per chunk) is precomputed by the packer.
//   warps     own contiguous chunk ranges cut at row boundaries (the tail of the last chunk is zero padding), so no
//             row ever spans two warps and no atomics or fix-up passes are needed.
//
// Per chunk a lane issues its E gathers, walks its run (acc += a * g; at an END bit: stage the sum, reset), a
// warp-level segmented scan carries partial rows from lane to lane (and chunk to chunk), the first END of every lane
// receives the carry, and then the finished rows -- consecutive indices -- are handed out one per lane for the fused
// row epilogue: every vector access of the epilogue is COALESCED.  Work per warp is equal by construction (power-law
// rows need no warp-row / CTA-row classes below LONG_ROW), sums have a fixed order: bitwise reproducible.
//
// Measured on config 2 (round 2): the walk's serial chain per chunk (gathers -> walk -> scan -> staging -> epilogue)
// with 12-16 warps per SM costs more than the coalesced epilogue saves -- K2 45 us against 37 us for SELL -- so the
// engine uses this layout only on request (PDLP_LAYOUT=run); it is the layout of choice for row-length distributions
// that defeat the slice padding of SELL.
#pragma once
#include "stream_spmv.cuh"

namespace pdlp {

#ifndef PDLP_RUN_W
#define PDLP_RUN_W 12          // warps per CTA
#endif
#ifndef PDLP_RUN_E
#define PDLP_RUN_E 8           // entries per lane and chunk
#endif
#ifndef PDLP_RUN_CAP
#define PDLP_RUN_CAP 128       // finished rows staged per epilogue round (a chunk can finish up to 32 E rows)
#endif
#ifndef PDLP_RUN_BUDGET
#define PDLP_RUN_BUDGET (120 * 1024)   // shared memory per SM given to the chunk rings (the rest stays L1 for the gathers)
#endif
#ifndef PDLP_RUN_PIPE
#define PDLP_RUN_PIPE 1        // 1: the next chunk's gathers are issued before the current chunk is multiplied
#endif

template <typename T, typename CT>
struct RunCfg {
  static constexpr int E = PDLP_RUN_E;
  static constexpr int CHUNK = 32 * E;
  static constexpr int CAP = PDLP_RUN_CAP;
  static constexpr int VALS_BYTES = CHUNK * (int)sizeof(T);
  static constexpr int COLS_BYTES = CHUNK * (int)sizeof(CT);
  static constexpr int META_BYTES = 32 * 4;                           // one word per lane, see RunMeta
  static constexpr int CHUNK_BYTES = VALS_BYTES + COLS_BYTES + META_BYTES + 16;
  static constexpr int STAGE_BYTES = (CHUNK_BYTES + 127) / 128 * 128;
  static constexpr CT END = (CT)((CT)1 << (8 * sizeof(CT) - 1));      // last entry of its row
  static constexpr CT MASK = (CT)(END - 1);
  static_assert(E <= 8, "the per-lane END mask has 8 bits");
  static_assert(CAP % 32 == 0 && CAP >= 32, "CAP must be a multiple of the warp size");
};

// Everything the chunk walk would otherwise derive from the END bits with ballots and popcounts is computed once by
// the packer.  Per lane (one word): bits 0-7 = END mask of the lane's E entries, bits 8-16 = number of ENDs in the
// lanes before it (= staging slot of the lane's first END).  Per chunk (the 16-byte trailer): the first row that ends
// in the chunk, how many rows end in it, the steps at which any lane has an END, the lanes that have one.
struct RunTrailer {
  int row0;
  int nrows;
  unsigned any_steps;
  unsigned end_lanes;
};
static_assert(sizeof(RunTrailer) == 16, "RunTrailer must be 16 bytes");

template <typename T>
struct RunView {
  const unsigned char* chunks;
  const int* warp_chunk_begin;   // grid * W + 1
  const int* long_rows;          // rows handled one per CTA from CSR (the stream leaves them out)
  int n_long;
  const int* rowptr;
  const int* colidx;
  const T* vals;
  double* block_partials;        // grid * MAX_RED
  unsigned int* ticket;
  int nstages;
  int grid;
};

// ---------------------------------------------------------------------------
// the chunk walk of one warp
// ---------------------------------------------------------------------------
// The ring position of a warp: stage index and mbarrier parity advance together (no division in the loop).
struct RingPos {
  int s;
  unsigned par;
  __device__ __forceinline__ void advance(int S) {
    if (++s == S) { s = 0; par ^= 1u; }
  }
};

template <typename T, typename CT>
__device__ __forceinline__ void run_ring_prologue(const unsigned char* chunks, const Ring& R, int cb, int ce, int lane) {
  using Cfg = RunCfg<T, CT>;
  if (lane == 0) {
    for (int i = 0; i < R.S && cb + i < ce; ++i) {
      mbar_arrive_expect_tx(&R.bars[i], Cfg::CHUNK_BYTES);
      tma_load_1d(R.mem + (size_t)i * R.stage_bytes, chunks + (size_t)(cb + i) * Cfg::CHUNK_BYTES, Cfg::CHUNK_BYTES, &R.bars[i]);
    }
    if (PDLP_PREFETCH > 0 && cb + R.S < ce)       // HBM -> L2 for the chunks behind the ring
      prefetch_l2(chunks + (size_t)(cb + R.S) * Cfg::CHUNK_BYTES, (uint32_t)(min(ce - cb - R.S, PDLP_PREFETCH) * Cfg::CHUNK_BYTES));
  }
}

// One pass of a warp over its chunks [cb, ce).  The ring must be empty and at stage 0 / parity 0 (one pass per
// kernel).  stg: CAP x NV values of per-warp shared scratch.
template <class Op, typename CT>
__device__ __forceinline__ void run_phase(const unsigned char* chunks, Op& op, double (&red)[Op::NRED > 0 ? Op::NRED : 1],
                                          const Ring& R, int cb, int ce, typename Op::value_type* stg) {
  using T = typename Op::value_type;
  using Cfg = RunCfg<T, CT>;
  constexpr int NV = Op::NV;
  constexpr int E = Cfg::E;
  constexpr int CAP = Cfg::CAP;
  constexpr bool PIPE = PDLP_RUN_PIPE == 2 || (PDLP_RUN_PIPE != 0 && Op::PIPE != 0);   // 2: also for wide gathers
  const int lane = threadIdx.x & 31;
  const int S = R.S;

  // wait for the chunk in stage p.s and put its E gathers in flight
  auto issue = [&](const RingPos& p, typename Op::Gather (&g)[E]) {
    mbar_wait(&R.bars[p.s], p.par);
    const CT* cols = reinterpret_cast<const CT*>(R.mem + (size_t)p.s * R.stage_bytes + Cfg::VALS_BYTES);
#pragma unroll
    for (int k = 0; k < E; ++k) g[k] = op.gather(PDLP_COL((int)(cols[k * 32 + lane] & Cfg::MASK)));
  };

  T carry[NV];            // the unfinished row's partial sum entering the next chunk (same in all lanes)
#pragma unroll
  for (int v = 0; v < NV; ++v) carry[v] = T(0);

  // one chunk (in stage p.s); g: its gathers (in flight), g_next: where the next chunk's gathers go
  auto step = [&](int c, const RingPos& p, typename Op::Gather (&g)[E], typename Op::Gather (&g_next)[E]) {
    const unsigned char* st = R.mem + (size_t)p.s * R.stage_bytes;
    const T* vals = reinterpret_cast<const T*>(st);
    const unsigned meta = reinterpret_cast<const unsigned*>(st + Cfg::VALS_BYTES + Cfg::COLS_BYTES)[lane];
    const int4 tr4 = *reinterpret_cast<const int4*>(st + Cfg::VALS_BYTES + Cfg::COLS_BYTES + Cfg::META_BYTES);
    const int row0 = tr4.x, nrows = tr4.y;
    const unsigned any_steps = (unsigned)tr4.z, F = (unsigned)tr4.w;
    const unsigned endm = meta & 0xffu;
    const int below = (int)((meta >> 8) & 0x1ffu);
    // epilogue operands of the first 32 finished rows: requested now, used after the walk
    typename Op::Ctx ctx0;
    if (lane < nrows) ctx0 = op.row_load(row0 + lane);
    if constexpr (PIPE) {
      if (c + 1 < ce) {
        RingPos q = p;
        q.advance(S);
        issue(q, g_next);
      }
    }
    T cin[NV];            // carry into this lane's first END
    for (int lo = 0; lo == 0 || lo < nrows; lo += CAP) {
      T acc[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) acc[v] = T(0);
      int slot = below - lo;
#pragma unroll
      for (int k = 0; k < E; ++k) {
        op.mac(acc, vals[k * 32 + lane], g[k]);
        if ((any_steps >> k) & 1u) {                  // warp-uniform
          if ((endm >> k) & 1u) {
            if ((unsigned)slot < (unsigned)CAP) {
#pragma unroll
              for (int v = 0; v < NV; ++v) stg[slot * NV + v] = acc[v];
            }
            ++slot;
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] = T(0);
          }
        }
      }
      if (lo == 0) {
        // segmented inclusive scan over the lanes: S_r = has_end_r ? tail_r : S_{r-1} + tail_r, S_{-1} = carry;
        // the carry into lane r is S_{r-1}
        T sv[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) sv[v] = (lane == 0 && !endm) ? carry[v] + acc[v] : acc[v];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          // lane r's window is (r - d, r]; it absorbs lane r - d's window unless an END lies in its own
          const unsigned wm = ((1u << d) - 1u) << ((lane - d + 1) & 31);
          const bool take = lane >= d && !(F & wm);
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            const T up = __shfl_up_sync(0xffffffffu, sv[v], d);
            if (take) sv[v] = up + sv[v];
          }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const T up = __shfl_up_sync(0xffffffffu, sv[v], 1);
          cin[v] = lane == 0 ? carry[v] : up;
          carry[v] = __shfl_sync(0xffffffffu, sv[v], 31);
        }
      }
      const int first = below - lo;                   // slot of this lane's first END (if it has one)
      if (endm && (unsigned)first < (unsigned)CAP) {
#pragma unroll
        for (int v = 0; v < NV; ++v) stg[first * NV + v] = cin[v] + stg[first * NV + v];
      }
      __syncwarp();
      // the finished rows of this round, one per lane, consecutive row indices: coalesced epilogue
      const int cnt = min(CAP, nrows - lo);
      for (int i = lane; i < cnt; i += 32) {
        T a[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) a[v] = stg[i * NV + v];
        const int row = row0 + lo + i;
        if (lo == 0 && i < 32) op.row_finish(row, ctx0, a, red);
        else op.row_finish(row, op.row_load(row), a, red);
      }
      __syncwarp();         // staging and (last round) the stage are free
    }
    if (c + S < ce && lane == 0) {
      mbar_arrive_expect_tx(&R.bars[p.s], Cfg::CHUNK_BYTES);
      tma_load_1d(R.mem + (size_t)p.s * R.stage_bytes, chunks + (size_t)(c + S) * Cfg::CHUNK_BYTES, Cfg::CHUNK_BYTES,
                  &R.bars[p.s]);
    }
    if (PDLP_PREFETCH > 0 && c + S + PDLP_PREFETCH < ce && lane == 0)
      prefetch_l2(chunks + (size_t)(c + S + PDLP_PREFETCH) * Cfg::CHUNK_BYTES, Cfg::CHUNK_BYTES);
  };

  typename Op::Gather ga[E], gb[E];
  RingPos p{0, 0u};
  if constexpr (PIPE) {
    if (cb < ce) issue(p, ga);
    int c = cb;
    for (; c + 1 < ce; c += 2) {
      step(c, p, ga, gb);
      p.advance(S);
      step(c + 1, p, gb, ga);
      p.advance(S);
    }
    if (c < ce) step(c, p, ga, gb);
  } else {
    for (int c = cb; c < ce; ++c) {
      issue(p, ga);
      step(c, p, ga, gb);
      p.advance(S);
    }
  }
}

// shared-memory geometry of run_kernel / panel_kernel: [W x S stages][W x CAP x NV staging values][W x CAP flags]
template <typename T, typename CT, int NV, int W>
struct RunSmem {
  using Cfg = RunCfg<T, CT>;
  __host__ __device__ static constexpr int staging_bytes() { return W * Cfg::CAP * NV * (int)sizeof(T); }
  __host__ __device__ static int total(int nstages) { return W * nstages * Cfg::STAGE_BYTES + staging_bytes() + 128; }
};

template <class Op>
__global__ void __launch_bounds__(PDLP_RUN_W * 32, 1)
run_kernel(const RunView<typename Op::value_type> A, Op op) {
  using T = typename Op::value_type;
  using CT = uint32_t;
  using Cfg = RunCfg<T, CT>;
  constexpr int W = PDLP_RUN_W;
  constexpr int NV = Op::NV;
  constexpr int NRED = Op::NRED;
  constexpr int NR = NRED > 0 ? NRED : 1;
  constexpr int SR = NR > NV ? NR : NV;

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[W][PDLP_SMAX];
  __shared__ double s_red[W][SR];
  __shared__ int s_last;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  Ring R{smem + (size_t)warp * A.nstages * Cfg::STAGE_BYTES, &full_bar[warp][0], A.nstages, Cfg::STAGE_BYTES, 0u};
  T* stg = reinterpret_cast<T*>(smem + (size_t)W * A.nstages * Cfg::STAGE_BYTES) + (size_t)warp * Cfg::CAP * NV;
  if (lane == 0) {
    for (int s = 0; s < R.S; ++s) mbar_init(&R.bars[s], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncwarp();

  double red[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) red[i] = 0.0;

  const int gw = blockIdx.x * W + warp;
  const int cb = A.warp_chunk_begin[gw];
  const int ce = A.warp_chunk_begin[gw + 1];
  run_ring_prologue<T, CT>(A.chunks, R, cb, ce, lane);

  // programmatic dependent launch: everything above touched only the constant matrix stream
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!op.begin()) {
    for (int s = 0; s < R.S && cb + s < ce; ++s) mbar_wait(&R.bars[s], 0);     // drain the copies in flight
    return;
  }

  if (A.n_long > 0) long_rows_phase<Op, SR>(A.long_rows, A.n_long, A.rowptr, A.colidx, A.vals, op, red, s_red);
  run_phase<Op, CT>(A.chunks, op, red, R, cb, ce, stg);

  if constexpr (NRED > 0) {
    block_reduce<NRED, SR>(red, s_red, &A.block_partials[(size_t)blockIdx.x * MAX_RED]);
    if (tid == 0) {
      __threadfence();
      const unsigned int ticket = atomicAdd(A.ticket, 1u);
      s_last = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last && warp == 0) {
      __threadfence();
      double tot[NR];
      grid_totals<NRED>(A.block_partials, MAX_RED, (int)gridDim.x, tot);
      if (lane == 0) {
        op.finalize(tot);
        *A.ticket = 0u;
      }
    }
  }
}

}  // namespace pdlp
