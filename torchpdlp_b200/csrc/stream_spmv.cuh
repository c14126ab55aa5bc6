// stream_spmv.cuh -- the streaming CSR kernel skeleton every hot kernel is built on.
//
// One persistent CTA per (SM x occupancy) slot walks a contiguous, weight-balanced
// range of "chunks" (row ranges of the CSR matrix whose nonzeros fit one shared
// memory stage).  A dedicated producer warp streams each chunk's values, column
// indices and row pointers into a STAGES-deep shared-memory ring with 1-D TMA
// bulk copies (cp.async.bulk + mbarrier complete_tx, SASS UBLKCP); CONSUMER warps
// wait on the stage's "full" barrier, reduce their rows with sub-warp groups
// (gathering the dense operand through the read-only path), run the fused
// epilogue for the rows they own, and release the stage through an "empty"
// mbarrier -- no CTA-wide barrier inside the loop, so warps drift across stages
// and gather latency overlaps the stream.  Every reduction (row sums, block
// partials, grid totals) has a fixed order: no atomics touch data, results are
// bitwise reproducible for a given plan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pdlp {

// ---------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------
struct Chunk {        // 16 bytes, copied into the stage with the data
  int row_begin;
  int nnz_begin;
  int nnz_end;
  uint32_t meta;      // bits 0..15 rows, 16..19 log2(threads per row), bit 31: long row (streamed from global)
};
static_assert(sizeof(Chunk) == 16, "Chunk must be 16 bytes");

constexpr uint32_t CHUNK_LONG = 0x80000000u;

template <typename T>
struct CsrView {
  const int* rowptr;      // padded by >= 8 entries
  const int* colidx;      // padded by >= 8 entries, 16B aligned
  const T* vals;          // padded by >= 8 entries, 16B aligned
  const Chunk* chunks;
  const int* cta_chunk_begin;   // gridDim.x + 1 entries
  double* block_partials;       // gridDim.x * MAX_RED
  unsigned int* ticket;
  int rows;
};

constexpr int MAX_RED = 24;

// ---------------------------------------------------------------------------
// kernel configuration
// ---------------------------------------------------------------------------
#ifndef PDLP_NWARPS
#define PDLP_NWARPS 7
#endif
#ifndef PDLP_CAP
#define PDLP_CAP 1024
#endif
#ifndef PDLP_RCAP
#define PDLP_RCAP 288
#endif
#ifndef PDLP_STAGES
#define PDLP_STAGES 6
#endif
#ifndef PDLP_MINB
#define PDLP_MINB 2
#endif
#ifndef PDLP_U
#define PDLP_U 4
#endif

template <typename T>
struct StreamCfg {
  static constexpr int NWARPS = PDLP_NWARPS;       // consumer warps
  static constexpr int NCONS = NWARPS * 32;
  static constexpr int NT = NCONS + 32;            // + producer warp
  static constexpr int CAP = PDLP_CAP;             // nonzeros per stage (incl. alignment slack)
  static constexpr int RCAP = PDLP_RCAP;           // rows per stage
  static constexpr int STAGES = PDLP_STAGES;
  static constexpr int MINB = PDLP_MINB;           // CTAs per SM the register budget is sized for
  static constexpr int U = PDLP_U;                 // gathers in flight per lane and batch
  static constexpr int DESC_OFF = 0;
  static constexpr int RPTR_OFF = 128;
  static constexpr int COLS_OFF = RPTR_OFF + ((RCAP + 8) * 4 + 127) / 128 * 128;
  static constexpr int VALS_OFF = COLS_OFF + CAP * 4;
  static constexpr int STAGE_BYTES = (VALS_OFF + CAP * (int)sizeof(T) + 127) / 128 * 128;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 128;
  static constexpr int MAX_CHUNK_NNZ = CAP - 8;
  static constexpr int MAX_CHUNK_ROWS = RCAP - 8;
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void consumer_barrier(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

template <typename T>
__device__ __forceinline__ T shfl_xor_t(T v, int mask) {
  return __shfl_xor_sync(0xffffffffu, v, mask);
}

// ---------------------------------------------------------------------------
// The kernel.  Op contract:
//   using value_type = T;   using Gather = ...;   using Ctx = ...;
//   static constexpr int NV;      number of accumulators per row
//   static constexpr int NRED;    number of grid-wide reductions (<= MAX_RED; 0 = none)
//   static constexpr int UDIV;    divides the per-lane gather batch (wide Gather types keep registers in check)
//   __device__ bool begin();                         called by every thread first; false -> whole grid exits
//   __device__ Gather gather(int col) const;         operand[col] through the read-only path
//   __device__ void mac(T (&acc)[NV], T a, const Gather& g) const;     acc += a * g
//   __device__ Ctx row_load(int row) const;          issue the loads the row's epilogue needs (no use yet)
//   __device__ void row_finish(int row, const Ctx&, const T (&acc)[NV], double (&red)[max(NRED,1)]);
//   __device__ void finalize(const double (&tot)[max(NRED,1)]);        one thread of the last CTA
//
// Consumer warps are software-pipelined over chunks: the gathers and the epilogue operands of chunk i+1 are in
// flight while chunk i is multiplied, reduced (sub-warp shuffles, fixed order) and finished, so every warp keeps
// two chunks' worth of L2 gathers outstanding.
// ---------------------------------------------------------------------------
template <class Op>
struct Work {
  typename Op::Gather g[StreamCfg<typename Op::value_type>::U / Op::UDIV];
  typename Op::Ctx ctx;
  int row;     // global row of this lane's group in round 0
  int i0;      // stage index of the lane's first element
  int e;       // stage index one past the row's last element
  bool valid;
};

template <class Op>
__global__ void __launch_bounds__(StreamCfg<typename Op::value_type>::NT, StreamCfg<typename Op::value_type>::MINB)
stream_kernel(const CsrView<typename Op::value_type> A, Op op) {
  using T = typename Op::value_type;
  using Cfg = StreamCfg<T>;
  constexpr int NV = Op::NV;
  constexpr int NRED = Op::NRED;
  constexpr int NR = NRED > 0 ? NRED : 1;
  constexpr int U = Cfg::U / Op::UDIV > 0 ? Cfg::U / Op::UDIV : 1;   // gathers in flight per lane and batch

  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[Cfg::STAGES];
  __shared__ uint64_t empty_bar[Cfg::STAGES];
  __shared__ double s_red[Cfg::NWARPS][NR > NV ? NR : NV];
  __shared__ int s_last;

  if (!op.begin()) return;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int cb = A.cta_chunk_begin[blockIdx.x];
  const int ce = A.cta_chunk_begin[blockIdx.x + 1];

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], Cfg::NWARPS);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();

  double red[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) red[i] = 0.0;

  if (warp == Cfg::NWARPS) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      Chunk next = cb < ce ? A.chunks[cb] : Chunk{0, 0, 0, 0};
      for (int ci = cb; ci < ce; ++ci) {
        const int it = ci - cb;
        const int s = it % Cfg::STAGES;
        const int use = it / Cfg::STAGES;
        const Chunk ch = next;
        if (ci + 1 < ce) next = A.chunks[ci + 1];   // hide the descriptor latency behind the wait below
        if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1);
        unsigned char* st = smem + (size_t)s * Cfg::STAGE_BYTES;
        uint32_t bytes = 16;
        const bool is_long = (ch.meta & CHUNK_LONG) != 0;
        const int nrows = ch.meta & 0xffff;
        const int r0 = ch.row_begin & ~3;
        const int r1 = (ch.row_begin + nrows + 1 + 3) & ~3;
        const int e0 = ch.nnz_begin & ~3;
        const int e1 = (ch.nnz_end + 3) & ~3;
        const uint32_t rbytes = (uint32_t)(r1 - r0) * 4u;
        const uint32_t cbytes = (uint32_t)(e1 - e0) * 4u;
        const uint32_t vbytes = (uint32_t)(e1 - e0) * (uint32_t)sizeof(T);
        if (!is_long) bytes += rbytes + cbytes + vbytes;
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        tma_load_1d(st + Cfg::DESC_OFF, A.chunks + ci, 16, &full_bar[s]);
        if (!is_long) {
          tma_load_1d(st + Cfg::RPTR_OFF, A.rowptr + r0, rbytes, &full_bar[s]);
          if (e1 > e0) {
            tma_load_1d(st + Cfg::COLS_OFF, A.colidx + e0, cbytes, &full_bar[s]);
            tma_load_1d(st + Cfg::VALS_OFF, A.vals + e0, vbytes, &full_bar[s]);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ consumers ------------------------------
    // issue(): wait for the chunk, start this lane's first batch of gathers and the epilogue operand loads
    auto issue = [&](int ci, Work<Op>& w) {
      const int it = ci - cb;
      const int s = it % Cfg::STAGES;
      mbar_wait(&full_bar[s], (it / Cfg::STAGES) & 1);
      const unsigned char* st = smem + (size_t)s * Cfg::STAGE_BYTES;
      const Chunk ch = *reinterpret_cast<const Chunk*>(st + Cfg::DESC_OFF);
      w.valid = false;
      w.i0 = 0;
      w.e = 0;
      w.row = 0;
      if (ch.meta & CHUNK_LONG) return;
      const int* rptr = reinterpret_cast<const int*>(st + Cfg::RPTR_OFF);
      const int* cols = reinterpret_cast<const int*>(st + Cfg::COLS_OFF);
      const int nrows = ch.meta & 0xffff;
      const int tl = (ch.meta >> 16) & 0xf;
      const int tpr = 1 << tl;
      const int rpw = (nrows + Cfg::NWARPS - 1) / Cfg::NWARPS;
      const int rw1 = min(nrows, warp * rpw + rpw);
      const int r = warp * rpw + (lane >> tl);
      const int li = lane & (tpr - 1);
      if (r < rw1) {
        const int roff = ch.row_begin - (ch.row_begin & ~3);
        const int ebase = ch.nnz_begin & ~3;
        w.valid = true;
        w.row = ch.row_begin + r;
        w.i0 = rptr[roff + r] - ebase + li;
        w.e = rptr[roff + r + 1] - ebase;
#ifndef PDLP_DBG_NOEPI
        if (li == 0) w.ctx = op.row_load(w.row);
#endif
      }
#pragma unroll
      for (int j = 0; j < U; ++j) {
        const int idx = w.i0 + j * tpr;
        if (idx < w.e) w.g[j] = op.gather(PDLP_COL(cols[idx]));
      }
    };

    // finish(): multiply, reduce, epilogue, release the stage
    auto finish = [&](int ci, Work<Op>& w) {
      const int it = ci - cb;
      const int s = it % Cfg::STAGES;
      const unsigned char* st = smem + (size_t)s * Cfg::STAGE_BYTES;
      const Chunk ch = *reinterpret_cast<const Chunk*>(st + Cfg::DESC_OFF);
      if (ch.meta & CHUNK_LONG) {
        // one row longer than a stage: every consumer thread strides over it in global memory
        T acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = T(0);
        for (int e = ch.nnz_begin + tid; e < ch.nnz_end; e += Cfg::NCONS) op.mac(acc, A.vals[e], op.gather(A.colidx[e]));
#pragma unroll
        for (int v = 0; v < NV; ++v) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc[v] += shfl_xor_t(acc[v], o);
        }
        consumer_barrier(Cfg::NCONS);      // s_red free (previous long row fully consumed)
        if (lane == 0) {
#pragma unroll
          for (int v = 0; v < NV; ++v) s_red[warp][v] = (double)acc[v];
        }
        consumer_barrier(Cfg::NCONS);
        if (tid == 0) {
          T tot[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            double t = 0.0;
            for (int ww = 0; ww < Cfg::NWARPS; ++ww) t += s_red[ww][v];
            tot[v] = (T)t;
          }
          op.row_finish(ch.row_begin, op.row_load(ch.row_begin), tot, red);
        }
      } else {
        const int* rptr = reinterpret_cast<const int*>(st + Cfg::RPTR_OFF);
        const int* cols = reinterpret_cast<const int*>(st + Cfg::COLS_OFF);
        const T* vals = reinterpret_cast<const T*>(st + Cfg::VALS_OFF);
        const int nrows = ch.meta & 0xffff;
        const int tl = (ch.meta >> 16) & 0xf;
        const int tpr = 1 << tl;
        const int groups = 32 >> tl;
        const int li = lane & (tpr - 1);
        // ---- round 0: gathers were issued by issue() ----
        T acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = T(0);
#pragma unroll
        for (int j = 0; j < U; ++j) {
          const int idx = w.i0 + j * tpr;
          if (idx < w.e) op.mac(acc, vals[idx], w.g[j]);
        }
        for (int i = w.i0 + U * tpr; i < w.e; i += U * tpr) {       // rows longer than U * tpr
          typename Op::Gather g[U];
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const int idx = i + j * tpr;
            if (idx < w.e) g[j] = op.gather(PDLP_COL(cols[idx]));
          }
#pragma unroll
          for (int j = 0; j < U; ++j) {
            const int idx = i + j * tpr;
            if (idx < w.e) op.mac(acc, vals[idx], g[j]);
          }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          for (int o = tpr >> 1; o > 0; o >>= 1) acc[v] += shfl_xor_t(acc[v], o);
        }
#ifndef PDLP_DBG_NOEPI
        if (w.valid && li == 0) op.row_finish(w.row, w.ctx, acc, red);
#else
        if (w.valid && li == 0) red[0] += (double)acc[0];
#endif
        // ---- further rounds (chunks with more rows than one pass of the CTA covers) ----
        const int rpw = (nrows + Cfg::NWARPS - 1) / Cfg::NWARPS;
        const int rw0 = warp * rpw;
        const int rw1 = min(nrows, rw0 + rpw);
        if (rw0 + groups < rw1) {
          const int roff = ch.row_begin - (ch.row_begin & ~3);
          const int ebase = ch.nnz_begin & ~3;
          for (int rr = rw0 + groups; rr < rw1; rr += groups) {
            const int r = rr + (lane >> tl);
            const bool valid = r < rw1;
            int b = 0, e = 0;
            typename Op::Ctx ctx;
            if (valid) {
              b = rptr[roff + r] - ebase;
              e = rptr[roff + r + 1] - ebase;
              if (li == 0) ctx = op.row_load(ch.row_begin + r);
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) acc[v] = T(0);
            for (int i = b + li; i < e; i += U * tpr) {
              typename Op::Gather g[U];
#pragma unroll
              for (int j = 0; j < U; ++j) {
                const int idx = i + j * tpr;
                if (idx < e) g[j] = op.gather(PDLP_COL(cols[idx]));
              }
#pragma unroll
              for (int j = 0; j < U; ++j) {
                const int idx = i + j * tpr;
                if (idx < e) op.mac(acc, vals[idx], g[j]);
              }
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
              for (int o = tpr >> 1; o > 0; o >>= 1) acc[v] += shfl_xor_t(acc[v], o);
            }
            if (valid && li == 0) op.row_finish(ch.row_begin + r, ctx, acc, red);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    };

#ifdef PDLP_DBG_NOWORK
    for (int ci = cb; ci < ce; ++ci) {
      const int it = ci - cb;
      const int s = it % Cfg::STAGES;
      mbar_wait(&full_bar[s], (it / Cfg::STAGES) & 1);
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    (void)issue; (void)finish;
#else
    Work<Op> cur, nxt;
    if (cb < ce) issue(cb, cur);
    for (int ci = cb; ci < ce; ++ci) {
      if (ci + 1 < ce) issue(ci + 1, nxt);
      finish(ci, cur);
      cur = nxt;
    }
#endif
  }

  // ---------------- deterministic grid reduction + finalize ----------------
  if constexpr (NRED > 0) {
    if (warp < Cfg::NWARPS) {
#pragma unroll
      for (int i = 0; i < NRED; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) red[i] += shfl_xor_t(red[i], o);
      }
    }
    __syncthreads();   // long-row users of s_red are done
    if (warp < Cfg::NWARPS && lane == 0) {
#pragma unroll
      for (int i = 0; i < NRED; ++i) s_red[warp][i] = red[i];
    }
    __syncthreads();
    if (tid == 0) {
      for (int i = 0; i < NRED; ++i) {
        double t = 0.0;
        for (int w = 0; w < Cfg::NWARPS; ++w) t += s_red[w][i];
        A.block_partials[(size_t)blockIdx.x * MAX_RED + i] = t;
      }
      __threadfence();
      const unsigned int ticket = atomicAdd(A.ticket, 1u);
      s_last = (ticket == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (s_last && warp == 0) {
      __threadfence();
      double tot[NR];
      for (int i = 0; i < NRED; ++i) {
        double t = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&A.block_partials[(size_t)b * MAX_RED + i]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += shfl_xor_t(t, o);
        tot[i] = t;
      }
      if (lane == 0) {
        op.finalize(tot);
        *A.ticket = 0u;
      }
    }
  }
}

}  // namespace pdlp
