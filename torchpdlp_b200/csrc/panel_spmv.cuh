// panel_spmv.cuh -- K^T y with the gathered vector staged in shared memory ("panels").
//
// On B200 the iteration kernels are bound by the 32-byte sectors that cross the L2 -> SM fabric: the matrix stream
// plus ONE sector per gathered operand element.  For K^T y the gathered vector y is small
// (config 2: 800 KB), so it can be taken off the fabric altogether: the columns of K^T (= elements of y) are cut
// into P panels of <= PDLP_PANEL_BYTES, the SMs into P groups, and group p streams only the entries of panel p with
// its slice of y resident in shared memory -- the per-nonzero gather becomes an LDS.  Every (panel, row) pair is a
// row of a P*n-row "panel matrix" stored in the ordinary SELL-32-sigma layout (stream_spmv.cuh: one row part per
// lane, no cross-lane sums), whose row epilogue stores the partial sum partial[p * n + row]; combine_kernel then adds
// the P partial sums of a row in panel order and runs the fused row epilogue of the op (PrimalOp) on rows in their
// natural order, i.e. with coalesced vector accesses.  Sums have a fixed order: bitwise reproducible.
#pragma once
#include "stream_spmv.cuh"

namespace pdlp {

#ifndef PDLP_PANEL_BYTES
#define PDLP_PANEL_BYTES (100 * 1024)   // largest slice of the gathered vector a CTA keeps in shared memory
#endif

struct NoPanelCtx {};

template <typename T>
struct PanelOp {
  using value_type = T;
  using Gather = T;
  using Ctx = NoPanelCtx;
  static constexpr int NV = 1;
  static constexpr int NRED = 0;
  static constexpr int UDIV = 1;
  static constexpr int PIPE = 1;
  const T* panel;     // shared memory: this CTA's slice of the gathered vector
  T* out;             // partial sums, indexed by the panel matrix's row id p * n + row
  __device__ bool begin() { return true; }
  __device__ T gather(int col) const { return panel[col]; }
  __device__ void mac(T (&acc)[1], T a, const T& g) const { acc[0] += a * g; }
  __device__ Ctx row_load(int) const { return Ctx{}; }
  __device__ void row_finish(int r, const Ctx&, const T (&acc)[1], double (&)[1]) { out[r] = acc[0]; }
  __device__ void finalize(const double (&)[1]) {}
};

struct PanelInfo {
  const int* cta_panel;          // grid entries: the panel each CTA serves
  const int* panel_begin;        // npanels + 1 offsets into the gathered vector (multiples of 4)
  int panel_cap;                 // elements of shared memory reserved for the panel
};

// A: the panel matrix (P * n rows, panel-relative columns) in SELL layout, its warp ranges laid out so that the warps
// of CTA b hold pieces of panel cta_panel[b] only.
template <typename T>
__global__ void __launch_bounds__(StreamCfg<T>::NT, 1)
panel_kernel(const SellView<T> A, const PanelInfo info, const T* in, T* out, const int* stop) {
  using Cfg = StreamCfg<T>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[Cfg::W][Cfg::SMAX];
  __shared__ uint64_t panel_bar;
  __shared__ double s_red[Cfg::W][1];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  T* panel = reinterpret_cast<T*>(smem);
  unsigned char* rings = smem + ((size_t)info.panel_cap * sizeof(T) + 127) / 128 * 128;
  Ring R{rings + (size_t)warp * A.nstages * A.stage_bytes, &full_bar[warp][0], A.nstages, A.stage_bytes, 0u};
  if (lane == 0) {
    for (int s = 0; s < R.S; ++s) mbar_init(&R.bars[s], 1);
    if (warp == 0) mbar_init(&panel_bar, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncthreads();

  const int gw = blockIdx.x * Cfg::W + warp;
  const int pb = A.warp_piece_begin[gw];
  const int pe = A.warp_piece_begin[gw + 1];
  ring_prologue<T>(A, R, pb, pe, lane);
  const int p = info.cta_panel[blockIdx.x];
  const int vb = info.panel_begin[p];
  const int vcount = info.panel_begin[p + 1] - vb;

  // programmatic dependent launch: everything above touched only the constant plan
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (stop && *stop) {
    for (int s = 0; s < R.S && pb + s < pe; ++s) mbar_wait(&R.bars[s], 0);     // drain the copies in flight
    return;
  }
  // this CTA's slice of the gathered vector (written by the previous kernel, hence only now): bulk copies of
  // <= 32 KB, all on one barrier
  if (tid == 0) {
    const uint32_t bytes = ((uint32_t)vcount * (uint32_t)sizeof(T) + 15u) & ~15u;     // `in` is padded by 16 bytes
    mbar_arrive_expect_tx(&panel_bar, bytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(in + vb);
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t len = min(32768u, bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(smem + off)),
                   "l"(src + off), "r"(len), "r"(smem_u32(&panel_bar))
                   : "memory");
    }
  }
  mbar_wait(&panel_bar, 0);

  PanelOp<T> op{panel, out};
  double red[1] = {0.0};
  stream_phase<PanelOp<T>, 1>(A, op, red, R, pb, pe, s_red);
}

// Row epilogue on the combined panel sums: acc = partial[0][r] + partial[1][r] + ... (panel order), then the op's
// row_load / row_finish / finalize exactly as in the streaming kernels.  Natural row order: coalesced.
constexpr int COMBINE_MAX_PANELS = 16;
constexpr int COMBINE_THREADS = 256;

template <class Op>
__global__ void __launch_bounds__(COMBINE_THREADS) combine_kernel(const typename Op::value_type* partial, int npanels,
                                                                  int rows, Op op, double* block_partials,
                                                                  unsigned int* ticket) {
  using T = typename Op::value_type;
  constexpr int NRED = Op::NRED;
  constexpr int NR = NRED > 0 ? NRED : 1;
  static_assert(Op::NV == 1, "combine_kernel: one accumulator per row");
  __shared__ double s_red[COMBINE_THREADS / 32][NR];
  __shared__ int s_last;
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (!op.begin()) return;
  double red[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) red[i] = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += stride) {
    const typename Op::Ctx ctx = op.row_load(r);
    T part[COMBINE_MAX_PANELS];              // all loads in flight before the first add
#pragma unroll
    for (int p = 0; p < COMBINE_MAX_PANELS; ++p) part[p] = p < npanels ? partial[(size_t)p * rows + r] : T(0);
    T acc[1] = {part[0]};
#pragma unroll
    for (int p = 1; p < COMBINE_MAX_PANELS; ++p)
      if (p < npanels) acc[0] += part[p];   // panel order
    op.row_finish(r, ctx, acc, red);
  }
  if constexpr (NRED > 0) {
    block_reduce<NRED, NR>(red, s_red, &block_partials[(size_t)blockIdx.x * MAX_RED]);
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
      __threadfence();
      double tot[NR];
      grid_totals<NRED>(block_partials, MAX_RED, (int)gridDim.x, tot);
      if (threadIdx.x == 0) {
        op.finalize(tot);
        *ticket = 0u;
      }
    }
  }
}

}  // namespace pdlp
