// pdlp_block.cuh -- one persistent cooperative kernel per block of PDHG iterations.
//
// The grid (one CTA per SM) stays resident for n_iters iterations; every iteration is two passes of
// stream_phase (PrimalOp over K^T, DualOp over K) separated by grid barriers instead of kernel boundaries.
// The TMA ring of the next pass is primed BEFORE the barrier (it only touches the constant plan), every CTA
// reduces the per-CTA partial norms itself (same fixed order as the multi-kernel path: the results are bitwise
// identical) and applies the step-size rule to its own copy of the scalars; CTA 0 writes them back at the end.
// Small LPs are no longer launch-bound, large ones lose the dependent-launch gap of 80 kernels per block.
#pragma once
#include "pdlp_ops.cuh"

namespace pdlp {

template <typename T>
struct BlockArgs {
  // vectors
  T* y; const T* c; const T* l; const T* u; Pair<T>* xp; T* xsum; Quad<T>* X4;
  const T* q; T* ysum; Quad<T>* Y4; T* kty_old;
  Scalars* S;
  double* partP;            // 2 x grid
  double* partD;            // 2 x grid x 2
  unsigned int* gbar;       // grid barrier counter, zeroed by the host before the launch
  int n_iters;
  int save_prev;            // keep the iterate before the LAST step (restart test)
  int ring_stages;          // common ring geometry of K and K^T
  int ring_stage_bytes;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// all CTAs of the (cooperatively launched, fully resident) grid
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire_u32(bar) < target) {
    }
    __threadfence();
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(StreamCfg<T>::NT, 1)
block_kernel(const SellView<T> Kt, const SellView<T> K, const BlockArgs<T> a) {
  using Cfg = StreamCfg<T>;
  constexpr int SR = 2;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[Cfg::W][Cfg::SMAX];
  __shared__ double s_red[Cfg::W][SR];
  __shared__ Scalars ls;

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int grid = (int)gridDim.x;

  Ring R{smem + (size_t)warp * a.ring_stages * a.ring_stage_bytes, &full_bar[warp][0], a.ring_stages, a.ring_stage_bytes, 0u};
  if (lane == 0) {
    for (int s = 0; s < R.S; ++s) mbar_init(&R.bars[s], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  if (tid == 0) ls = *a.S;
  __syncthreads();

  const int gw = blockIdx.x * Cfg::W + warp;
  const bool inT = (int)blockIdx.x < Kt.grid, inK = (int)blockIdx.x < K.grid;
  const int pbT = inT ? Kt.warp_piece_begin[gw] : 0, peT = inT ? Kt.warp_piece_begin[gw + 1] : 0;
  const int pbK = inK ? K.warp_piece_begin[gw] : 0, peK = inK ? K.warp_piece_begin[gw + 1] : 0;

  unsigned int bar_target = 0;
  int par = 0;
  ring_prologue<T>(Kt, R, pbT, peT, lane);
  for (int it = 0; it < a.n_iters; ++it) {
    const bool sp = a.save_prev && it == a.n_iters - 1;
    // ---------------- primal pass over K^T ----------------
    {
      double red[1] = {0.0};
      const T tau = (T)(ls.eta / ls.omega), wp = (T)ls.w_pending;
      PrimalOp<T, true, false> p{a.y, a.c, a.l, a.u, a.xp, a.xsum, a.X4, a.S, a.kty_old, tau, wp, T(0), sp};
      stream_phase<PrimalOp<T, true, false>, SR>(Kt, p, red, R, pbT, peT, s_red);
      block_reduce<1, SR>(red, s_red, &a.partP[par * grid + blockIdx.x]);
    }
    ring_prologue<T>(K, R, pbK, peK, lane);           // flies while we wait for the other CTAs
    grid_barrier(a.gbar, bar_target);
    // ---------------- dual pass over K ----------------
    {
      double red[2] = {0.0, 0.0};
      const T sigma = (T)(ls.eta * ls.omega), th = (T)ls.theta, wp = (T)ls.w_pending;
      DualOp<T, true, false> d{a.xp, a.q, a.y, a.ysum, a.Y4, a.S, sigma, th, wp, T(0), ls.m_ineq, sp};
      stream_phase<DualOp<T, true, false>, SR>(K, d, red, R, pbK, peK, s_red);
      block_reduce<2, SR>(red, s_red, &a.partD[(par * grid + blockIdx.x) * 2]);
    }
    if (it + 1 < a.n_iters) ring_prologue<T>(Kt, R, pbT, peT, lane);
    grid_barrier(a.gbar, bar_target);
    // ---------------- every CTA: grid totals (fixed order) and the step rule on its own scalars ----------------
    if (warp == 0) {
      double t1[1], t2[2];
      grid_totals<1, false>(a.partP + par * grid, 1, grid, t1);
      grid_totals<2, false>(a.partD + par * grid * 2, 2, grid, t2);
      if (lane == 0) {
        ls.dx2 = t1[0];
        ls.dy2 = t2[0];
        ls.inter = t2[1];
        double w, next;
        step_rule(ls, t1[0], t2[0], t2[1], w, next);
        commit_step(ls, w, next);
      }
    }
    __syncthreads();
    par ^= 1;
  }
  if (blockIdx.x == 0 && tid == 0) *a.S = ls;
}

}  // namespace pdlp
