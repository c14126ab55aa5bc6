// pdlp_ops.cuh -- fused epilogues ("ops") plugged into stream_kernel, the device
// scalar block, and the small element-wise kernels around the hot loop.
//
// Reference semantics (paths under /root/reference/PDLP):
//   PrimalOp      primal_dual_hybrid_gradient_step.py:25-27 / :68-77  (x' = clip(x - tau (c - K^T y)))
//   DualOp        step.py:30-38 / :80-108 (y' = proj(y + sigma (q - K xbar)), step rule)
//   KktColsOp     helpers.py:75-85,:94   (grad, lambda, objective terms, dual residual)
//   KktRowsOp     helpers.py:77,:88-91   (q^T y, primal residual)
//   decide/apply  primal_dual_hybrid_gradient.py:115-177
#pragma once
#include <math.h>

#include "stream_spmv.cuh"

namespace pdlp {

template <typename T> struct Pair;
template <> struct __align__(16) Pair<double> { double a, b; };
template <> struct __align__(8) Pair<float> { float a, b; };

template <typename T> struct __align__(32) Quad { T v[4]; };
template <> struct __align__(16) Quad<float> { float v[4]; };

// 8 right-hand sides side by side (fishnet batches, spectral_casting.py:254-293)
template <typename T> struct __align__(64) Oct { T v[8]; };
template <> struct __align__(32) Oct<float> { float v[8]; };

// slots of Quad: the three iterates the restart test evaluates (:121-126)
enum { IT_CUR = 0, IT_AVG = 1, IT_PREV = 2 };

// raw sums per iterate kept in Scalars::sums[v][*]
enum { S_PR2 = 0, S_DR2, S_CTX, S_LPOS, S_UNEG, S_QTY, S_PR2U, S_DR2U, S_NSUM };

struct Resid {
  double pr, dr, gap, prim_obj, adj_dual, kkt, pr_u, dr_u;
};

struct Scalars {
  // step state
  double eta, omega, theta, w_pending, eta_sum;
  double dx2, dy2, inter, last_weight;
  double kkt_first, q_norm, c_norm, tol, infeas_tol;
  long long k;
  int t, n_restarts, adaptive, primal_update;
  int infeas_detect, infeas, infeas_k, solved;
  int stop, m_ineq;
  int sharded;        // row-partitioned engine: sums over rows are partial until the host allreduces them
  int rule_pending;   // sharded: the step rule of the latest iteration waits for the allreduced norms
  // restart check
  double sums[3][S_NSUM];
  int chk_kind, chk_use_avg, chk_t, chk_solved;
  long long chk_k;
  double chk_omega, kkt_cur, kkt_avg, kkt_prev, kkt_first_new;
  Resid point;
  double restart_dx2, restart_dy2;
  // infeasibility detection (enhancements.py:106-161): sums of the latest step
  double inf_ctdx, inf_bound_viol;                 // c^T dx, #variables failing the bound test (:126-138)
  double inf_eq2, inf_ineq_viol, inf_qtdy, inf_dy_viol;
  double inf_dres2, inf_lsum, inf_usum;
  // scratch reductions: power iteration, Ruiz, norms
  double scratch[8];
  // fishnet: per point c^T x, l_dual^T max(lam,0), u_dual^T min(lam,0) for a batch of 8 points
  double batch[24];
  // peer-memory exchange of the row-partitioned engine: sequence number of the next exchange, sticky error
  unsigned long long xchg_seq;
  int xchg_error;
  // pipelined solve (pdlp_solve): a restart point that meets the termination test raises `stop`, so the block and
  // the restart test the host issued ahead of reading this test's record run as no-ops
  int stop_on_solved;
};

template <typename T>
__device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }
__device__ __forceinline__ Pair<double> ldg(const Pair<double>* p) {
  const double2 v = __ldg(reinterpret_cast<const double2*>(p));
  return Pair<double>{v.x, v.y};
}
__device__ __forceinline__ Pair<float> ldg(const Pair<float>* p) {
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  return Pair<float>{v.x, v.y};
}

template <typename T> __device__ __forceinline__ T clampv(T v, T lo, T hi) { return fmin(fmax(v, lo), hi); }

#ifdef PDLP_GATHER_EVICT_LAST
// gathers of the dense operand with an L2 evict-last hint (developer experiment)
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ double ldg_keep(const double* p) {
  double v;
  asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(l2_evict_last_policy()));
  return v;
}
__device__ __forceinline__ Pair<double> ldg_keep(const Pair<double>* p) {
  Pair<double> v;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.a), "=d"(v.b) : "l"(p), "l"(l2_evict_last_policy()));
  return v;
}
__device__ __forceinline__ float ldg_keep(const float* p) { return __ldg(p); }
__device__ __forceinline__ Pair<float> ldg_keep(const Pair<float>* p) { return ldg(p); }
#else
template <typename P> __device__ __forceinline__ auto ldg_keep(const P* p) { return ldg(p); }
#endif

// The step-size rule: step.py:96-115 (adaptive) / :40 (fixed).  Increments k (pdlp :76 counts the iteration
// before the rule reads it) and returns the averaging weight of the new iterate and the next step size.
__device__ inline void rule_values(int adaptive, long long k, double eta, double omega, double dx2, double dy2,
                                   double inter, double& w, double& next) {   // k: already incremented
  w = eta;
  next = eta;
  if (adaptive) {
    const double den = 2.0 * inter;
    double eta_bar, t1;
    if (den != 0.0) {
      eta_bar = (omega * dx2 + dy2 / omega) / fabs(den);
      t1 = (1.0 - pow((double)(k + 1), -0.3)) * eta_bar;
    } else {
      eta_bar = INFINITY;
      t1 = INFINITY;
    }
    const double t2 = (1.0 + pow((double)(k + 1), -0.6)) * eta;
    const double ep = fmin(t1, t2);
    if (eta <= eta_bar) { w = eta; next = ep; } else { w = ep; next = ep; }
  }
}
__device__ inline void step_rule(Scalars& s, double dx2, double dy2, double inter, double& w, double& next) {
  const long long k = s.k + 1;
  s.k = k;
  rule_values(s.adaptive, k, s.eta, s.omega, dx2, dy2, inter, w, next);
}
// averaging bookkeeping of pdlp :104-112 (the weight is folded into the sums by the next kernels that read the iterate)
__device__ inline void commit_step(Scalars& s, double w, double next) {
  s.eta_sum += w;
  s.w_pending = w;
  s.last_weight = w;
  s.eta = next;
  s.t += 1;
}

struct NoCtx {};

// ---------------------------------------------------------------------------
// Plain SpMV: out[row] = sum, optional sum of squares -> Scalars::scratch[slot]
// (power iteration helpers.py:41-51, pdlp_spmv).
// ---------------------------------------------------------------------------
template <typename T, bool SUMSQ>
struct SpmvOp {
  using value_type = T;
  using Gather = T;
  using Ctx = NoCtx;
  static constexpr int NV = 1;
  static constexpr int NRED = SUMSQ ? 1 : 0;
  static constexpr int UDIV = 1;
  static constexpr int PIPE = 1;
  const T* in;
  T* out;
  Scalars* S;
  int slot;
  // div_slot >= 0: out = (A in) / sqrt(scratch[div_slot]) -- the power iteration's b = v / ||v|| (helpers.py:50)
  // applied once per ROW of the product instead of once per element of v in a pass of its own
  int div_slot = -1;
  T nrm = T(1);
  __device__ bool begin() {
    if (div_slot >= 0) nrm = (T)sqrt(S->scratch[div_slot]);
    return true;
  }
  __device__ T gather(int col) const { return ldg(in + col); }
  __device__ void mac(T (&acc)[1], T a, const T& g) const { acc[0] += a * g; }
  __device__ Ctx row_load(int) const { return Ctx{}; }
  __device__ void row_finish(int r, const Ctx&, const T (&acc)[1], double (&red)[1]) {
    const T o = div_slot >= 0 ? acc[0] / nrm : acc[0];
    out[r] = o;
    if (SUMSQ) red[0] += (double)o * (double)o;
  }
  __device__ void finalize(const double (&tot)[1]) { S->scratch[slot] = tot[0]; }
};

// ---------------------------------------------------------------------------
// K1: rows of K^T (one per variable).  g = c - K^T y; x' = clip(x - tau g, l, u);
// dx = x' - x; deferred averaging x_sum += w_pending * x; ||dx||^2.
// ---------------------------------------------------------------------------
template <typename T, bool SAVE_PREV, bool INFEAS = false>
struct PrimalOp {
  using value_type = T;
  using Gather = T;
  struct Ctx { T c, l, u, xo, xs; };
  static constexpr int NV = 1;
  static constexpr int NRED = INFEAS ? 3 : 1;
  static constexpr int UDIV = 1;
  static constexpr int PIPE = 1;
  const T* y;
  const T* c;
  const T* l;
  const T* u;
  Pair<T>* xp;       // (x, dx)
  T* xsum;
  Quad<T>* X4;
  Scalars* S;
  T* kty_old;        // INFEAS: K^T y of the iterate the step started from
  T tau, wpend, itol;
  bool do_save = true;   // runtime gate of SAVE_PREV (the persistent kernel saves only on its last iteration)
  const Quad<T>* clu = nullptr;   // (c, l, u, 0) records: one 32-byte sector per row instead of three (sigma-sorted rows
                                  // make every epilogue access its own sector, section 4b of DESIGN.md)
  T* xb = nullptr;                // dense xbar = x' + theta dx for DualBarOp (8-byte gathers), null otherwise
  T theta = T(1);
  bool defer = false;             // stop after the per-CTA partial ||dx||^2: the DualOp that follows adds them up
                                  // (DualOp::dx2_partials); not with INFEAS, whose finalize stores more
  __device__ bool begin() {
    if (S->stop) return false;
    tau = (T)(S->eta / S->omega);
    wpend = (T)S->w_pending;
    itol = (T)S->infeas_tol;
    theta = (T)S->theta;
    return true;
  }
  __device__ T gather(int col) const { return ldg_keep(y + col); }
  __device__ void mac(T (&acc)[1], T a, const T& g) const { acc[0] += a * g; }
  __device__ Ctx row_load(int j) const {
    if (clu) {
      const Quad<T> r = clu[j];
      return Ctx{r.v[0], r.v[1], r.v[2], xp[j].a, xsum[j]};
    }
    return Ctx{c[j], l[j], u[j], xp[j].a, xsum[j]};
  }
  __device__ void row_finish(int j, const Ctx& k, const T (&acc)[1], double (&red)[NRED]) {
    const T g = k.c - acc[0];
    const T xn = clampv(k.xo - tau * g, k.l, k.u);
    const T dx = xn - k.xo;
    xp[j] = Pair<T>{xn, dx};
    if (xb) xb[j] = xn + theta * dx;          // the value DualOp::mac forms from the pair
    xsum[j] = k.xs + wpend * k.xo;
    if (SAVE_PREV && do_save) X4[j].v[IT_PREV] = k.xo;
    red[0] += (double)dx * (double)dx;
    if (INFEAS) {
      kty_old[j] = acc[0];
      red[1] += (double)k.c * (double)dx;
      // enhancements.py:132-136
      const bool ok = (!isinf(k.l) && !isinf(k.u) && fabs(dx) <= itol) ||
                      (k.u == (T)INFINITY && k.c >= T(0) && dx >= -itol) ||
                      (k.l == -(T)INFINITY && k.c <= T(0) && dx <= itol);
      red[2] += ok ? 0.0 : 1.0;
    }
  }
  __device__ void finalize(const double (&tot)[NRED]) {
    S->dx2 = tot[0];
    if (INFEAS) { S->inf_ctdx = tot[1]; S->inf_bound_viol = tot[2]; }
  }
};

// ---------------------------------------------------------------------------
// K2: rows of K.  xbar = x' + theta dx gathered as a pair; y' = y + sigma (q - K xbar),
// projected on rows < m_ineq; dy; deferred averaging; ||dy||^2 and dy^T (K dx); then the
// step-size rule on the last CTA.
// ---------------------------------------------------------------------------
template <typename T, bool SAVE_PREV, bool INFEAS = false>
struct DualOp {
  using value_type = T;
  using Gather = Pair<T>;
  struct Ctx { T yo, q, ys; };
  static constexpr int NV = 2;
  static constexpr int NRED = INFEAS ? 6 : 2;
  static constexpr int UDIV = 1;
  static constexpr int PIPE = 0;   // 16-byte gathers: a second batch buffer costs 32 registers for ~2 %
  const Pair<T>* xp;
  const T* q;
  T* y;
  T* ysum;
  Quad<T>* Y4;
  Scalars* S;
  T sigma, theta, wpend, itol;
  int m_ineq;
  bool do_save = true;
  // ||dx||^2 left as per-CTA partial sums by a PrimalOp with `defer` (its kernel's block_partials, stride MAX_RED):
  // summed here in grid_totals' fixed order, by the warp that runs the finalize; null: Scalars::dx2 is current
  const double* dx2_partials = nullptr;
  int dx2_blocks = 0;
  double dx2_tot = 0.0;
  double dxv[5];
  // loads issued before the kernel's own partial sums are read (one round trip for everything the finalize needs):
  // the first 160 partial ||dx||^2 and the two cache lines of Scalars that hold the step state
  __device__ void pre_finalize_issue() {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S));
    asm volatile("prefetch.global.L1 [%0];" ::"l"((const char*)S + 128));
    if (dx2_partials) {
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int b = lane + 32 * j;
        dxv[j] = b < dx2_blocks ? __ldcg(dx2_partials + (size_t)b * MAX_RED) : 0.0;
      }
    }
  }
  __device__ void pre_finalize() {          // grid_totals' order: lane-strided sums from +0.0, then the xor tree
    if (dx2_partials) {
      const int lane = threadIdx.x & 31;
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < 5; ++j) t += dxv[j];
      for (int b = lane + 160; b < dx2_blocks; b += 32) t += __ldcg(dx2_partials + (size_t)b * MAX_RED);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += shfl_xor_t(t, o);
      dx2_tot = t;
    }
  }
  __device__ bool begin() {
    if (S->stop) return false;
    sigma = (T)(S->eta * S->omega);
    theta = (T)S->theta;
    wpend = (T)S->w_pending;
    itol = (T)S->infeas_tol;
    m_ineq = S->m_ineq;
    return true;
  }
  __device__ Pair<T> gather(int col) const { return ldg_keep(xp + col); }
  __device__ void mac(T (&acc)[2], T a, const Pair<T>& p) const {
    acc[0] += a * (p.a + theta * p.b);
    acc[1] += a * p.b;
  }
  __device__ Ctx row_load(int i) const { return Ctx{y[i], q[i], ysum[i]}; }
  __device__ void row_finish(int i, const Ctx& k, const T (&acc)[2], double (&red)[NRED]) {
    T yn = k.yo + sigma * (k.q - acc[0]);
    if (i < m_ineq) yn = fmax(yn, T(0));
    const T dy = yn - k.yo;
    y[i] = yn;
    ysum[i] = k.ys + wpend * k.yo;
    if (SAVE_PREV && do_save) Y4[i].v[IT_PREV] = k.yo;
    red[0] += (double)dy * (double)dy;
    red[1] += (double)dy * (double)acc[1];
    if (INFEAS) {
      const T kdx = acc[1];                                   // (K dx)_i
      if (i < m_ineq) {
        red[3] += kdx >= -itol ? 0.0 : 1.0;                   // enhancements.py:119
        red[5] += dy >= -itol ? 0.0 : 1.0;                    // :147
      } else {
        red[2] += (double)kdx * (double)kdx;                  // :116
      }
      red[4] += (double)k.q * (double)dy;                     // :148
    }
  }
  // step.py:96-115 (adaptive) / :40 (fixed) and primal_dual_hybrid_gradient.py:104-112
  __device__ void finalize(const double (&tot)[NRED]) {
    Scalars& s = *S;
    if (dx2_partials) s.dx2 = dx2_tot;
    s.dy2 = tot[0];
    s.inter = tot[1];
    if (s.sharded) {          // partial sums: the rule runs after the host allreduce (shard_rule_kernel)
      s.rule_pending = 1;
      return;
    }
    double w, next;
    step_rule(s, s.dx2, tot[0], tot[1], w, next);
    if (INFEAS) {
      s.inf_eq2 = tot[2]; s.inf_ineq_viol = tot[3]; s.inf_qtdy = tot[4]; s.inf_dy_viol = tot[5];
      // dual infeasibility test, enhancements.py:116-141 (only from the second iteration on, pdlp :89)
      if (s.k > 1 && sqrt(tot[2]) < s.infeas_tol && tot[3] == 0.0 && s.inf_ctdx < s.infeas_tol &&
          s.inf_bound_viol == 0.0) {
        s.infeas = 1;
        s.infeas_k = (int)s.k;
        s.stop = 1;
        s.last_weight = w;
        return;   // the reference returns before the averaging update (:94-100)
      }
    }
    commit_step(s, w, next);
  }
};

// ---------------------------------------------------------------------------
// K2 in column windows (operands larger than L2, config 5): K is cut into P column windows whose slice of the
// (x, dx) pair vector stays L2-resident; window p is one launch over the sub-matrix K[:, window p] (columns relative
// to the window) that adds its two partial row sums (K xbar, K dx) to the pair carried in part[] -- read unless
// FIRST, written unless LAST; the LAST window runs DualOp's epilogue on the totals.  Fixed window order: reproducible.
// ---------------------------------------------------------------------------
template <typename T, bool FIRST, bool LAST, bool SAVE_PREV>
struct DualWindowOp {
  using value_type = T;
  using Gather = Pair<T>;
  using Inner = DualOp<T, SAVE_PREV, false>;
  struct Ctx { Pair<T> part; typename Inner::Ctx in; };
  static constexpr int NV = 2;
  static constexpr int NRED = LAST ? Inner::NRED : 0;
  static constexpr int UDIV = 1;
  static constexpr int PIPE = 0;
  Inner d;               // d.xp already points at the window's first pair
  Pair<T>* part;
  __device__ bool begin() { return d.begin(); }
  __device__ Pair<T> gather(int col) const { return d.gather(col); }
  __device__ void mac(T (&acc)[2], T a, const Pair<T>& p) const { d.mac(acc, a, p); }
  __device__ Ctx row_load(int i) const {
    Ctx c;
    if (!FIRST) c.part = ldg(part + i);
    if (LAST) c.in = d.row_load(i);
    return c;
  }
  __device__ void row_finish(int i, const Ctx& k, const T (&acc)[2], double (&red)[NRED > 0 ? NRED : 1]) {
    T a[2] = {acc[0], acc[1]};
    if (!FIRST) { a[0] = k.part.a + a[0]; a[1] = k.part.b + a[1]; }
    if constexpr (LAST) d.row_finish(i, k.in, a, red);
    else part[i] = Pair<T>{a[0], a[1]};
  }
  __device__ void pre_finalize_issue() {
    if constexpr (LAST) d.pre_finalize_issue();
  }
  __device__ void pre_finalize() {
    if constexpr (LAST) d.pre_finalize();
  }
  __device__ void finalize(const double (&tot)[NRED > 0 ? NRED : 1]) {
    if constexpr (LAST) d.finalize(tot);
  }
};

// ---------------------------------------------------------------------------
// K2 on the dense xbar vector (operands beyond L2: configs 4 and 5).  The (x, dx) pair costs 16 bytes per column of
// L2 footprint only so that K dx -- needed by the adaptive rule's dy^T K dx -- comes out of the same gather.  Here
// K1 also writes xbar = x' + theta dx as a dense vector (PrimalOp::xb), K2 gathers 8 bytes (half the table: config 4's
// 32 MB stay L2-resident where 64 MB of pairs do not), forms u = K xbar exactly as DualOp does -- so the dual update
// is the same to the bit -- and takes K dx from the recursion on v = K x kept per row:
//     xbar = (1 + theta) x' - theta x   =>   K dx = (u - v) / (1 + theta),   v' = K x' = v + K dx.
// An error in v is damped by theta / (1 + theta) per iteration; v is recomputed from x at every restart and whenever
// the iterate or the scaling is set from outside (KxRefreshOp).  In column windows (FIRST / LAST as DualWindowOp) the
// carried partial sum is one value per row instead of a pair.
// ---------------------------------------------------------------------------
template <typename T, bool FIRST, bool LAST, bool SAVE_PREV>
struct DualBarOp {
  using value_type = T;
  using Gather = T;
  using Inner = DualOp<T, SAVE_PREV, false>;
  struct Ctx { T part, kx; typename Inner::Ctx in; };
  static constexpr int NV = 1;
  static constexpr int NRED = LAST ? Inner::NRED : 0;
  static constexpr int UDIV = 1;
  static constexpr int PIPE = 1;
  Inner d;               // d.xp is not used
  const T* xb;           // already points at the window's first column
  T* part;
  T* kx;
  T inv1t = T(1);
  __device__ bool begin() {
    if (!d.begin()) return false;
    inv1t = T(1) / (T(1) + d.theta);
    return true;
  }
  __device__ T gather(int col) const { return ldg_keep(xb + col); }
  __device__ void mac(T (&acc)[1], T a, const T& g) const { acc[0] += a * g; }
  __device__ Ctx row_load(int i) const {
    Ctx c;
    if (!FIRST) c.part = ldg(part + i);
    if (LAST) { c.kx = kx[i]; c.in = d.row_load(i); }
    return c;
  }
  __device__ void row_finish(int i, const Ctx& k, const T (&acc)[1], double (&red)[NRED > 0 ? NRED : 1]) {
    T u = acc[0];
    if (!FIRST) u = k.part + u;
    if constexpr (LAST) {
      const T kdx = (u - k.kx) * inv1t;
      kx[i] = k.kx + kdx;
      const T a[2] = {u, kdx};
      d.row_finish(i, k.in, a, red);
    } else {
      part[i] = u;
    }
  }
  __device__ void pre_finalize_issue() {
    if constexpr (LAST) d.pre_finalize_issue();
  }
  __device__ void pre_finalize() {
    if constexpr (LAST) d.pre_finalize();
  }
  __device__ void finalize(const double (&tot)[NRED > 0 ? NRED : 1]) {
    if constexpr (LAST) d.finalize(tot);
  }
};

// v = K x from the pair vector (full K): after a restart (conditional: nothing to do when the test did not restart or
// the solve has stopped) and when the iterate or the scaling was set from outside
template <typename T>
struct KxRefreshOp {
  using value_type = T;
  using Gather = Pair<T>;
  using Ctx = NoCtx;
  static constexpr int NV = 1;
  static constexpr int NRED = 0;
  static constexpr int UDIV = 1;
  static constexpr int PIPE = 0;
  const Pair<T>* xp;
  T* kx;
  const Scalars* S;
  int after_restart;
  __device__ bool begin() { return !(after_restart && (S->stop || S->chk_kind == 0)); }
  __device__ Pair<T> gather(int col) const { return ldg(xp + col); }
  __device__ void mac(T (&acc)[1], T a, const Pair<T>& p) const { acc[0] += a * p.a; }
  __device__ Ctx row_load(int) const { return Ctx{}; }
  __device__ void row_finish(int i, const Ctx&, const T (&acc)[1], double (&)[1]) { kx[i] = acc[0]; }
  __device__ void finalize(const double (&)[1]) {}
};

// ---------------------------------------------------------------------------
// Infeasibility, column side (after K2): lam = proj(c - K^T y'), dlam = lam - lam_prev,
// dual_res = K^T dy - dlam, certificate sums (enhancements.py:106-112,145-159; call site pdlp :90).
// ---------------------------------------------------------------------------
template <typename T>
struct InfeasColsOp {
  using value_type = T;
  using Gather = T;
  struct Ctx { T c, l, u, kty, lamp; };
  static constexpr int NV = 1;
  static constexpr int NRED = 3;
  static constexpr int UDIV = 1;
  static constexpr int PIPE = 1;
  const T* y;
  const T* c;
  const T* l;
  const T* u;
  const T* kty_old;
  T* lam_prev;
  Scalars* S;
  __device__ bool begin() { return !S->stop && S->k >= 2; }
  __device__ T gather(int col) const { return ldg(y + col); }
  __device__ void mac(T (&acc)[1], T a, const T& g) const { acc[0] += a * g; }
  __device__ Ctx row_load(int j) const { return Ctx{c[j], l[j], u[j], kty_old[j], lam_prev[j]}; }
  __device__ void row_finish(int j, const Ctx& k, const T (&acc)[1], double (&red)[3]) {
    const bool lo_inf = isinf(k.l) && k.l < T(0);
    const bool up_inf = isinf(k.u) && k.u > T(0);
    const T g = k.c - acc[0];
    T lam;
    if (lo_inf && up_inf) lam = T(0);
    else if (lo_inf) lam = fmin(g, T(0));
    else if (up_inf) lam = fmax(g, T(0));
    else lam = g;
    const T dlam = lam - k.lamp;
    lam_prev[j] = lam;                                        // pdlp :101
    const T dres = (acc[0] - k.kty) - dlam;                   // K^T dy - dlam, :145
    red[0] += (double)dres * (double)dres;
    if (!isinf(k.l) && k.l != T(0)) red[1] += (double)k.l * (double)fmax(dlam, T(0));    // :112,:154
    if (!isinf(k.u) && k.u != T(0)) red[2] += (double)k.u * (double)fmax(-dlam, T(0));   // :111,:156
  }
  __device__ void finalize(const double (&tot)[3]) {
    Scalars& s = *S;
    s.inf_dres2 = tot[0]; s.inf_lsum = tot[1]; s.inf_usum = tot[2];
    if (sqrt(tot[0]) < s.infeas_tol && s.inf_dy_viol == 0.0) {
      const double combo = s.inf_qtdy - tot[1] - tot[2];
      if (combo > -s.infeas_tol) {
        s.infeas = 2;
        s.infeas_k = (int)s.k;
        s.stop = 1;
        // undo the averaging bookkeeping of this iteration: the reference returns before it (:94-100)
        s.eta_sum -= s.w_pending;
        s.w_pending = 0.0;
        s.t -= 1;
      }
    }
  }
};

// ---------------------------------------------------------------------------
// KKT, column side: rows of K^T, three iterates at once (Y4 gathered).
// ---------------------------------------------------------------------------
template <typename T>
struct KktColsOp {
  using value_type = T;
  using Gather = Quad<T>;
  struct Ctx { T c, l, u, inv; Quad<T> x; };
  static constexpr int NV = 3;
  static constexpr int NRED = 15;   // per iterate: dr2, cTx, lpos, uneg, dr2u
  static constexpr int UDIV = 2;
  static constexpr int PIPE = 0;
  const Quad<T>* Y4;
  const Quad<T>* X4;
  const T* c;
  const T* l;
  const T* u;
  const T* dcol;     // nullptr when unscaled
  Scalars* S;
  __device__ bool begin() { return true; }
  __device__ Quad<T> gather(int col) const { return Y4[col]; }
  __device__ void mac(T (&acc)[3], T a, const Quad<T>& g) const {
#pragma unroll
    for (int v = 0; v < 3; ++v) acc[v] += a * g.v[v];
  }
  __device__ Ctx row_load(int j) const { return Ctx{c[j], l[j], u[j], dcol ? T(1) / dcol[j] : T(1), X4[j]}; }
  __device__ void row_finish(int, const Ctx& k, const T (&acc)[3], double (&red)[15]) {
    const bool lo_inf = isinf(k.l) && k.l < T(0);
    const bool up_inf = isinf(k.u) && k.u > T(0);
    const T ld = lo_inf ? T(0) : k.l;      // l_dual, :14-17
    const T ud = up_inf ? T(0) : k.u;
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const T g = k.c - acc[v];
      T lam;                               // helpers.py:3-39
      if (lo_inf && up_inf) lam = T(0);
      else if (lo_inf) lam = fmin(g, T(0));
      else if (up_inf) lam = fmax(g, T(0));
      else lam = g;
      const T d = g - lam;
      red[5 * v + 0] += (double)d * (double)d;
      red[5 * v + 1] += (double)k.c * (double)k.x.v[v];
      red[5 * v + 2] += (double)ld * (double)fmax(lam, T(0));
      red[5 * v + 3] += (double)ud * (double)fmin(lam, T(0));
      const T du = d * k.inv;
      red[5 * v + 4] += (double)du * (double)du;
    }
  }
  __device__ void finalize(const double (&tot)[15]) {
    for (int v = 0; v < 3; ++v) {
      S->sums[v][S_DR2] = tot[5 * v + 0];
      S->sums[v][S_CTX] = tot[5 * v + 1];
      S->sums[v][S_LPOS] = tot[5 * v + 2];
      S->sums[v][S_UNEG] = tot[5 * v + 3];
      S->sums[v][S_DR2U] = tot[5 * v + 4];
    }
    S->w_pending = 0.0;   // the flush kernel folded the pending weight into the sums
  }
};

// ---------------------------------------------------------------------------
// KKT, row side: rows of K, three iterates at once (X4 gathered).
// ---------------------------------------------------------------------------
template <typename T>
struct KktRowsOp {
  using value_type = T;
  using Gather = Quad<T>;
  struct Ctx { T q, inv; Quad<T> y; };
  static constexpr int NV = 3;
  static constexpr int NRED = 9;    // per iterate: pr2, qTy, pr2u
  static constexpr int UDIV = 2;
  static constexpr int PIPE = 0;
  const Quad<T>* X4;
  const Quad<T>* Y4;
  const T* q;
  const T* drow;     // nullptr when unscaled
  Scalars* S;
  int m_ineq;
  __device__ bool begin() { m_ineq = S->m_ineq; return true; }
  __device__ Quad<T> gather(int col) const { return X4[col]; }
  __device__ void mac(T (&acc)[3], T a, const Quad<T>& g) const {
#pragma unroll
    for (int v = 0; v < 3; ++v) acc[v] += a * g.v[v];
  }
  __device__ Ctx row_load(int i) const { return Ctx{q[i], drow ? T(1) / drow[i] : T(1), Y4[i]}; }
  __device__ void row_finish(int i, const Ctx& k, const T (&acc)[3], double (&red)[9]) {
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      T r = acc[v] - k.q;                      // helpers.py:88
      if (i < m_ineq) r = fmin(r, T(0));       // :89
      red[3 * v + 0] += (double)r * (double)r;
      red[3 * v + 1] += (double)k.q * (double)k.y.v[v];
      const T ru = r * k.inv;
      red[3 * v + 2] += (double)ru * (double)ru;
    }
  }
  __device__ void finalize(const double (&tot)[9]) {
    for (int v = 0; v < 3; ++v) {
      S->sums[v][S_PR2] = tot[3 * v + 0];
      S->sums[v][S_QTY] = tot[3 * v + 1];
      S->sums[v][S_PR2U] = tot[3 * v + 2];
    }
  }
};

// ---------------------------------------------------------------------------
// Fishnet ("spectral casting", spectral_casting.py): fixed-step PDHG on 8 points at once.  The matrix is
// streamed once per 8 right-hand sides; a gather brings the 8 values of a column (64 contiguous bytes).
// ---------------------------------------------------------------------------
// MODE 0: Y8 = K X8 (spectral_casting.py:99).  MODE 1 (rows of K^T): X8 <- clip(X8 - tau (c - K^T Y8)), XB8 = 2 X8' - X8
// (:272-281).  MODE 2 (rows of K): Y8 <- proj(Y8 + sigma (q - K XB8)) (:285-291).
template <typename T, int MODE>
struct BatchOp {
  using value_type = T;
  using Gather = Oct<T>;
  struct Ctx { T a, b, c; };
  static constexpr int NV = 8;
  static constexpr int NRED = 0;
  static constexpr int UDIV = 4;
  static constexpr int PIPE = 0;
  const Oct<T>* in;      // gathered operand
  Oct<T>* out;           // MODE 0: Y8; MODE 1: X8 (updated in place); MODE 2: Y8 (in place)
  Oct<T>* aux;           // MODE 1: XB8
  const T* v0;           // MODE 1: c; MODE 2: q
  const T* v1;           // MODE 1: l
  const T* v2;           // MODE 1: u
  T step;                // tau or sigma
  int m_ineq;
  __device__ bool begin() { return true; }
  __device__ Oct<T> gather(int col) const { return in[col]; }
  __device__ void mac(T (&acc)[8], T a, const Oct<T>& g) const {
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[v] += a * g.v[v];
  }
  __device__ Ctx row_load(int r) const {
    if (MODE == 1) return Ctx{v0[r], v1[r], v2[r]};
    if (MODE == 2) return Ctx{v0[r], T(0), T(0)};
    return Ctx{T(0), T(0), T(0)};
  }
  __device__ void row_finish(int r, const Ctx& k, const T (&acc)[8], double (&)[1]) {
    Oct<T> o;
    if (MODE == 0) {
#pragma unroll
      for (int v = 0; v < 8; ++v) o.v[v] = acc[v];
      out[r] = o;
    } else if (MODE == 1) {
      const Oct<T> xo = out[r];
      Oct<T> xb;
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        const T xn = clampv(xo.v[v] - step * (k.a - acc[v]), k.b, k.c);
        o.v[v] = xn;
        xb.v[v] = T(2) * xn - xo.v[v];
      }
      out[r] = o;
      aux[r] = xb;
    } else {
      const Oct<T> yo = out[r];
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        T yn = yo.v[v] + step * (k.a - acc[v]);
        if (r < m_ineq) yn = fmax(yn, T(0));
        o.v[v] = yn;
      }
      out[r] = o;
    }
  }
  __device__ void finalize(const double (&)[1]) {}
};

// signed duality-gap terms of 8 points (spectral_casting.py:217-234): rows of K^T, gathers Y8
template <typename T>
struct BatchGapOp {
  using value_type = T;
  using Gather = Oct<T>;
  struct Ctx { T c, l, u; Oct<T> x; };
  static constexpr int NV = 8;
  static constexpr int NRED = 24;
  static constexpr int UDIV = 4;
  static constexpr int PIPE = 0;
  const Oct<T>* Y8;
  const Oct<T>* X8;
  const T* c;
  const T* l;
  const T* u;
  Scalars* S;
  __device__ bool begin() { return true; }
  __device__ Oct<T> gather(int col) const { return Y8[col]; }
  __device__ void mac(T (&acc)[8], T a, const Oct<T>& g) const {
#pragma unroll
    for (int v = 0; v < 8; ++v) acc[v] += a * g.v[v];
  }
  __device__ Ctx row_load(int j) const { return Ctx{c[j], l[j], u[j], X8[j]}; }
  __device__ void row_finish(int, const Ctx& k, const T (&acc)[8], double (&red)[24]) {
    const bool lo_inf = isinf(k.l) && k.l < T(0);
    const bool up_inf = isinf(k.u) && k.u > T(0);
    const T ld = lo_inf ? T(0) : k.l;
    const T ud = up_inf ? T(0) : k.u;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const T g = k.c - acc[v];
      T lam;
      if (lo_inf && up_inf) lam = T(0);
      else if (lo_inf) lam = fmin(g, T(0));
      else if (up_inf) lam = fmax(g, T(0));
      else lam = g;
      red[3 * v + 0] += (double)k.c * (double)k.x.v[v];
      red[3 * v + 1] += (double)ld * (double)fmax(lam, T(0));
      red[3 * v + 2] += (double)ud * (double)fmin(lam, T(0));
    }
  }
  __device__ void finalize(const double (&tot)[24]) {
    for (int i = 0; i < 24; ++i) S->batch[i] = tot[i];
  }
};

// ---------------------------------------------------------------------------
// Row-partitioned (multi-GPU) pieces: SURVEY.md section 8e.  K^T y = sum over ranks of K_g^T y_g, so the column
// side splits into "partial SpMV -> allreduce (host, NCCL) -> element-wise epilogue".
// ---------------------------------------------------------------------------
// out3[3 row + v] = (K_g^T Y4)[row][v] for the three iterates of the restart test
template <typename T>
struct Spmv3Op {
  using value_type = T;
  using Gather = Quad<T>;
  using Ctx = NoCtx;
  static constexpr int NV = 3;
  static constexpr int NRED = 0;
  static constexpr int UDIV = 2;
  static constexpr int PIPE = 0;
  const Quad<T>* Y4;
  T* out3;
  __device__ bool begin() { return true; }
  __device__ Quad<T> gather(int col) const { return Y4[col]; }
  __device__ void mac(T (&acc)[3], T a, const Quad<T>& g) const {
#pragma unroll
    for (int v = 0; v < 3; ++v) acc[v] += a * g.v[v];
  }
  __device__ Ctx row_load(int) const { return Ctx{}; }
  __device__ void row_finish(int r, const Ctx&, const T (&acc)[3], double (&)[1]) {
#pragma unroll
    for (int v = 0; v < 3; ++v) out3[3 * (size_t)r + v] = acc[v];
  }
  __device__ void finalize(const double (&)[1]) {}
};

// tail[0..1] = this rank's partial ||dy||^2, dy^T K dx of the latest iteration (0 when none is pending)
__global__ void shard_tail_pack_kernel(const Scalars* S, double* tail) {
  tail[0] = S->rule_pending ? S->dy2 : 0.0;
  tail[1] = S->rule_pending ? S->inter : 0.0;
  tail[2] = 0.0;
  tail[3] = 0.0;
}
// after the allreduce: the step rule of the pending iteration with the global norms
__global__ void shard_rule_kernel(Scalars* S, const double* tail) {
  Scalars& s = *S;
  if (!s.rule_pending) return;
  s.dy2 = tail[0];
  s.inter = tail[1];
  double w, next;
  step_rule(s, s.dx2, tail[0], tail[1], w, next);
  commit_step(s, w, next);
  s.rule_pending = 0;
}

// ---------------------------------------------------------------------------
// Peer-memory exchange (row-partitioned engine, NVLink / NVSwitch).  Every rank owns one block
//   part[L] doubles | recv[world][seg] LL words | tot[L] LL words          (L = 4 + n rounded up to even)
// where element 0..3 of the exchanged vector is the step-rule tail (it is needed first) and element 4 + j is column j,
// mapped into every peer (CUDA IPC).  An LL word is 16 bytes {lo32, seq, hi32, seq}: one double travelling with
// the sequence number of its exchange in both 8-byte halves, so a reader that sees seq twice has the value --
// no flags, fences or acquire/release round trips, and nothing to reset (seq only grows).  (The accessors are
// volatile asm without a memory clobber: ordered among themselves, but ordinary loads may be scheduled around them.)  One exchange is a
// two-shot allreduce:
//   shard_fused_kernel   its SpMV epilogue (ShardPushOp) stores the partial sum of column j straight into recv[rank]
//                        of the rank that owns j -- NVLink stores issued while the SpMV is still running;
//                 then   every rank polls its recv[*] slots, sums them in rank order and stores the sums into tot[]
//                        of every rank (NVLink stores), evaluates the step rule from the tail and updates x as its
//                        share of tot[] arrives (xchg_reduce_rule_primal).
//   xchg_allreduce_kernel  the same two phases for a caller's buffer (restart test, power iteration).
// With TWO ranks the fused kernel uses a one-shot variant instead (same bytes on the wire, one NVLink hop less): each
// rank stores its partial sums into its slot on BOTH ranks and every rank adds the two slots itself, in rank order.  The
// slots are double-buffered by the parity of seq: a rank can run at most one exchange ahead of its peer (it needs the
// peer's pushes of exchange s + 1 to finish s + 1), so slot (s & 1) is rewritten only after the peer has read it.
// Each segment is summed by exactly one rank, so all ranks see bit-identical totals.  The grid of xchg_kernel is
// co-resident (phase 2 of one GPU waits for phase 1 of the others, never for its own later CTAs).  A poll that
// lasts longer than XCHG_TIMEOUT_NS sets Scalars::xchg_error + stop instead of hanging the GPU.
// ---------------------------------------------------------------------------
constexpr int MAX_PEERS = 16;
constexpr int XCHG_TAIL = 4;
constexpr unsigned long long XCHG_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;
struct Xchg {
  uint4* recv[MAX_PEERS];      // every rank's recv area (world x seg words)
  uint4* tot[MAX_PEERS];       // every rank's tot area (L words)
  uint4* os[MAX_PEERS];        // one-shot exchange (two ranks): every rank's 2 parities x 2 slots x L words
  double* part;                // this rank's partial vector
  int world, rank;
  int L, seg;                  // padded length, elements per segment
  int n, oneshot;
};

__device__ __forceinline__ void ll_store(uint4* p, double v, unsigned int seq) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)b), "r"(seq),
               "r"((unsigned int)(b >> 32)), "r"(seq));
}
__device__ __forceinline__ uint4 ll_load(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ double ll_value(const uint4& v) {
  return __longlong_as_double((long long)(((unsigned long long)v.z << 32) | (unsigned long long)v.x));
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// cooperative copy of the descriptor into shared memory (it is read by every thread afterwards)
__device__ __forceinline__ void xchg_load(Xchg& X, const Xchg* Xp) {
  static_assert(sizeof(Xchg) % 8 == 0, "Xchg is copied in 8-byte words");
  if (threadIdx.x < sizeof(Xchg) / 8)
    reinterpret_cast<unsigned long long*>(&X)[threadIdx.x] = reinterpret_cast<const unsigned long long*>(Xp)[threadIdx.x];
}
// value of an LL word once it carries `seq`; on timeout (or after an earlier one) 0 and the sticky error
__device__ __noinline__ double ll_wait_slow(const uint4* p, unsigned int seq, Scalars* S) {
  unsigned long long t0 = 0;
  unsigned int polls = 0;
  for (;;) {
    const uint4 v = ll_load(p);
    if (v.y == seq && v.w == seq) return ll_value(v);
    if ((++polls & 0xffu) == 0) {
      if (*(volatile int*)&S->xchg_error) return 0.0;
      const unsigned long long now = global_timer_ns();
      if (t0 == 0) t0 = now;
      else if (now - t0 > XCHG_TIMEOUT_NS) {
        S->xchg_error = 1;
        S->stop = 1;
        return 0.0;
      }
    }
  }
}
// v: a speculative load of *p issued earlier (the common case: it already carries seq)
__device__ __forceinline__ double ll_wait(const uint4* p, unsigned int seq, Scalars* S, const uint4& v) {
  if (v.y == seq && v.w == seq) return ll_value(v);
  return ll_wait_slow(p, seq, S);
}
__device__ __forceinline__ double ll_wait(const uint4* p, unsigned int seq, Scalars* S) {
  return ll_wait(p, seq, S, ll_load(p));
}

// Phase 1 of an exchange of src[0 .. count): every element goes to the recv[] slot of its owner.
// All loads of a batch are issued before the first store (volatile accesses are not reordered by the compiler).
__device__ __forceinline__ void xchg_scatter(const Xchg& X, const double* src, int count, unsigned int seq, int tid,
                                             int nthr) {
  constexpr int NB = 4;
  for (int e0 = tid; e0 < count; e0 += NB * nthr) {
    double v[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (e0 + b * nthr < count) v[b] = src[e0 + b * nthr];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const int e = e0 + b * nthr;
      if (e < count) {
        const int s = e / X.seg;
        ll_store(X.recv[s] + (size_t)X.rank * X.seg + (e - s * X.seg), v[b], seq);
      }
    }
  }
}
// Phase 2: this rank sums, in rank order, the slots of the segment it owns as they arrive and stores the totals into
// tot[] of every rank.  All loads of a batch are issued before the first one is inspected (a loop of ll_wait would
// pay one L2 / NVLink latency per peer).
__device__ __forceinline__ void xchg_reduce_owned(const Xchg& X, int count, unsigned int seq, Scalars* S, int tid, int nthr) {
  const int s0 = X.rank * X.seg;
  const int len = max(0, min(count, s0 + X.seg) - s0);
  const uint4* mine = X.recv[X.rank];
  for (int i = tid; i < len; i += nthr) {
    double a = 0.0;
    for (int g0 = 0; g0 < X.world; g0 += 8) {
      uint4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (g0 + j < X.world) v[j] = ll_load(mine + (size_t)(g0 + j) * X.seg + i);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (g0 + j < X.world) a += ll_wait(mine + (size_t)(g0 + j) * X.seg + i, seq, S, v[j]);     // rank order
    }
    for (int p = 0; p < X.world; ++p) ll_store(X.tot[p] + s0 + i, a, seq);
  }
}
__device__ __forceinline__ void xchg_scatter_reduce(const Xchg& X, const double* src, int count, unsigned int seq,
                                                    Scalars* S, int tid, int nthr) {
  xchg_scatter(X, src, count, seq, tid, nthr);
  xchg_reduce_owned(X, count, seq, S, tid, nthr);
}

// stand-alone allreduce (sum) of buf[0 .. count) over the same exchange block, count <= L: the small collectives of
// the restart test and the power iteration (pdlp_shard_allreduce_p2p).  Co-resident grid; the last CTA advances seq.
__global__ void __launch_bounds__(256) xchg_allreduce_kernel(const Xchg* Xp, Scalars* S, double* buf, int count,
                                                             unsigned int* ticket) {
  __shared__ Xchg X;
  __shared__ int s_last;
  xchg_load(X, Xp);
  __syncthreads();
  const unsigned int seq = (unsigned int)S->xchg_seq;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  xchg_scatter_reduce(X, buf, count, seq, S, tid, nthr);
  const uint4* tot = X.tot[X.rank];
  for (int e = tid; e < count; e += nthr) buf[e] = ll_wait(tot + e, seq, S);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    S->xchg_seq = (unsigned long long)seq + 1;
    *ticket = 0u;
  }
}

// element-wise K1 epilogue on the allreduced K^T y (same arithmetic as PrimalOp::row_finish)
template <typename T, bool SAVE_PREV>
__global__ void shard_primal_kernel(const double* kty, const T* c, const T* l, const T* u, Pair<T>* xp, T* xsum,
                                    Quad<T>* X4, int n, Scalars* S, double* partials, unsigned int* ticket, T* xb) {
  if (S->stop) return;
  const T tau = (T)(S->eta / S->omega);
  const T wpend = (T)S->w_pending;
  const T theta = (T)S->theta;
  const int stride = gridDim.x * blockDim.x;
  double dx2 = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const T g = c[j] - (T)kty[j];
    const T xo = xp[j].a;
    const T xn = clampv(xo - tau * g, l[j], u[j]);
    const T dx = xn - xo;
    xp[j] = Pair<T>{xn, dx};
    if (xb) xb[j] = xn + theta * dx;
    xsum[j] += wpend * xo;
    if (SAVE_PREV) X4[j].v[IT_PREV] = xo;
    dx2 += (double)dx * (double)dx;
  }
  __shared__ double sh[32];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dx2 += __shfl_xor_sync(0xffffffffu, dx2, o);
  if (lane == 0) sh[warp] = dx2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sh[w];
    partials[blockIdx.x] = a;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && warp == 0) {           // one warp, batched loads (grid_totals): not a serial chain over the blocks
    __threadfence();
    double a[1];
    grid_totals<1>(partials, 1, (int)gridDim.x, a);
    if (lane == 0) {
      S->dx2 = a[0];
      *ticket = 0u;
    }
  }
}

// Everything of a sharded iteration's column side that follows the scatter of the partial K_g^T y_g: reduce the owned
// segment and broadcast it, then every CTA polls the tail of tot[], evaluates the pending step rule for itself
// (read-only on Scalars), polls its share of tot[] and updates x; the last CTA -- every reader of Scalars is done by
// then -- commits the rule, ||dx||^2 and the sequence number.  The grid must be co-resident.  sh: >= blockDim / 32
// doubles, s_rule: 6 doubles, s_int: 3 ints of shared memory.
template <typename T, bool SAVE_PREV, bool ONESHOT>
__device__ __forceinline__ void xchg_reduce_rule_primal(const Xchg& X, unsigned int seq, const T* c, const T* l, const T* u,
                                                        Pair<T>* xp, T* xsum, Quad<T>* X4, Scalars* S, double* partials,
                                                        unsigned int* ticket, double* sh, double* s_rule, int* s_int,
                                                        T* xb) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthr = gridDim.x * blockDim.x;
  if (!ONESHOT) xchg_reduce_owned(X, X.L, seq, S, tid, nthr);
  // two-shot: the totals arrive in tot[]; one-shot (two ranks): both ranks' partial vectors arrive in this rank's two
  // slots of parity seq & 1 and are added here, in rank order, by every rank for itself (same bits on both)
  const uint4* tot = ONESHOT ? X.os[X.rank] + (size_t)((seq & 1u) * 2u) * X.L : X.tot[X.rank];
  const uint4* tot1 = tot + X.L;          // one-shot: rank 1's slot
  if (threadIdx.x == 0) {
    const uint4 v0 = ll_load(tot + 0), v1 = ll_load(tot + 1);
    double dy2 = ll_wait(tot + 0, seq, S, v0);
    double inter = ll_wait(tot + 1, seq, S, v1);
    if (ONESHOT) {
      const uint4 w0 = ll_load(tot1 + 0), w1 = ll_load(tot1 + 1);
      dy2 = dy2 + ll_wait(tot1 + 0, seq, S, w0);
      inter = inter + ll_wait(tot1 + 1, seq, S, w1);
    }
    const int pending = S->rule_pending;
    double eta = S->eta, wp = S->w_pending, w = 0.0, next = 0.0;
    if (pending) {
      rule_values(S->adaptive, S->k + 1, eta, S->omega, S->dx2, dy2, inter, w, next);
      eta = next;
      wp = w;
    }
    s_rule[0] = eta; s_rule[1] = wp; s_rule[2] = dy2; s_rule[3] = inter; s_rule[4] = w; s_rule[5] = next;
    s_int[1] = pending;
    s_int[2] = S->stop;
  }
  __syncthreads();
  const int s_stop = s_int[2];
  double dx2 = 0.0;
  if (!s_stop) {
    const T tau = (T)(s_rule[0] / S->omega);
    const T wpend = (T)s_rule[1];
    const T theta = (T)S->theta;
    const uint4* col = tot + XCHG_TAIL;
    const uint4* col1 = tot1 + XCHG_TAIL;
    constexpr int NB = 4;
    for (int j0 = tid; j0 < X.n; j0 += NB * nthr) {
      uint4 v[NB], w[ONESHOT ? NB : 1];
      T cj[NB], lj[NB], uj[NB], sj[NB];
      Pair<T> pj[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int j = j0 + b * nthr;
        if (j < X.n) {
          v[b] = ll_load(col + j);
          if (ONESHOT) w[ONESHOT ? b : 0] = ll_load(col1 + j);
          cj[b] = c[j]; lj[b] = l[j]; uj[b] = u[j]; sj[b] = xsum[j]; pj[b] = xp[j];
        }
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const int j = j0 + b * nthr;
        if (j < X.n) {
          double kty = ll_wait(col + j, seq, S, v[b]);
          if (ONESHOT) kty = (0.0 + kty) + ll_wait(col1 + j, seq, S, w[ONESHOT ? b : 0]);     // rank order, as the owner sums
          const T g = cj[b] - (T)kty;
          const T xo = pj[b].a;
          const T xn = clampv(xo - tau * g, lj[b], uj[b]);
          const T dx = xn - xo;
          xp[j] = Pair<T>{xn, dx};
          if (xb) xb[j] = xn + theta * dx;
          xsum[j] = sj[b] + wpend * xo;
          if (SAVE_PREV) X4[j].v[IT_PREV] = xo;
          dx2 += (double)dx * (double)dx;
        }
      }
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dx2 += __shfl_xor_sync(0xffffffffu, dx2, o);
  __syncthreads();          // sh may still be in use by the caller's earlier phase
  if (lane == 0) sh[warp] = dx2;
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sh[w];
    partials[blockIdx.x] = a;
    __threadfence();
    s_int[0] = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_int[0] && warp == 0) {
    __threadfence();
    double a[1] = {0.0};
    if (!s_stop) grid_totals<1>(partials, 1, (int)gridDim.x, a);      // one warp, batched loads
    if (lane == 0) {
      Scalars& s = *S;
      if (s_int[1]) {           // commit the rule every CTA evaluated (shard_rule_kernel's arithmetic)
        s.k += 1;
        s.dy2 = s_rule[2];
        s.inter = s_rule[3];
        commit_step(s, s_rule[4], s_rule[5]);
        s.rule_pending = 0;
      }
      if (!s_stop) s.dx2 = a[0];
      s.xchg_seq = (unsigned long long)seq + 1;
      *ticket = 0u;
    }
  }
}

// rows of K_g^T: the partial sum of column j goes, as an LL word, straight into the recv[] slot of the rank that owns
// it -- over NVLink while the SpMV is still running, no staging copy in HBM
template <typename T>
struct ShardPushOp {
  using value_type = T;
  using Gather = T;
  using Ctx = NoCtx;
  static constexpr int NV = 1;
  static constexpr int NRED = 0;
  static constexpr int UDIV = 1;
  static constexpr int PIPE = 1;
  const T* in;
  const Xchg* X;          // shared memory
  unsigned int seq;
  __device__ bool begin() { return true; }
  __device__ T gather(int col) const { return ldg(in + col); }
  __device__ void mac(T (&acc)[1], T a, const T& g) const { acc[0] += a * g; }
  __device__ Ctx row_load(int) const { return Ctx{}; }
  __device__ void row_finish(int r, const Ctx&, const T (&acc)[1], double (&)[1]) {
    const int e = XCHG_TAIL + r;
    if (X->oneshot) {         // two ranks: the partial sum goes to this rank's slot on BOTH ranks
      const size_t off = (size_t)((seq & 1u) * 2u + (unsigned)X->rank) * X->L + e;
      ll_store(X->os[0] + off, (double)acc[0], seq);
      ll_store(X->os[1] + off, (double)acc[0], seq);
      return;
    }
    const int s = e / X->seg;
    ll_store(X->recv[s] + (size_t)X->rank * X->seg + (e - s * X->seg), (double)acc[0], seq);
  }
  __device__ void finalize(const double (&)[1]) {}
};

// ONE kernel for the column side of a sharded iteration (pdlp_shard_iterate_p2p): the K_g^T y_g SpMV whose row
// epilogue pushes the partial sums to their owners, then xchg_reduce_rule_primal.  Persistent, co-resident grid
// (one CTA per SM on every rank); CTAs poll only after their own pushes are out, so there is no circular wait.
template <typename T, bool SAVE_PREV>
__global__ void __launch_bounds__(StreamCfg<T>::NT, 1)
shard_fused_kernel(const SellView<T> A, const Xchg* Xp, const T* y, const T* c, const T* l, const T* u, Pair<T>* xp,
                   T* xsum, Quad<T>* X4, Scalars* S, double* partials, unsigned int* ticket, T* xb) {
  using Cfg = StreamCfg<T>;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t full_bar[Cfg::W][Cfg::SMAX];
  __shared__ double s_red[Cfg::W][1];
  __shared__ Xchg X;
  __shared__ double sh[32];
  __shared__ double s_rule[6];
  __shared__ int s_int[3];

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  Ring R{smem + (size_t)warp * A.nstages * A.stage_bytes, &full_bar[warp][0], A.nstages, A.stage_bytes, 0u};
  if (lane == 0) {
    for (int s = 0; s < R.S; ++s) mbar_init(&R.bars[s], 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  xchg_load(X, Xp);
  __syncthreads();
  const int gw = blockIdx.x * Cfg::W + warp;
  const int pb = A.warp_piece_begin[gw];
  const int pe = A.warp_piece_begin[gw + 1];
  ring_prologue<T>(A, R, pb, pe, lane);

  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const unsigned int seq = (unsigned int)S->xchg_seq;
  if (blockIdx.x == 0 && tid == 0) {       // the step-rule partials of the previous iteration travel in the tail
    const double t0 = S->rule_pending ? S->dy2 : 0.0, t1 = S->rule_pending ? S->inter : 0.0;
    for (int p = 0; p < (X.oneshot ? 2 : 1); ++p) {
      uint4* dst = X.oneshot ? X.os[p] + (size_t)((seq & 1u) * 2u + (unsigned)X.rank) * X.L
                             : X.recv[0] + (size_t)X.rank * X.seg;      // elements 0 .. 3 belong to rank 0's segment
      ll_store(dst + 0, t0, seq);
      ll_store(dst + 1, t1, seq);
      ll_store(dst + 2, 0.0, seq);
      ll_store(dst + 3, 0.0, seq);
    }
  }
  ShardPushOp<T> op{y, &X, seq};
  double red[1] = {0.0};
  stream_phase<ShardPushOp<T>, 1>(A, op, red, R, pb, pe, s_red);
  if (X.oneshot)
    xchg_reduce_rule_primal<T, SAVE_PREV, true>(X, seq, c, l, u, xp, xsum, X4, S, partials, ticket, sh, s_rule, s_int, xb);
  else
    xchg_reduce_rule_primal<T, SAVE_PREV, false>(X, seq, c, l, u, xp, xsum, X4, S, partials, ticket, sh, s_rule, s_int, xb);
}

// element-wise KKT column side on the allreduced K^T Y4 (same arithmetic as KktColsOp::row_finish); one block
// per slab, fixed-order reduction of the 15 sums
template <typename T>
__global__ void shard_kkt_cols_kernel(const double* kty3, const Quad<T>* X4, const T* c, const T* l, const T* u,
                                      const T* dcol, int n, Scalars* S, double* partials, unsigned int* ticket) {
  double red[15];
#pragma unroll
  for (int i = 0; i < 15; ++i) red[i] = 0.0;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const T cj = c[j], lj = l[j], uj = u[j];
    const bool lo_inf = isinf(lj) && lj < T(0);
    const bool up_inf = isinf(uj) && uj > T(0);
    const T ld = lo_inf ? T(0) : lj;
    const T ud = up_inf ? T(0) : uj;
    const T inv = dcol ? T(1) / dcol[j] : T(1);
    const Quad<T> x = X4[j];
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      const T g = cj - (T)kty3[3 * (size_t)j + v];
      T lam;
      if (lo_inf && up_inf) lam = T(0);
      else if (lo_inf) lam = fmin(g, T(0));
      else if (up_inf) lam = fmax(g, T(0));
      else lam = g;
      const T d = g - lam;
      red[5 * v + 0] += (double)d * (double)d;
      red[5 * v + 1] += (double)cj * (double)x.v[v];
      red[5 * v + 2] += (double)ld * (double)fmax(lam, T(0));
      red[5 * v + 3] += (double)ud * (double)fmin(lam, T(0));
      const T du = d * inv;
      red[5 * v + 4] += (double)du * (double)du;
    }
  }
  __shared__ double sh[32][15];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 15; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) red[i] += __shfl_xor_sync(0xffffffffu, red[i], o);
  }
  if (lane == 0)
    for (int i = 0; i < 15; ++i) sh[warp][i] = red[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < 15; ++i) {
      double a = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sh[w][i];
      partials[(size_t)blockIdx.x * 16 + i] = a;
    }
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && (threadIdx.x >> 5) == 0) {
    __threadfence();
    double t[15];
    grid_totals<15>(partials, 16, (int)gridDim.x, t);       // one warp, batched loads
    if ((threadIdx.x & 31) == 0) {
      for (int v = 0; v < 3; ++v) {
        S->sums[v][S_DR2] = t[5 * v + 0];
        S->sums[v][S_CTX] = t[5 * v + 1];
        S->sums[v][S_LPOS] = t[5 * v + 2];
        S->sums[v][S_UNEG] = t[5 * v + 3];
        S->sums[v][S_DR2U] = t[5 * v + 4];
      }
      S->w_pending = 0.0;
      *ticket = 0u;
    }
  }
}

// the row-side KKT sums (partial over this rank's rows) <-> a device buffer the host allreduces
__global__ void shard_rows_pack_kernel(const Scalars* S, double* buf) {
  for (int v = 0; v < 3; ++v) {
    buf[3 * v + 0] = S->sums[v][S_PR2];
    buf[3 * v + 1] = S->sums[v][S_QTY];
    buf[3 * v + 2] = S->sums[v][S_PR2U];
  }
}
__global__ void shard_rows_unpack_kernel(Scalars* S, const double* buf) {
  for (int v = 0; v < 3; ++v) {
    S->sums[v][S_PR2] = buf[3 * v + 0];
    S->sums[v][S_QTY] = buf[3 * v + 1];
    S->sums[v][S_PR2U] = buf[3 * v + 2];
  }
}
__global__ void shard_dy_pack_kernel(const Scalars* S, double* buf) { buf[0] = S->restart_dy2; }
__global__ void shard_dy_unpack_kernel(Scalars* S, const double* buf) { S->restart_dy2 = buf[0]; }

// ---------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------
__host__ __device__ inline Resid make_resid(const double (&s)[S_NSUM], double omega) {
  Resid r;
  r.pr = sqrt(s[S_PR2]);
  r.dr = sqrt(s[S_DR2]);
  r.prim_obj = s[S_CTX];
  r.adj_dual = s[S_QTY] + s[S_LPOS] + s[S_UNEG];     // helpers.py:84
  r.gap = r.adj_dual - r.prim_obj;                   // :85 (signed)
  const double w2 = omega * omega;
  r.kkt = sqrt(w2 * r.pr * r.pr + r.dr * r.dr / w2 + r.gap * r.gap);   // :105
  r.pr_u = sqrt(s[S_PR2U]);
  r.dr_u = sqrt(s[S_DR2U]);
  return r;
}

// restart criteria, primal_dual_hybrid_gradient.py:121-141
__global__ void decide_kernel(Scalars* S, int force) {
  Scalars& s = *S;
  if (s.stop) return;
  const Resid rc = make_resid(s.sums[IT_CUR], s.omega);
  const Resid ra = make_resid(s.sums[IT_AVG], s.omega);
  const Resid rp = make_resid(s.sums[IT_PREV], s.omega);
  s.kkt_cur = rc.kkt;
  s.kkt_avg = ra.kkt;
  s.kkt_prev = rp.kkt;
  const double kmin = fmin(rc.kkt, ra.kkt);
  int kind = 0;
  int use_avg = rc.kkt >= ra.kkt ? 1 : 0;
  if (force) {
    kind = 4;
    use_avg = 0;
  } else if (kmin <= 0.2 * s.kkt_first) {
    kind = 1;
  } else if (kmin <= 0.8 * s.kkt_first && kmin > rp.kkt) {
    kind = 2;
  } else if ((double)s.t >= 0.36 * (double)s.k) {
    kind = 3;
  }
  s.chk_kind = kind;
  s.chk_use_avg = use_avg;
  s.chk_t = s.t;
  s.chk_k = s.k;
  s.chk_solved = 0;
  s.point = use_avg ? ra : rc;
  s.restart_dx2 = 0.0;
  s.restart_dy2 = 0.0;
}

// after apply_restart: primal weight (enhancements.py:73-78), KKT_first (:153), termination (:173)
__global__ void post_restart_kernel(Scalars* S, int scaled) {
  Scalars& s = *S;
  if (s.stop || s.chk_kind == 0) return;
  s.n_restarts += 1;
  if (s.primal_update) {
    const double dxn = sqrt(s.restart_dx2), dyn = sqrt(s.restart_dy2);
    if (dxn > 0.0 && dyn > 0.0) s.omega = exp(0.5 * log(dyn / dxn) + 0.5 * log(s.omega));
  }
  const double w2 = s.omega * s.omega;
  Resid& p = s.point;
  p.kkt = sqrt(w2 * p.pr * p.pr + p.dr * p.dr / w2 + p.gap * p.gap);
  s.kkt_first = p.kkt;
  s.kkt_first_new = p.kkt;
  s.chk_omega = s.omega;
  const double pr = scaled ? p.pr_u : p.pr;
  const double dr = scaled ? p.dr_u : p.dr;
  const bool ok = pr <= s.tol * (1.0 + s.q_norm) && dr <= s.tol * (1.0 + s.c_norm) &&
                  p.gap <= s.tol * (1.0 + fabs(p.prim_obj) + fabs(p.adj_dual));     // helpers.py:125-128
  s.chk_solved = ok ? 1 : 0;
  s.solved = ok ? 1 : 0;
  if (ok && s.stop_on_solved) s.stop = 1;
  s.t = 0;
  s.eta_sum = 0.0;
  s.w_pending = 0.0;
}

// ---------------------------------------------------------------------------
// element-wise kernels
// ---------------------------------------------------------------------------
// fold the pending weight into the sums and lay out the iterates for the KKT pass (:118-119)
template <typename T>
__global__ void flush_kernel(const Pair<T>* xp, const T* y, T* xsum, T* ysum, Quad<T>* X4, Quad<T>* Y4, int n, int m,
                             const Scalars* S, int have_prev) {
  if (S->stop) return;
  const T wp = (T)S->w_pending;
  const T es = (T)S->eta_sum;
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const T xo = xp[j].a;
    const T sm = xsum[j] + wp * xo;
    xsum[j] = sm;
    Quad<T> qd = X4[j];
    qd.v[IT_CUR] = xo;
    qd.v[IT_AVG] = es > T(0) ? sm / es : xo;
    if (!have_prev) qd.v[IT_PREV] = xo;
    qd.v[3] = T(0);
    X4[j] = qd;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const T yo = y[i];
    const T sm = ysum[i] + wp * yo;
    ysum[i] = sm;
    Quad<T> qd = Y4[i];
    qd.v[IT_CUR] = yo;
    qd.v[IT_AVG] = es > T(0) ? sm / es : yo;
    if (!have_prev) qd.v[IT_PREV] = yo;
    qd.v[3] = T(0);
    Y4[i] = qd;
  }
}

// restart at the chosen candidate: copy, zero the sums, ||x - x_last_restart||^2 (:58-64,:133,:150)
template <typename T>
__global__ void apply_restart_kernel(Pair<T>* xp, T* y, T* xsum, T* ysum, const Quad<T>* X4, const Quad<T>* Y4,
                                     T* xlast, T* ylast, int n, int m, Scalars* S, double* partials,
                                     unsigned int* ticket) {
  if (S->stop || S->chk_kind == 0) return;
  const int slot = S->chk_use_avg ? IT_AVG : IT_CUR;
  const int stride = gridDim.x * blockDim.x;
  double dx2 = 0.0, dy2 = 0.0;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const T xc = X4[j].v[slot];
    const T d = xlast[j] - xc;
    dx2 += (double)d * (double)d;
    xp[j] = Pair<T>{xc, T(0)};
    xlast[j] = xc;
    xsum[j] = T(0);
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const T yc = Y4[i].v[slot];
    const T d = ylast[i] - yc;
    dy2 += (double)d * (double)d;
    y[i] = yc;
    ylast[i] = yc;
    ysum[i] = T(0);
  }
  // fixed-order block + grid reduction
  __shared__ double sh[2][32];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dx2 += __shfl_xor_sync(0xffffffffu, dx2, o);
    dy2 += __shfl_xor_sync(0xffffffffu, dy2, o);
  }
  if (lane == 0) { sh[0][warp] = dx2; sh[1][warp] = dy2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += sh[0][w]; b += sh[1][w]; }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && warp == 0) {
    __threadfence();
    double a = 0.0, b = 0.0;
    for (int bl = lane; bl < (int)gridDim.x; bl += 32) { a += __ldcg(&partials[2 * bl]); b += __ldcg(&partials[2 * bl + 1]); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      S->restart_dx2 = a;
      S->restart_dy2 = b;
      *ticket = 0u;
    }
  }
}

template <typename T>
__global__ void set_iterate_kernel(const T* x, const T* y, Pair<T>* xp, T* yy, T* xsum, T* ysum, T* xlast, T* ylast,
                                   int n, int m) {
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const T v = x ? x[j] : T(0);
    xp[j] = Pair<T>{v, T(0)};
    xsum[j] = T(0);
    xlast[j] = v;
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
    const T v = y ? y[i] : T(0);
    yy[i] = v;
    ysum[i] = T(0);
    ylast[i] = v;
  }
}

template <typename T>
__global__ void get_iterate_kernel(const Pair<T>* xp, const T* yy, T* x, T* y, int n, int m) {
  const int stride = gridDim.x * blockDim.x;
  if (x) for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) x[j] = xp[j].a;
  if (y) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) y[i] = yy[i];
}

template <typename T>
__global__ void get_average_kernel(const Pair<T>* xp, const T* yy, const T* xsum, const T* ysum, T* x, T* y, int n,
                                   int m, const Scalars* S) {
  const T wp = (T)S->w_pending, es = (T)S->eta_sum;
  const int stride = gridDim.x * blockDim.x;
  if (x) for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride)
    x[j] = es > T(0) ? (xsum[j] + wp * xp[j].a) / es : xp[j].a;
  if (y) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride)
    y[i] = es > T(0) ? (ysum[i] + wp * yy[i]) / es : yy[i];
}

// X4[j] = (x, x, x, 0): standalone KKT of caller vectors
template <typename T>
__global__ void pack3_kernel(const T* x, Quad<T>* X4, int n) {
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const T v = x[j];
    X4[j] = Quad<T>{{v, v, v, T(0)}};
  }
}

// c^T x of the current iterate -> scratch[slot] (single block, fixed order)
template <typename T>
__global__ void objective_kernel(const T* c, const Pair<T>* xp, int n, Scalars* S, int slot) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) a += (double)c[j] * (double)xp[j].a;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    S->scratch[slot] = t;
  }
}

// sum of squares of a vector -> scratch[slot] (single block, fixed order)
template <typename T>
__global__ void sumsq_kernel(const T* v, int n, Scalars* S, int slot) {
  __shared__ double sh[32];
  double a = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) a += (double)v[j] * (double)v[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    S->scratch[slot] = t;
  }
}

// ---------------------------------------------------------------------------
// preconditioning (setup time; enhancements.py:43-71)
// ---------------------------------------------------------------------------
// out[row] = max |vals| (MODE 0) or sum |vals| (MODE 1) over the row; one warp per row
template <typename T, int MODE>
__global__ void row_absreduce_kernel(const int* rowptr, const T* vals, int rows, T* out) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    T a = T(0);
    for (int e = rowptr[r] + lane; e < rowptr[r + 1]; e += 32) {
      const T v = fabs(vals[e]);
      a = MODE == 0 ? fmax(a, v) : a + v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const T b = __shfl_xor_sync(0xffffffffu, a, o);
      a = MODE == 0 ? fmax(a, b) : a + b;
    }
    if (lane == 0) out[r] = a;
  }
}

// norms <- sqrt(norms), entries < eps -> 1, D /= norms; scratch[slot] = max |1 - norms| (single block)
template <typename T>
__global__ void ruiz_fix_kernel(T* norms, T* D, int len, T eps, Scalars* S, int slot) {
  __shared__ double sh[32];
  double mx = 0.0;
  for (int i = threadIdx.x; i < len; i += blockDim.x) {
    T r = sqrt(norms[i]);
    if (r < eps) r = T(1);
    norms[i] = r;
    D[i] = D[i] / r;
    mx = fmax(mx, fabs(1.0 - (double)r));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t = fmax(t, sh[w]);
    S->scratch[slot] = t;
  }
}

template <typename T>
__global__ void scale_rows_kernel(const int* rowptr, T* vals, int rows, const T* div) {
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const T d = div[r];
    for (int e = rowptr[r] + lane; e < rowptr[r + 1]; e += 32) vals[e] = vals[e] / d;
  }
}

template <typename T>
__global__ void scale_cols_kernel(const int* colidx, T* vals, int nnz, const T* div) {
  const int stride = gridDim.x * blockDim.x;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += stride) vals[e] = vals[e] / div[colidx[e]];
}

// v = v * s (MODE 0) or v / s (MODE 1), element-wise
template <typename T, int MODE>
__global__ void vec_scale_kernel(T* v, const T* s, int len) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) v[i] = MODE == 0 ? v[i] * s[i] : v[i] / s[i];
}

// clu[j] = (c_j, l_j, u_j, 0): the constant operands of the primal epilogue in one sector
template <typename T>
__global__ void pack_clu_kernel(const T* c, const T* l, const T* u, Quad<T>* clu, int n) {
  const int stride = gridDim.x * blockDim.x;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) clu[j] = Quad<T>{{c[j], l[j], u[j], T(0)}};
}

template <typename T>
__global__ void fill_kernel(T* v, T val, int len) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) v[i] = val;
}

}  // namespace pdlp
