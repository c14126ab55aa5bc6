// pdlp_b200.cu -- engine, plan builder and C ABI (include/pdlp_b200.h).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <mutex>
#include <queue>
#include <utility>
#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "../../include/pdlp_b200.h"
#include "pdlp_block.cuh"
#include "pdlp_ops.cuh"
#include "panel_spmv.cuh"
#include "run_spmv.cuh"

using namespace pdlp;

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CU(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(PDLP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                     std::to_string(__LINE__) + ")");                              \
  } while (0)

extern "C" const char* pdlp_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------------------
// plan: SELL-32-sigma slices, pieces, per-warp piece ranges (layout described in stream_spmv.cuh)
// ---------------------------------------------------------------------------
struct HostSell {
  std::vector<Piece> pieces;
  std::vector<int> perm;               // row per (slice, lane), -1 = padding lane
  std::vector<int> warp_begin;         // grid * W + 1
  std::vector<int> long_rows;
  size_t packed_bytes = 0;
  int grid = 1;
  int max_piece_bytes = 16;
  int sigma = 0;                       // sorting window the cost model chose
  int balanced = 0;                    // 1: slices dealt out to the warps longest-first (LPT) instead of contiguous ranges
};

static HostSell build_sell_plan(const int* rowptr, int rows, int value_bytes, int W, int KB, int sigma,
                                int long_row, int warp_row, int max_grid, double epi_sectors = 6.0) {
  HostSell P;
  auto len = [&](int r) { return rowptr[r + 1] - rowptr[r]; };
  // three classes of rows: one per lane (SELL slices), one per warp (32 equal segments, one per lane: the heavy
  // tail of a power-law matrix would otherwise pad its slice mates and make single slices larger than a warp's
  // share of the matrix), one per CTA (the few giants, straight from CSR)
  std::vector<int> order, warp_rows;
  order.reserve(rows);
  for (int r = 0; r < rows; ++r) {
    if (len(r) > long_row) P.long_rows.push_back(r);
    else if (len(r) > warp_row) warp_rows.push_back(r);
    else order.push_back(r);
  }
  // Sort by length (descending, stable) inside windows of sigma rows: neighbours in a slice get equal lengths.
  // sigma is chosen per matrix by a cost model (below): heavy-tailed row lengths need a wide window to avoid padding,
  // near-uniform ones keep a narrow window and with it the locality of the row epilogue.
  // (lengths are small integers, <= long_row: histograms and a stable counting sort instead of comparison sorts --
  // the planner is on the end-to-end path of every solve)
  std::vector<int> hist, tmp;
  auto window_hist = [&](size_t w0, size_t w1, int& lo, int& hi) {
    lo = 0x7fffffff;
    hi = 0;
    for (size_t i = w0; i < w1; ++i) {
      const int L = len(order[i]);
      lo = std::min(lo, L);
      hi = std::max(hi, L);
    }
    hist.assign((size_t)(hi - lo + 1), 0);
    for (size_t i = w0; i < w1; ++i) ++hist[(size_t)(len(order[i]) - lo)];
  };
  {
    // sigma minimises  padded slots x (index + value bytes)  +  rows x epilogue sector bytes x g(sigma):
    // a wide window removes padding, a narrow one keeps the rows of a slice (and of the next slices of the same
    // warp) within a few cache lines, so that the scattered epilogue accesses share sectors in L1.  g is the
    // measured share of epilogue accesses that fetch their own 32-byte sector from L2 (configs 2 and 4, K and K^T:
    // sigma = 128 / 256 / 1024 -> 36.7 / 39.1 / 41.2 us for K1 of config 2, 282 / 293 / 363 us for K1 of config 4).
    auto g = [](int cand) { return cand <= 128 ? 0.25 : cand <= 256 ? 0.35 : cand <= 512 ? 0.6 : 1.0; };
    const double epi_bytes = epi_sectors * 32.0;    // scattered vector accesses of the row epilogue
    double best_cost = 0.0;
    int best = sigma;
    long long nnz_short = 0;
    for (int r : order) nnz_short += len(r);
    for (int cand = sigma; cand <= (1 << 18); cand = cand < 1024 ? cand * 2 : cand * 4) {
      long long padded = 0;
      for (size_t w0 = 0; w0 < order.size(); w0 += (size_t)cand) {
        const size_t w1 = std::min(order.size(), w0 + (size_t)cand);
        int lo, hi;
        window_hist(w0, w1, lo, hi);
        size_t placed = 0;      // rows of the window already placed in descending order; a slice starts every 32
        for (int L = hi; L >= lo; --L) {
          const size_t cnt = (size_t)hist[(size_t)(L - lo)];
          if (!cnt) continue;
          padded += 32LL * L * (long long)((placed + cnt + 31) / 32 - (placed + 31) / 32);
          placed += cnt;
        }
      }
      const double cost = (double)padded * (4 + value_bytes) + (double)order.size() * epi_bytes * g(cand);
      if (cand == sigma || cost < best_cost) {
        best_cost = cost;
        best = cand;
      }
      if ((size_t)cand >= order.size()) break;
      // no wider window can win once even zero padding would not pay for its worse locality
      const int next = cand < 1024 ? cand * 2 : cand * 4;
      if ((double)nnz_short * (4 + value_bytes) + (double)order.size() * epi_bytes * g(next) >= best_cost) break;
    }
    sigma = best;
    P.sigma = sigma;
    if (getenv("PDLP_TRACE_SETUP")) fprintf(stderr, "[pdlp setup] sorting window %d rows\n", sigma);
  }
  for (size_t w0 = 0; w0 < order.size(); w0 += (size_t)sigma) {     // stable, descending by length
    const size_t w1 = std::min(order.size(), w0 + (size_t)sigma);
    int lo, hi;
    window_hist(w0, w1, lo, hi);
    int run = 0;
    for (int L = hi; L >= lo; --L) {                                  // hist -> start offset of each length
      const int cnt = hist[(size_t)(L - lo)];
      hist[(size_t)(L - lo)] = run;
      run += cnt;
    }
    tmp.assign(order.begin() + w0, order.begin() + w1);
    for (int r : tmp) order[w0 + (size_t)hist[(size_t)(len(r) - lo)]++] = r;
  }
  while (order.size() % 32) order.push_back(-1);
  const int nsl_lane = (int)(order.size() / 32);
  for (int r : warp_rows) order.insert(order.end(), 32, r);       // a warp-row slice: the same row in all 32 lanes
  int nsl = (int)(order.size() / 32);
  std::vector<int> klen(nsl, 0);
  std::vector<char> wr(nsl, 0), cut;                               // warp-row slice; LPT: a warp's range starts here
  for (int i = 0; i < nsl_lane; ++i)
    for (int r = 0; r < 32; ++r) {
      const int row = order[(size_t)i * 32 + r];
      if (row >= 0) klen[i] = std::max(klen[i], len(row));
    }
  for (int i = nsl_lane; i < nsl; ++i) {
    klen[i] = (len(order[(size_t)i * 32]) + 31) / 32;
    wr[i] = 1;
  }
  const int MAXSL = 16;     // StreamCfg::MAXSL
  std::vector<int> kbase;   // first step (position inside the row) covered by each piece
  std::vector<int> piece_of_slice;
  // pieces of the slices in their current order (short slices are merged, never across a cut)
  auto make_pieces = [&]() {
    P.pieces.clear();
    kbase.clear();
    piece_of_slice.assign((size_t)nsl, 0);
    for (int i = 0; i < nsl;) {
      piece_of_slice[i] = (int)P.pieces.size();
      if (klen[i] > KB || wr[i]) {
        const uint32_t w = wr[i] ? PIECE_WARP_ROW : 0u;
        for (int k0 = 0; k0 < klen[i]; k0 += KB) {
          Piece pc{0, i * 32, (uint16_t)std::min(KB, klen[i] - k0), 1,
                   (k0 == 0 ? PIECE_FIRST : 0u) | (k0 + KB >= klen[i] ? PIECE_LAST : 0u) | w};
          P.pieces.push_back(pc);
          kbase.push_back(k0);
        }
        ++i;
      } else {
        int j = i, kmax = klen[i];
        while (j + 1 < nsl && !wr[j + 1] && (cut.empty() || !cut[j + 1]) && (j + 1 - i) < MAXSL && klen[j + 1] <= KB &&
               std::max(kmax, klen[j + 1]) * (j + 2 - i) <= KB) {
          kmax = std::max(kmax, klen[j + 1]);
          ++j;
          piece_of_slice[j] = (int)P.pieces.size();
        }
        Piece pc{0, i * 32, (uint16_t)kmax, (uint16_t)(j - i + 1), PIECE_FIRST | PIECE_LAST};
        P.pieces.push_back(pc);
        kbase.push_back(0);
        i = j + 1;
      }
    }
  };
  // cost of a range of pieces: bytes + a fixed per-piece cost + the row epilogues
  std::vector<double> pre;
  auto prefix_costs = [&]() {
    const int np = (int)P.pieces.size();
    pre.assign((size_t)np + 1, 0.0);
    for (int p = 0; p < np; ++p)
      pre[p + 1] = pre[p] + (double)P.pieces[p].klen * P.pieces[p].nslices * 32 * (4 + value_bytes) + 256.0 +
                   ((P.pieces[p].flags & PIECE_LAST) ? 32.0 * 40.0 * P.pieces[p].nslices : 0.0);
  };
  make_pieces();
  prefix_costs();
  int np = (int)P.pieces.size();
  int starts = 0;
  for (int p = 0; p < np; ++p) starts += (P.pieces[p].flags & PIECE_FIRST) ? 1 : 0;
  P.grid = std::max(1, std::min(max_grid, (starts + W - 1) / W));
  if (!P.long_rows.empty()) P.grid = std::max(P.grid, std::min(max_grid, (int)P.long_rows.size()));
  const int nw = P.grid * W;
  // warps: contiguous piece ranges balanced by cost, cut only where rows start
  P.warp_begin.assign(nw + 1, np);
  P.warp_begin[0] = 0;
  for (int w = 1; w < nw; ++w) {
    const double target = pre[np] * w / nw;
    int lo = (int)(std::lower_bound(pre.begin(), pre.end(), target) - pre.begin());
    lo = std::min(std::max(lo, P.warp_begin[w - 1]), np);
    while (lo < np && !(P.pieces[lo].flags & PIECE_FIRST)) ++lo;
    P.warp_begin[w] = lo;
  }
  P.warp_begin[nw] = np;
  // Contiguous ranges can only be cut where rows start, so a warp may exceed its share by one slice -- for power-law
  // rows up to 256 steps x 32 lanes (98 KB) against a share of a few hundred KB (config 4's K: the longest range was
  // 1.41x the mean, and the kernel ran 1.43x longer than K^T's with the same bytes and sectors).  Where the window is
  // wide anyway (no epilogue locality to keep) the slices are dealt out instead: longest first, each to the warp with
  // the least work so far (LPT), the CTAs that start with a one-row-per-CTA giant handicapped by its cost; every warp's
  // slices are then stored contiguously, in descending length, so short ones still merge into pieces.
  {
    double worst = 0.0;
    for (int w = 0; w < nw; ++w) worst = std::max(worst, pre[P.warp_begin[w + 1]] - pre[P.warp_begin[w]]);
    double giants = 0.0;
    for (int r : P.long_rows) giants += (double)len(r) * (4 + value_bytes);
    const double mean = (pre[np] + giants) / nw;
    const bool lpt = sigma >= 1024 && nsl > nw && worst > 1.08 * mean && !getenv("PDLP_NO_LPT");
    if (lpt) {
      std::vector<double> cost((size_t)nsl);
      for (int i = 0; i < nsl; ++i)
        cost[i] = (double)klen[i] * 32 * (4 + value_bytes) + 256.0 * std::max(1.0 / MAXSL, (double)klen[i] / KB) + 32.0 * 40.0;
      std::vector<int> by_cost((size_t)nsl);
      for (int i = 0; i < nsl; ++i) by_cost[i] = i;
      std::stable_sort(by_cost.begin(), by_cost.end(), [&](int a, int b) { return cost[a] > cost[b]; });
      std::vector<double> load((size_t)nw, 0.0);
      for (size_t li = 0; li < P.long_rows.size(); ++li) {           // long_rows_phase: row li goes to CTA li % grid
        const int cta = (int)(li % (size_t)P.grid);
        // straight from CSR, no ring: costed at twice its bytes, shared by the W warps of the CTA
        for (int w = 0; w < W; ++w) load[(size_t)cta * W + w] += 2.0 * len(P.long_rows[li]) * (4 + value_bytes) / W;
      }
      using Item = std::pair<double, int>;
      std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
      for (int w = 0; w < nw; ++w) heap.push(Item{load[w], w});
      std::vector<std::vector<int>> mine((size_t)nw);
      for (int i : by_cost) {
        Item it = heap.top();
        heap.pop();
        mine[(size_t)it.second].push_back(i);
        it.first += cost[i];
        heap.push(it);
      }
      std::vector<int> order2, klen2;
      std::vector<char> wr2;
      order2.reserve(order.size());
      klen2.reserve((size_t)nsl);
      wr2.reserve((size_t)nsl);
      cut.assign((size_t)nsl + 1, 0);
      std::vector<int> first_slice((size_t)nw + 1, 0);
      for (int w = 0; w < nw; ++w) {
        first_slice[w] = (int)klen2.size();
        cut[klen2.size()] = 1;
        for (int i : mine[(size_t)w]) {
          order2.insert(order2.end(), order.begin() + (size_t)i * 32, order.begin() + (size_t)i * 32 + 32);
          klen2.push_back(klen[i]);
          wr2.push_back(wr[i]);
        }
      }
      first_slice[nw] = nsl;
      order.swap(order2);
      klen.swap(klen2);
      wr.swap(wr2);
      make_pieces();
      prefix_costs();
      np = (int)P.pieces.size();
      for (int w = 0; w <= nw; ++w) P.warp_begin[w] = first_slice[w] < nsl ? piece_of_slice[first_slice[w]] : np;
      P.balanced = 1;
    }
  }
  P.perm = order;
  // block offsets; the first step of a split slice's piece travels in the upper bits of flags (pack_kernel)
  size_t bytes = 0;
  for (int p = 0; p < np; ++p) {
    Piece& pc = P.pieces[p];
    pc.data_off16 = (uint32_t)(bytes / 16);
    pc.flags |= (uint32_t)kbase[p] << 8;
    const long long steps = (long long)pc.klen * pc.nslices;
    const size_t pb = 16 + ((pc.flags & PIECE_FIRST) ? (size_t)pc.nslices * 128 : 0) + (size_t)steps * 32 * (4 + value_bytes);
    P.max_piece_bytes = std::max(P.max_piece_bytes, (int)pb);
    bytes += pb;
  }
  P.packed_bytes = bytes;
  if (getenv("PDLP_TRACE_SETUP")) {
    double worst = 0.0;
    for (int w = 0; w < nw; ++w) worst = std::max(worst, pre[P.warp_begin[w + 1]] - pre[P.warp_begin[w]]);
    fprintf(stderr, "[pdlp setup] %d warps, %s ranges: longest %.0f KB, mean %.0f KB\n", nw,
            P.balanced ? "dealt (LPT)" : "contiguous", worst / 1e3, pre[np] / nw / 1e3);
  }
  return P;
}

// fill the piece blocks: header, row ids, and for every slot (slice sl, step k, lane r) the k-th entry of the
// lane's row (column index and value), zero beyond the row's end; one CTA per piece
template <typename T>
__global__ void pack_kernel(const Piece* pieces, const int* perm, int npieces, const int* rowptr, const int* colidx,
                            const T* vals, unsigned char* packed) {
  for (int p = blockIdx.x; p < npieces; p += gridDim.x) {
    const Piece pc = pieces[p];
    unsigned char* blk = packed + (size_t)pc.data_off16 * 16u;
    const int klen = pc.klen;
    const int nslots = klen * pc.nslices * 32;
    const int nrows = (pc.flags & PIECE_FIRST) ? (int)pc.nslices * 32 : 0;
    const int kbase = (int)(pc.flags >> 8);
    if (threadIdx.x == 0) {
      Piece hdr = pc;
      hdr.flags &= 0xffu;
      *reinterpret_cast<Piece*>(blk) = hdr;
    }
    int* rdst = reinterpret_cast<int*>(blk + 16);
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) rdst[i] = perm[pc.perm_off + i];
    int* cdst = rdst + nrows;
    T* vdst = reinterpret_cast<T*>(cdst + nslots);
    for (int i = threadIdx.x; i < nslots; i += blockDim.x) {
      const int r = i & 31;
      const int k = (i >> 5) % klen;
      const int sl = (i >> 5) / klen;
      const int row = perm[pc.perm_off + sl * 32 + r];
      int src = -1;
      if (row >= 0) {
        const int b = rowptr[row], e = rowptr[row + 1];
        if (pc.flags & PIECE_WARP_ROW) {          // lane r owns entries [r * seg, (r + 1) * seg) of the row
          const int seg = (e - b + 31) / 32;
          if (kbase + k < seg && b + r * seg + kbase + k < e) src = b + r * seg + kbase + k;
        } else if (b + kbase + k < e) {
          src = b + kbase + k;
        }
      }
      cdst[i] = src >= 0 ? colidx[src] : 0;
      vdst[i] = src >= 0 ? vals[src] : T(0);
    }
  }
}

// ---------------------------------------------------------------------------
// plan of the run layout (run_spmv.cuh): stream positions, chunk descriptors, per-warp chunk ranges
// ---------------------------------------------------------------------------
// A chunk covers stream positions [gp0, gp0 + CHUNK) of its (panel's) stream, clipped to gp_limit: the tail behind
// gp_limit is zero padding.  Chunks end where a warp's row range ends and where a long row (one row per CTA, not in
// the stream) interrupts the sequence of rows, so that the rows finishing inside a chunk are always consecutive.
struct ChunkDesc {
  int gp0, gp_limit, panel, pad;
};

struct HostRun {
  std::vector<int> sptr;               // rows + 1: stream position of each row's first entry
  std::vector<int> warp_chunk_begin;   // grid * W + 1
  std::vector<ChunkDesc> chunks;
  std::vector<int> long_rows;
  int grid = 1;
  bool ok = true;
};

// Warps [w0, w0 + nwarps) share rows [0, rows) of one stream (sptr): contiguous row ranges balanced by entries, cut
// into chunks.  `breaks`: sorted rows that interrupt the stream (long rows), or nullptr.
static void run_cut_warps(const int* sptr, int rows, int chunk, int w0, int nwarps, int nused, int panel,
                          const std::vector<int>* breaks, std::vector<int>& wcb, std::vector<ChunkDesc>& chunks) {
  const long long total = sptr[rows];
  int r_prev = 0;
  size_t bi = 0;
  for (int w = 0; w < nwarps; ++w) {
    int r = rows;
    if (w + 1 < nused) {
      const long long target = total * (w + 1) / nused;
      r = (int)(std::lower_bound(sptr, sptr + rows + 1, (int)target) - sptr);
      r = std::min(std::max(r, r_prev), rows);
    } else if (w + 1 < nwarps && w + 1 > nused) {
      r = rows;
    }
    wcb[w0 + w] = (int)chunks.size();
    // segments of [r_prev, r) between the long rows
    int seg_begin = r_prev;
    while (seg_begin < r) {
      while (breaks && bi < breaks->size() && (*breaks)[bi] < seg_begin) ++bi;
      int seg_end = r;
      if (breaks && bi < breaks->size() && (*breaks)[bi] < r) seg_end = (*breaks)[bi];
      for (int gp = sptr[seg_begin]; gp < sptr[seg_end]; gp += chunk) chunks.push_back(ChunkDesc{gp, sptr[seg_end], panel, 0});
      seg_begin = seg_end == r ? r : seg_end + 1;
    }
    r_prev = r;
  }
  wcb[w0 + nwarps] = (int)chunks.size();
}

static HostRun build_run_plan(const int* rowptr, int rows, int long_row, int max_grid, int W, int chunk) {
  HostRun P;
  P.sptr.resize((size_t)rows + 1);
  long long pos = 0;
  for (int r = 0; r < rows; ++r) {
    P.sptr[r] = (int)pos;
    const int len = rowptr[r + 1] - rowptr[r];
    if (len > long_row) P.long_rows.push_back(r);      // not in the stream
    else pos += std::max(1, len);                       // an empty row is one zero entry
  }
  if (pos >= 0x7fffffffLL - chunk) { P.ok = false; return P; }
  P.sptr[rows] = (int)pos;
  const int nused = (int)std::max<long long>(1, std::min<long long>((long long)max_grid * W, (pos + chunk - 1) / chunk));
  P.grid = std::max(1, (nused + W - 1) / W);
  if (!P.long_rows.empty()) P.grid = std::max(P.grid, std::min(max_grid, (int)P.long_rows.size()));
  const int nw = P.grid * W;
  P.warp_chunk_begin.assign((size_t)nw + 1, 0);
  run_cut_warps(P.sptr.data(), rows, chunk, 0, nw, nused, 0, &P.long_rows, P.warp_chunk_begin, P.chunks);
  return P;
}

// the row that holds stream position gp: sptr[row] <= gp < sptr[row + 1] (rows outside the stream have no positions)
__host__ __device__ inline int run_row_of(const int* sptr, int rows, int gp) {
  int lo = 0, hi = rows;             // invariant: sptr[lo] <= gp < sptr[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (sptr[mid] <= gp) lo = mid; else hi = mid;
  }
  return lo;
}

// One slot of a chunk of the plain layout.  Returns whether the slot is the last entry of its row; *row = its row
// (-1 for padding).  Host and device: the device packers call it per thread, pdlp_run_emulate on the CPU.
template <typename T, typename CT>
__host__ __device__ inline bool run_pack_slot(const ChunkDesc& d, int s, int rows, const int* sptr, const int* rowptr,
                                              const int* colidx, const T* vals, unsigned char* blk, int* row) {
  using Cfg = RunCfg<T, CT>;
  T* vdst = reinterpret_cast<T*>(blk);
  CT* cdst = reinterpret_cast<CT*>(blk + Cfg::VALS_BYTES);
  const int gp = d.gp0 + s;
  const int at = (s % Cfg::E) * 32 + s / Cfg::E;          // lane s / E, step s % E
  *row = -1;
  if (gp >= d.gp_limit) {                                // zero padding
    vdst[at] = T(0);
    cdst[at] = (CT)0;
    return false;
  }
  const int r = run_row_of(sptr, rows, gp), off = gp - sptr[r];
  const int len = rowptr[r + 1] - rowptr[r];
  *row = r;
  if (len == 0) {                                        // empty row: one zero entry
    vdst[at] = T(0);
    cdst[at] = Cfg::END;
    return true;
  }
  const int src = rowptr[r] + off;
  const bool end = off == len - 1;
  vdst[at] = vals[src];
  cdst[at] = (CT)((CT)colidx[src] | (end ? Cfg::END : (CT)0));
  return end;
}

// per-lane words and the trailer of a chunk from the END flags of its CHUNK slots (stream order); host version
template <int E>
static void run_pack_meta_host(const unsigned char* ends, int row0, unsigned* meta, RunTrailer* tr) {
  int below = 0;
  unsigned any_steps = 0, F = 0;
  for (int lane = 0; lane < 32; ++lane) {
    unsigned endm = 0;
    for (int k = 0; k < E; ++k) endm |= (ends[lane * E + k] ? 1u : 0u) << k;
    meta[lane] = endm | ((unsigned)below << 8);
    below += __builtin_popcount(endm);
    any_steps |= endm;
    if (endm) F |= 1u << lane;
  }
  *tr = RunTrailer{row0, below, any_steps, F};
}

// device: threads 0..31 of the block (one warp), after the END flags of all slots are in shared memory
template <int E>
__device__ inline void run_pack_meta_dev(const unsigned char* ends, int row0, unsigned* meta, RunTrailer* tr) {
  const int lane = threadIdx.x;
  unsigned endm = 0;
#pragma unroll
  for (int k = 0; k < E; ++k) endm |= (ends[lane * E + k] ? 1u : 0u) << k;
  const int cnt = __popc(endm);
  int incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  unsigned any_steps = endm;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) any_steps |= __shfl_xor_sync(0xffffffffu, any_steps, o);
  const unsigned F = __ballot_sync(0xffffffffu, endm != 0);
  const int nrows = __shfl_sync(0xffffffffu, incl, 31);
  meta[lane] = endm | ((unsigned)(incl - cnt) << 8);
  if (lane == 0) *tr = RunTrailer{row0, nrows, any_steps, F};
}

// one CTA (CHUNK threads) per chunk
template <typename T>
__global__ void run_pack_kernel(const ChunkDesc* desc, int nchunks, int rows, const int* sptr, const int* rowptr,
                                const int* colidx, const T* vals, unsigned char* chunks) {
  using CT = uint32_t;
  using Cfg = RunCfg<T, CT>;
  __shared__ unsigned char s_end[Cfg::CHUNK];
  __shared__ int s_row0;
  for (int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const ChunkDesc d = desc[c];
    unsigned char* blk = chunks + (size_t)c * Cfg::CHUNK_BYTES;
    for (int s = threadIdx.x; s < Cfg::CHUNK; s += blockDim.x) {
      int row;
      const bool end = run_pack_slot<T, CT>(d, s, rows, sptr, rowptr, colidx, vals, blk, &row);
      s_end[s] = end ? 1 : 0;
      if (s == 0) s_row0 = row;
    }
    __syncthreads();
    if (threadIdx.x < 32)
      run_pack_meta_dev<Cfg::E>(s_end, s_row0, reinterpret_cast<unsigned*>(blk + Cfg::VALS_BYTES + Cfg::COLS_BYTES),
                                reinterpret_cast<RunTrailer*>(blk + Cfg::VALS_BYTES + Cfg::COLS_BYTES + Cfg::META_BYTES));
    __syncthreads();
  }
}

// ---- shared-memory panels of K^T (panel_spmv.cuh) ----
constexpr int PANEL_MAX = 16;
struct HostPanel {
  int P = 0, grid = 0, panel_cap = 0;
  std::vector<int> panel_begin;        // P + 1
  std::vector<int> cta_panel;          // grid
  std::vector<int> rowptr;             // P * rows + 1: the panel matrix (row p * rows + r = entries of row r in panel p)
  HostSell sell;                       // its SELL plan, every panel planned for its own CTAs
};

// cnt[p * rows + r] = entries of row r whose column lies in panel p
__global__ void panel_count_kernel(const int* rowptr, const int* colidx, int rows, const int* panel_begin, int P, int* cnt) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    int c[PANEL_MAX];
#pragma unroll
    for (int p = 0; p < PANEL_MAX; ++p) c[p] = 0;
    for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const int col = colidx[e];
#pragma unroll
      for (int p = 0; p < PANEL_MAX; ++p)
        if (p < P && col >= panel_begin[p] && col < panel_begin[p + 1]) ++c[p];
    }
#pragma unroll
    for (int p = 0; p < PANEL_MAX; ++p)
      if (p < P) cnt[(size_t)p * rows + r] = c[p];
  }
}

// CSR arrays of the panel matrix from the matrix's own: entries keep their order inside a row, columns become
// panel-relative.  colidx_out may be null (values only: after a scaling pass).
template <typename T>
__global__ void panel_fill_kernel(const int* rowptr, const int* colidx, const T* vals, int rows, const int* panel_begin,
                                  int P, const int* prowptr, int* colidx_out, T* vals_out) {
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x) {
    int cur[PANEL_MAX];
#pragma unroll
    for (int p = 0; p < PANEL_MAX; ++p) cur[p] = p < P ? prowptr[(size_t)p * rows + r] : 0;
    for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) {
      const int col = colidx[e];
      const T v = vals[e];
#pragma unroll
      for (int p = 0; p < PANEL_MAX; ++p)
        if (p < P && col >= panel_begin[p] && col < panel_begin[p + 1]) {
          if (colidx_out) colidx_out[cur[p]] = col - panel_begin[p];
          vals_out[cur[p]] = v;
          ++cur[p];
        }
    }
  }
}

// panel boundaries: P panels over `cols` gathered elements, CTA counts as even as possible, boundaries (multiples of
// 4) placed so that a panel's share of the nonzeros matches its share of the CTAs.  col_nnz_prefix: cols + 1 entries.
static bool panel_boundaries(const int* col_nnz_prefix, int cols, int value_bytes, int sms, HostPanel& H) {
  const int cap = std::min(32760, (int)(PDLP_PANEL_BYTES / value_bytes)) / 4 * 4;
  for (int P = std::max(1, (cols + cap - 1) / cap); P <= PANEL_MAX && P <= sms; ++P) {
    std::vector<int> ctas(P, sms / P);
    for (int p = 0; p < sms % P; ++p) ++ctas[p];
    std::vector<int> pb(P + 1, 0);
    pb[P] = cols;
    const long long nnz = col_nnz_prefix[cols];
    long long cum = 0;
    bool fits = true;
    for (int p = 1; p < P; ++p) {
      cum += ctas[p - 1];
      const long long target = nnz * cum / sms;
      int b = (int)(std::lower_bound(col_nnz_prefix, col_nnz_prefix + cols + 1, (int)target) - col_nnz_prefix);
      b = std::min(cols, (b + 2) / 4 * 4);
      pb[p] = std::max(b, pb[p - 1]);
    }
    for (int p = 0; p < P; ++p) fits = fits && (pb[p + 1] - pb[p] <= cap);
    if (!fits) {                            // skewed columns: equal-width panels instead
      const int width = ((cols + P - 1) / P + 3) / 4 * 4;
      fits = width <= cap;
      for (int p = 1; p < P; ++p) pb[p] = std::min(cols, p * width);
    }
    if (!fits) continue;
    H.P = P;
    H.grid = sms;
    H.panel_begin = pb;
    H.cta_panel.clear();
    for (int p = 0; p < P; ++p) H.cta_panel.insert(H.cta_panel.end(), ctas[p], p);
    int widest = 0;
    for (int p = 0; p < P; ++p) widest = std::max(widest, pb[p + 1] - pb[p]);
    H.panel_cap = (widest + 3) / 4 * 4 + 4;
    return true;
  }
  return false;
}

// The panel matrix and its SELL plan from the per-(panel, row) entry counts.  Every panel is planned on its own (its
// rows sorted and cut into slices, its pieces split over the warps of ITS CTAs) and the plans are concatenated, so
// that CTA b only ever touches pieces of panel cta_panel[b].  The row epilogue of a panel row is one scattered
// 8-byte store, hence the small epilogue weight: the cost model goes for wide sorting windows (little padding).
static bool panel_plan(const int* cnt, int rows, int value_bytes, int W, int KB, HostPanel& H) {
  const int P = H.P;
  H.rowptr.assign((size_t)P * rows + 1, 0);
  long long pos = 0;
  for (size_t i = 0; i < (size_t)P * rows; ++i) {
    H.rowptr[i] = (int)pos;
    pos += cnt[i];
  }
  if (pos >= 0x7fffffffLL) return false;
  H.rowptr[(size_t)P * rows] = (int)pos;
  HostSell& S = H.sell;
  S = HostSell();
  S.warp_begin.assign((size_t)H.grid * W + 1, 0);
  int cta = 0;
  for (int p = 0; p < P; ++p) {
    int n = 0;
    while (cta + n < H.grid && H.cta_panel[cta + n] == p) ++n;
    const char* sg = getenv("PDLP_PANEL_SIGMA");      // developer knob: smallest sorting window considered
    HostSell Q = build_sell_plan(H.rowptr.data() + (size_t)p * rows, rows, value_bytes, W, KB, sg ? atoi(sg) : PDLP_SIGMA,
                                 PDLP_LONG_ROW, PDLP_WARP_ROW, n, 1.0);
    if (!Q.long_rows.empty() || Q.packed_bytes % 16 != 0) return false;      // (a row part of > LONG_ROW entries: not here)
    const int piece0 = (int)S.pieces.size(), perm0 = (int)S.perm.size();
    const size_t byte0 = S.packed_bytes;
    if ((byte0 + Q.packed_bytes) / 16 >= 0xffffffffull) return false;
    for (Piece pc : Q.pieces) {
      pc.data_off16 += (uint32_t)(byte0 / 16);
      pc.perm_off += perm0;
      S.pieces.push_back(pc);
    }
    for (int r : Q.perm) S.perm.push_back(r < 0 ? r : r + p * rows);          // global row id of the panel matrix
    for (int w = 0; w < n * W; ++w)
      S.warp_begin[(size_t)cta * W + w] = piece0 + (w <= Q.grid * W ? Q.warp_begin[std::min(w, Q.grid * W)] : (int)Q.pieces.size());
    S.packed_bytes += Q.packed_bytes;
    S.max_piece_bytes = std::max(S.max_piece_bytes, Q.max_piece_bytes);
    S.sigma = std::max(S.sigma, Q.sigma);
    cta += n;
  }
  S.warp_begin[(size_t)H.grid * W] = (int)S.pieces.size();
  S.grid = H.grid;
  return true;
}

// ---------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------
template <typename T>
struct DeviceCsr {
  // CSR master copy (setup-time kernels, long rows)
  int* rowptr = nullptr;
  int* colidx = nullptr;
  T* vals = nullptr;
  // SELL-32-sigma packed copy (hot kernels)
  unsigned char* packed = nullptr;
  Piece* pieces = nullptr;
  int* perm = nullptr;
  int* warp_begin = nullptr;
  int* long_rows = nullptr;
  double* partials = nullptr;
  unsigned int* ticket = nullptr;
  int rows = 0, cols = 0, nnz = 0, grid = 0, npieces = 0, nlong = 0;
  int stage_bytes = 0, nstages = 2, smem_bytes = 0;
  size_t packed_bytes = 0;
  // run layout (run_spmv.cuh): row-major chunk stream, natural row order
  std::vector<int> h_rowptr;             // host copy of rowptr (plans; column counts of the transpose)
  unsigned char* run_chunks = nullptr;
  int *run_sptr = nullptr, *run_wcb = nullptr;
  ChunkDesc* run_desc = nullptr;
  int run_grid = 0, run_nchunks = 0, run_nstages = 2;
  bool run_on = false;
  RunView<T> run_view() const {
    return RunView<T>{run_chunks, run_wcb, long_rows, nlong, rowptr, colidx, vals, partials, ticket, run_nstages, run_grid};
  }
  SellView<T> view() const {
    return SellView<T>{packed, pieces, warp_begin, long_rows, nlong, rowptr, colidx, vals, partials, ticket,
                       stage_bytes, nstages, grid};
  }
};

struct pdlp_engine {
  int dtype = PDLP_F64;
  int device = 0;
  cudaStream_t stream = nullptr;
  int64_t m = 0, n = 0, m_ineq = 0, nnz = 0;
  int sms = 0, ctas_per_sm = 1;
  bool scaled = false;
  bool have_prev = false;
  bool infeas = false;    // host mirror of Scalars::infeas_detect
  void* impl = nullptr;   // EngineT<T>*
  // peer-memory exchange of the row-partitioned engine (pdlp_shard_exchange_*)
  unsigned char* xchg_block = nullptr;          // this rank's part | tot | flags block (cudaMalloc, IPC-exported)
  void* xchg_peer[MAX_PEERS] = {};              // peers' blocks mapped here (nullptr for this rank)
  Xchg* xchg_dev = nullptr;
  int xchg_world = 0, xchg_rank = 0, xchg_len = 0;
  bool xchg_connected = false;
};

static void xchg_release(pdlp_engine* B);

// pinned read-back buffers are recycled across engines: cudaMallocHost / cudaFreeHost are synchronous and cost
// milliseconds with a large variance, which would sit on the end-to-end path of every solve
static std::mutex g_pinned_mutex;
static std::vector<void*> g_pinned_free;
static void* pinned_scalars_get() {
  {
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    if (!g_pinned_free.empty()) {
      void* p = g_pinned_free.back();
      g_pinned_free.pop_back();
      return p;
    }
  }
  void* p = nullptr;
  if (cudaMallocHost(&p, sizeof(Scalars)) != cudaSuccess) return nullptr;
  return p;
}
static void pinned_scalars_put(void* p) {
  std::lock_guard<std::mutex> lock(g_pinned_mutex);
  g_pinned_free.push_back(p);
}
// pinned staging buffers of arbitrary size (host ingest: the column counts come back while the transpose still runs);
// recycled across engines like the read-back buffers above -- cudaMallocHost costs a fraction of a millisecond
static std::vector<std::pair<void*, size_t>> g_pinned_bufs;
static void* pinned_buf_get(size_t bytes, size_t* cap) {
  {
    std::lock_guard<std::mutex> lock(g_pinned_mutex);
    for (size_t i = 0; i < g_pinned_bufs.size(); ++i)
      if (g_pinned_bufs[i].second >= bytes) {
        void* p = g_pinned_bufs[i].first;
        *cap = g_pinned_bufs[i].second;
        g_pinned_bufs.erase(g_pinned_bufs.begin() + (long)i);
        return p;
      }
  }
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  *cap = bytes;
  return p;
}
static void pinned_buf_put(void* p, size_t cap) {
  std::lock_guard<std::mutex> lock(g_pinned_mutex);
  if (g_pinned_bufs.size() < 4) g_pinned_bufs.emplace_back(p, cap);
  else cudaFreeHost(p);
}

template <typename T>
struct EngineT {
  pdlp_engine* base;
  DeviceCsr<T> K, Kt;
  T *c = nullptr, *q = nullptr, *l = nullptr, *u = nullptr;
  Quad<T>* clu = nullptr;                 // (c, l, u, 0) records for the primal epilogue
  Pair<T>* xp = nullptr;
  T *y = nullptr, *xsum = nullptr, *ysum = nullptr, *xlast = nullptr, *ylast = nullptr;
  Quad<T>*X4 = nullptr, *Y4 = nullptr;
  T *drow = nullptr, *dcol = nullptr;     // accumulated scalings (ones until pdlp_ruiz)
  T *tmp_n = nullptr, *tmp_m = nullptr;
  T *kty_old = nullptr, *lam_prev = nullptr;   // infeasibility detection state
  Scalars* S = nullptr;
  double* ew_partials = nullptr;
  double* kkt_partials = nullptr;   // sharded KKT column sums: ew_grid x 16
  double *blk_partP = nullptr, *blk_partD = nullptr;   // persistent block kernel: per-CTA partial norms (double-buffered)
  unsigned int* blk_bar = nullptr;
  unsigned int* ew_ticket = nullptr;
  int ew_grid = 0;
  // cached graphs for pdlp_iterate
  struct GraphEntry { int n_iters; int save_prev; int infeas; cudaGraphExec_t exec; };
  std::vector<GraphEntry> graphs;
  cudaGraphExec_t check_graph[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [force][have_prev]
  bool check_warm[2][2] = {{false, false}, {false, false}};
  Scalars* h_scalars = nullptr;          // pinned read-back buffer
  // pipelined solve: the steady-state restart test (force = 0, have_prev = 1) without its read-back, the records of two
  // tests in flight, and the events that mark their arrival
  cudaGraphExec_t check_graph_async = nullptr;
  Scalars* h_rec[2] = {nullptr, nullptr};
  cudaEvent_t rec_ev[2] = {nullptr, nullptr};
  std::vector<void*> allocs;
  // layouts of the two iteration kernels.  k1 (K^T y): 0 = SELL (stream_spmv.cuh), 1 = run layout (run_spmv.cuh),
  // 2 = shared-memory panels (panel_spmv.cuh; falls back to 0 when y does not fit).  k2 (K x): 0 = SELL, 1 = run.
  int k1 = 0, k2 = 0;
  struct Panel {
    bool on = false;
    int P = 0, panel_cap = 0, smem = 0, cgrid = 0;
    DeviceCsr<T> M;                      // the panel matrix (P * n rows, panel-relative columns): CSR + SELL copy
    int *cta_panel = nullptr, *panel_begin = nullptr;
    T* partial = nullptr;                // P x n partial sums of K^T y
    double* cpartials = nullptr;         // combine kernel: per-CTA reduction slots
    unsigned int* cticket = nullptr;
    PanelInfo info() const { return PanelInfo{cta_panel, panel_begin, panel_cap}; }
  } panel;
  // K in column windows for operands larger than L2 (DualWindowOp, pdlp_ops.cuh): the window matrix (P * m rows,
  // window-relative columns) in CSR, one SELL copy per window (each planned for the whole grid: the windows are
  // separate launches), the partial (K xbar, K dx) pairs carried between them
  struct Windows {
    bool on = false;
    int P = 0;
    std::vector<int> begin;              // P + 1 column offsets
    int* d_begin = nullptr;
    int *rowptr = nullptr, *colidx = nullptr;
    T* vals = nullptr;
    std::vector<DeviceCsr<T>> M;
    Pair<T>* part = nullptr;
  } win;
  // K2 on the dense xbar vector (DualBarOp, pdlp_ops.cuh): 8-byte gathers instead of (x, dx) pairs for operands whose
  // pair vector does not stay L2-resident; K dx comes from the recursion on kx = K x.  `dirty`: kx must be recomputed
  // from x before the next iteration (iterate or scaling set from outside, iterations that did not maintain it)
  struct Bar {
    bool on = false, dirty = true;
    T *xb = nullptr, *kx = nullptr;
  } bar;
};

// stream-ordered allocation from the device's default pool (kept cached: building an engine right after another
// one was destroyed does not go back to the driver)
template <typename T>
static int dev_alloc(EngineT<T>* E, void** p, size_t bytes) {
  CU(cudaMallocAsync(p, bytes ? bytes : 16, E->base->stream));
  E->allocs.push_back(*p);
  return 0;
}
#define ALLOC(ptr, count)                                                                  \
  do {                                                                                     \
    int rc_ = dev_alloc(E, (void**)&(ptr), sizeof(*(ptr)) * (size_t)(count));              \
    if (rc_) return rc_;                                                                   \
  } while (0)

// L2 residency of the gathered operand.  The matrix stream carries an evict-first hint (stream_spmv.cuh), which
// is what helps (config 4: 1142 -> 987 us per iteration, config 2: 65.4 -> 63.7 us).  A per-launch access-policy
// window marking the operand "persisting" on top of it was measured too and made things WORSE (config 4:
// 993 -> 1195 us), so it is compiled in only with -DPDLP_L2_WINDOW.
static size_t g_l2_window_max = 0;
static void init_l2_persistence() {
  static bool done = false;
  if (done) return;
  done = true;
#ifdef PDLP_L2_WINDOW
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return;
  if (prop.persistingL2CacheMaxSize <= 0) return;
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)prop.persistingL2CacheMaxSize) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  g_l2_window_max = std::min((size_t)prop.accessPolicyMaxWindowSize, (size_t)prop.persistingL2CacheMaxSize);
#endif
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remember per kernel (its address -- two
// instantiations may share a function type) AND device
static cudaError_t ensure_smem_attr_impl(const void* func, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& d : done)
    if (d.first == func && d.second == dev) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done.emplace_back(func, dev);
  return e;
}
template <typename F>
static cudaError_t ensure_smem_attr(F* func, int bytes) {
  return ensure_smem_attr_impl(reinterpret_cast<const void*>(func), bytes);
}

template <class Op>
static cudaError_t launch_stream(const DeviceCsr<typename Op::value_type>& A, Op op, cudaStream_t st,
                                 const void* gathered = nullptr, size_t gathered_bytes = 0,
                                 unsigned long long* stamp = nullptr) {
  using T = typename Op::value_type;
  {
    cudaError_t e = ensure_smem_attr(stream_kernel<Op>, StreamCfg<T>::MAX_SMEM_BYTES - 4096);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(A.grid);
  cfg.blockDim = dim3(StreamCfg<T>::NT);
  cfg.dynamicSmemBytes = A.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous kernel's tail
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (gathered && gathered_bytes && g_l2_window_max) {
    attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[1].val.accessPolicyWindow.base_ptr = const_cast<void*>(gathered);
    attr[1].val.accessPolicyWindow.num_bytes = std::min(gathered_bytes, g_l2_window_max);
    attr[1].val.accessPolicyWindow.hitRatio = 1.0f;
    attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.numAttrs = 2;
  }
  SellView<T> v = A.view();
  v.stamp = stamp;
  return cudaLaunchKernelEx(&cfg, stream_kernel<Op>, v, op);
}

static inline int ew_blocks(int64_t len, int sms) {
  int64_t b = (len + 255) / 256;
  return (int)std::max<int64_t>(1, std::min<int64_t>(b, (int64_t)sms * 8));
}

static void pdl_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int grid, int block, size_t smem, cudaStream_t st) {
  cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: prologue overlaps the previous kernel's tail
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

constexpr int RUN_MAX_DYN_SMEM = 227 * 1024 - 4096;

template <class Op>
static cudaError_t prep_run_kernel() {
  return ensure_smem_attr(run_kernel<Op>, RUN_MAX_DYN_SMEM);
}

template <class Op>
static cudaError_t launch_run(const DeviceCsr<typename Op::value_type>& A, Op op, cudaStream_t st) {
  using T = typename Op::value_type;
  cudaError_t e = prep_run_kernel<Op>();
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  pdl_config(cfg, attr, A.run_grid, PDLP_RUN_W * 32, (size_t)RunSmem<T, uint32_t, Op::NV, PDLP_RUN_W>::total(A.run_nstages), st);
  return cudaLaunchKernelEx(&cfg, run_kernel<Op>, A.run_view(), op);
}

// K^T y through the shared-memory panels, then the combine kernel with the op's row epilogue
template <typename T, class Op>
static cudaError_t launch_panel_combine(EngineT<T>* E, Op op, const T* in, cudaStream_t st) {
  const auto& P = E->panel;
  const int rows = (int)E->base->n;
  cudaError_t e = ensure_smem_attr(panel_kernel<T>, RUN_MAX_DYN_SMEM);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  pdl_config(cfg, attr, P.M.grid, StreamCfg<T>::NT, (size_t)P.smem, st);
  e = cudaLaunchKernelEx(&cfg, panel_kernel<T>, P.M.view(), P.info(), in, P.partial, (const int*)&E->S->stop);
  if (e != cudaSuccess) return e;
  pdl_config(cfg, attr, P.cgrid, COMBINE_THREADS, 0, st);
  return cudaLaunchKernelEx(&cfg, combine_kernel<Op>, (const T*)P.partial, P.P, rows, op, P.cpartials, P.cticket);
}

// PDLP_TRACE_SETUP=1: wall-clock of the setup stages on stderr (developer aid)
struct SetupTrace {
  bool on = getenv("PDLP_TRACE_SETUP") != nullptr;
  std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
  void lap(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[pdlp setup] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
    t = now;
  }
};

// fill the run-layout chunks / the panel matrix from the CSR master copy
template <typename T>
static int pack_run(EngineT<T>* E, DeviceCsr<T>& A) {
  pdlp_engine* B = E->base;
  if (!A.run_on || A.run_nchunks == 0) return 0;
  run_pack_kernel<T><<<std::min(A.run_nchunks, B->sms * 32), RunCfg<T, uint32_t>::CHUNK, 0, B->stream>>>(
      A.run_desc, A.run_nchunks, A.rows, A.run_sptr, A.rowptr, A.colidx, A.vals, A.run_chunks);
  CU(cudaGetLastError());
  return 0;
}
template <typename T>
static int pack_panels(EngineT<T>* E, bool with_columns) {
  pdlp_engine* B = E->base;
  auto& P = E->panel;
  const DeviceCsr<T>& A = E->Kt;
  if (!P.on) return 0;
  panel_fill_kernel<T><<<ew_blocks(A.rows, B->sms), 256, 0, B->stream>>>(A.rowptr, A.colidx, A.vals, A.rows, P.panel_begin,
                                                                        P.P, P.M.rowptr, with_columns ? P.M.colidx : nullptr,
                                                                        P.M.vals);
  CU(cudaGetLastError());
  if (P.M.npieces > 0) {
    pack_kernel<T><<<std::min(P.M.npieces, B->sms * 16), 256, 0, B->stream>>>(P.M.pieces, P.M.perm, P.M.npieces, P.M.rowptr,
                                                                             P.M.colidx, P.M.vals, P.M.packed);
    CU(cudaGetLastError());
  }
  return 0;
}

template <typename T>
static int pack_windows(EngineT<T>* E, bool with_columns);

// (re)fill the packed copy from the CSR master copy (after create and after every scaling pass)
template <typename T>
static int repack(EngineT<T>* E, DeviceCsr<T>& A) {
  pdlp_engine* B = E->base;
  if (A.npieces == 0) return 0;
  pack_kernel<T><<<std::min(A.npieces, B->sms * 16), 256, 0, B->stream>>>(A.pieces, A.perm, A.npieces, A.rowptr,
                                                                         A.colidx, A.vals, A.packed);
  CU(cudaGetLastError());
  int rc = pack_run<T>(E, A);
  if (rc) return rc;
  if (&A == &E->Kt) return pack_panels<T>(E, false);
  return pack_windows<T>(E, false);
}

// PDLP_LAYOUT selects the layout of the two iteration kernels when the engine is created:
//   "sell"  (default) both kernels on SELL-32-sigma slices (stream_spmv.cuh)
//   "run"   both on the row-major run layout (run_spmv.cuh)
//   "panel" K^T y through shared-memory panels when y fits (panel_spmv.cuh), K x on SELL; "runpanel": panels + run
// Measured on config 2 (round 2, us per launch K1 / K2): sell 35.8 / 37.0, run 45.2 / 45.1, panel 45.8 / 36.3 -- the
// alternatives are parity-tested opt-ins, DESIGN.md section 4 explains what bounds each.
static void layout_from_env(int* k1, int* k2) {
  const char* v = getenv("PDLP_LAYOUT");
  *k1 = 0; *k2 = 0;
  if (!v || !*v || !strcmp(v, "auto") || !strcmp(v, "sell")) return;
  if (!strcmp(v, "run")) { *k1 = 1; *k2 = 1; }
  else if (!strcmp(v, "panel")) { *k1 = 2; *k2 = 0; }
  else if (!strcmp(v, "runpanel")) { *k1 = 2; *k2 = 1; }
}

// run-layout copy of a matrix (plan on the host from the row lengths, packing on the device)
template <typename T>
static int build_run(EngineT<T>* E, DeviceCsr<T>& A) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  using Cfg = RunCfg<T, uint32_t>;
  if ((unsigned)A.cols >= (unsigned)Cfg::END) return 0;             // column indices need the END bit
  HostRun P = build_run_plan(A.h_rowptr.data(), A.rows, PDLP_LONG_ROW, B->sms, PDLP_RUN_W, Cfg::CHUNK);
  if (!P.ok) return 0;                                              // falls back to SELL
  A.run_grid = P.grid;
  A.run_nchunks = (int)P.chunks.size();
  A.run_nstages = std::max(2, std::min((int)PDLP_SMAX, (int)(PDLP_RUN_BUDGET / ((size_t)PDLP_RUN_W * Cfg::STAGE_BYTES))));
  ALLOC(A.run_chunks, P.chunks.size() * Cfg::CHUNK_BYTES + 16);
  ALLOC(A.run_sptr, P.sptr.size());
  ALLOC(A.run_desc, P.chunks.size() + 1);
  ALLOC(A.run_wcb, P.warp_chunk_begin.size());
  CU(cudaMemcpyAsync(A.run_sptr, P.sptr.data(), sizeof(int) * P.sptr.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(A.run_desc, P.chunks.data(), sizeof(ChunkDesc) * P.chunks.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(A.run_wcb, P.warp_chunk_begin.data(), sizeof(int) * P.warp_chunk_begin.size(), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));      // the host vectors go out of scope
  if (P.grid > A.grid) {              // per-CTA reduction slots are shared with the SELL kernels
    ALLOC(A.partials, (size_t)P.grid * MAX_RED);
  }
  A.run_on = true;
  return pack_run<T>(E, A);
}

// K^T as shared-memory panels (only when the gathered vector is small enough and no row needs the CTA path)
template <typename T>
static int build_panels(EngineT<T>* E) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  DeviceCsr<T>& A = E->Kt;
  using Cfg = StreamCfg<T>;
  auto& D = E->panel;
  const char* force = getenv("PDLP_PANEL_MIN_NNZ");
  const long long min_nnz = force ? atoll(force) : 500000;
  if (A.nlong > 0 || A.nnz < min_nnz || E->K.h_rowptr.empty()) return 0;
  HostPanel H;
  if (!panel_boundaries(E->K.h_rowptr.data(), A.cols, (int)sizeof(T), B->sms, H)) return 0;
  ALLOC(D.panel_begin, H.panel_begin.size());
  CU(cudaMemcpyAsync(D.panel_begin, H.panel_begin.data(), sizeof(int) * H.panel_begin.size(), cudaMemcpyHostToDevice, st));
  int* d_cnt = nullptr;
  ALLOC(d_cnt, (size_t)H.P * A.rows);
  panel_count_kernel<<<ew_blocks(A.rows, B->sms), 256, 0, st>>>(A.rowptr, A.colidx, A.rows, D.panel_begin, H.P, d_cnt);
  CU(cudaGetLastError());
  std::vector<int> cnt((size_t)H.P * A.rows);
  CU(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * cnt.size(), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (!panel_plan(cnt.data(), A.rows, (int)sizeof(T), Cfg::W, Cfg::KB, H)) return 0;
  const HostSell& S = H.sell;
  // shared memory: the panel, then the rings (as many stages as fit)
  const int panel_bytes = (int)(((size_t)H.panel_cap * sizeof(T) + 127) / 128 * 128);
  DeviceCsr<T>& M = D.M;
  M.rows = H.P * A.rows; M.cols = H.panel_cap; M.nnz = A.nnz;
  M.stage_bytes = (S.max_piece_bytes + 127) / 128 * 128;
  M.nstages = std::min((int)Cfg::SMAX, (RUN_MAX_DYN_SMEM - panel_bytes - 128) / (Cfg::W * M.stage_bytes));
  if (M.nstages < 2) return 0;
  M.grid = S.grid; M.npieces = (int)S.pieces.size(); M.nlong = 0; M.packed_bytes = S.packed_bytes;
  D.P = H.P; D.panel_cap = H.panel_cap;
  D.smem = panel_bytes + Cfg::W * M.nstages * M.stage_bytes + 128;
  D.cgrid = std::max(1, std::min(B->sms * 8, (A.rows + COMBINE_THREADS - 1) / COMBINE_THREADS));
  ALLOC(M.rowptr, H.rowptr.size());
  ALLOC(M.colidx, (size_t)A.nnz + 1);
  ALLOC(M.vals, (size_t)A.nnz + 1);
  ALLOC(M.packed, S.packed_bytes + 16);
  ALLOC(M.pieces, S.pieces.size() + 1);
  ALLOC(M.perm, S.perm.size() + 32);
  ALLOC(M.warp_begin, S.warp_begin.size());
  ALLOC(M.long_rows, 1);
  ALLOC(M.partials, (size_t)S.grid * MAX_RED);
  ALLOC(M.ticket, 1);
  ALLOC(D.cta_panel, H.cta_panel.size());
  ALLOC(D.partial, (size_t)H.P * A.rows);
  ALLOC(D.cpartials, (size_t)D.cgrid * MAX_RED);
  ALLOC(D.cticket, 1);
  CU(cudaMemsetAsync(M.ticket, 0, sizeof(unsigned int), st));
  CU(cudaMemsetAsync(D.cticket, 0, sizeof(unsigned int), st));
  CU(cudaMemcpyAsync(M.rowptr, H.rowptr.data(), sizeof(int) * H.rowptr.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(M.pieces, S.pieces.data(), sizeof(Piece) * S.pieces.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(M.perm, S.perm.data(), sizeof(int) * S.perm.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(M.warp_begin, S.warp_begin.data(), sizeof(int) * S.warp_begin.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(D.cta_panel, H.cta_panel.data(), sizeof(int) * H.cta_panel.size(), cudaMemcpyHostToDevice, st));
  D.on = true;
  int rc = pack_panels<T>(E, true);
  if (rc) return rc;
  CU(cudaStreamSynchronize(st));      // the host vectors go out of scope
  if (getenv("PDLP_TRACE_SETUP"))
    fprintf(stderr, "[pdlp setup] K^T panels: %d panels (cap %d elements), sigma %d, %d pieces, %.1f MB packed, %d ring stages of %d B, "
            "%d B shared memory\n", D.P, D.panel_cap, S.sigma, M.npieces, S.packed_bytes / 1e6, M.nstages, M.stage_bytes, D.smem);
  return 0;
}

// K in column windows: only when the gathered pair vector is well beyond L2 (config 5); PDLP_WINDOW_MB overrides the
// window size (MB of pairs per window, 0 = never)
template <typename T>
static int pack_windows(EngineT<T>* E, bool with_columns) {
  pdlp_engine* B = E->base;
  auto& W = E->win;
  const DeviceCsr<T>& A = E->K;
  if (!W.on) return 0;
  panel_fill_kernel<T><<<ew_blocks(A.rows, B->sms), 256, 0, B->stream>>>(A.rowptr, A.colidx, A.vals, A.rows, W.d_begin, W.P,
                                                                        W.rowptr, with_columns ? W.colidx : nullptr, W.vals);
  CU(cudaGetLastError());
  for (auto& M : W.M)
    if (M.npieces > 0) {
      pack_kernel<T><<<std::min(M.npieces, B->sms * 16), 256, 0, B->stream>>>(M.pieces, M.perm, M.npieces, M.rowptr, M.colidx,
                                                                             M.vals, M.packed);
      CU(cudaGetLastError());
    }
  return 0;
}

template <typename T>
static int build_windows(EngineT<T>* E) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  DeviceCsr<T>& A = E->K;
  using Cfg = StreamCfg<T>;
  auto& W = E->win;
  const char* env = getenv("PDLP_WINDOW_MB");
  const double window_mb = env ? atof(env) : 64.0;
  const size_t elem = E->bar.on ? sizeof(T) : sizeof(Pair<T>);       // what K2 gathers per column
  const double table_mb = (double)A.cols * elem / 1e6;
  // a pair vector of about one window already misses L2 in 60 % of its gathers (config 4: 64 MB, measured): from
  // 0.9 windows on the matrix is cut in at least two
  if (window_mb <= 0.0 || table_mb <= 0.9 * window_mb) return 0;
  const int P = std::max(2, (int)std::ceil(table_mb / window_mb));
  if (P > PANEL_MAX) return 0;
  const int width = ((A.cols + P - 1) / P + 3) / 4 * 4;
  W.begin.assign((size_t)P + 1, A.cols);
  for (int p = 0; p < P; ++p) W.begin[p] = std::min(A.cols, p * width);
  W.P = P;
  SetupTrace tr;
  ALLOC(W.d_begin, W.begin.size());
  CU(cudaMemcpyAsync(W.d_begin, W.begin.data(), sizeof(int) * W.begin.size(), cudaMemcpyHostToDevice, st));
  int* d_cnt = nullptr;
  ALLOC(d_cnt, (size_t)P * A.rows);
  panel_count_kernel<<<ew_blocks(A.rows, B->sms), 256, 0, st>>>(A.rowptr, A.colidx, A.rows, W.d_begin, P, d_cnt);
  CU(cudaGetLastError());
  std::vector<int> cnt((size_t)P * A.rows);
  CU(cudaMemcpyAsync(cnt.data(), d_cnt, sizeof(int) * cnt.size(), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  std::vector<int> rowptr((size_t)P * A.rows + 1);
  long long pos = 0;
  for (size_t i = 0; i < cnt.size(); ++i) {
    rowptr[i] = (int)pos;
    pos += cnt[i];
  }
  rowptr[cnt.size()] = (int)pos;
  if (pos != A.nnz) return fail(PDLP_ERR_STATE, "column windows do not cover the matrix");
  ALLOC(W.rowptr, rowptr.size());
  ALLOC(W.colidx, (size_t)A.nnz + 1);
  ALLOC(W.vals, (size_t)A.nnz + 1);
  ALLOC(W.part, A.rows);
  CU(cudaMemcpyAsync(W.rowptr, rowptr.data(), sizeof(int) * rowptr.size(), cudaMemcpyHostToDevice, st));
  tr.lap("windows: counts + row pointers");
  W.M.resize(P);
  for (int p = 0; p < P; ++p) {
    // every window is an ordinary SELL matrix of A.rows rows on a slice of the window matrix's CSR arrays; a row part
    // reads and writes the carried pair (2 sector operations) instead of the full dual epilogue
    HostSell S = build_sell_plan(rowptr.data() + (size_t)p * A.rows, A.rows, (int)sizeof(T), Cfg::W, Cfg::KB, PDLP_SIGMA,
                                 PDLP_LONG_ROW, PDLP_WARP_ROW, B->sms * PDLP_CTAS * PDLP_OVERSUB, 2.0);
    DeviceCsr<T>& M = W.M[p];
    M.rows = A.rows; M.cols = W.begin[p + 1] - W.begin[p]; M.nnz = rowptr[(size_t)(p + 1) * A.rows] - rowptr[(size_t)p * A.rows];
    M.rowptr = W.rowptr + (size_t)p * A.rows;          // entries are offsets into the shared colidx / vals arrays
    M.colidx = W.colidx;
    M.vals = W.vals;
    M.stage_bytes = (S.max_piece_bytes + 127) / 128 * 128;
    M.nstages = std::max(2, std::min(Cfg::SMAX, (int)(PDLP_SMEM_BUDGET / PDLP_CTAS / ((size_t)Cfg::W * M.stage_bytes))));
    M.smem_bytes = Cfg::W * M.nstages * M.stage_bytes + 128;
    if (S.packed_bytes / 16 >= 0xffffffffull) return fail(PDLP_ERR_INVALID, "packed window exceeds 64 GiB");
    M.grid = S.grid; M.npieces = (int)S.pieces.size(); M.nlong = (int)S.long_rows.size(); M.packed_bytes = S.packed_bytes;
    ALLOC(M.packed, S.packed_bytes + 16);
    ALLOC(M.pieces, S.pieces.size() + 1);
    ALLOC(M.perm, S.perm.size() + 32);
    ALLOC(M.warp_begin, S.warp_begin.size());
    ALLOC(M.long_rows, S.long_rows.size() + 1);
    ALLOC(M.partials, (size_t)S.grid * MAX_RED);
    ALLOC(M.ticket, 1);
    CU(cudaMemsetAsync(M.ticket, 0, sizeof(unsigned int), st));
    CU(cudaMemcpyAsync(M.pieces, S.pieces.data(), sizeof(Piece) * S.pieces.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(M.perm, S.perm.data(), sizeof(int) * S.perm.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(M.warp_begin, S.warp_begin.data(), sizeof(int) * S.warp_begin.size(), cudaMemcpyHostToDevice, st));
    if (!S.long_rows.empty())
      CU(cudaMemcpyAsync(M.long_rows, S.long_rows.data(), sizeof(int) * S.long_rows.size(), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));    // S goes out of scope
  }
  W.on = true;
  int rc = pack_windows<T>(E, true);
  if (rc) return rc;
  CU(cudaStreamSynchronize(st));
  tr.lap("windows: plans + pack");
  if (getenv("PDLP_TRACE_SETUP"))
    fprintf(stderr, "[pdlp setup] K in %d column windows of %d columns (%.0f MB of %s each)\n", P, width,
            (double)width * elem / 1e6, E->bar.on ? "xbar" : "pairs");
  return 0;
}

// SELL plan of a matrix from the host copy of its row pointers (host work only) + the ring geometry that follows from it
template <typename T>
static HostSell plan_csr(EngineT<T>* E, DeviceCsr<T>& A) {
  using Cfg = StreamCfg<T>;
  HostSell P = build_sell_plan(A.h_rowptr.data(), A.rows, (int)sizeof(T), Cfg::W, Cfg::KB, PDLP_SIGMA, PDLP_LONG_ROW,
                               PDLP_WARP_ROW, E->base->sms * PDLP_CTAS * PDLP_OVERSUB);
  // ring geometry: a stage holds the largest piece of THIS matrix; as many stages as the budget allows
  A.stage_bytes = (P.max_piece_bytes + 127) / 128 * 128;
  A.nstages = std::max(2, std::min(Cfg::SMAX, (int)(PDLP_SMEM_BUDGET / PDLP_CTAS / ((size_t)Cfg::W * A.stage_bytes))));
  A.smem_bytes = Cfg::W * A.nstages * A.stage_bytes + 128;
  A.grid = P.grid;
  A.npieces = (int)P.pieces.size();
  A.nlong = (int)P.long_rows.size();
  A.packed_bytes = P.packed_bytes;
  return P;
}

// device side of a plan: allocate the packed copy, upload the plan, fill the packed copy from the CSR master copy.
// Asynchronous on the engine's stream; the caller synchronises before P goes out of scope.
template <typename T>
static int upload_plan_and_pack(EngineT<T>* E, DeviceCsr<T>& A, const HostSell& P) {
  cudaStream_t st = E->base->stream;
  if (P.packed_bytes / 16 >= 0xffffffffull) return fail(PDLP_ERR_INVALID, "packed matrix exceeds 64 GiB");
  ALLOC(A.packed, P.packed_bytes + 16);
  ALLOC(A.pieces, P.pieces.size() + 1);
  ALLOC(A.perm, P.perm.size() + 32);
  ALLOC(A.warp_begin, P.warp_begin.size());
  ALLOC(A.long_rows, P.long_rows.size() + 1);
  ALLOC(A.partials, (size_t)P.grid * MAX_RED);
  ALLOC(A.ticket, 1);
  CU(cudaMemcpyAsync(A.pieces, P.pieces.data(), sizeof(Piece) * P.pieces.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(A.perm, P.perm.data(), sizeof(int) * P.perm.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(A.warp_begin, P.warp_begin.data(), sizeof(int) * P.warp_begin.size(), cudaMemcpyHostToDevice, st));
  if (!P.long_rows.empty())
    CU(cudaMemcpyAsync(A.long_rows, P.long_rows.data(), sizeof(int) * P.long_rows.size(), cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(A.ticket, 0, sizeof(unsigned int), st));
  return repack<T>(E, A);
}

template <typename T>
static int upload_csr(EngineT<T>* E, DeviceCsr<T>& A, int rows, int cols, int nnz, const int32_t* rowptr,
                      const int32_t* colidx, const void* vals) {
  cudaStream_t st = E->base->stream;
  SetupTrace tr;
  A.rows = rows;
  A.cols = cols;
  A.nnz = nnz;
  ALLOC(A.rowptr, rows + 1);
  ALLOC(A.colidx, nnz);
  ALLOC(A.vals, nnz);
  CU(cudaMemcpyAsync(A.rowptr, rowptr, sizeof(int) * (rows + 1), cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(A.colidx, colidx, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemcpyAsync(A.vals, vals, sizeof(T) * (size_t)nnz, cudaMemcpyDeviceToDevice, st));
  std::vector<int>& h_rowptr = A.h_rowptr;
  h_rowptr.resize((size_t)rows + 1);
  CU(cudaMemcpyAsync(h_rowptr.data(), A.rowptr, sizeof(int) * (rows + 1), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (h_rowptr[0] != 0 || h_rowptr[rows] != nnz) return fail(PDLP_ERR_INVALID, "rowptr does not span [0, nnz]");
  for (int r = 0; r < rows; ++r)
    if (h_rowptr[r + 1] < h_rowptr[r]) return fail(PDLP_ERR_INVALID, "rowptr is not monotone");
  tr.lap("csr copy + rowptr to host");
  HostSell P = plan_csr<T>(E, A);
  tr.lap("plan (host)");
  int rc = upload_plan_and_pack<T>(E, A, P);
  if (rc) return rc;
  CU(cudaStreamSynchronize(st));
  tr.lap("alloc + plan upload + pack");
  return 0;
}

// ---------------------------------------------------------------------------
// ingest of a HOST CSR matrix (pdlp_create_host): H2D copies, 32-bit indices, canonical-form check, and the CSR of
// K^T built on the device, overlapped with the host-side planning.  The reference does this with ATen calls spread
// over PDLP/util.py:239-269 and K.to(device) (primal_dual_hybrid_gradient.py:7 receives the device tensor).
// ---------------------------------------------------------------------------
// one thread per entry: 32-bit column, owning row (binary search in rowptr), canonical-form check (columns strictly
// increasing inside a row, inside [0, ncols)), identity permutation for the sort, histogram of the columns
template <typename IDX>
__global__ void ingest_entries_kernel(const int* __restrict__ rowptr, int rows, const IDX* __restrict__ col_in, int nnz,
                                      int ncols, int* __restrict__ col_out, int* __restrict__ row_out,
                                      int* __restrict__ idx_out, int* __restrict__ col_count, int* __restrict__ err) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    const long long cl = (long long)col_in[e];
    if (cl < 0 || cl >= ncols) {
      err[0] = 1;
      col_out[e] = 0;
      row_out[e] = 0;
      idx_out[e] = (int)e;
      continue;
    }
    int lo = 0, hi = rows;                 // last row r with rowptr[r] <= e
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (rowptr[mid] <= (int)e) lo = mid; else hi = mid;
    }
    if ((int)e > rowptr[lo] && (long long)col_in[e - 1] >= cl) err[1] = 1;
    col_out[e] = (int)cl;
    row_out[e] = lo;
    idx_out[e] = (int)e;
    atomicAdd(&col_count[(int)cl], 1);
  }
}

// entries of K^T in (column of K, row of K) order: perm = entry indices of K stably sorted by column
template <typename T>
__global__ void transpose_fill_kernel(const int* __restrict__ perm, const int* __restrict__ row_of, const T* __restrict__ vals,
                                      int nnz, int* __restrict__ t_colidx, T* __restrict__ t_vals) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
    const int e = perm[p];
    t_colidx[p] = row_of[e];
    t_vals[p] = vals[e];
  }
}

template <typename T>
static int ingest_host_csr(EngineT<T>* E, const void* h_rowptr, const void* h_colidx, int index_bytes, const void* h_vals) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const int m = (int)B->m, n = (int)B->n, nnz = (int)B->nnz;
  DeviceCsr<T>&K = E->K, &Kt = E->Kt;
  SetupTrace tr;
  K.rows = m; K.cols = n; K.nnz = nnz;
  Kt.rows = n; Kt.cols = m; Kt.nnz = nnz;
  // row pointers: validated and narrowed on the host (the planner needs them there anyway)
  K.h_rowptr.resize((size_t)m + 1);
  for (int r = 0; r <= m; ++r) {
    const long long v = index_bytes == 8 ? (long long)((const int64_t*)h_rowptr)[r] : (long long)((const int32_t*)h_rowptr)[r];
    if (v < 0 || v > nnz || (r > 0 && v < K.h_rowptr[(size_t)r - 1])) return fail(PDLP_ERR_INVALID, "rowptr is not monotone inside [0, nnz]");
    K.h_rowptr[(size_t)r] = (int)v;
  }
  if (K.h_rowptr[0] != 0 || K.h_rowptr[(size_t)m] != nnz) return fail(PDLP_ERR_INVALID, "rowptr does not span [0, nnz]");
  ALLOC(K.rowptr, m + 1); ALLOC(K.colidx, nnz); ALLOC(K.vals, nnz);
  ALLOC(Kt.rowptr, n + 1); ALLOC(Kt.colidx, nnz); ALLOC(Kt.vals, nnz);
  // temporaries (returned to the pool at the end, stream-ordered)
  void* raw_col = nullptr;
  int *row_of = nullptr, *idx = nullptr, *perm = nullptr, *keys_out = nullptr, *cnt = nullptr, *err = nullptr;
  void* cub_tmp = nullptr;
  std::vector<void*> tmps;
  auto tmp_alloc = [&](void** p, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 16, st);
    if (e == cudaSuccess) tmps.push_back(*p);
    return e;
  };
  auto tmp_free = [&]() {
    for (void* p : tmps) cudaFreeAsync(p, st);
    tmps.clear();
  };
#define TMP(ptr, bytes)                                                      \
  do {                                                                       \
    cudaError_t e_ = tmp_alloc((void**)&(ptr), (bytes));                     \
    if (e_ != cudaSuccess) {                                                 \
      tmp_free();                                                            \
      return fail(PDLP_ERR_CUDA, std::string("cudaMallocAsync: ") + cudaGetErrorString(e_)); \
    }                                                                        \
  } while (0)
#define CUT(expr)                                                            \
  do {                                                                       \
    cudaError_t e_ = (expr);                                                 \
    if (e_ != cudaSuccess) {                                                 \
      tmp_free();                                                            \
      return fail(PDLP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    }                                                                        \
  } while (0)
  TMP(raw_col, (size_t)nnz * (size_t)index_bytes);
  TMP(row_of, sizeof(int) * (size_t)nnz);
  TMP(idx, sizeof(int) * (size_t)nnz);
  TMP(perm, sizeof(int) * (size_t)nnz);
  TMP(keys_out, sizeof(int) * (size_t)nnz);
  TMP(cnt, sizeof(int) * ((size_t)n + 1));
  TMP(err, sizeof(int) * 2);
  int end_bit = 1;
  while (end_bit < 31 && (1ll << end_bit) < (long long)n) ++end_bit;
  size_t scan_bytes = 0, sort_bytes = 0;
  CUT(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, cnt, Kt.rowptr, n + 1, st));
  CUT(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const int*)K.colidx, keys_out, (const int*)idx, perm, nnz, 0, end_bit, st));
  TMP(cub_tmp, std::max(scan_bytes, sort_bytes));
  // H2D: pinned sources are DMA'd asynchronously (the host plans K meanwhile), pageable ones are staged by the driver
  CUT(cudaMemcpyAsync(K.rowptr, K.h_rowptr.data(), sizeof(int) * ((size_t)m + 1), cudaMemcpyHostToDevice, st));
  CUT(cudaMemcpyAsync(raw_col, h_colidx, (size_t)nnz * (size_t)index_bytes, cudaMemcpyHostToDevice, st));
  // (the values follow further down: the column counts, which the K^T plan waits for, need the indices only)
  CUT(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)n + 1), st));
  CUT(cudaMemsetAsync(err, 0, sizeof(int) * 2, st));
  if (nnz > 0) {
    const int blocks = (int)std::min<int64_t>(((int64_t)nnz + 255) / 256, (int64_t)B->sms * 32);
    if (index_bytes == 8)
      ingest_entries_kernel<int64_t><<<blocks, 256, 0, st>>>(K.rowptr, m, (const int64_t*)raw_col, nnz, n, K.colidx, row_of, idx, cnt, err);
    else
      ingest_entries_kernel<int32_t><<<blocks, 256, 0, st>>>(K.rowptr, m, (const int32_t*)raw_col, nnz, n, K.colidx, row_of, idx, cnt, err);
    CUT(cudaGetLastError());
  }
  CUT(cub::DeviceScan::ExclusiveSum(cub_tmp, scan_bytes, cnt, Kt.rowptr, n + 1, st));
  // K^T's row pointers are final here, long before the transpose itself: they come back through a pinned buffer and
  // a helper thread plans K^T from them while this thread plans K (both are host work of a millisecond or two)
  size_t pin_cap = 0;
  int* pin_rowptr = (int*)pinned_buf_get(sizeof(int) * ((size_t)n + 1), &pin_cap);
  cudaEvent_t ev_counts = nullptr;
  if (pin_rowptr && cudaEventCreateWithFlags(&ev_counts, cudaEventDisableTiming) == cudaSuccess) {
    if (cudaMemcpyAsync(pin_rowptr, Kt.rowptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaEventRecord(ev_counts, st) != cudaSuccess) {
      cudaGetLastError();
      cudaEventDestroy(ev_counts);
      ev_counts = nullptr;
    }
  }
  CUT(cudaMemcpyAsync(K.vals, h_vals, sizeof(T) * (size_t)nnz, cudaMemcpyHostToDevice, st));
  if (nnz > 0) {
    CUT(cub::DeviceRadixSort::SortPairs(cub_tmp, sort_bytes, (const int*)K.colidx, keys_out, (const int*)idx, perm, nnz, 0, end_bit, st));
    const int blocks = (int)std::min<int64_t>(((int64_t)nnz + 255) / 256, (int64_t)B->sms * 32);
    transpose_fill_kernel<T><<<blocks, 256, 0, st>>>(perm, row_of, K.vals, nnz, Kt.colidx, Kt.vals);
    CUT(cudaGetLastError());
  }
  tr.lap("host csr: copies + kernels issued");
  Kt.h_rowptr.resize((size_t)n + 1);
  HostSell PT;
  bool planned_t = false;
  std::thread helper;
  if (ev_counts && !getenv("PDLP_NO_PLAN_THREAD")) {
    const int device = B->device;
    helper = std::thread([&, device]() {
      if (cudaSetDevice(device) != cudaSuccess || cudaEventSynchronize(ev_counts) != cudaSuccess) return;
      memcpy(Kt.h_rowptr.data(), pin_rowptr, sizeof(int) * ((size_t)n + 1));
      if (Kt.h_rowptr[(size_t)n] != nnz) return;        // reported below, from the main thread
      PT = plan_csr<T>(E, Kt);
      planned_t = true;
    });
  }
  // while the device copies and transposes: plan K on the host
  HostSell PK = plan_csr<T>(E, K);
  tr.lap("plan K (host, overlapped)");
  if (helper.joinable()) helper.join();
  tr.lap("plan K^T (helper thread, overlapped)");
  if (ev_counts) cudaEventDestroy(ev_counts);
  int h_err[2] = {0, 0};
  cudaError_t e_back = cudaSuccess;
  if (!planned_t) e_back = cudaMemcpyAsync(Kt.h_rowptr.data(), Kt.rowptr, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st);
  if (e_back == cudaSuccess) e_back = cudaMemcpyAsync(h_err, err, sizeof(h_err), cudaMemcpyDeviceToHost, st);
  if (e_back == cudaSuccess) e_back = cudaStreamSynchronize(st);
  if (pin_rowptr) pinned_buf_put(pin_rowptr, pin_cap);
  CUT(e_back);
  tmp_free();
#undef TMP
#undef CUT
  if (h_err[0]) return fail(PDLP_ERR_INVALID, "column index outside [0, n)");
  if (h_err[1]) return fail(PDLP_ERR_NOT_CANONICAL, "K is not a canonical CSR matrix (columns must be strictly increasing inside every row)");
  if (Kt.h_rowptr[(size_t)n] != nnz) return fail(PDLP_ERR_STATE, "transpose does not cover the matrix");
  tr.lap("wait: H2D + transpose");
  int rc = upload_plan_and_pack<T>(E, K, PK);          // the device packs K while the host plans K^T (if not done yet)
  if (rc) return rc;
  if (!planned_t) {
    PT = plan_csr<T>(E, Kt);
    tr.lap("plan K^T (host, overlapped)");
  }
  rc = upload_plan_and_pack<T>(E, Kt, PT);
  if (rc) return rc;
  CU(cudaStreamSynchronize(st));                        // PK, PT go out of scope
  tr.lap("plan uploads + packs");
  return 0;
}

template <typename T>
static int engine_create(pdlp_engine* B, const int32_t* K_rowptr, const int32_t* K_colidx, const void* K_vals,
                         const int32_t* Kt_rowptr, const int32_t* Kt_colidx, const void* Kt_vals, const void* c,
                         const void* q, const void* l, const void* u, int host_index_bytes = 0) {
  // host_index_bytes != 0: K_rowptr / K_colidx / K_vals are HOST arrays with 4- or 8-byte indices and K^T is built
  // here (pdlp_create_host); c, q, l, u may then live on either side (cudaMemcpyDefault)
  EngineT<T>* E = new EngineT<T>();
  E->base = B;
  B->impl = E;
  const int m = (int)B->m, n = (int)B->n, nnz = (int)B->nnz;
  cudaStream_t st = B->stream;
  SetupTrace tr;
  // per-row epilogue traffic expressed in nonzero-equivalents (12 or 8 bytes per nonzero)
  int rc;
  if (host_index_bytes) {
    rc = ingest_host_csr<T>(E, K_rowptr, K_colidx, host_index_bytes, K_vals);
    if (rc) return rc;
  } else {
    rc = upload_csr<T>(E, E->K, m, n, nnz, K_rowptr, K_colidx, K_vals);
    if (rc) return rc;
    rc = upload_csr<T>(E, E->Kt, n, m, nnz, Kt_rowptr, Kt_colidx, Kt_vals);
    if (rc) return rc;
  }
  layout_from_env(&E->k1, &E->k2);
  if (E->k2 == 1) {
    rc = build_run<T>(E, E->K);
    if (rc) return rc;
    if (!E->K.run_on) E->k2 = 0;
  }
  if (E->k1 == 1) {
    rc = build_run<T>(E, E->Kt);
    if (rc) return rc;
    if (!E->Kt.run_on) E->k1 = 0;
  }
  const char* kt_env = getenv("PDLP_KT_RUN_MIN_NNZ");       // developer / test knob: smallest nnz for this rule
  if (E->k1 != 1 && nnz >= (kt_env ? atoll(kt_env) : 10000000LL) && (double)nnz <= 8.0 * (double)n) {
    // very short rows of K^T (a row block of a much larger LP): the plain K_g^T y_g product of the row-partitioned
    // engine (pdlp_shard_kty) runs faster on the run layout; the fused PrimalOp keeps SELL
    rc = build_run<T>(E, E->Kt);
    if (rc) return rc;
  }
  if (E->k1 == 1 || E->k2 == 1 || E->Kt.run_on) tr.lap("run layout");
  {
    // dense-xbar K2 (DualBarOp): by default exactly where the pair vector would otherwise be cut into column windows
    // (it does not stay L2-resident); PDLP_XBAR=1 / 0 forces it on / off.  SELL layout of K only.
    const char* xenv = getenv("PDLP_XBAR");
    const char* wenv = getenv("PDLP_WINDOW_MB");
    const double window_mb = wenv ? atof(wenv) : 64.0;
    const bool beyond = window_mb > 0.0 && (double)n * sizeof(Pair<T>) / 1e6 > 0.9 * window_mb;
    E->bar.on = E->k2 == 0 && (xenv ? atoi(xenv) != 0 : beyond);
    if (E->bar.on && getenv("PDLP_TRACE_SETUP"))
      fprintf(stderr, "[pdlp setup] K2 gathers the dense xbar vector (%.1f MB) instead of (x, dx) pairs (%.1f MB)\n",
              (double)n * sizeof(T) / 1e6, (double)n * sizeof(Pair<T>) / 1e6);
  }
  rc = build_windows<T>(E);
  if (rc) return rc;
  if (E->k1 == 2) {
    rc = build_panels<T>(E);
    if (rc) return rc;
    if (!E->panel.on) E->k1 = 0;
    tr.lap("K^T panels");
  }
  ALLOC(E->c, n); ALLOC(E->l, n); ALLOC(E->u, n); ALLOC(E->q, m); ALLOC(E->clu, n);
  ALLOC(E->xp, n); ALLOC(E->xsum, n); ALLOC(E->xlast, n); ALLOC(E->X4, n);
  ALLOC(E->y, m + 4);      // the panel kernel's bulk copies round the last panel up to 16 bytes
  ALLOC(E->ysum, m); ALLOC(E->ylast, m); ALLOC(E->Y4, m);
  ALLOC(E->drow, m); ALLOC(E->dcol, n);
  ALLOC(E->tmp_n, n); ALLOC(E->tmp_m, m);
  ALLOC(E->kty_old, n); ALLOC(E->lam_prev, n);
  if (E->bar.on) {
    ALLOC(E->bar.xb, n); ALLOC(E->bar.kx, m);
    CU(cudaMemsetAsync(E->bar.xb, 0, sizeof(T) * n, st));
    CU(cudaMemsetAsync(E->bar.kx, 0, sizeof(T) * m, st));
  }
  CU(cudaMemsetAsync(E->kty_old, 0, sizeof(T) * n, st));
  CU(cudaMemsetAsync(E->lam_prev, 0, sizeof(T) * n, st));
  ALLOC(E->S, 1);
  // element-wise kernels (restart, sharded primal update): 2 CTAs of 256 threads per SM, up to 8 per SM for vectors long
  // enough to be bandwidth-bound (config 5: n = 40M; measured 0.78 ms for the 3.2 GB of the sharded primal update
  // with 2 per SM)
  E->ew_grid = (int)std::min<int64_t>((int64_t)B->sms * 8, std::max<int64_t>((int64_t)B->sms * 2, std::max<int64_t>(n, m) / 2048));
  E->h_scalars = (Scalars*)pinned_scalars_get();
  if (!E->h_scalars) return fail(PDLP_ERR_CUDA, "cudaMallocHost failed");
  ALLOC(E->ew_partials, 2 * E->ew_grid);
  ALLOC(E->kkt_partials, 16 * E->ew_grid);
  ALLOC(E->blk_partP, 2 * B->sms * 4);
  ALLOC(E->blk_partD, 4 * B->sms * 4);
  ALLOC(E->blk_bar, 4);
  ALLOC(E->ew_ticket, 1);
  CU(cudaMemsetAsync(E->ew_ticket, 0, sizeof(unsigned int), st));
  const cudaMemcpyKind vkind = host_index_bytes ? cudaMemcpyDefault : cudaMemcpyDeviceToDevice;
  CU(cudaMemcpyAsync(E->c, c, sizeof(T) * n, vkind, st));
  CU(cudaMemcpyAsync(E->l, l, sizeof(T) * n, vkind, st));
  CU(cudaMemcpyAsync(E->u, u, sizeof(T) * n, vkind, st));
  CU(cudaMemcpyAsync(E->q, q, sizeof(T) * m, vkind, st));
  pack_clu_kernel<T><<<ew_blocks(n, B->sms), 256, 0, st>>>(E->c, E->l, E->u, E->clu, n);
  CU(cudaMemsetAsync(E->X4, 0, sizeof(Quad<T>) * n, st));
  CU(cudaMemsetAsync(E->Y4, 0, sizeof(Quad<T>) * m, st));
  fill_kernel<T><<<ew_blocks(m, B->sms), 256, 0, st>>>(E->drow, T(1), m);
  fill_kernel<T><<<ew_blocks(n, B->sms), 256, 0, st>>>(E->dcol, T(1), n);
  set_iterate_kernel<T><<<ew_blocks(std::max(m, n), B->sms), 256, 0, st>>>(nullptr, nullptr, E->xp, E->y, E->xsum,
                                                                          E->ysum, E->xlast, E->ylast, n, m);
  CU(cudaGetLastError());
  Scalars h;
  memset(&h, 0, sizeof(h));
  h.eta = 1.0; h.omega = 1.0; h.theta = 1.0; h.tol = 1e-4; h.infeas_tol = 1e-4;
  h.m_ineq = (int)B->m_ineq;
  CU(cudaMemcpyAsync(E->S, &h, sizeof(h), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  tr.lap("engine_create total");
  return 0;
}

template <typename T>
static void engine_destroy(pdlp_engine* B) {
  EngineT<T>* E = (EngineT<T>*)B->impl;
  if (!E) return;
  for (auto& g : E->graphs) cudaGraphExecDestroy(g.exec);
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b)
      if (E->check_graph[a][b]) cudaGraphExecDestroy(E->check_graph[a][b]);
  if (E->check_graph_async) cudaGraphExecDestroy(E->check_graph_async);
  if (E->h_scalars) pinned_scalars_put(E->h_scalars);
  for (int i = 0; i < 2; ++i) {
    if (E->h_rec[i]) pinned_scalars_put(E->h_rec[i]);
    if (E->rec_ev[i]) cudaEventDestroy(E->rec_ev[i]);
  }
  xchg_release(B);
  for (void* p : E->allocs) cudaFreeAsync(p, B->stream);
  cudaStreamSynchronize(B->stream);
  delete E;
  B->impl = nullptr;
}

// ||q||, ||c|| of the CURRENT (possibly scaled) problem -> Scalars (primal_dual_hybrid_gradient.py:19-20)
template <typename T>
static int refresh_norms(EngineT<T>* E) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  sumsq_kernel<T><<<1, 1024, 0, st>>>(E->q, (int)B->m, E->S, 0);
  sumsq_kernel<T><<<1, 1024, 0, st>>>(E->c, (int)B->n, E->S, 1);
  CU(cudaGetLastError());
  double sc[2];
  CU(cudaMemcpyAsync(sc, (char*)E->S + offsetof(Scalars, scratch), sizeof(sc), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  double nrm[2] = {sqrt(sc[0]), sqrt(sc[1])};
  CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, q_norm), &nrm[0], sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, c_norm), &nrm[1], sizeof(double), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------
// hot loop
// ---------------------------------------------------------------------------
// kx = K x for DualBarOp: recomputed from the pair vector after a restart (after_restart = 1: the kernel looks at the
// restart decision on the device and retires at once when there was none) or, after_restart = 0, unconditionally
template <typename T>
static cudaError_t launch_kx_refresh(EngineT<T>* E, int after_restart, cudaStream_t st) {
  return launch_stream(E->K, KxRefreshOp<T>{E->xp, E->bar.kx, E->S, after_restart}, st);
}
// before iterations are issued: bring kx up to date if something outside the iteration kernels moved x or K
template <typename T>
static cudaError_t ensure_kx(EngineT<T>* E, cudaStream_t st) {
  if (!E->bar.on || !E->bar.dirty) return cudaSuccess;
  E->bar.dirty = false;
  return launch_kx_refresh<T>(E, 0, st);
}

// K2 (K xbar + dual update + step rule) in whichever form the engine holds K: the dense xbar vector (DualBarOp, single
// pass or column windows), column windows of pairs (operand beyond L2; not with infeasibility detection), the run
// layout, or SELL
// dx2_deferred: the PrimalOp before this launch left ||dx||^2 as per-CTA partial sums (primal_defers)
template <typename T, bool SP, bool INF>
static cudaError_t launch_dual(EngineT<T>* E, cudaStream_t st, bool dx2_deferred = false) {
  auto take_dx2 = [&](auto& d) {
    if (dx2_deferred) {
      d.dx2_partials = E->Kt.partials;
      d.dx2_blocks = E->Kt.grid;
    }
  };
  if (!INF && E->bar.on) {
    auto& W = E->win;
    DualOp<T, SP, false> d{nullptr, E->q, E->y, E->ysum, E->Y4, E->S};
    take_dx2(d);
    if (!W.on)
      return launch_stream(E->K, DualBarOp<T, true, true, SP>{d, E->bar.xb, nullptr, E->bar.kx}, st, E->bar.xb,
                           sizeof(T) * (size_t)E->base->n);
    cudaError_t e = cudaSuccess;
    T* part = reinterpret_cast<T*>(W.part);
    for (int p = 0; p < W.P && e == cudaSuccess; ++p) {
      const T* xb = E->bar.xb + W.begin[p];
      if (p == 0) e = launch_stream(W.M[p], DualBarOp<T, true, false, SP>{d, xb, part, E->bar.kx}, st);
      else if (p + 1 < W.P) e = launch_stream(W.M[p], DualBarOp<T, false, false, SP>{d, xb, part, E->bar.kx}, st);
      else e = launch_stream(W.M[p], DualBarOp<T, false, true, SP>{d, xb, part, E->bar.kx}, st);
    }
    return e;
  }
  if (!INF && E->win.on) {
    auto& W = E->win;
    cudaError_t e = cudaSuccess;
    for (int p = 0; p < W.P && e == cudaSuccess; ++p) {
      DualOp<T, SP, false> d{E->xp + W.begin[p], E->q, E->y, E->ysum, E->Y4, E->S};
      take_dx2(d);
      if (p == 0) e = launch_stream(W.M[p], DualWindowOp<T, true, false, SP>{d, W.part}, st);
      else if (p + 1 < W.P) e = launch_stream(W.M[p], DualWindowOp<T, false, false, SP>{d, W.part}, st);
      else e = launch_stream(W.M[p], DualWindowOp<T, false, true, SP>{d, W.part}, st);
    }
    return e;
  }
  DualOp<T, SP, INF> d{E->xp, E->q, E->y, E->ysum, E->Y4, E->S};
  take_dx2(d);
  if (E->k2 == 1) return launch_run(E->K, d, st);
  return launch_stream(E->K, d, st, E->xp, sizeof(Pair<T>) * (size_t)E->base->n);
}

// K1's grid reduction can be left to K2's finalize when both run in stream_kernel (SELL layouts) and K1's finalize
// stores nothing else (no infeasibility detection)
template <typename T>
static bool primal_defers(const EngineT<T>* E, bool inf) {
  return !inf && E->k1 == 0 && E->k2 == 0 && !getenv("PDLP_NO_DEFER");
}

template <typename T, bool SP, bool INF>
static cudaError_t enqueue_one(EngineT<T>* E, cudaStream_t st) {
  PrimalOp<T, SP, INF> p{E->y, E->c, E->l, E->u, E->xp, E->xsum, E->X4, E->S, E->kty_old};
  p.clu = E->clu;
  p.xb = E->bar.xb;
  p.defer = primal_defers<T>(E, INF);
  cudaError_t e;
  if (E->k1 == 2) e = launch_panel_combine<T>(E, p, (const T*)E->y, st);
  else if (E->k1 == 1) e = launch_run(E->Kt, p, st);
  else e = launch_stream(E->Kt, p, st, E->y, sizeof(T) * (size_t)E->base->m);
  if (e != cudaSuccess) return e;
  e = launch_dual<T, SP, INF>(E, st, p.defer);
  if (e != cudaSuccess) return e;
  if (INF) {
    InfeasColsOp<T> f{E->y, E->c, E->l, E->u, E->kty_old, E->lam_prev, E->S};
    e = launch_stream(E->Kt, f, st);
  }
  return e;
}

template <typename T>
static cudaError_t enqueue_iterations(EngineT<T>* E, int n_iters, int save_prev, cudaStream_t st) {
  const bool inf = E->base->infeas;
  for (int i = 0; i < n_iters; ++i) {
    const bool sp = save_prev && (i == n_iters - 1);
    cudaError_t e;
    if (inf) e = sp ? enqueue_one<T, true, true>(E, st) : enqueue_one<T, false, true>(E, st);
    else e = sp ? enqueue_one<T, true, false>(E, st) : enqueue_one<T, false, false>(E, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

template <class Op>
static cudaError_t prep_kernel() {
  return ensure_smem_attr(stream_kernel<Op>, StreamCfg<typename Op::value_type>::MAX_SMEM_BYTES - 4096);
}

// the CUDA graph of a block of n_iters iterations (cached per engine; host work only: nothing is launched)
template <typename T>
static int get_block_graph(EngineT<T>* E, int n_iters, int save_prev, cudaGraphExec_t* out) {
  pdlp_engine* B = E->base;
  const int inf = B->infeas ? 1 : 0;
  for (auto& g : E->graphs)
    if (g.n_iters == n_iters && g.save_prev == (save_prev ? 1 : 0) && g.infeas == inf) {
      *out = g.exec;
      return 0;
    }
  // first use: set the kernels' attributes outside capture, then capture the whole block
  cudaStream_t cap;
  CU(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  CU((prep_kernel<PrimalOp<T, true, false>>()));
  CU((prep_kernel<PrimalOp<T, false, false>>()));
  CU((prep_kernel<PrimalOp<T, true, true>>()));
  CU((prep_kernel<PrimalOp<T, false, true>>()));
  CU((prep_kernel<DualOp<T, true, false>>()));
  CU((prep_kernel<DualOp<T, false, false>>()));
  CU((prep_kernel<DualOp<T, true, true>>()));
  CU((prep_kernel<DualOp<T, false, true>>()));
  CU((prep_kernel<InfeasColsOp<T>>()));
  if (E->win.on) {
    CU((prep_kernel<DualWindowOp<T, true, false, true>>()));
    CU((prep_kernel<DualWindowOp<T, true, false, false>>()));
    CU((prep_kernel<DualWindowOp<T, false, false, true>>()));
    CU((prep_kernel<DualWindowOp<T, false, false, false>>()));
    CU((prep_kernel<DualWindowOp<T, false, true, true>>()));
    CU((prep_kernel<DualWindowOp<T, false, true, false>>()));
  }
  if (E->k1 || E->k2) {
    CU((prep_run_kernel<PrimalOp<T, true, false>>()));
    CU((prep_run_kernel<PrimalOp<T, false, false>>()));
    CU((prep_run_kernel<PrimalOp<T, true, true>>()));
    CU((prep_run_kernel<PrimalOp<T, false, true>>()));
    CU((prep_run_kernel<DualOp<T, true, false>>()));
    CU((prep_run_kernel<DualOp<T, false, false>>()));
    CU((prep_run_kernel<DualOp<T, true, true>>()));
    CU((prep_run_kernel<DualOp<T, false, true>>()));
    CU(ensure_smem_attr(panel_kernel<T>, RUN_MAX_DYN_SMEM));
  }
  CU(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
  cudaError_t e = enqueue_iterations<T>(E, n_iters, save_prev, cap);
  cudaError_t e2 = cudaStreamEndCapture(cap, &graph);
  if (e != cudaSuccess || e2 != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    cudaStreamDestroy(cap);
    return fail(PDLP_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
  }
  cudaGraphExec_t exec = nullptr;
  CU(cudaGraphInstantiate(&exec, graph, 0));
  cudaGraphDestroy(graph);
  cudaStreamDestroy(cap);
  E->graphs.push_back({n_iters, save_prev ? 1 : 0, inf, exec});
  *out = exec;
  return 0;
}

template <typename T>
static int engine_iterate(EngineT<T>* E, int n_iters, int save_prev, int use_graph) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  if (n_iters <= 0) return 0;
  if (save_prev) B->have_prev = true;
  if (B->infeas) E->bar.dirty = true;        // the kernels with the certificate sums gather pairs and leave kx alone
  else CU(ensure_kx<T>(E, st));
  if (use_graph == 2 && !B->infeas && E->k1 == 0 && E->k2 == 0 && !E->win.on && !E->bar.on) {
    // one persistent cooperative kernel for the whole block (SELL layout only; otherwise the CUDA graph below)
    const int grid = std::max(E->K.grid, E->Kt.grid);
    const int stages = std::min(E->K.nstages, E->Kt.nstages);
    const int stage_bytes = std::max(E->K.stage_bytes, E->Kt.stage_bytes);
    const int smem = StreamCfg<T>::W * stages * stage_bytes + 128;
    CU(ensure_smem_attr(block_kernel<T>, StreamCfg<T>::MAX_SMEM_BYTES - 4096));
    CU(cudaMemsetAsync(E->blk_bar, 0, sizeof(unsigned int), st));
    SellView<T> vKt = E->Kt.view(), vK = E->K.view();
    BlockArgs<T> a{E->y, E->c, E->l, E->u, E->xp, E->xsum, E->X4, E->q, E->ysum, E->Y4, E->kty_old, E->S,
                   E->blk_partP, E->blk_partD, E->blk_bar, n_iters, save_prev ? 1 : 0, stages, stage_bytes};
    void* args[] = {(void*)&vKt, (void*)&vK, (void*)&a};
    CU(cudaLaunchCooperativeKernel((const void*)block_kernel<T>, dim3(grid), dim3(StreamCfg<T>::NT), args, (size_t)smem, st));
    return 0;
  }
  if (!use_graph || n_iters < 2) {
    CU(enqueue_iterations<T>(E, n_iters, save_prev, st));
    return 0;
  }
  // (mode 2 with infeasibility detection lands here: CUDA graph)
  cudaGraphExec_t exec = nullptr;
  int rc = get_block_graph<T>(E, n_iters, save_prev, &exec);
  if (rc) return rc;
  CU(cudaGraphLaunch(exec, st));
  return 0;
}

template <typename T>
static int engine_time_kernels(EngineT<T>* E, int n_iters, float* primal_ms, float* dual_ms) {
  cudaStream_t st = E->base->stream;
  std::vector<cudaEvent_t> ev(3 * (size_t)n_iters);
  for (auto& e : ev) CU(cudaEventCreate(&e));
  CU(ensure_kx<T>(E, st));
  for (int i = 0; i < n_iters; ++i) {
    PrimalOp<T, false, false> p{E->y, E->c, E->l, E->u, E->xp, E->xsum, E->X4, E->S, E->kty_old};
    p.clu = E->clu;
    p.xb = E->bar.xb;
    p.defer = primal_defers<T>(E, false);
    CU(cudaEventRecord(ev[3 * i], st));
    if (E->k1 == 2) CU(launch_panel_combine<T>(E, p, (const T*)E->y, st));
    else if (E->k1 == 1) CU(launch_run(E->Kt, p, st));
    else CU(launch_stream(E->Kt, p, st));
    CU(cudaEventRecord(ev[3 * i + 1], st));
    CU((launch_dual<T, false, false>(E, st, p.defer)));
    CU(cudaEventRecord(ev[3 * i + 2], st));
  }
  CU(cudaStreamSynchronize(st));
  double tp = 0.0, td = 0.0;
  for (int i = 0; i < n_iters; ++i) {
    float a = 0.f, b = 0.f;
    CU(cudaEventElapsedTime(&a, ev[3 * i], ev[3 * i + 1]));
    CU(cudaEventElapsedTime(&b, ev[3 * i + 1], ev[3 * i + 2]));
    tp += a;
    td += b;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  *primal_ms = (float)(tp / n_iters);
  *dual_ms = (float)(td / n_iters);
  return 0;
}

// The two iteration kernels timed AS THEY RUN IN THE LOOP: n_iters iterations launched back to back with programmatic
// dependent launch (no events between the kernels, so the overlap of a kernel's prologue with its predecessor's tail
// and the warm L2 are what the CUDA graph of a block sees), every launch writing its own {start, end} %globaltimer
// stamps: start when CTA 0 leaves griddepcontrol.wait, end after the last CTA's finalize.  SELL layout, single pass
// only.  (pdlp_time_kernels brackets every launch with events: isolated launches, cold start, no overlap.)
template <typename T>
static int engine_time_kernels_in_loop(EngineT<T>* E, int n_iters, float* primal_ms, float* dual_ms, float* iter_ms) {
  cudaStream_t st = E->base->stream;
  if (E->k1 != 0 || E->k2 != 0 || E->win.on || E->bar.on || E->base->infeas)
    return fail(PDLP_ERR_STATE, "in-loop kernel timing covers the default (SELL, single pass, pair gathers) kernels only");
  if (n_iters < 1) return fail(PDLP_ERR_INVALID, "n_iters must be positive");
  unsigned long long* d = nullptr;
  CU(cudaMallocAsync((void**)&d, sizeof(unsigned long long) * 4 * (size_t)n_iters, st));
  CU(cudaMemsetAsync(d, 0, sizeof(unsigned long long) * 4 * (size_t)n_iters, st));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  cudaError_t err = cudaEventRecord(e0, st);
  for (int i = 0; i < n_iters && err == cudaSuccess; ++i) {
    PrimalOp<T, false, false> p{E->y, E->c, E->l, E->u, E->xp, E->xsum, E->X4, E->S, E->kty_old};
    p.clu = E->clu;
    p.defer = primal_defers<T>(E, false);
    err = launch_stream(E->Kt, p, st, nullptr, 0, d + 4 * (size_t)i);
    if (err != cudaSuccess) break;
    DualOp<T, false, false> q{E->xp, E->q, E->y, E->ysum, E->Y4, E->S};
    if (p.defer) {
      q.dx2_partials = E->Kt.partials;
      q.dx2_blocks = E->Kt.grid;
    }
    err = launch_stream(E->K, q, st, nullptr, 0, d + 4 * (size_t)i + 2);
  }
  if (err == cudaSuccess) err = cudaEventRecord(e1, st);
  std::vector<unsigned long long> h(4 * (size_t)n_iters);
  if (err == cudaSuccess) err = cudaMemcpyAsync(h.data(), d, sizeof(unsigned long long) * h.size(), cudaMemcpyDeviceToHost, st);
  if (err == cudaSuccess) err = cudaStreamSynchronize(st);
  float ms = 0.f;
  if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFreeAsync(d, st);
  if (err != cudaSuccess) return fail(PDLP_ERR_CUDA, std::string("in-loop timing: ") + cudaGetErrorString(err));
  double tp = 0.0, td = 0.0;
  for (int i = 0; i < n_iters; ++i) {
    if (!h[4 * (size_t)i] || !h[4 * (size_t)i + 1] || !h[4 * (size_t)i + 2] || !h[4 * (size_t)i + 3])
      return fail(PDLP_ERR_STATE, "in-loop timing: a launch left no stamp (solver stopped?)");
    tp += (double)(h[4 * (size_t)i + 1] - h[4 * (size_t)i]);
    td += (double)(h[4 * (size_t)i + 3] - h[4 * (size_t)i + 2]);
  }
  *primal_ms = (float)(tp / n_iters * 1e-6);
  *dual_ms = (float)(td / n_iters * 1e-6);
  *iter_ms = ms / (float)n_iters;
  return 0;
}

template <typename T>
static int enqueue_kkt(EngineT<T>* E, cudaStream_t st) {
  pdlp_engine* B = E->base;
  KktColsOp<T> a{E->Y4, E->X4, E->c, E->l, E->u, B->scaled ? E->dcol : nullptr, E->S};
  CU(launch_stream(E->Kt, a, st));
  KktRowsOp<T> b{E->X4, E->Y4, E->q, B->scaled ? E->drow : nullptr, E->S, 0};
  CU(launch_stream(E->K, b, st));
  return 0;
}

static void fill_residuals(pdlp_residuals* o, const Resid& r) {
  o->pr = r.pr; o->dr = r.dr; o->gap = r.gap; o->prim_obj = r.prim_obj; o->adj_dual = r.adj_dual; o->kkt = r.kkt;
  o->pr_unscaled = r.pr_u; o->dr_unscaled = r.dr_u;
}
static void fill_check(pdlp_check* out, const Scalars& h) {
  out->kind = h.chk_kind;
  out->use_average = h.chk_use_avg;
  out->t = h.chk_t;
  out->solved = h.chk_solved;
  out->k = h.chk_k;
  out->omega = h.omega;
  out->kkt_current = h.kkt_cur;
  out->kkt_average = h.kkt_avg;
  out->kkt_previous = h.kkt_prev;
  out->kkt_first_new = h.kkt_first;
  fill_residuals(&out->point, h.point);
}

// the cached restart-check graphs bake in the scaling state (pointers and the `scaled` flag): drop them when it changes
template <typename T>
static void drop_check_graphs(EngineT<T>* E) {
  E->bar.dirty = true;                 // called whenever the scaling (hence K) changes: kx = K x is stale
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      if (E->check_graph[a][b]) cudaGraphExecDestroy(E->check_graph[a][b]);
      E->check_graph[a][b] = nullptr;
      E->check_warm[a][b] = false;
    }
  if (E->check_graph_async) cudaGraphExecDestroy(E->check_graph_async);
  E->check_graph_async = nullptr;
}

template <typename T>
static int enqueue_check(EngineT<T>* E, int force, int have_prev, cudaStream_t st, bool read_back = true) {
  pdlp_engine* B = E->base;
  const int n = (int)B->n, m = (int)B->m;
  flush_kernel<T><<<ew_blocks(std::max(n, m), B->sms), 256, 0, st>>>(E->xp, E->y, E->xsum, E->ysum, E->X4, E->Y4, n, m,
                                                                     E->S, have_prev);
  CU(cudaGetLastError());
  int rc = enqueue_kkt<T>(E, st);
  if (rc) return rc;
  decide_kernel<<<1, 1, 0, st>>>(E->S, force);
  apply_restart_kernel<T><<<E->ew_grid, 256, 0, st>>>(E->xp, E->y, E->xsum, E->ysum, E->X4, E->Y4, E->xlast, E->ylast,
                                                      n, m, E->S, E->ew_partials, E->ew_ticket);
  if (E->bar.on) CU(launch_kx_refresh<T>(E, 1, st));       // kx = K x at the restart point (no-op without a restart)
  post_restart_kernel<<<1, 1, 0, st>>>(E->S, B->scaled ? 1 : 0);
  CU(cudaGetLastError());
  if (read_back) CU(cudaMemcpyAsync(E->h_scalars, E->S, sizeof(Scalars), cudaMemcpyDeviceToHost, st));
  return 0;
}

// the kernels of a restart test captured into a graph (their attributes are set first: not capturable)
template <typename T>
static int capture_check_graph(EngineT<T>* E, int force, int have_prev, bool read_back, cudaGraphExec_t* out) {
  CU((prep_kernel<KktColsOp<T>>()));
  CU((prep_kernel<KktRowsOp<T>>()));
  cudaStream_t cap;
  CU(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  cudaError_t e1 = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
  int rc = e1 == cudaSuccess ? enqueue_check<T>(E, force, have_prev, cap, read_back) : 0;
  cudaError_t e2 = e1 == cudaSuccess ? cudaStreamEndCapture(cap, &graph) : e1;
  if (rc || e2 != cudaSuccess) {
    if (graph) cudaGraphDestroy(graph);
    cudaStreamDestroy(cap);
    return rc ? rc : fail(PDLP_ERR_CUDA, std::string("check capture failed: ") + cudaGetErrorString(e2));
  }
  cudaError_t e3 = cudaGraphInstantiate(out, graph, 0);
  cudaGraphDestroy(graph);
  cudaStreamDestroy(cap);
  if (e3 != cudaSuccess) return fail(PDLP_ERR_CUDA, std::string("check graph: ") + cudaGetErrorString(e3));
  return 0;
}

// the restart test is one cached CUDA graph (7 kernels + the read-back copy) per (force, have_prev) pair
template <typename T>
static int engine_restart_check(EngineT<T>* E, int force, pdlp_check* out) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const int have_prev = (B->have_prev && !force) ? 1 : 0;
  const int f = force ? 1 : 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  CU(cudaStreamIsCapturing(st, &cs));
  if (cs != cudaStreamCaptureStatusNone) return fail(PDLP_ERR_STATE, "pdlp_restart_check synchronises: not capturable");
  if (!E->check_graph[f][have_prev]) {
    // first use: run eagerly once (sets kernel attributes), capture for the following calls
    if (E->check_warm[f][have_prev]) {
      int rc = capture_check_graph<T>(E, force, have_prev, true, &E->check_graph[f][have_prev]);
      if (rc) return rc;
    } else {
      E->check_warm[f][have_prev] = true;
      int rc = enqueue_check<T>(E, force, have_prev, st);
      if (rc) return rc;
    }
  }
  if (E->check_graph[f][have_prev]) CU(cudaGraphLaunch(E->check_graph[f][have_prev], st));
  B->have_prev = false;
  CU(cudaStreamSynchronize(st));
  if (out) fill_check(out, *E->h_scalars);
  return 0;
}

template <typename T>
static int engine_kkt(EngineT<T>* E, const void* x, const void* y, double omega, pdlp_residuals* out) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const int n = (int)B->n, m = (int)B->m;
  pack3_kernel<T><<<ew_blocks(n, B->sms), 256, 0, st>>>((const T*)x, E->X4, n);
  pack3_kernel<T><<<ew_blocks(m, B->sms), 256, 0, st>>>((const T*)y, E->Y4, m);
  CU(cudaGetLastError());
  // the column-side finalize zeroes w_pending (it normally follows a flush): preserve it
  double wp;
  CU(cudaMemcpyAsync(&wp, (char*)E->S + offsetof(Scalars, w_pending), sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  int rc = enqueue_kkt<T>(E, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, w_pending), &wp, sizeof(double), cudaMemcpyHostToDevice, st));
  Scalars h;
  CU(cudaMemcpyAsync(&h, E->S, sizeof(h), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  B->have_prev = false;
  const Resid r = make_resid(h.sums[IT_CUR], omega);
  fill_residuals(out, r);
  return 0;
}

template <typename T>
static int enqueue_power_rounds(EngineT<T>* E, int iters, cudaStream_t st) {
  T* v = E->tmp_n;         // un-normalised iterate; ||v||^2 lives in scratch[2]
  T* w = E->tmp_m;
  // helpers.py:48-50: b = K^T (K b); b /= ||b||.  The division is applied to the ROWS of the next product,
  // K (v / ||v||) = (K v) / ||v||, so a round is two kernels and v is never rewritten (a separate normalisation pass
  // cost ~18 of 88 us per round on config 2; dividing every GATHERED element instead made the round slower, 110 us).
  // A zero iterate (K = 0) divides by zero and yields NaN, like the reference.
  for (int i = 0; i < iters; ++i) {
    SpmvOp<T, false> f{v, w, E->S, 2};
    f.div_slot = i > 0 ? 2 : -1;
    CU(launch_stream(E->K, f, st));
    SpmvOp<T, true> g{w, v, E->S, 2};
    CU(launch_stream(E->Kt, g, st));
  }
  SpmvOp<T, true> f{v, w, E->S, 3};
  f.div_slot = iters > 0 ? 2 : -1;
  CU(launch_stream(E->K, f, st));
  return 0;
}

// The graphs a solve with the default restart period replays -- the block of 40 iterations and the steady-state
// restart test -- built ahead of time.  Capture and instantiation are host work only (about a millisecond); called
// while the GPU runs the power iteration, they leave the critical path of pdlp_algorithm.  Best effort: a failure
// here is not an error, the solve builds what it misses.
#ifndef PDLP_DEFAULT_PERIOD
#define PDLP_DEFAULT_PERIOD 40
#endif
template <typename T>
static void prebuild_graphs(EngineT<T>* E) {
  pdlp_engine* B = E->base;
  if (B->infeas || getenv("PDLP_NO_PREBUILD")) return;
  const std::string keep = g_err;
  cudaGraphExec_t exec = nullptr;
  bool ok = get_block_graph<T>(E, PDLP_DEFAULT_PERIOD, 1, &exec) == 0;
  if (ok && !E->check_graph_async) ok = capture_check_graph<T>(E, 0, 1, false, &E->check_graph_async) == 0;
  if (!ok) {
    cudaGetLastError();
    g_err = keep;
  }
}

template <typename T>
static int engine_power_iteration(EngineT<T>* E, const void* b0, int iters, double* sigma) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const int n = (int)B->n;
  CU(cudaMemcpyAsync(E->tmp_n, b0, sizeof(T) * n, cudaMemcpyDeviceToDevice, st));
  // (replaying the rounds as one CUDA graph was measured in round 2: capture + instantiation of the 301-node graph per
  // engine costs more than the launch gaps it removes -- 10.2 ms against 9.0 ms for the reference's 100 rounds)
  {
    const int rc = enqueue_power_rounds<T>(E, iters, st);
    if (rc) return rc;
  }
  if (iters >= 10) prebuild_graphs<T>(E);       // the GPU is busy for milliseconds: use the host meanwhile
  double ss;
  CU(cudaMemcpyAsync(&ss, (char*)E->S + offsetof(Scalars, scratch) + 3 * sizeof(double), sizeof(double),
                     cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  *sigma = sqrt(ss);
  return 0;
}

template <typename T>
static int read_scratch(EngineT<T>* E, int slot, double* out) {
  cudaStream_t st = E->base->stream;
  CU(cudaMemcpyAsync(out, (char*)E->S + offsetof(Scalars, scratch) + slot * sizeof(double), sizeof(double),
                     cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

template <typename T>
static int apply_vector_scaling(EngineT<T>* E, const T* drow_step, const T* dcol_step) {
  // c *= D, q *= D, l /= D, u /= D for the scaling accumulated in THIS call (enhancements.py:64-67)
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const int n = (int)B->n, m = (int)B->m;
  vec_scale_kernel<T, 0><<<ew_blocks(n, B->sms), 256, 0, st>>>(E->c, dcol_step, n);
  vec_scale_kernel<T, 0><<<ew_blocks(m, B->sms), 256, 0, st>>>(E->q, drow_step, m);
  vec_scale_kernel<T, 1><<<ew_blocks(n, B->sms), 256, 0, st>>>(E->l, dcol_step, n);
  vec_scale_kernel<T, 1><<<ew_blocks(n, B->sms), 256, 0, st>>>(E->u, dcol_step, n);
  pack_clu_kernel<T><<<ew_blocks(n, B->sms), 256, 0, st>>>(E->c, E->l, E->u, E->clu, n);
  CU(cudaGetLastError());
  return 0;
}

template <typename T>
static int engine_ruiz(EngineT<T>* E, int max_iter, double eps, int* rounds_out) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const int n = (int)B->n, m = (int)B->m, nnz = (int)B->nnz;
  const int wb_m = ew_blocks((int64_t)m * 32, B->sms), wb_n = ew_blocks((int64_t)n * 32, B->sms);
  const int eb = ew_blocks(nnz, B->sms);
  // this call's own accumulators (the reference starts from ones, :46-47)
  T* drs = nullptr; T* dcs = nullptr;
  CU(cudaMalloc(&drs, sizeof(T) * std::max(m, 1)));
  CU(cudaMalloc(&dcs, sizeof(T) * std::max(n, 1)));
  fill_kernel<T><<<ew_blocks(m, B->sms), 256, 0, st>>>(drs, T(1), m);
  fill_kernel<T><<<ew_blocks(n, B->sms), 256, 0, st>>>(dcs, T(1), n);
  int rounds = 0;
  for (int it = 0; it < max_iter; ++it) {
    ++rounds;
    row_absreduce_kernel<T, 0><<<wb_m, 256, 0, st>>>(E->K.rowptr, E->K.vals, m, E->tmp_m);
    ruiz_fix_kernel<T><<<1, 1024, 0, st>>>(E->tmp_m, drs, m, (T)eps, E->S, 4);
    scale_rows_kernel<T><<<wb_m, 256, 0, st>>>(E->K.rowptr, E->K.vals, m, E->tmp_m);
    scale_cols_kernel<T><<<eb, 256, 0, st>>>(E->Kt.colidx, E->Kt.vals, nnz, E->tmp_m);
    row_absreduce_kernel<T, 0><<<wb_n, 256, 0, st>>>(E->Kt.rowptr, E->Kt.vals, n, E->tmp_n);
    ruiz_fix_kernel<T><<<1, 1024, 0, st>>>(E->tmp_n, dcs, n, (T)eps, E->S, 5);
    scale_rows_kernel<T><<<wb_n, 256, 0, st>>>(E->Kt.rowptr, E->Kt.vals, n, E->tmp_n);
    scale_cols_kernel<T><<<eb, 256, 0, st>>>(E->K.colidx, E->K.vals, nnz, E->tmp_n);
    CU(cudaGetLastError());
    double dev;
    int rc = read_scratch<T>(E, 4, &dev);   // max |1 - row_norms| (:60-61 tests the row norms twice)
    if (rc) return rc;
    if (dev < eps) break;
  }
  int rc = apply_vector_scaling<T>(E, drs, dcs);
  if (rc) return rc;
  vec_scale_kernel<T, 0><<<ew_blocks(m, B->sms), 256, 0, st>>>(E->drow, drs, m);
  vec_scale_kernel<T, 0><<<ew_blocks(n, B->sms), 256, 0, st>>>(E->dcol, dcs, n);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));
  cudaFree(drs);
  cudaFree(dcs);
  B->scaled = true;
  drop_check_graphs<T>(E);
  if (rounds_out) *rounds_out = rounds;
  rc = repack<T>(E, E->K);
  if (!rc) rc = repack<T>(E, E->Kt);
  if (rc) return rc;
  return refresh_norms<T>(E);
}

template <typename T>
static int engine_pock_chambolle(EngineT<T>* E) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const int n = (int)B->n, m = (int)B->m, nnz = (int)B->nnz;
  const int wb_m = ew_blocks((int64_t)m * 32, B->sms), wb_n = ew_blocks((int64_t)n * 32, B->sms);
  const int eb = ew_blocks(nnz, B->sms);
  T* drs = nullptr; T* dcs = nullptr;
  CU(cudaMalloc(&drs, sizeof(T) * std::max(m, 1)));
  CU(cudaMalloc(&dcs, sizeof(T) * std::max(n, 1)));
  fill_kernel<T><<<ew_blocks(m, B->sms), 256, 0, st>>>(drs, T(1), m);
  fill_kernel<T><<<ew_blocks(n, B->sms), 256, 0, st>>>(dcs, T(1), n);
  // both L1 norms are taken on the input matrix, then applied
  row_absreduce_kernel<T, 1><<<wb_m, 256, 0, st>>>(E->K.rowptr, E->K.vals, m, E->tmp_m);
  row_absreduce_kernel<T, 1><<<wb_n, 256, 0, st>>>(E->Kt.rowptr, E->Kt.vals, n, E->tmp_n);
  ruiz_fix_kernel<T><<<1, 1024, 0, st>>>(E->tmp_m, drs, m, (T)1e-6, E->S, 4);
  ruiz_fix_kernel<T><<<1, 1024, 0, st>>>(E->tmp_n, dcs, n, (T)1e-6, E->S, 5);
  scale_rows_kernel<T><<<wb_m, 256, 0, st>>>(E->K.rowptr, E->K.vals, m, E->tmp_m);
  scale_cols_kernel<T><<<eb, 256, 0, st>>>(E->Kt.colidx, E->Kt.vals, nnz, E->tmp_m);
  scale_rows_kernel<T><<<wb_n, 256, 0, st>>>(E->Kt.rowptr, E->Kt.vals, n, E->tmp_n);
  scale_cols_kernel<T><<<eb, 256, 0, st>>>(E->K.colidx, E->K.vals, nnz, E->tmp_n);
  CU(cudaGetLastError());
  int rc = apply_vector_scaling<T>(E, drs, dcs);
  if (rc) return rc;
  vec_scale_kernel<T, 0><<<ew_blocks(m, B->sms), 256, 0, st>>>(E->drow, drs, m);
  vec_scale_kernel<T, 0><<<ew_blocks(n, B->sms), 256, 0, st>>>(E->dcol, dcs, n);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(st));
  cudaFree(drs);
  cudaFree(dcs);
  B->scaled = true;
  drop_check_graphs<T>(E);
  rc = repack<T>(E, E->K);
  if (!rc) rc = repack<T>(E, E->Kt);
  if (rc) return rc;
  return refresh_norms<T>(E);
}

// ---------------------------------------------------------------------------
// fishnet batches (8 right-hand sides per pass over the matrix)
// ---------------------------------------------------------------------------
template <typename T>
static int batch_spmv(EngineT<T>* E, const void* X8, void* Y8) {
  BatchOp<T, 0> f{(const Oct<T>*)X8, (Oct<T>*)Y8, nullptr, nullptr, nullptr, nullptr, T(0), 0};
  CU(launch_stream(E->K, f, E->base->stream));
  return 0;
}
template <typename T>
static int batch_step(EngineT<T>* E, void* X8, void* XB8, void* Y8, double eta, double omega) {
  cudaStream_t st = E->base->stream;
  BatchOp<T, 1> p{(const Oct<T>*)Y8, (Oct<T>*)X8, (Oct<T>*)XB8, E->c, E->l, E->u, (T)(eta / omega), 0};
  CU(launch_stream(E->Kt, p, st));
  BatchOp<T, 2> d{(const Oct<T>*)XB8, (Oct<T>*)Y8, nullptr, E->q, nullptr, nullptr, (T)(eta * omega), (int)E->base->m_ineq};
  CU(launch_stream(E->K, d, st));
  return 0;
}
template <typename T>
static int batch_gap(EngineT<T>* E, const void* X8, const void* Y8, double* out24) {
  cudaStream_t st = E->base->stream;
  BatchGapOp<T> g{(const Oct<T>*)Y8, (const Oct<T>*)X8, E->c, E->l, E->u, E->S};
  CU(launch_stream(E->Kt, g, st));
  CU(cudaMemcpyAsync(out24, (char*)E->S + offsetof(Scalars, batch), 24 * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

// ---------------------------------------------------------------------------
// row-partitioned engine: the pieces between the host-side allreduces (see include/pdlp_b200.h)
// ---------------------------------------------------------------------------
template <typename T>
static int shard_kty(EngineT<T>* E, double* buf) {
  pdlp_engine* B = E->base;
  if (sizeof(T) != sizeof(double)) return fail(PDLP_ERR_INVALID, "the sharded engine is fp64 only");
  SpmvOp<T, false> f{E->y, (T*)buf, E->S, 6};
  // K_g^T in the run layout when it is there: PDLP_LAYOUT=run, or rows so short that SELL's per-slice cost and slot
  // padding dominate (config 5 over 2 / 4 / 8 GPUs: 6.3 / 3.1 / 1.6 entries per row: 1.76 -> 1.70, 0.77 -> 0.65,
  // 0.54 -> 0.45 ms)
  if (E->Kt.run_on) CU(launch_run(E->Kt, f, B->stream));
  else CU(launch_stream(E->Kt, f, B->stream));
  shard_tail_pack_kernel<<<1, 1, 0, B->stream>>>(E->S, buf + B->n);
  CU(cudaGetLastError());
  return 0;
}

template <typename T>
static int shard_step(EngineT<T>* E, const double* buf, int save_prev) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const int n = (int)B->n;
  CU(ensure_kx<T>(E, st));
  shard_rule_kernel<<<1, 1, 0, st>>>(E->S, buf + B->n);
  if (save_prev) {
    shard_primal_kernel<T, true><<<E->ew_grid, 256, 0, st>>>(buf, E->c, E->l, E->u, E->xp, E->xsum, E->X4, n, E->S,
                                                             E->ew_partials, E->ew_ticket, E->bar.xb);
    CU((launch_dual<T, true, false>(E, st)));
    B->have_prev = true;
  } else {
    shard_primal_kernel<T, false><<<E->ew_grid, 256, 0, st>>>(buf, E->c, E->l, E->u, E->xp, E->xsum, E->X4, n, E->S,
                                                              E->ew_partials, E->ew_ticket, E->bar.xb);
    CU((launch_dual<T, false, false>(E, st)));
  }
  CU(cudaGetLastError());
  return 0;
}

// ---- peer-memory exchange: one iteration = SpMV -> exchange kernel -> rule -> primal (polls) -> dual ----
static int xchg_seg(int L, int world) { return ((L + world - 1) / world + 1) / 2 * 2; }
static size_t xchg_block_bytes(int L, int world) {
  // part | recv | tot | (two ranks) the one-shot slots: 2 parities x 2 ranks x L words
  return sizeof(double) * (size_t)L + 16 * ((size_t)world * xchg_seg(L, world) + (size_t)L + (world == 2 ? 4 * (size_t)L : 0));
}

// unmap the peers' blocks (every rank does this, then a host barrier, before any rank frees its own block)
static void xchg_disconnect(pdlp_engine* B) {
  for (int p = 0; p < MAX_PEERS; ++p)
    if (B->xchg_peer[p]) {
      cudaIpcCloseMemHandle(B->xchg_peer[p]);
      B->xchg_peer[p] = nullptr;
    }
  B->xchg_connected = false;
  cudaGetLastError();
}

static void xchg_release(pdlp_engine* B) {
  xchg_disconnect(B);
  if (B->xchg_dev) cudaFree(B->xchg_dev);
  if (B->xchg_block) cudaFree(B->xchg_block);
  B->xchg_dev = nullptr;
  B->xchg_block = nullptr;
  B->xchg_connected = false;
  cudaGetLastError();
}

static int xchg_create(pdlp_engine* B, int world, int rank, unsigned char* handle_out) {
  if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world) return fail(PDLP_ERR_INVALID, "bad world / rank");
  xchg_release(B);
  const int L = ((int)B->n + 4 + 1) / 2 * 2;
  B->xchg_world = world;
  B->xchg_rank = rank;
  B->xchg_len = L;
  CU(cudaMalloc((void**)&B->xchg_block, xchg_block_bytes(L, world)));   // not from the async pool: IPC needs cudaMalloc
  CU(cudaMemset(B->xchg_block, 0, xchg_block_bytes(L, world)));
  CU(cudaMalloc((void**)&B->xchg_dev, sizeof(Xchg)));
  cudaIpcMemHandle_t hd;
  memset(&hd, 0, sizeof(hd));
  if (world > 1) CU(cudaIpcGetMemHandle(&hd, B->xchg_block));
  static_assert(sizeof(hd) == PDLP_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &hd, sizeof(hd));
  CU(cudaDeviceSynchronize());
  return 0;
}

template <typename T>
static int xchg_connect(EngineT<T>* E, const unsigned char* handles) {
  pdlp_engine* B = E->base;
  if (!B->xchg_block) return fail(PDLP_ERR_INVALID, "pdlp_shard_exchange_create has not been called");
  const int world = B->xchg_world, rank = B->xchg_rank, L = B->xchg_len;
  const int seg = xchg_seg(L, world);
  Xchg X;
  memset(&X, 0, sizeof(X));
  for (int p = 0; p < world; ++p) {
    unsigned char* base = B->xchg_block;
    if (p != rank) {
      cudaIpcMemHandle_t hd;
      memcpy(&hd, handles + (size_t)p * PDLP_IPC_HANDLE_BYTES, sizeof(hd));
      void* ptr = nullptr;
      CU(cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess));
      B->xchg_peer[p] = ptr;
      base = (unsigned char*)ptr;
    }
    X.recv[p] = (uint4*)(base + sizeof(double) * (size_t)L);
    X.tot[p] = X.recv[p] + (size_t)world * seg;
    X.os[p] = X.tot[p] + (size_t)L;
  }
  {
    const char* v = getenv("PDLP_XCHG_ONESHOT");      // developer knob: 0 keeps the two-shot protocol with two ranks
    X.oneshot = (world == 2 && !(v && v[0] == '0')) ? 1 : 0;
  }
  X.part = (double*)B->xchg_block;
  X.world = world;
  X.rank = rank;
  X.L = L;
  X.seg = seg;
  X.n = (int)B->n;
  CU(cudaMemcpy(B->xchg_dev, &X, sizeof(X), cudaMemcpyHostToDevice));
  Scalars h;
  CU(cudaMemcpyAsync(&h, E->S, sizeof(h), cudaMemcpyDeviceToHost, B->stream));
  CU(cudaStreamSynchronize(B->stream));
  h.xchg_seq = 1;
  h.xchg_error = 0;
  CU(cudaMemcpyAsync(E->S, &h, sizeof(h), cudaMemcpyHostToDevice, B->stream));
  CU(cudaStreamSynchronize(B->stream));
  B->xchg_connected = true;
  return 0;
}

template <typename T>
static int shard_allreduce_p2p(EngineT<T>* E, double* buf, int64_t count) {
  pdlp_engine* B = E->base;
  if (!B->xchg_connected) return fail(PDLP_ERR_INVALID, "pdlp_shard_exchange_connect has not been called");
  const int64_t cap = B->xchg_len;
  for (int64_t off = 0; off < count; off += cap) {       // longer vectors go through the block in chunks
    const int c = (int)std::min<int64_t>(cap, count - off);
    const int grid = std::max(1, std::min(B->sms * 2, (c + 255) / 256));
    xchg_allreduce_kernel<<<grid, 256, 0, B->stream>>>(B->xchg_dev, E->S, buf + off, c, E->ew_ticket);
    CU(cudaGetLastError());
  }
  return 0;
}

template <typename T>
static int shard_iterate_p2p(EngineT<T>* E, int n_iters, int save_prev) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  if (sizeof(T) != sizeof(double)) return fail(PDLP_ERR_INVALID, "the sharded engine is fp64 only");
  if (!B->xchg_connected) return fail(PDLP_ERR_INVALID, "pdlp_shard_exchange_connect has not been called");
  if (E->Kt.grid > B->sms) return fail(PDLP_ERR_STATE, "the fused exchange needs a co-resident grid");
  CU(ensure_smem_attr(shard_fused_kernel<T, true>, StreamCfg<T>::MAX_SMEM_BYTES - 4096));
  CU(ensure_smem_attr(shard_fused_kernel<T, false>, StreamCfg<T>::MAX_SMEM_BYTES - 4096));
  CU(ensure_kx<T>(E, st));
  for (int i = 0; i < n_iters; ++i) {
    const bool save = save_prev && i == n_iters - 1;
    // column side: SpMV + exchange + step rule + primal update in ONE kernel; row side: the ordinary DualOp
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    pdl_config(cfg, attr, E->Kt.grid, StreamCfg<T>::NT, (size_t)E->Kt.smem_bytes, st);
    const Xchg* xd = B->xchg_dev;
    if (save) {
      CU(cudaLaunchKernelEx(&cfg, shard_fused_kernel<T, true>, E->Kt.view(), xd, (const T*)E->y, (const T*)E->c,
                            (const T*)E->l, (const T*)E->u, E->xp, E->xsum, E->X4, E->S, E->ew_partials, E->ew_ticket,
                            E->bar.xb));
      CU((launch_dual<T, true, false>(E, st)));
      B->have_prev = true;
    } else {
      CU(cudaLaunchKernelEx(&cfg, shard_fused_kernel<T, false>, E->Kt.view(), xd, (const T*)E->y, (const T*)E->c,
                            (const T*)E->l, (const T*)E->u, E->xp, E->xsum, E->X4, E->S, E->ew_partials, E->ew_ticket,
                            E->bar.xb));
      CU((launch_dual<T, false, false>(E, st)));
    }
    CU(cudaGetLastError());
  }
  return 0;
}

template <typename T>
static int shard_check_begin(EngineT<T>* E, double* buf3, int have_prev) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const int n = (int)B->n, m = (int)B->m;
  if (sizeof(T) != sizeof(double)) return fail(PDLP_ERR_INVALID, "the sharded engine is fp64 only");
  flush_kernel<T><<<ew_blocks(std::max(n, m), B->sms), 256, 0, st>>>(E->xp, E->y, E->xsum, E->ysum, E->X4, E->Y4, n, m,
                                                                     E->S, have_prev ? 1 : 0);
  CU(cudaGetLastError());
  Spmv3Op<T> f{E->Y4, (T*)buf3};
  CU(launch_stream(E->Kt, f, st));
  return 0;
}

template <typename T>
static int shard_check_mid(EngineT<T>* E, const double* buf3, double* buf9) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  shard_kkt_cols_kernel<T><<<E->ew_grid, 256, 0, st>>>(buf3, E->X4, E->c, E->l, E->u, B->scaled ? E->dcol : nullptr,
                                                        (int)B->n, E->S, E->kkt_partials, E->ew_ticket);
  KktRowsOp<T> b{E->X4, E->Y4, E->q, B->scaled ? E->drow : nullptr, E->S, 0};
  CU(launch_stream(E->K, b, st));
  shard_rows_pack_kernel<<<1, 1, 0, st>>>(E->S, buf9);
  CU(cudaGetLastError());
  return 0;
}

template <typename T>
static int shard_check_decide(EngineT<T>* E, const double* buf9, int force, double* buf1) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  shard_rows_unpack_kernel<<<1, 1, 0, st>>>(E->S, buf9);
  decide_kernel<<<1, 1, 0, st>>>(E->S, force);
  apply_restart_kernel<T><<<E->ew_grid, 256, 0, st>>>(E->xp, E->y, E->xsum, E->ysum, E->X4, E->Y4, E->xlast, E->ylast,
                                                      (int)B->n, (int)B->m, E->S, E->ew_partials, E->ew_ticket);
  if (E->bar.on) CU(launch_kx_refresh<T>(E, 1, st));
  shard_dy_pack_kernel<<<1, 1, 0, st>>>(E->S, buf1);
  CU(cudaGetLastError());
  return 0;
}

template <typename T>
static int shard_check_end(EngineT<T>* E, const double* buf1, pdlp_check* out) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  shard_dy_unpack_kernel<<<1, 1, 0, st>>>(E->S, buf1);
  post_restart_kernel<<<1, 1, 0, st>>>(E->S, B->scaled ? 1 : 0);
  CU(cudaGetLastError());
  B->have_prev = false;
  Scalars h;
  CU(cudaMemcpyAsync(&h, E->S, sizeof(h), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (h.xchg_error) return fail(PDLP_ERR_CUDA, "peer-memory exchange timed out waiting for another rank");
  if (out) fill_check(out, h);
  return 0;
}

template <typename T>
static int engine_objective(EngineT<T>* E, double* out) {
  pdlp_engine* B = E->base;
  objective_kernel<T><<<1, 1024, 0, B->stream>>>(E->c, E->xp, (int)B->n, E->S, 7);
  CU(cudaGetLastError());
  return read_scratch<T>(E, 7, out);
}

// ---------------------------------------------------------------------------
// whole solve: primal_dual_hybrid_gradient.py:54-181 with the reference's k / n / j accounting
// ---------------------------------------------------------------------------
template <typename T>
static int engine_solve(EngineT<T>* E, const pdlp_solve_options* o, pdlp_solve_result* res, pdlp_check* trace) {
  pdlp_engine* B = E->base;
  cudaStream_t st = B->stream;
  const auto t0 = std::chrono::steady_clock::now();
  auto elapsed = [&]() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() + o->time_used;
  };
  struct Events {             // destroyed on every return path
    cudaEvent_t a = nullptr, b = nullptr;
    ~Events() {
      if (a) cudaEventDestroy(a);
      if (b) cudaEventDestroy(b);
    }
  } evs;
  CU(cudaEventCreate(&evs.a));
  CU(cudaEventCreate(&evs.b));
  const cudaEvent_t ev0 = evs.a, ev1 = evs.b;
  CU(cudaEventRecord(ev0, st));
  const int period = o->restart_period > 0 ? o->restart_period : 40;
  Scalars h;
  CU(cudaMemcpyAsync(&h, E->S, sizeof(h), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  int64_t k = h.k, j = 0;
  int n_restart = 0, status = 1, n_trace = 0;
  const bool infeas = h.infeas_detect != 0;
  double prim_obj = NAN;
  bool timed_out = false, detected = false;
  // Pipelined loop (graph mode, no infeasibility detection).  In the steady state every unit of work is the same pair
  // of graphs -- a block of `period` iterations that saves the previous iterate, then the restart test -- whatever the
  // test decides: the restart itself is applied on the device.  So the host issues unit i + 1 BEFORE it reads the
  // record of unit i: the GPU never waits for the host round trip (read-back, wake-up, next launch).  A restart point
  // that meets the termination test raises Scalars::stop on the device (stop_on_solved), which turns the unit issued
  // ahead into no-ops; the records of two tests in flight land in two pinned buffers.  A unit is issued ahead only when
  // the KKT-pass budget covers it in the worst case (this test restarts), so the end game runs the sequential code.
  const bool pipelined = !infeas && o->use_graph == 1 && !getenv("PDLP_NO_PIPELINE");
  if (pipelined) {
    for (int i = 0; i < 2; ++i) {
      if (!E->h_rec[i]) E->h_rec[i] = (Scalars*)pinned_scalars_get();
      if (!E->h_rec[i]) return fail(PDLP_ERR_CUDA, "cudaMallocHost failed");
      if (!E->rec_ev[i]) CU(cudaEventCreateWithFlags(&E->rec_ev[i], cudaEventDisableTiming));
    }
    if (!E->check_graph_async) {
      int rc = capture_check_graph<T>(E, 0, 1, false, &E->check_graph_async);
      if (rc) return rc;
    }
    const int one = 1;
    CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, stop_on_solved), &one, sizeof(int), cudaMemcpyHostToDevice, st));
  }
  // restart test of a steady-state unit, asynchronous: its record lands in h_rec[slot], rec_ev[slot] marks the arrival
  auto issue_check = [&](int slot) -> int {
    CU(cudaGraphLaunch(E->check_graph_async, st));
    CU(cudaMemcpyAsync(E->h_rec[slot], E->S, sizeof(Scalars), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(E->rec_ev[slot], st));
    B->have_prev = false;
    return 0;
  };
  int slot = 0;
  bool in_flight = false;           // the next unit (block + test) is already issued; its record goes to h_rec[slot]
  while (j < o->max_kkt && status == 1) {
    int t = 0;
    bool restarted = false;
    while (j < o->max_kkt) {
      if (!in_flight && elapsed() >= o->time_limit) { timed_out = true; break; }
      // iterations until the next restart test, limited by the KKT-pass budget (:67: guard before every iteration)
      int want = period - (t % period);
      int iters = 0;
      int64_t jj = j, kk = k;
      while (iters < want && jj < o->max_kkt) {
        ++kk;
        jj += 1 + ((infeas && kk > 1) ? 1 : 0);
        ++iters;
      }
      const int at_check = ((t + iters) % period == 0) ? 1 : 0;
      const bool steady = pipelined && iters == period && at_check;
      int rc = 0;
      if (!in_flight) {
        rc = engine_iterate<T>(E, iters, at_check, o->use_graph);
        if (rc) return rc;
      } else if (!steady) {
        return fail(PDLP_ERR_STATE, "pipelined solve: a unit was issued ahead of a non-steady step");
      }
      if (infeas) {
        // :89-101: the device raised `stop` at the iteration where a certificate test passed
        Scalars hs;
        CU(cudaMemcpyAsync(&hs, E->S, sizeof(hs), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (hs.infeas) {
          while (k < hs.infeas_k) { ++k; j += 1 + (k > 1 ? 1 : 0); }
          status = hs.infeas == 2 ? 3 : 4;
          rc = engine_objective<T>(E, &prim_obj);
          if (rc) return rc;
          detected = true;
          break;
        }
      }
      j = jj; k = kk; t += iters;
      if (at_check) {
        pdlp_check chk;
        if (steady) {
          if (!in_flight) {
            rc = issue_check(slot);
            if (rc) return rc;
          }
          in_flight = false;
          // the next unit, ahead of this test's record (worst case for the budget: this test restarts, +3 +2)
          const bool ahead = j + 5 + period <= o->max_kkt && elapsed() < o->time_limit;
          if (ahead) {
            rc = engine_iterate<T>(E, period, 1, o->use_graph);
            if (!rc) rc = issue_check(slot ^ 1);
            if (rc) return rc;
          }
          CU(cudaEventSynchronize(E->rec_ev[slot]));
          fill_check(&chk, *E->h_rec[slot]);
          if (ahead) {
            in_flight = true;
            slot ^= 1;
          }
        } else {
          rc = engine_restart_check<T>(E, 0, &chk);
          if (rc) return rc;
        }
        j += 3;                                    // :128
        if (chk.kind != PDLP_RESTART_NONE) {
          restarted = true;
          ++n_restart;
          j += 2;                                  // :154, :165
          prim_obj = chk.point.prim_obj;
          if (trace && n_trace < o->max_trace) trace[n_trace++] = chk;
          if (chk.solved) status = 0;
          break;
        }
      }
    }
    if (timed_out) { status = 2; break; }
    if (detected) break;
    if (!restarted) {
      // loop left through the guard: the epilogue runs on the current iterate (:148-177)
      pdlp_check chk;
      int rc = engine_restart_check<T>(E, 1, &chk);
      if (rc) return rc;
      ++n_restart;
      j += 2;
      prim_obj = chk.point.prim_obj;
      if (trace && n_trace < o->max_trace) trace[n_trace++] = chk;
      if (chk.solved) status = 0;
    }
  }
  if (pipelined) {
    // disarm; a unit issued ahead of the final record has run as no-ops behind `stop`, which is lowered again
    const int zero2[2] = {0, 0};
    CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, stop_on_solved), zero2, sizeof(int), cudaMemcpyHostToDevice, st));
    if (status == 0)
      CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, stop), zero2, sizeof(int), cudaMemcpyHostToDevice, st));
  }
  CU(cudaEventRecord(ev1, st));
  CU(cudaEventSynchronize(ev1));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, ev0, ev1));
  res->k = k; res->j = j; res->n = n_restart; res->status = status; res->prim_obj = prim_obj;
  res->seconds = elapsed(); res->device_ms = ms; res->n_trace = n_trace; res->reserved = 0;
  return 0;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
#define DISPATCH(h, fn, ...)                                                         \
  ((h)->dtype == PDLP_F64 ? fn<double>((EngineT<double>*)(h)->impl, ##__VA_ARGS__)   \
                          : fn<float>((EngineT<float>*)(h)->impl, ##__VA_ARGS__))

#define CHECK_HANDLE(h)                                              \
  do {                                                               \
    if (!(h) || !(h)->impl) return fail(PDLP_ERR_INVALID, "null handle"); \
    cudaError_t e_ = cudaSetDevice((h)->device);                     \
    if (e_ != cudaSuccess) return fail(PDLP_ERR_CUDA, cudaGetErrorString(e_)); \
  } while (0)

static int create_common(pdlp_handle* out, int dtype, int64_t m, int64_t n, int64_t m_ineq, int64_t nnz,
                         const void* K_rowptr, const void* K_colidx, const void* K_vals, const int32_t* Kt_rowptr,
                         const int32_t* Kt_colidx, const void* Kt_vals, const void* c, const void* q, const void* l,
                         const void* u, void* stream, int host_index_bytes) {
  if (!out) return fail(PDLP_ERR_INVALID, "out is null");
  *out = nullptr;
  if (dtype != PDLP_F64 && dtype != PDLP_F32) return fail(PDLP_ERR_INVALID, "dtype must be PDLP_F64 or PDLP_F32");
  if (m <= 0 || n <= 0 || m_ineq < 0 || m_ineq > m || nnz < 0) return fail(PDLP_ERR_INVALID, "bad dimensions");
  if (m >= INT32_MAX - 64 || n >= INT32_MAX - 64 || nnz >= INT32_MAX - 64)
    return fail(PDLP_ERR_INVALID, "dimensions exceed int32 indexing");
  const bool host = host_index_bytes != 0;
  if (host && host_index_bytes != 4 && host_index_bytes != 8) return fail(PDLP_ERR_INVALID, "index_bytes must be 4 or 8");
  if (!K_rowptr || !c || !q || !l || !u || (nnz > 0 && (!K_colidx || !K_vals)))
    return fail(PDLP_ERR_INVALID, "null problem pointer");
  if (!host && (!Kt_rowptr || (nnz > 0 && (!Kt_colidx || !Kt_vals)))) return fail(PDLP_ERR_INVALID, "null problem pointer");
  SetupTrace tr;
  int dev = 0, cc_major = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  // (two attribute queries, not cudaGetDeviceProperties: that call asks the kernel driver for every property of the
  // device and takes milliseconds -- tens of them when other processes keep the driver busy -- on every engine build)
  CU(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major < 9) return fail(PDLP_ERR_CUDA, "pdlp_b200 needs TMA bulk copies (built for sm_100a)");
  pdlp_engine* B = new pdlp_engine();
  B->dtype = dtype; B->device = dev; B->stream = (cudaStream_t)stream;
  B->m = m; B->n = n; B->m_ineq = m_ineq; B->nnz = nnz;
  B->sms = sms;
  init_l2_persistence();
  {
    // keep freed blocks in the device's default pool: the next engine build does not go back to the driver
    static std::mutex mu;
    static std::vector<int> done;
    std::lock_guard<std::mutex> lock(mu);
    if (std::find(done.begin(), done.end(), dev) == done.end()) {
      cudaMemPool_t pool;
      CU(cudaDeviceGetDefaultMemPool(&pool, dev));
      uint64_t keep = UINT64_MAX;
      CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
      done.push_back(dev);
    }
  }
  tr.lap("create: device queries");
  const int32_t *kr = (const int32_t*)K_rowptr, *kc = (const int32_t*)K_colidx;   // (host path: reinterpreted by index_bytes)
  int rc = dtype == PDLP_F64
               ? engine_create<double>(B, kr, kc, K_vals, Kt_rowptr, Kt_colidx, Kt_vals, c, q, l, u, host_index_bytes)
               : engine_create<float>(B, kr, kc, K_vals, Kt_rowptr, Kt_colidx, Kt_vals, c, q, l, u, host_index_bytes);
  if (!rc) rc = dtype == PDLP_F64 ? refresh_norms<double>((EngineT<double>*)B->impl)
                                  : refresh_norms<float>((EngineT<float>*)B->impl);
  tr.lap("create: engine + norms");
  if (rc) {
    const std::string msg = g_err;                    // (the teardown must not replace the message)
    if (dtype == PDLP_F64) engine_destroy<double>(B); else engine_destroy<float>(B);
    delete B;
    g_err = msg;
    return rc;
  }
  *out = B;
  return 0;
}

extern "C" int pdlp_create(pdlp_handle* out, int dtype, int64_t m, int64_t n, int64_t m_ineq, int64_t nnz,
                           const int32_t* K_rowptr, const int32_t* K_colidx, const void* K_vals,
                           const int32_t* Kt_rowptr, const int32_t* Kt_colidx, const void* Kt_vals, const void* c,
                           const void* q, const void* l, const void* u, void* stream) {
  return create_common(out, dtype, m, n, m_ineq, nnz, K_rowptr, K_colidx, K_vals, Kt_rowptr, Kt_colidx, Kt_vals, c, q, l,
                       u, stream, 0);
}

extern "C" int pdlp_create_host(pdlp_handle* out, int dtype, int64_t m, int64_t n, int64_t m_ineq, int64_t nnz,
                                const void* host_K_rowptr, const void* host_K_colidx, int32_t index_bytes,
                                const void* host_K_vals, const void* c, const void* q, const void* l, const void* u,
                                void* stream) {
  return create_common(out, dtype, m, n, m_ineq, nnz, host_K_rowptr, host_K_colidx, host_K_vals, nullptr, nullptr, nullptr,
                       c, q, l, u, stream, index_bytes ? index_bytes : -1);
}

extern "C" int pdlp_get_matrix_structure(pdlp_handle h, int32_t* K_rowptr, int32_t* K_colidx) {
  if (!h || !h->impl) return fail(PDLP_ERR_INVALID, "null handle");
  cudaError_t e_ = cudaSetDevice(h->device);
  if (e_ != cudaSuccess) return fail(PDLP_ERR_CUDA, cudaGetErrorString(e_));
  const int* rp; const int* ci;
  if (h->dtype == PDLP_F64) { auto* E = (EngineT<double>*)h->impl; rp = E->K.rowptr; ci = E->K.colidx; }
  else { auto* E = (EngineT<float>*)h->impl; rp = E->K.rowptr; ci = E->K.colidx; }
  if (K_rowptr) CU(cudaMemcpyAsync(K_rowptr, rp, sizeof(int) * ((size_t)h->m + 1), cudaMemcpyDeviceToDevice, h->stream));
  if (K_colidx) CU(cudaMemcpyAsync(K_colidx, ci, sizeof(int) * (size_t)h->nnz, cudaMemcpyDeviceToDevice, h->stream));
  return 0;
}

extern "C" int pdlp_plan_inspect(const int32_t* host_rowptr, int64_t rows, int32_t value_bytes, int32_t sms,
                                 pdlp_plan_info* o, int8_t* row_class) {
  if (!host_rowptr || !o || rows < 0 || rows > 0x7fffffff || (value_bytes != 4 && value_bytes != 8) || sms < 1)
    return fail(PDLP_ERR_INVALID, "bad argument");
  std::vector<int> rowptr(host_rowptr, host_rowptr + rows + 1);
  for (int64_t r = 0; r < rows; ++r)
    if (rowptr[r + 1] < rowptr[r]) return fail(PDLP_ERR_INVALID, "rowptr is not monotone");
  const int W = PDLP_W, KB = PDLP_KB;
  HostSell P = build_sell_plan(rowptr.data(), (int)rows, value_bytes, W, KB, PDLP_SIGMA, PDLP_LONG_ROW, PDLP_WARP_ROW,
                               sms * PDLP_CTAS * PDLP_OVERSUB);
  memset(o, 0, sizeof(*o));
  o->rows = rows;
  o->nnz = rowptr[rows] - rowptr[0];
  o->cta_rows = (int64_t)P.long_rows.size();
  o->sigma = P.sigma;
  o->grid = P.grid;
  o->warps = P.grid * W;
  o->pieces = (int32_t)P.pieces.size();
  o->max_piece_bytes = P.max_piece_bytes;
  o->packed_bytes = (int64_t)P.packed_bytes;
  const int stage = (P.max_piece_bytes + 127) / 128 * 128;
  o->stages = std::max(2, std::min((int)PDLP_SMAX, (int)(PDLP_SMEM_BUDGET / PDLP_CTAS / ((size_t)W * stage))));
  // classes and coverage
  std::vector<int> seen(rows, 0);
  std::vector<int8_t> cls(rows, -1);
  bool ok = true;
  for (int r : P.long_rows) {
    if (r < 0 || r >= rows) { ok = false; continue; }
    ++seen[r];
    cls[r] = 2;
  }
  for (const Piece& pc : P.pieces) {
    o->slots += (int64_t)pc.klen * pc.nslices * 32;
    if (!(pc.flags & PIECE_FIRST)) continue;
    for (int i = 0; i < (int)pc.nslices * 32; ++i) {
      const int r = P.perm[(size_t)pc.perm_off + i];
      if (r < 0) continue;
      if (r >= rows) { ok = false; continue; }
      if (pc.flags & PIECE_WARP_ROW) {
        if (i == 0) { ++seen[r]; cls[r] = 1; }
        else if (r != P.perm[(size_t)pc.perm_off]) ok = false;
      } else {
        ++seen[r];
        cls[r] = 0;
      }
    }
  }
  for (int64_t r = 0; r < rows; ++r) {
    if (seen[r] != 1) ok = false;
    if (cls[r] == 0) ++o->lane_rows;
    if (cls[r] == 1) ++o->warp_rows;
    if (row_class) row_class[r] = cls[r];
  }
  o->covered_once = ok ? 1 : 0;
  // warp ranges
  bool rok = (int)P.warp_begin.size() == o->warps + 1 && P.warp_begin.front() == 0 && P.warp_begin.back() == o->pieces;
  int64_t maxb = 0, sumb = 0;
  for (int w = 0; rok && w < o->warps; ++w) {
    const int b = P.warp_begin[w], e = P.warp_begin[w + 1];
    if (b > e) { rok = false; break; }
    if (b < e && !(P.pieces[b].flags & PIECE_FIRST)) rok = false;
    int64_t bytes = 0;
    for (int p = b; p < e; ++p) bytes += 16 + ((P.pieces[p].flags & PIECE_FIRST) ? (int64_t)P.pieces[p].nslices * 128 : 0) +
                                         (int64_t)P.pieces[p].klen * P.pieces[p].nslices * 32 * (4 + value_bytes);
    maxb = std::max(maxb, bytes);
    sumb += bytes;
  }
  o->ranges_ok = rok ? 1 : 0;
  o->max_warp_bytes = maxb;
  o->mean_warp_bytes = o->warps ? sumb / o->warps : 0;
  return 0;
}

// ---------------------------------------------------------------------------
// Host-side checker of the run layout (tests/test_cpu_plan.py): plans and packs a CSR matrix with the SAME planner and
// per-slot packer the device path uses, then walks the chunks with a 32-lane emulation of run_phase (ballots, slots,
// segmented carry scan, staging rounds) and returns out = A x.  mode 0: plain layout (rows above long_row are summed
// directly, as long_rows_phase does); mode 1: shared-memory panels + combine.  No GPU involved.
// ---------------------------------------------------------------------------
template <typename CT>
static void emulate_walk(const unsigned char* chunks, int cb, int ce, const double* operand, double* out, int cap,
                         int64_t* rows_done) {
  using Cfg = RunCfg<double, CT>;
  constexpr int E = Cfg::E;
  double carry = 0.0;
  std::vector<double> stg(cap);
  for (int c = cb; c < ce; ++c) {
    const unsigned char* st = chunks + (size_t)c * Cfg::CHUNK_BYTES;
    const double* vals = reinterpret_cast<const double*>(st);
    const CT* cols = reinterpret_cast<const CT*>(st + Cfg::VALS_BYTES);
    const unsigned* meta = reinterpret_cast<const unsigned*>(st + Cfg::VALS_BYTES + Cfg::COLS_BYTES);
    const RunTrailer tr = *reinterpret_cast<const RunTrailer*>(st + Cfg::VALS_BYTES + Cfg::COLS_BYTES + Cfg::META_BYTES);
    const int nrows = tr.nrows;
    double g[32][E];
    for (int lane = 0; lane < 32; ++lane)
      for (int k = 0; k < E; ++k) g[lane][k] = operand[(int)(cols[k * 32 + lane] & Cfg::MASK)];
    double cin[32], tail[32];
    for (int lo = 0; lo == 0 || lo < nrows; lo += cap) {
      for (int lane = 0; lane < 32; ++lane) {
        const unsigned endm = meta[lane] & 0xffu;
        double acc = 0.0;
        int slot = (int)((meta[lane] >> 8) & 0x1ffu) - lo;
        for (int k = 0; k < E; ++k) {
          acc += vals[k * 32 + lane] * g[lane][k];
          if ((tr.any_steps >> k) & 1u) {
            if ((endm >> k) & 1u) {
              if ((unsigned)slot < (unsigned)cap) stg[slot] = acc;
              ++slot;
              acc = 0.0;
            }
          }
        }
        tail[lane] = acc;
      }
      if (lo == 0) {
        const unsigned F = tr.end_lanes;
        double sv[32];
        for (int lane = 0; lane < 32; ++lane) sv[lane] = (lane == 0 && !(meta[0] & 0xffu)) ? carry + tail[0] : tail[lane];
        for (int d = 1; d < 32; d <<= 1) {
          double nv[32];
          for (int lane = 0; lane < 32; ++lane) {
            const unsigned wm = ((1u << d) - 1u) << ((lane - d + 1) & 31);
            const bool take = lane >= d && !(F & wm);
            nv[lane] = take ? sv[lane - d] + sv[lane] : sv[lane];
          }
          for (int lane = 0; lane < 32; ++lane) sv[lane] = nv[lane];
        }
        for (int lane = 0; lane < 32; ++lane) cin[lane] = lane == 0 ? carry : sv[lane - 1];
        carry = sv[31];
      }
      for (int lane = 0; lane < 32; ++lane) {
        const int first = (int)((meta[lane] >> 8) & 0x1ffu) - lo;
        if ((meta[lane] & 0xffu) && (unsigned)first < (unsigned)cap) stg[first] = cin[lane] + stg[first];
      }
      const int cnt = std::min(cap, nrows - lo);
      for (int i = 0; i < cnt; ++i) {
        out[tr.row0 + lo + i] = stg[i];
        ++*rows_done;
      }
    }
  }
}

extern "C" int pdlp_run_emulate(const int32_t* rowptr, const int32_t* colidx, const double* vals, int64_t rows,
                                int64_t cols, int32_t sms, int32_t long_row, int32_t mode, const double* x, double* out,
                                pdlp_run_info* info) {
  if (!rowptr || !x || !out || !info || rows < 1 || cols < 1 || sms < 1 || rows > 0x7ffffff0 || cols > 0x3ffffff0)
    return fail(PDLP_ERR_INVALID, "bad argument");
  memset(info, 0, sizeof(*info));
  const int R = (int)rows;
  int64_t rows_done = 0;
  for (int r = 0; r < R; ++r) out[r] = NAN;
  if (mode == 0) {
    using Cfg = RunCfg<double, uint32_t>;
    HostRun P = build_run_plan(rowptr, R, long_row, sms, PDLP_RUN_W, Cfg::CHUNK);
    if (!P.ok) return fail(PDLP_ERR_INVALID, "stream too long");
    const int nchunks = (int)P.chunks.size();
    std::vector<unsigned char> chunks((size_t)nchunks * Cfg::CHUNK_BYTES + 16, 0xcd);
    const int nw = P.grid * PDLP_RUN_W;
    for (int c = 0; c < nchunks; ++c) {
      unsigned char* blk = chunks.data() + (size_t)c * Cfg::CHUNK_BYTES;
      unsigned char ends[Cfg::CHUNK];
      int row0 = -1;
      for (int s = 0; s < Cfg::CHUNK; ++s) {
        int row;
        ends[s] = run_pack_slot<double, uint32_t>(P.chunks[c], s, R, P.sptr.data(), rowptr, colidx, vals, blk, &row) ? 1 : 0;
        if (s == 0) row0 = row;
      }
      run_pack_meta_host<Cfg::E>(ends, row0, reinterpret_cast<unsigned*>(blk + Cfg::VALS_BYTES + Cfg::COLS_BYTES),
                                 reinterpret_cast<RunTrailer*>(blk + Cfg::VALS_BYTES + Cfg::COLS_BYTES + Cfg::META_BYTES));
    }
    for (int w = 0; w < nw; ++w)
      emulate_walk<uint32_t>(chunks.data(), P.warp_chunk_begin[w], P.warp_chunk_begin[w + 1], x, out, Cfg::CAP, &rows_done);
    for (int r : P.long_rows) {
      double a = 0.0;
      for (int e = rowptr[r]; e < rowptr[r + 1]; ++e) a += vals[e] * x[colidx[e]];
      out[r] = a;
      ++rows_done;
    }
    info->grid = P.grid; info->warps = nw; info->chunks = nchunks; info->long_rows = (int32_t)P.long_rows.size();
    info->stream_entries = P.sptr[R];
    info->chunk_bytes = Cfg::CHUNK_BYTES;
    int64_t mx = 0;
    for (int w = 0; w < nw; ++w) mx = std::max<int64_t>(mx, P.warp_chunk_begin[w + 1] - P.warp_chunk_begin[w]);
    info->max_warp_chunks = (int32_t)mx;
  } else {
    // nonzeros per column (= per element of the gathered vector)
    std::vector<int> cpre((size_t)cols + 1, 0);
    for (int e = rowptr[0]; e < rowptr[R]; ++e) ++cpre[(size_t)colidx[e] + 1];
    for (int64_t j = 0; j < cols; ++j) cpre[j + 1] += cpre[j];
    HostPanel H;
    if (!panel_boundaries(cpre.data(), (int)cols, 8, sms, H)) return fail(PDLP_ERR_INVALID, "the vector does not fit the panels");
    std::vector<int> cnt((size_t)H.P * R, 0);
    for (int r = 0; r < R; ++r)
      for (int e = rowptr[r]; e < rowptr[r + 1]; ++e)
        for (int p = 0; p < H.P; ++p)
          if (colidx[e] >= H.panel_begin[p] && colidx[e] < H.panel_begin[p + 1]) ++cnt[(size_t)p * R + r];
    if (!panel_plan(cnt.data(), R, 8, PDLP_W, PDLP_KB, H)) return fail(PDLP_ERR_INVALID, "a row part needs the CTA path");
    // the panel matrix on the host (what panel_fill_kernel writes)
    std::vector<int> pcol((size_t)rowptr[R] - rowptr[0]);
    std::vector<double> pval(pcol.size());
    {
      std::vector<int> cur(H.rowptr.begin(), H.rowptr.end() - 1);
      for (int r = 0; r < R; ++r)
        for (int e = rowptr[r]; e < rowptr[r + 1]; ++e)
          for (int p = 0; p < H.P; ++p)
            if (colidx[e] >= H.panel_begin[p] && colidx[e] < H.panel_begin[p + 1]) {
              int& c = cur[(size_t)p * R + r];
              pcol[c] = colidx[e] - H.panel_begin[p];
              pval[c] = vals[e];
              ++c;
            }
    }
    // walk the plan the way the kernel's CTAs do: CTA b serves panel cta_panel[b] and must only meet rows of it
    const HostSell& S = H.sell;
    std::vector<double> partial((size_t)H.P * R, NAN);
    std::vector<int> seen((size_t)H.P * R, 0);
    bool ok = (int)S.warp_begin.size() == H.grid * PDLP_W + 1;
    for (int b = 0; ok && b < H.grid; ++b) {
      const int p = H.cta_panel[b];
      for (int pc = S.warp_begin[(size_t)b * PDLP_W]; pc < S.warp_begin[(size_t)(b + 1) * PDLP_W]; ++pc) {
        const Piece& d = S.pieces[pc];
        if (!(d.flags & PIECE_FIRST)) continue;
        for (int i = 0; i < (int)d.nslices * 32; ++i) {
          const int row = S.perm[(size_t)d.perm_off + i];
          if (row < 0 || ((d.flags & PIECE_WARP_ROW) && i > 0)) continue;
          if (row / R != p) { ok = false; continue; }
          ++seen[row];
          double a = 0.0;
          for (int e = H.rowptr[row]; e < H.rowptr[row + 1]; ++e) a += pval[e] * x[H.panel_begin[p] + pcol[e]];
          partial[row] = a;
        }
      }
    }
    for (size_t i = 0; i < seen.size(); ++i) {
      if (seen[i] != 1) ok = false;
      else ++rows_done;
    }
    if (!ok) return fail(PDLP_ERR_STATE, "panel plan does not cover every (panel, row) exactly once inside its panel's CTAs");
    for (int r = 0; r < R; ++r) {
      double a = partial[r];
      for (int p = 1; p < H.P; ++p) a += partial[(size_t)p * R + r];
      out[r] = a;
    }
    rows_done /= H.P;
    info->grid = H.grid; info->warps = H.grid * PDLP_W; info->chunks = (int32_t)S.pieces.size(); info->panels = H.P;
    info->panel_cap = H.panel_cap;
    info->stream_entries = 0;
    for (const Piece& d : S.pieces) info->stream_entries += (int64_t)d.klen * d.nslices * 32;     // padded slots
    info->chunk_bytes = S.max_piece_bytes;
    info->max_warp_chunks = S.sigma;
  }
  info->rows_finished = rows_done;
  return 0;
}

extern "C" int pdlp_destroy(pdlp_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->dtype == PDLP_F64) engine_destroy<double>(h); else engine_destroy<float>(h);
  delete h;
  return 0;
}

extern "C" int pdlp_set_stream(pdlp_handle h, void* stream) {
  CHECK_HANDLE(h);
  h->stream = (cudaStream_t)stream;
  return 0;
}

extern "C" int64_t pdlp_bytes_per_iteration(pdlp_handle h) {
  if (!h) return 0;
  const int64_t vb = h->dtype == PDLP_F64 ? 8 : 4;
  return 2 * (vb + 4) * h->nnz + 4 * (h->m + h->n + 2) + 9 * vb * h->n + 6 * vb * h->m;
}

template <typename T>
static int t_set_iterate(EngineT<T>* E, const void* x, const void* y) {
  pdlp_engine* B = E->base;
  const int n = (int)B->n, m = (int)B->m;
  set_iterate_kernel<T><<<ew_blocks(std::max(m, n), B->sms), 256, 0, B->stream>>>(
      (const T*)x, (const T*)y, E->xp, E->y, E->xsum, E->ysum, E->xlast, E->ylast, n, m);
  CU(cudaGetLastError());
  E->bar.dirty = true;
  // :58-64 / :104-112: fresh epoch
  const double z[2] = {0.0, 0.0};
  CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, w_pending), z, 2 * sizeof(double), cudaMemcpyHostToDevice, B->stream));
  const int zi = 0;
  CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, t), &zi, sizeof(int), cudaMemcpyHostToDevice, B->stream));
  CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, stop), &zi, sizeof(int), cudaMemcpyHostToDevice, B->stream));
  CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, stop_on_solved), &zi, sizeof(int), cudaMemcpyHostToDevice, B->stream));
  CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, infeas), &zi, sizeof(int), cudaMemcpyHostToDevice, B->stream));
  CU(cudaMemcpyAsync((char*)E->S + offsetof(Scalars, infeas_k), &zi, sizeof(int), cudaMemcpyHostToDevice, B->stream));
  CU(cudaMemsetAsync(E->lam_prev, 0, sizeof(T) * (size_t)n, B->stream));   // pdlp :39-40
  CU(cudaStreamSynchronize(B->stream));
  B->have_prev = false;
  return 0;
}
extern "C" int pdlp_set_iterate(pdlp_handle h, const void* x, const void* y) {
  CHECK_HANDLE(h);
  return DISPATCH(h, t_set_iterate, x, y);
}

template <typename T>
static int t_get_iterate(EngineT<T>* E, void* x, void* y) {
  pdlp_engine* B = E->base;
  get_iterate_kernel<T><<<ew_blocks(std::max(B->m, B->n), B->sms), 256, 0, B->stream>>>(E->xp, E->y, (T*)x, (T*)y,
                                                                                        (int)B->n, (int)B->m);
  CU(cudaGetLastError());
  return 0;
}
extern "C" int pdlp_get_iterate(pdlp_handle h, void* x, void* y) {
  CHECK_HANDLE(h);
  return DISPATCH(h, t_get_iterate, x, y);
}

template <typename T>
static int t_get_average(EngineT<T>* E, void* x, void* y) {
  pdlp_engine* B = E->base;
  get_average_kernel<T><<<ew_blocks(std::max(B->m, B->n), B->sms), 256, 0, B->stream>>>(
      E->xp, E->y, E->xsum, E->ysum, (T*)x, (T*)y, (int)B->n, (int)B->m, E->S);
  CU(cudaGetLastError());
  return 0;
}
extern "C" int pdlp_get_average(pdlp_handle h, void* x, void* y) {
  CHECK_HANDLE(h);
  return DISPATCH(h, t_get_average, x, y);
}

template <typename T>
static int t_get_scalars(EngineT<T>* E, Scalars* hs) {
  CU(cudaMemcpyAsync(hs, E->S, sizeof(Scalars), cudaMemcpyDeviceToHost, E->base->stream));
  CU(cudaStreamSynchronize(E->base->stream));
  return 0;
}
template <typename T>
static int t_put_scalars(EngineT<T>* E, const Scalars* hs) {
  CU(cudaMemcpyAsync(E->S, hs, sizeof(Scalars), cudaMemcpyHostToDevice, E->base->stream));
  CU(cudaStreamSynchronize(E->base->stream));
  return 0;
}

extern "C" int pdlp_set_step(pdlp_handle h, double eta, double omega, double theta, int adaptive) {
  CHECK_HANDLE(h);
  if (!(eta > 0.0) || !(omega > 0.0)) return fail(PDLP_ERR_INVALID, "eta and omega must be positive");
  Scalars s;
  int rc = DISPATCH(h, t_get_scalars, &s);
  if (rc) return rc;
  s.eta = eta; s.omega = omega; s.theta = theta; s.adaptive = adaptive ? 1 : 0;
  return DISPATCH(h, t_put_scalars, &s);
}

extern "C" int pdlp_set_options(pdlp_handle h, double tol, int primal_update, int infeasibility_detect,
                                double infeas_tol) {
  CHECK_HANDLE(h);
  Scalars s;
  int rc = DISPATCH(h, t_get_scalars, &s);
  if (rc) return rc;
  s.tol = tol; s.primal_update = primal_update ? 1 : 0; s.infeas_detect = infeasibility_detect ? 1 : 0;
  s.infeas_tol = infeas_tol;
  h->infeas = infeasibility_detect != 0;
  return DISPATCH(h, t_put_scalars, &s);
}

extern "C" int pdlp_get_state(pdlp_handle h, pdlp_state* o) {
  CHECK_HANDLE(h);
  if (!o) return fail(PDLP_ERR_INVALID, "host_out is null");
  Scalars s;
  int rc = DISPATCH(h, t_get_scalars, &s);
  if (rc) return rc;
  o->eta = s.eta; o->omega = s.omega; o->theta = s.theta; o->eta_sum = s.eta_sum; o->last_weight = s.last_weight;
  o->dx2 = s.dx2; o->dy2 = s.dy2; o->inter = s.inter; o->kkt_first = s.kkt_first; o->q_norm = s.q_norm;
  o->c_norm = s.c_norm; o->k = s.k; o->t = s.t; o->n_restarts = s.n_restarts; o->infeas = s.infeas;
  o->infeas_k = s.infeas_k; o->solved = s.solved; o->reserved = 0;
  return 0;
}

extern "C" int pdlp_set_counters(pdlp_handle h, int64_t k, int32_t t) {
  CHECK_HANDLE(h);
  Scalars s;
  int rc = DISPATCH(h, t_get_scalars, &s);
  if (rc) return rc;
  s.k = k; s.t = t;
  return DISPATCH(h, t_put_scalars, &s);
}

extern "C" int pdlp_iterate(pdlp_handle h, int n_iters, int save_previous, int use_graph) {
  CHECK_HANDLE(h);
  if (n_iters < 0) return fail(PDLP_ERR_INVALID, "n_iters < 0");
  return DISPATCH(h, engine_iterate, n_iters, save_previous, use_graph);
}

extern "C" int pdlp_time_kernels(pdlp_handle h, int n_iters, float* host_primal_ms, float* host_dual_ms) {
  CHECK_HANDLE(h);
  if (n_iters <= 0 || !host_primal_ms || !host_dual_ms) return fail(PDLP_ERR_INVALID, "bad argument");
  return DISPATCH(h, engine_time_kernels, n_iters, host_primal_ms, host_dual_ms);
}

extern "C" int pdlp_time_kernels_in_loop(pdlp_handle h, int n_iters, float* host_primal_ms, float* host_dual_ms,
                                         float* host_iteration_ms) {
  CHECK_HANDLE(h);
  if (!host_primal_ms || !host_dual_ms || !host_iteration_ms) return fail(PDLP_ERR_INVALID, "null output");
  return DISPATCH(h, engine_time_kernels_in_loop, n_iters, host_primal_ms, host_dual_ms, host_iteration_ms);
}

extern "C" int pdlp_restart_check(pdlp_handle h, int force, pdlp_check* host_out) {
  CHECK_HANDLE(h);
  return DISPATCH(h, engine_restart_check, force, host_out);
}

extern "C" int pdlp_kkt(pdlp_handle h, const void* x, const void* y, double omega, pdlp_residuals* host_out) {
  CHECK_HANDLE(h);
  if (!x || !y || !host_out) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, engine_kkt, x, y, omega, host_out);
}

extern "C" int pdlp_power_iteration(pdlp_handle h, const void* b0, int iters, double* host_sigma) {
  CHECK_HANDLE(h);
  if (!b0 || !host_sigma || iters < 0) return fail(PDLP_ERR_INVALID, "bad argument");
  return DISPATCH(h, engine_power_iteration, b0, iters, host_sigma);
}

template <typename T>
static int t_spmv(EngineT<T>* E, int transpose, const void* in, void* out) {
  SpmvOp<T, false> f{(const T*)in, (T*)out, E->S, 6};
  CU(launch_stream(transpose ? E->Kt : E->K, f, E->base->stream));
  return 0;
}
extern "C" int pdlp_spmv(pdlp_handle h, int transpose, const void* in, void* out) {
  CHECK_HANDLE(h);
  if (!in || !out) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, t_spmv, transpose, in, out);
}

extern "C" int pdlp_ruiz(pdlp_handle h, int max_iter, double eps, int* host_rounds) {
  CHECK_HANDLE(h);
  return DISPATCH(h, engine_ruiz, max_iter, eps, host_rounds);
}
extern "C" int pdlp_pock_chambolle(pdlp_handle h) {
  CHECK_HANDLE(h);
  return DISPATCH(h, engine_pock_chambolle);
}

template <typename T>
static int t_copy_out(EngineT<T>* E, void* dst, const T* src, int64_t len) {
  if (!dst) return 0;
  CU(cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)len, cudaMemcpyDeviceToDevice, E->base->stream));
  return 0;
}
extern "C" int pdlp_get_scaling(pdlp_handle h, void* D_row, void* D_col) {
  CHECK_HANDLE(h);
  int rc = h->dtype == PDLP_F64 ? t_copy_out<double>((EngineT<double>*)h->impl, D_row, ((EngineT<double>*)h->impl)->drow, h->m)
                                : t_copy_out<float>((EngineT<float>*)h->impl, D_row, ((EngineT<float>*)h->impl)->drow, h->m);
  if (rc) return rc;
  return h->dtype == PDLP_F64 ? t_copy_out<double>((EngineT<double>*)h->impl, D_col, ((EngineT<double>*)h->impl)->dcol, h->n)
                              : t_copy_out<float>((EngineT<float>*)h->impl, D_col, ((EngineT<float>*)h->impl)->dcol, h->n);
}

template <typename T>
static int t_set_scaling(EngineT<T>* E, const void* drow, const void* dcol) {
  pdlp_engine* B = E->base;
  drop_check_graphs<T>(E);
  if (!drow && !dcol) {       // back to the unscaled termination test
    fill_kernel<T><<<ew_blocks(B->m, B->sms), 256, 0, B->stream>>>(E->drow, T(1), (int)B->m);
    fill_kernel<T><<<ew_blocks(B->n, B->sms), 256, 0, B->stream>>>(E->dcol, T(1), (int)B->n);
    CU(cudaGetLastError());
    B->scaled = false;
    return 0;
  }
  CU(cudaMemcpyAsync(E->drow, drow, sizeof(T) * (size_t)B->m, cudaMemcpyDeviceToDevice, B->stream));
  CU(cudaMemcpyAsync(E->dcol, dcol, sizeof(T) * (size_t)B->n, cudaMemcpyDeviceToDevice, B->stream));
  B->scaled = true;
  return 0;
}
extern "C" int pdlp_set_scaling(pdlp_handle h, const void* D_row, const void* D_col) {
  CHECK_HANDLE(h);
  if ((D_row == nullptr) != (D_col == nullptr)) return fail(PDLP_ERR_INVALID, "D_row and D_col must both be given or both be null");
  return DISPATCH(h, t_set_scaling, D_row, D_col);
}

template <typename T>
static int t_get_problem(EngineT<T>* E, void* c, void* q, void* l, void* u) {
  int rc = t_copy_out<T>(E, c, E->c, E->base->n);
  if (!rc) rc = t_copy_out<T>(E, q, E->q, E->base->m);
  if (!rc) rc = t_copy_out<T>(E, l, E->l, E->base->n);
  if (!rc) rc = t_copy_out<T>(E, u, E->u, E->base->n);
  return rc;
}
extern "C" int pdlp_get_problem(pdlp_handle h, void* c, void* q, void* l, void* u) {
  CHECK_HANDLE(h);
  return DISPATCH(h, t_get_problem, c, q, l, u);
}

template <typename T>
static int t_get_matrix_values(EngineT<T>* E, void* vals) {
  return t_copy_out<T>(E, vals, E->K.vals, E->base->nnz);
}
extern "C" int pdlp_get_matrix_values(pdlp_handle h, void* K_vals) {
  CHECK_HANDLE(h);
  return DISPATCH(h, t_get_matrix_values, K_vals);
}

extern "C" int pdlp_batch_spmv(pdlp_handle h, const void* X8, void* Y8) {
  CHECK_HANDLE(h);
  if (!X8 || !Y8) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, batch_spmv, X8, Y8);
}
extern "C" int pdlp_batch_step(pdlp_handle h, void* X8, void* XB8, void* Y8, double eta, double omega) {
  CHECK_HANDLE(h);
  if (!X8 || !XB8 || !Y8) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, batch_step, X8, XB8, Y8, eta, omega);
}
extern "C" int pdlp_batch_gap(pdlp_handle h, const void* X8, const void* Y8, double* host_out24) {
  CHECK_HANDLE(h);
  if (!X8 || !Y8 || !host_out24) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, batch_gap, X8, Y8, host_out24);
}

extern "C" int pdlp_shard_enable(pdlp_handle h, double q_norm_global, double c_norm) {
  CHECK_HANDLE(h);
  if (h->dtype != PDLP_F64) return fail(PDLP_ERR_INVALID, "the sharded engine is fp64 only");
  Scalars s;
  int rc = DISPATCH(h, t_get_scalars, &s);
  if (rc) return rc;
  s.sharded = 1; s.rule_pending = 0; s.q_norm = q_norm_global; s.c_norm = c_norm;
  return DISPATCH(h, t_put_scalars, &s);
}
extern "C" int pdlp_shard_kty(pdlp_handle h, double* buf) {
  CHECK_HANDLE(h);
  if (!buf) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, shard_kty, buf);
}
extern "C" int pdlp_shard_step(pdlp_handle h, const double* buf, int save_previous) {
  CHECK_HANDLE(h);
  if (!buf) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, shard_step, buf, save_previous);
}
extern "C" int pdlp_shard_exchange_create(pdlp_handle h, int world, int rank, unsigned char* handle_out) {
  CHECK_HANDLE(h);
  if (!handle_out) return fail(PDLP_ERR_INVALID, "null argument");
  if (h->dtype != PDLP_F64) return fail(PDLP_ERR_INVALID, "the sharded engine is fp64 only");
  cudaSetDevice(h->device);
  return xchg_create(h, world, rank, handle_out);
}
extern "C" int pdlp_shard_exchange_connect(pdlp_handle h, const unsigned char* handles) {
  CHECK_HANDLE(h);
  if (!handles) return fail(PDLP_ERR_INVALID, "null argument");
  cudaSetDevice(h->device);
  return DISPATCH(h, xchg_connect, handles);
}
extern "C" int pdlp_shard_exchange_disconnect(pdlp_handle h) {
  CHECK_HANDLE(h);
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
  xchg_disconnect(h);
  return 0;
}
extern "C" int pdlp_shard_iterate_p2p(pdlp_handle h, int n_iters, int save_previous) {
  CHECK_HANDLE(h);
  if (n_iters < 0) return fail(PDLP_ERR_INVALID, "n_iters < 0");
  return DISPATCH(h, shard_iterate_p2p, n_iters, save_previous);
}
extern "C" int pdlp_shard_allreduce_p2p(pdlp_handle h, double* buf, int64_t count) {
  CHECK_HANDLE(h);
  if (!buf || count < 0) return fail(PDLP_ERR_INVALID, "bad argument");
  return DISPATCH(h, shard_allreduce_p2p, buf, count);
}
extern "C" int pdlp_shard_rule_pack(pdlp_handle h, double* buf4) {
  CHECK_HANDLE(h);
  if (!buf4) return fail(PDLP_ERR_INVALID, "null argument");
  Scalars* S = h->dtype == PDLP_F64 ? ((EngineT<double>*)h->impl)->S : ((EngineT<float>*)h->impl)->S;
  shard_tail_pack_kernel<<<1, 1, 0, h->stream>>>(S, buf4);
  CU(cudaGetLastError());
  return 0;
}
extern "C" int pdlp_shard_rule_apply(pdlp_handle h, const double* buf4) {
  CHECK_HANDLE(h);
  if (!buf4) return fail(PDLP_ERR_INVALID, "null argument");
  Scalars* S = h->dtype == PDLP_F64 ? ((EngineT<double>*)h->impl)->S : ((EngineT<float>*)h->impl)->S;
  shard_rule_kernel<<<1, 1, 0, h->stream>>>(S, buf4);
  CU(cudaGetLastError());
  return 0;
}
extern "C" int pdlp_shard_check_begin(pdlp_handle h, double* buf3, int have_previous) {
  CHECK_HANDLE(h);
  if (!buf3) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, shard_check_begin, buf3, have_previous);
}
extern "C" int pdlp_shard_check_mid(pdlp_handle h, const double* buf3, double* buf9) {
  CHECK_HANDLE(h);
  if (!buf3 || !buf9) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, shard_check_mid, buf3, buf9);
}
extern "C" int pdlp_shard_check_decide(pdlp_handle h, const double* buf9, int force, double* buf1) {
  CHECK_HANDLE(h);
  if (!buf9 || !buf1) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, shard_check_decide, buf9, force, buf1);
}
extern "C" int pdlp_shard_check_end(pdlp_handle h, const double* buf1, pdlp_check* host_out) {
  CHECK_HANDLE(h);
  if (!buf1) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, shard_check_end, buf1, host_out);
}

extern "C" int pdlp_primal_objective(pdlp_handle h, double* host_out) {
  CHECK_HANDLE(h);
  if (!host_out) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, engine_objective, host_out);
}

extern "C" int pdlp_solve(pdlp_handle h, const pdlp_solve_options* opts, pdlp_solve_result* host_out,
                          pdlp_check* host_trace) {
  CHECK_HANDLE(h);
  if (!opts || !host_out) return fail(PDLP_ERR_INVALID, "null argument");
  return DISPATCH(h, engine_solve, opts, host_out, host_trace);
}
