// inflate_propagate.cuh -- inflation mode 1: the reference's nearest-source PROPAGATION, level-synchronous.
//
// InflationLayer::updateCosts (costmap_2d/plugins/inflation_layer.cpp:172-266) pops (cell, source) entries from a
// priority queue ordered by the cached distance cell <-> source, marks the cell seen, writes its cost and offers the
// SAME source to the four neighbours (enqueue, :277-293: not seen, cached distance <= cell_inflation_radius_).  A cell
// therefore gets the cost of the nearest source that REACHES it through cells that source owns -- not always its
// nearest source (mode 0 computes that) -- and where several queue entries have equal distance the winner depends on
// libstdc++'s heap history.
//
// This kernel executes the same loop with one explicit, order-independent choice for equal distances:
//   * per cell a 32-bit state = rank (14 bits) | direction (2) | carried source as an offset code (16).  rank 0 =
//     popped ("seen"), all ones = no pending entry, otherwise the dense RANK of the cached distance of the cell's best
//     pending entry (ranks are built on the host from the reference's hypot table: equal doubles <=> equal ranks;
//     rank 1 = distance 0) and the direction the entry came from (0: from the -x neighbour, 1: -y, 2: +x, 3: +y);
//   * a round pops ALL cells whose rank equals the smallest pending rank k of the whole grid.  Entries pushed by
//     those pops can never have rank k themselves (dx^2 + dy^2 changes parity between 4-neighbours), so the round is
//     a legal sequence of reference pops;
//   * a pop offers its source to the four neighbours with one atomicMin on the packed state: smallest cached
//     distance wins, first of -x, -y, +x, +y on equal distance, independent of the order in which the pops of one
//     round (or of several rounds) arrive.  Neighbours that are popped, or popping in this very round, are skipped
//     (the reference's `seen_` test at enqueue time).
// The CPU checker states exactly this (oracle variant 5) and certifies, by replaying the reference's sequential
// priority-queue loop with ties resolved towards this kernel's sources (variant 6), that the result is one the
// reference's own code produces under a legal heap order.  Wherever the heap order cannot matter the result IS the
// reference's (tests/test_oracle_tie_variants.py).
//
// Shape: ONE persistent cooperative kernel, one grid-wide barrier per round (about 50 rounds for R = 11, 145 for
// R = 20).  The cells with a pending entry -- the frontier, about the perimeter of what has been inflated so far --
// live in a compact list that every round re-emits (cells that pop drop out, cells that receive their first entry
// join), spread evenly over all threads of the grid; list appends are staged per CTA in shared memory so that a round
// costs a handful of global atomics per CTA.  A round is a chain of about six dependent L2 round trips plus the
// barrier: the kernel is latency-bound by construction (rounds x ~8 us), its traffic is a few MB per round.
#pragma once

#include "costmap_kernels.cuh"

namespace navgpu {

constexpr int kPT = 64;         // init phase: tile edge (cells)
constexpr int kPThreads = 512;
constexpr int kPBuf = 6144;     // per-CTA staging of list appends (entries); a pass adds at most 4 per thread
constexpr uint32_t kPNone = 0xffffffffu;  // unseen, no pending entry
constexpr uint32_t kPInf = 0x3fffu;       // rank field of kPNone
constexpr int kPMaxRank = 0x3ffe;

struct PropCtl {
  unsigned count, gen;  // grid barrier
  unsigned kmin[4];     // smallest pending rank of round r in kmin[r & 3]
  unsigned n_list[3];   // entries of the frontier list of round r in n_list[r % 3]
  unsigned rounds;      // rounds of the last run (diagnostics)
  unsigned pad_[2];
};

struct PropArgs {
  uint8_t* master;
  unsigned sx, sy, pitch;
  const DevWindow* win;
  int R;
  const uint32_t* seeds;  // k_merge_seed's bitmask as 32-bit words: cell x of row y = word y * sp32 + 1 + (x >> 5), bit x & 31
  uint32_t* state;        // sy x pitch
  uint32_t* list[2];      // frontier lists (cell offsets y * pitch + x), sy x pitch entries each
  const uint16_t* rank;   // (R+2)^2, [dx][dy]: rank of cached_distances_[dx][dy], 0xffff where it exceeds R
  const uint8_t* cost;    // (R+2)^2, [dx][dy]: cached_costs_
  PropCtl* ctl;
};

__device__ __forceinline__ void prop_grid_barrier(PropCtl* ctl, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0 && nblocks > 1) {
    volatile unsigned* gen = &ctl->gen;
    const unsigned g = *gen;
    __threadfence();
    if (atomicAdd(&ctl->count, 1u) == nblocks - 1) {
      ctl->count = 0;
      __threadfence();
      atomicAdd(&ctl->gen, 1u);
    } else {
      while (*gen == g) {
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// CTA-staged append to a global list: threads add to a shared buffer; prop_flush reserves a range with one atomicAdd
struct PropAppender {
  uint32_t* buf;
  unsigned* n;
  __device__ __forceinline__ void add(uint32_t v) const { buf[atomicAdd(n, 1u)] = v; }
};
__device__ __forceinline__ void prop_flush(PropAppender ap, uint32_t* list, unsigned* count, unsigned* base_smem) {
  // called by all threads of the CTA
  __syncthreads();
  const unsigned m = *ap.n;
  if (m == 0) return;  // uniform
  if (threadIdx.x == 0) *base_smem = atomicAdd(count, m);
  __syncthreads();
  const unsigned base = *base_smem;
  for (unsigned i = threadIdx.x; i < m; i += kPThreads) list[base + i] = ap.buf[i];
  __syncthreads();
  if (threadIdx.x == 0) *ap.n = 0;
  __syncthreads();
}

__global__ void __launch_bounds__(kPThreads, 2) k_inflate_propagate(PropArgs a) {
  extern __shared__ __align__(16) uint8_t prop_smem[];
  const int n = a.R + 2;
  uint32_t* const buf = reinterpret_cast<uint32_t*>(prop_smem);          // kPBuf staged appends
  uint16_t* const rank = reinterpret_cast<uint16_t*>(buf + kPBuf);       // n * n
  uint8_t* const cost = reinterpret_cast<uint8_t*>(rank + ((n * n + 1) & ~1));  // n * n
  __shared__ unsigned buf_n, buf_base, cta_min;

  const int tid = threadIdx.x;
  const DevWindow w = *a.win;
  if (!w.valid) return;  // uniform over the grid: nobody reaches a barrier
  for (int i = tid; i < n * n; i += kPThreads) {
    rank[i] = a.rank[i];
    cost[i] = a.cost[i];
  }
  if (tid == 0) {
    buf_n = 0;
    cta_min = kPInf;
  }
  __syncthreads();
  const PropAppender ap{buf, &buf_n};
  const int R = a.R;
  // tiles that inflation can touch: window +- 2R (seeds sit within window +- R, each reaches R further)
  const int tx_lo = max(0, w.x0 - 2 * R) / kPT, tx_hi = (min((int)a.sx, w.xn + 2 * R) - 1) / kPT;
  const int ty_lo = max(0, w.y0 - 2 * R) / kPT, ty_hi = (min((int)a.sy, w.yn + 2 * R) - 1) / kPT;
  const int rtx = tx_hi - tx_lo + 1, rty = ty_hi - ty_lo + 1, nt = rtx * rty;
  const int sx0 = max(0, w.x0 - R), sxn = min((int)a.sx, w.xn + R);  // seed region (:203-211)
  const int sy0 = max(0, w.y0 - R), syn = min((int)a.sy, w.yn + R);
  const unsigned sp32 = seed_pitch16(a.pitch) / 2;
  const int pitch = (int)a.pitch;

  // ---- initial states.  A seed enters the queue with distance 0 and itself as source (:213-224); its pop leaves the
  // LETHAL value as it is and offers the seed to its neighbours -- which, for a seed whose four neighbours are seeds
  // as well, are all seen or popping: such a seed is marked popped right away and never enters the list.
  for (int rt = blockIdx.x; rt < nt; rt += gridDim.x) {
    const int tx = tx_lo + rt % rtx, ty = ty_lo + rt / rtx;
    for (int i = tid; i < kPT * kPT; i += kPThreads) {
      const int x = tx * kPT + (i & (kPT - 1)), y = ty * kPT + (i >> 6);
      if (x >= (int)a.sx || y >= (int)a.sy) continue;
      auto seed = [&](int xx, int yy) -> bool {
        if (xx < sx0 || xx >= sxn || yy < sy0 || yy >= syn) return false;
        return (a.seeds[(size_t)yy * sp32 + 1 + (xx >> 5)] >> (xx & 31)) & 1u;
      };
      uint32_t st = kPNone;
      if (seed(x, y)) {
        // a map edge counts as "nothing to offer to" just like a seed neighbour
        const bool inner = (x == 0 || seed(x - 1, y)) && (y == 0 || seed(x, y - 1)) &&
                           (x + 1 >= (int)a.sx || seed(x + 1, y)) && (y + 1 >= (int)a.sy || seed(x, y + 1));
        st = inner ? 0x8080u : ((1u << 18) | 0x8080u);
        if (!inner) ap.add((uint32_t)(y * pitch + x));
      }
      a.state[(size_t)y * pitch + x] = st;
    }
    prop_flush(ap, a.list[0], &a.ctl->n_list[0], &buf_base);
  }
  prop_grid_barrier(a.ctl, gridDim.x);

  const unsigned total_threads = gridDim.x * kPThreads;
  unsigned round = 0;
  unsigned k = __ldcg(&a.ctl->n_list[0]) ? 1u : kPInf;  // every listed seed pops in round 0
  for (;; ++round) {
    if (k >= kPInf) break;
    const unsigned n_in = __ldcg(&a.ctl->n_list[round % 3]);
    const uint32_t* lin = a.list[round & 1];
    uint32_t* lout = a.list[(round + 1) & 1];
    unsigned* n_out = &a.ctl->n_list[(round + 1) % 3];
    if (blockIdx.x == 0 && tid == 0) {
      a.ctl->kmin[(round + 2) & 3] = kPInf;
      a.ctl->n_list[(round + 2) % 3] = 0;
    }
    unsigned my_min = kPInf;
    for (unsigned base = blockIdx.x * kPThreads; base < n_in; base += total_threads) {
      const unsigned i = base + tid;
      if (i < n_in) {
        const uint32_t c = __ldcg(lin + i);
        const uint32_t st = __ldcg(a.state + c);
        const unsigned rk = st >> 18;
        if (rk != k) {  // stays pending
          ap.add(c);
          my_min = min(my_min, rk);
        } else {        // pops now: seen, cost of the carried source (:236-254), source offered to the neighbours
          const uint32_t code = st & 0xffffu;
          a.state[c] = code;
          const int ox = (int)(code >> 8) - 128, oy = (int)(code & 0xff) - 128;  // source - cell
          {
            const uint8_t old = a.master[c], nw = inflate_combine(old, cost[abs(ox) * n + abs(oy)]);
            if (nw != old) a.master[c] = nw;
          }
          const int y = (int)(c / a.pitch), x = (int)(c - (unsigned)y * a.pitch);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            // q = the direction this cell lies in as seen from the receiver: receiver = cell + (1,0), (0,1), (-1,0), (0,-1)
            const int rx = x + (q == 0 ? 1 : q == 2 ? -1 : 0), ry = y + (q == 1 ? 1 : q == 3 ? -1 : 0);
            if (rx < 0 || ry < 0 || rx >= (int)a.sx || ry >= (int)a.sy) continue;
            const uint32_t rc = (uint32_t)(ry * pitch + rx);
            const uint32_t rst = __ldcg(a.state + rc);
            const unsigned rr = rst >> 18;
            if (rr == 0 || rr == k) continue;  // seen, or popping in this round
            const int sx_ = ox - (rx - x), sy_ = oy - (ry - y);  // source - receiver
            const int dx = abs(sx_), dy = abs(sy_);
            const unsigned r = rank[dx * n + dy];
            if (r == 0xffffu) continue;  // cached distance beyond cell_inflation_radius_ (:284-287)
            const uint32_t cand = (r << 18) | ((unsigned)q << 16) | ((unsigned)(sx_ + 128) << 8) | (unsigned)(sy_ + 128);
            if (cand < rst) {
              const uint32_t old = atomicMin(a.state + rc, cand);
              if (old == kPNone) ap.add(rc);  // first entry of this cell: it joins the frontier
              my_min = min(my_min, r);
            }
          }
        }
      }
      __syncthreads();
      const bool nearly_full = buf_n > kPBuf - 4 * kPThreads;  // uniform: read between two barriers
      __syncthreads();
      if (nearly_full) prop_flush(ap, lout, n_out, &buf_base);
    }
    prop_flush(ap, lout, n_out, &buf_base);
    my_min = __reduce_min_sync(0xffffffffu, my_min);
    if ((tid & 31) == 0 && my_min != kPInf) atomicMin(&cta_min, my_min);
    __syncthreads();
    if (tid == 0) {
      if (cta_min != kPInf) atomicMin(&a.ctl->kmin[(round + 1) & 3], cta_min);
      cta_min = kPInf;
    }
    prop_grid_barrier(a.ctl, gridDim.x);
    k = __ldcg(&a.ctl->kmin[(round + 1) & 3]);
  }
  if (blockIdx.x == 0 && tid == 0) a.ctl->rounds = round;
}

inline size_t propagate_smem(int R) {
  const size_t n = R + 2;
  return size_t(kPBuf) * 4 + ((n * n + 1) & ~size_t(1)) * 2 + n * n + 16;
}

}  // namespace navgpu
