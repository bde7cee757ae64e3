// graph-embed_b200 :: ball radii + rescale between levels on the device.
//
// Replaces the host loop of partition::embedMultilevel, /root/reference/src/embed.cpp:615-778
// (duplicated at :166-329): the radius r_A of every vertex of level l+1 -- the ball inside which
// its members will be placed at level l -- and the shrink of that level's coordinates / radii into
// the balls of level l+2.
//
// The reference keeps, per family (= the members of one aggregate of level l+2, or every vertex in
// the base case), a vector of events (t, i, j) "balls i and j touch", t = -|xi - xj| / 2, sorts it,
// pops the back, freezes the endpoint(s) that are still growing at radius -t, re-keys the events
// that touch a ball that just froze (t' = -(2(-t) - radius): the partner covers the rest alone) and
// sorts again.  Here: one CTA per family, the events in global memory, and per pop ONE pass of the
// CTA over the family's live events that applies the pending re-keys, drops events whose two ends
// are frozen, and finds the next event to pop as the lexicographic maximum of (t, i, j) -- exactly
// the element std::sort puts at the back.  Same arithmetic (IEEE add / mul / sqrt, no FMA
// contraction), same tie-breaking, same `r <= 0 means still growing` test as the host restatement
// ge_level_radii, which stays as the checker: the results are bit-identical.
//
// The coordinates never leave the device: ge_embed chains  solve(l+1) -> radii -> solve(l).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ge_context.h"

namespace ge {

namespace {

__device__ __forceinline__ double dist_rn(const double* a, const double* b, int dim) {
  double sum = 0.0;
  for (int k = 0; k < dim; ++k) {
    const double d = __dsub_rn(b[k], a[k]);
    sum = __dadd_rn(sum, __dmul_rn(d, d));
  }
  return __dsqrt_rn(sum);
}

// ---- event lists -------------------------------------------------------------------------------
// Base case (:616-634): every pair i < j of the m vertices.
__global__ void k_radii_events_base(int m, int dim, const double* __restrict__ x,
                                    double* __restrict__ ev_t, int2* __restrict__ ev_ij) {
  const int i = blockIdx.x;
  const long long off = (long long)i * (2LL * m - i - 1) / 2;
  for (int j = i + 1 + threadIdx.x; j < m; j += blockDim.x) {
    const long long e = off + (j - i - 1);
    ev_t[e] = -dist_rn(x + (size_t)i * dim, x + (size_t)j * dim, dim) / 2;
    ev_ij[e] = make_int2(i, j);
  }
}

// General case (:686-706): the edges a < j of A_c whose two ends have the same parent.  One warp
// per member position c of P_T_c (family-major order, so a family's events are contiguous).
template <bool FILL>
__global__ void __launch_bounds__(256) k_radii_events(int m, int dim, const int* __restrict__ I,
                                                      const int* __restrict__ J,
                                                      const int* __restrict__ parent,
                                                      const int* __restrict__ PJ,
                                                      const double* __restrict__ x,
                                                      int* __restrict__ count,
                                                      const int* __restrict__ ev_off,
                                                      double* __restrict__ ev_t,
                                                      int2* __restrict__ ev_ij) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= m) return;
  const int a = PJ[c];
  const int pa = parent[a];
  const int rb = I[a], re = I[a + 1];
  int cnt = 0;
  const int base = FILL ? ev_off[c] : 0;
  for (int e0 = rb; e0 < re; e0 += 32) {
    const int e = e0 + lane;
    const int j = e < re ? J[e] : -1;
    const bool take = j > a && parent[j] == pa;
    const unsigned mask = __ballot_sync(0xffffffffu, take);
    if (FILL && take) {
      const int dst = base + cnt + __popc(mask & ((1u << lane) - 1u));
      ev_t[dst] = -dist_rn(x + (size_t)a * dim, x + (size_t)j * dim, dim) / 2;
      ev_ij[dst] = make_int2(a, j);
    }
    cnt += __popc(mask);
  }
  if (!FILL && lane == 0) count[c] = cnt;
}

// ---- exclusive scan (three small kernels; m is at most a few hundred thousand) -----------------
constexpr int kScanBlock = 1024;
__device__ __forceinline__ int block_scan_excl(int v, int* scratch, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
  for (int off = 1; off < 32; off <<= 1) {
    const int o = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += o;
  }
  if (lane == 31) scratch[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (int)(blockDim.x >> 5) ? scratch[lane] : 0;
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, w, off);
      if (lane >= off) w += o;
    }
    scratch[lane] = w;
  }
  __syncthreads();
  const int before = warp > 0 ? scratch[warp - 1] : 0;
  total = scratch[(blockDim.x >> 5) - 1];
  __syncthreads();
  return before + inc - v;
}
__global__ void __launch_bounds__(kScanBlock) k_rscan_blocks(const int* __restrict__ in, int n,
                                                            int* __restrict__ out, int* __restrict__ sums) {
  __shared__ int scratch[32];
  const int i = blockIdx.x * kScanBlock + threadIdx.x;
  const int v = i < n ? in[i] : 0;
  int total;
  const int ex = block_scan_excl(v, scratch, total);
  if (i < n) out[i] = ex;
  if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
__global__ void __launch_bounds__(kScanBlock) k_rscan_sums(int* sums, int nb) {  // one CTA
  __shared__ int scratch[32];
  int carry = 0;
  for (int b0 = 0; b0 < nb; b0 += kScanBlock) {
    const int i = b0 + threadIdx.x;
    const int v = i < nb ? sums[i] : 0;
    int total;
    const int ex = block_scan_excl(v, scratch, total);
    if (i < nb) sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) sums[nb] = carry;  // grand total
}
__global__ void __launch_bounds__(kScanBlock) k_rscan_add(int* __restrict__ out, int n,
                                                         const int* __restrict__ sums, int nb) {
  const int i = blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n) out[i] += sums[blockIdx.x];
  if (i == 0) out[n] = sums[nb];
}

// ---- the event loop ----------------------------------------------------------------------------
struct GrowArgs {
  int dim;
  int general;             // 0: base case, one family = all vertices, no shrink
  int m_limit;             // the reference's loop bound `count < m` (m = vertices of the level)
  int nfam;
  const int* PI;           // general: families = rows of P_T_c
  const int* PJ;
  const int* ev_off;       // general: per member position; base: {0, E}
  double* ev_t;
  int2* ev_ij;             // .x < 0: served or dead
  double* x;               // [m][dim] in/out (shrunk in the general case)
  double* r;               // [m] zero on entry
  const double* xc;        // [mc][dim]
  const double* rc;        // [mc]
  long long E_base;
  int class_split;         // families with more events than this run in the wide kernel
  int batched;             // wide kernel: serve vertex-disjoint runs of events together (GE_RADII_BATCH=0: off)
  int smem_events;         // wide kernel: stage a family's events in shared memory when they fit
};
constexpr int kRadiiSmemEvents = 8192;  // 16 bytes each: 128 KB of dynamic shared memory

__device__ __forceinline__ bool key_greater(double t, int i, int j, double ot, int oi, int oj) {
  return t > ot || (t == ot && (i > oi || (i == oi && j > oj)));
}

// Batched event loop for large families.  The reference pops one event at a time; but an event
// that shares no vertex with the events popped just before it is untouched by their re-keys (which
// only ever move keys DOWN: t' = -(2 reach_old - r) <= t because reach_old >= r for an event that
// has not been popped yet), so the sorted prefix of pairwise vertex-disjoint events can be served
// together.  Per batch: one pass over the family's events that applies the re-keys owed to balls
// frozen since the event's key was last written (the key of an event with one frozen end is a
// function of its length and that end's radius alone), then the exact top of the order -- every
// event not below the 16th best of the per-warp maxima, at most 64 -- is sorted, cut at the first
// vertex conflict and served in parallel.  Same pops, same order-dependent arithmetic, ~10x fewer
// passes.  Families with a zero-length event (coincident points: a ball of radius 0 counts as
// still growing in the reference, :640-642) return false untouched and take the one-pop-per-pass
// loop below.
template <int THREADS>
__device__ int grow_batched(const GrowArgs& g, double* evt, int2* evij, const long long e0, const long long e1) {
  constexpr int NW = THREADS / 32, CAP = 64, K = 16;
  __shared__ double w_t[NW];
  __shared__ int w_i[NW], w_j[NW], w_ok[NW];
  __shared__ double l_t[CAP], s_t[CAP];
  __shared__ int l_i[CAP], l_j[CAP], s_i[CAP], s_j[CAP];
  __shared__ long long l_e[CAP], s_e[CAP];
  __shared__ double tau_t;
  __shared__ int tau_i, tau_j, cnt, firstconf, degenerate, nvalid;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ int n_batches, n_pops;
  if (tid == 0) {
    degenerate = 0;
    n_batches = 0;
    n_pops = 0;
  }
  __syncthreads();
  for (bool first = true;; first = false) {
    // Families where nearly every event touches one hub (stars) give batches of one: after 16
    // batches averaging fewer than four pops the owed re-keys are applied once more and the
    // one-pop-per-pass loop takes over (return 2).
    const bool hand_over = n_batches >= 16 && n_pops < 4 * n_batches;  // (a batched pass costs ~3 one-pop passes)
    // ---- pass 1: owed re-keys, dead events, per-thread best --------------------------------------
    double bt = 0.0;
    int bi = -1, bj = -1;
    bool have = false, zero = false;
    for (long long e = e0 + tid; e < e1; e += THREADS) {
      const int2 ij = evij[e];
      if (ij.x < 0) continue;
      const bool flag = ij.y < 0;  // key already carries one frozen end
      const int i = ij.x, j = flag ? ~ij.y : ij.y;
      double t = evt[e];
      const double ri = __ldcg(&g.r[i]), rj = __ldcg(&g.r[j]);
      const bool fi = ri > 0.0, fj = rj > 0.0;
      if (fi && fj) {
        evij[e] = make_int2(-1, -1);
        continue;
      }
      if ((fi || fj) && !flag) {
        t = -__dsub_rn(__dmul_rn(2.0, -t), fi ? ri : rj);
        evt[e] = t;
        evij[e] = make_int2(i, ~j);
      }
      zero |= t == 0.0;
      if (!have || key_greater(t, i, j, bt, bi, bj)) {
        bt = t;
        bi = i;
        bj = j;
        have = true;
      }
    }
    if (first && zero) degenerate = 1;  // (benign race: every writer stores 1)
    double wt = bt;
    int wi = bi, wj = bj, wh = have ? 1 : 0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ot = __shfl_xor_sync(0xffffffffu, wt, off);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, off);
      const int oj = __shfl_xor_sync(0xffffffffu, wj, off);
      const int oh = __shfl_xor_sync(0xffffffffu, wh, off);
      if (oh && (!wh || key_greater(ot, oi, oj, wt, wi, wj))) {
        wt = ot;
        wi = oi;
        wj = oj;
        wh = 1;
      }
    }
    if (lane == 0) {
      w_t[warp] = wt;
      w_i[warp] = wi;
      w_j[warp] = wj;
      w_ok[warp] = wh;
    }
    if (tid == 0) {
      cnt = 0;
      firstconf = CAP;
    }
    __syncthreads();
    if (first && degenerate) return 0;  // nothing was modified: no ball is frozen yet
    if (hand_over) return 2;            // every owed re-key has just been applied
    // ---- threshold: the K-th best of the per-warp maxima ---------------------------------------
    if (warp == 0) {
      const bool ok = lane < NW && w_ok[lane];
      const double mt = ok ? w_t[lane] : 0.0;
      const int mi = ok ? w_i[lane] : -1, mj = ok ? w_j[lane] : -1;
      int rank = 0;
      for (int o = 0; o < NW; ++o)
        if (w_ok[o] && key_greater(w_t[o], w_i[o], w_j[o], mt, mi, mj)) ++rank;
      const int nv = __popc(__ballot_sync(0xffffffffu, ok));
      const int pick = min(K, nv) - 1;
      if (lane == 0) nvalid = nv;
      if (ok && rank == pick) {
        tau_t = mt;
        tau_i = mi;
        tau_j = mj;
      }
    }
    __syncthreads();
    if (nvalid == 0) return 1;  // no live event left
    // ---- pass 2: threads whose best reaches the threshold list their events at or above it -------
    const double tt = tau_t;
    const int ti = tau_i, tj = tau_j;
    if (have && !key_greater(tt, ti, tj, bt, bi, bj)) {
      for (long long e = e0 + tid; e < e1; e += THREADS) {
        const int2 ij = evij[e];
        if (ij.x < 0) continue;
        const int i = ij.x, j = ij.y < 0 ? ~ij.y : ij.y;
        const double t = evt[e];
        if (key_greater(tt, ti, tj, t, i, j)) continue;
        const int slot = atomicAdd(&cnt, 1);
        if (slot < CAP) {
          l_t[slot] = t;
          l_i[slot] = i;
          l_j[slot] = j;
          l_e[slot] = e;
        }
      }
    }
    __syncthreads();
    int n_list = cnt;
    if (n_list > CAP) {  // (rare) too many ties at the threshold: serve the single best event
      if (tid == 0) {
        int best = 0;
        for (int o = 1; o < NW; ++o)
          if (w_ok[o] && (!w_ok[best] || key_greater(w_t[o], w_i[o], w_j[o], w_t[best], w_i[best], w_j[best]))) best = o;
        tau_t = w_t[best];
        tau_i = w_i[best];
        tau_j = w_j[best];
        cnt = 0;
      }
      __syncthreads();
      const double t1 = tau_t;
      const int i1 = tau_i, j1 = tau_j;
      if (have && bt == t1 && bi == i1 && bj == j1) {
        for (long long e = e0 + tid; e < e1; e += THREADS) {
          const int2 ij = evij[e];
          if (ij.x != i1) continue;
          const int j = ij.y < 0 ? ~ij.y : ij.y;
          if (j != j1 || evt[e] != t1) continue;
          l_t[0] = t1;
          l_i[0] = i1;
          l_j[0] = j1;
          l_e[0] = e;
          cnt = 1;
          break;
        }
      }
      __syncthreads();
      n_list = cnt;
    }
    // ---- sort by rank (descending key), cut at the first vertex conflict, serve ------------------
    if (tid < n_list) {
      int rank = 0;
      for (int o = 0; o < n_list; ++o)
        if (key_greater(l_t[o], l_i[o], l_j[o], l_t[tid], l_i[tid], l_j[tid])) ++rank;
      s_t[rank] = l_t[tid];
      s_i[rank] = l_i[tid];
      s_j[rank] = l_j[tid];
      s_e[rank] = l_e[tid];
    }
    __syncthreads();
    if (tid < n_list) {
      const int i = s_i[tid], j = s_j[tid];
      for (int o = 0; o < tid; ++o)
        if (s_i[o] == i || s_i[o] == j || s_j[o] == i || s_j[o] == j) {
          atomicMin(&firstconf, tid);
          break;
        }
    }
    __syncthreads();
    if (tid == 0) {
      n_batches += 1;
      n_pops += min(n_list, firstconf);
    }
    if (tid < min(n_list, firstconf)) {
      const int i = s_i[tid], j = s_j[tid];
      const double reach = -s_t[tid];
      evij[s_e[tid]] = make_int2(-1, -1);  // popped
      if (g.r[i] <= 0.0) g.r[i] = reach;
      if (g.r[j] <= 0.0) g.r[j] = reach;
    }
    __syncthreads();
  }
}

template <int THREADS, bool WIDE>
__global__ void __launch_bounds__(THREADS) k_radii_grow(const GrowArgs g) {
  constexpr int NW = THREADS / 32;
  __shared__ double s_t[NW];
  __shared__ int s_i[NW], s_j[NW];
  __shared__ long long s_e[NW];
  __shared__ double sh_reach;
  __shared__ int sh_f0, sh_f1, sh_stop;
  __shared__ double s_red[NW];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int c0 = 0, c1 = 0;
  long long e0 = 0, e1 = 0;
  if (g.general) {
    c0 = g.PI[b];
    c1 = g.PI[b + 1];
    e0 = g.ev_off[c0];
    e1 = g.ev_off[c1];
  } else {
    e1 = g.E_base;
  }
  const long long E = e1 - e0;
  if ((E > g.class_split) != WIDE) return;
  // Events of one family: in global memory, or -- for the wide kernel, when they fit -- staged in
  // shared memory for the whole loop.  The hub families of power-law hierarchies are near-stars
  // (R-MAT-20: 2 698 members, 7 661 events): one pop per pass, thousands of passes, each pass a few
  // events per thread -- with the events in shared memory a pass costs a few hundred cycles
  // instead of three dependent L2 round trips.
  extern __shared__ __align__(16) unsigned char radii_smem[];
  double* evt = g.ev_t;
  int2* evij = g.ev_ij;
  bool staged = false;
  if (WIDE && E > 0 && E <= kRadiiSmemEvents && g.smem_events) {
    double* st = reinterpret_cast<double*>(radii_smem);
    int2* sij = reinterpret_cast<int2*>(st + kRadiiSmemEvents);
    for (long long e = tid; e < E; e += THREADS) {
      st[e] = g.ev_t[e0 + e];
      sij[e] = g.ev_ij[e0 + e];
    }
    __syncthreads();
    evt = st - e0;   // indexed with the same global event numbers below
    evij = sij - e0;
    staged = true;
  }
  (void)staged;
  const int s = c1 - c0;
  if (g.general && s == 0) return;

  int batched = 0;  // 0: not run / declined, 1: finished, 2: handed over to the one-pop loop
  // Batching pays on mesh-like families (consecutive events rarely share a vertex; Delaunay
  // hierarchy 27 -> 11 ms per embed).  In the hub families of power-law graphs (near-stars) nearly
  // every event touches the hub and batches have length one: grow_batched notices and hands over
  // to the one-pop-per-pass loop.
  const long long members = g.general ? s : g.m_limit;
  (void)members;
  if (WIDE && !(g.general && s == 1) && E > 0 && g.batched)
    batched = grow_batched<THREADS>(g, evt, evij, e0, e1);
  if (g.general && s == 1) {  // :687-691
    if (tid == 0) g.r[g.PJ[c0]] = g.rc[b];
  } else if (E > 0 && batched != 1) {
    int f0 = -1, f1 = -1;
    double reach = 0.0;
    long long count = 0;
    for (;;) {
      double bt = 0.0;
      int bi = -1, bj = -1;
      long long be = -1;
      for (long long e = e0 + tid; e < e1; e += THREADS) {
        int2 ij = evij[e];
        if (ij.x < 0) continue;
        if (ij.y < 0) ij.y = ~ij.y;  // (flag of the batched loop: the key already carries one frozen end)
        double t = evt[e];
        const bool hit0 = (ij.x == f0) | (ij.y == f0), hit1 = (ij.x == f1) | (ij.y == f1);
        if (hit0 | hit1) {  // a ball this event touches froze at the last pop (:655-676, :732-753)
          const int fz = hit0 ? f0 : f1;
          const int other = ij.x == fz ? ij.y : ij.x;
          t = -__dsub_rn(__dmul_rn(2.0, -t), reach);
          if (__ldcg(&g.r[other]) > 0.0) {  // both ends frozen: can never act again
            evij[e] = make_int2(-1, -1);
            continue;
          }
          evt[e] = t;
        }
        const bool better = be < 0 || t > bt || (t == bt && (ij.x > bi || (ij.x == bi && ij.y > bj)));
        if (better) {
          bt = t;
          bi = ij.x;
          bj = ij.y;
          be = e;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ot = __shfl_xor_sync(0xffffffffu, bt, off);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
        const long long oe = __shfl_xor_sync(0xffffffffu, be, off);
        const bool better = oe >= 0 && (be < 0 || ot > bt || (ot == bt && (oi > bi || (oi == bi && oj > bj))));
        if (better) {
          bt = ot;
          bi = oi;
          bj = oj;
          be = oe;
        }
      }
      if (lane == 0) {
        s_t[warp] = bt;
        s_i[warp] = bi;
        s_j[warp] = bj;
        s_e[warp] = be;
      }
      __syncthreads();
      if (warp == 0) {  // the per-warp maxima meet in warp 0 (a serial scan by one thread cost ~1000 cycles per pop)
        bt = lane < NW ? s_t[lane] : 0.0;
        bi = lane < NW ? s_i[lane] : -1;
        bj = lane < NW ? s_j[lane] : -1;
        be = lane < NW ? s_e[lane] : -1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double ot = __shfl_xor_sync(0xffffffffu, bt, off);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
          const int oj = __shfl_xor_sync(0xffffffffu, bj, off);
          const long long oe = __shfl_xor_sync(0xffffffffu, be, off);
          const bool better = oe >= 0 && (be < 0 || ot > bt || (ot == bt && (oi > bi || (oi == bi && oj > bj))));
          if (better) {
            bt = ot;
            bi = oi;
            bj = oj;
            be = oe;
          }
        }
      }
      if (tid == 0) {
        int stop = 0, nf0 = -1, nf1 = -1;
        double nreach = 0.0;
        if (be < 0) {
          stop = 1;  // no live event left
        } else {
          evij[be] = make_int2(-1, -1);  // popped
          const bool live_i = g.r[bi] <= 0.0, live_j = g.r[bj] <= 0.0;
          if (live_i || live_j) {
            nreach = -bt;
            if (live_i) {
              g.r[bi] = nreach;
              nf0 = bi;
            }
            if (live_j) {
              g.r[bj] = nreach;
              nf1 = bj;
            }
            count += (live_i ? 1 : 0) + (live_j ? 1 : 0);
            if (count >= g.m_limit) stop = 1;
          }
        }
        sh_f0 = nf0;
        sh_f1 = nf1;
        sh_reach = nreach;
        sh_stop = stop;
      }
      __syncthreads();
      f0 = sh_f0;
      f1 = sh_f1;
      reach = sh_reach;
      if (sh_stop) break;
      // (the next pass starts with global loads; the barrier above also orders thread 0's writes)
    }
  }
  if (!g.general) return;
  __syncthreads();

  // :757-777: shrink the family into the ball of its parent
  const double* cb = g.xc + (size_t)b * g.dim;
  double alpha = 0.0;
  for (int c = c0 + tid; c < c1; c += THREADS) {
    const int a = g.PJ[c];
    const double v = __dadd_rn(dist_rn(cb, g.x + (size_t)a * g.dim, g.dim), __ldcg(&g.r[a]));
    alpha = fmax(alpha, v);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) alpha = fmax(alpha, __shfl_xor_sync(0xffffffffu, alpha, off));
  if (lane == 0) s_red[warp] = alpha;
  __syncthreads();
  alpha = s_red[0];
  for (int w = 1; w < NW; ++w) alpha = fmax(alpha, s_red[w]);
  if (alpha < 0.000001) alpha = 0.000001;
  const double scale = __ddiv_rn(g.rc[b], alpha);
  for (int c = c0 + tid; c < c1; c += THREADS) {
    const int a = g.PJ[c];
    for (int k = 0; k < g.dim; ++k) {
      const double xv = g.x[(size_t)a * g.dim + k];
      g.x[(size_t)a * g.dim + k] = __dadd_rn(cb[k], __dmul_rn(scale, __dsub_rn(xv, cb[k])));
    }
    g.r[a] = __dmul_rn(scale, __ldcg(&g.r[a]));
  }
}

}  // namespace

void level_radii_device(ge_context* ctx, int m, int dim, double* d_x, double* d_r,
                        const RadiiLevel* lv, const double* d_xc, const double* d_rc) {
  if (m <= 0) return;
  cudaStream_t st = ctx->stream;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  GE_CUDA(cudaEventCreate(&ev0));
  GE_CUDA(cudaEventCreate(&ev1));
  GE_CUDA(cudaEventRecord(ev0, st));
  GE_CUDA(cudaMemsetAsync(d_r, 0, sizeof(double) * (size_t)m, st));
  GrowArgs g{};
  g.dim = dim;
  g.m_limit = m;
  g.x = d_x;
  g.r = d_r;
  g.class_split = 2048;
  {
    const char* e = std::getenv("GE_RADII_BATCH");
    g.batched = e ? std::atoi(e) : 1;
    const char* e2 = std::getenv("GE_RADII_SMEM");
    g.smem_events = e2 ? std::atoi(e2) : 1;
  }
  const size_t wide_smem = (size_t)kRadiiSmemEvents * 16;
  GE_CUDA(cudaFuncSetAttribute(k_radii_grow<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem));
  DevBuf<double> ev_t;
  DevBuf<int2> ev_ij;
  DevBuf<int> cnt, off, sums;
  if (lv == nullptr) {  // base case: all pairs
    const long long E = (long long)m * (m - 1) / 2;
    GE_REQUIRE(m <= kRadiiBaseMax, "coarsest level too large for the all-pairs radii step on the device");
    if (E > 0) {
      ev_t.alloc(ctx, (size_t)E);
      ev_ij.alloc(ctx, (size_t)E);
      k_radii_events_base<<<m, 128, 0, st>>>(m, dim, d_x, ev_t.get(), ev_ij.get());
      GE_CUDA(cudaGetLastError());
      ctx->launches++;
    }
    g.general = 0;
    g.nfam = 1;
    g.ev_t = ev_t.get();
    g.ev_ij = ev_ij.get();
    g.E_base = E;
    g.class_split = -1;  // always the wide kernel
    k_radii_grow<1024, true><<<1, 1024, wide_smem, st>>>(g);
    GE_CUDA(cudaGetLastError());
    ctx->launches++;
  } else {
    const int mc = lv->mc;
    const int nb = (m + kScanBlock - 1) / kScanBlock;
    cnt.alloc(ctx, (size_t)m);
    off.alloc(ctx, (size_t)m + 1);
    sums.alloc(ctx, (size_t)nb + 1);
    const unsigned wgrid = (unsigned)(((long long)m * 32 + 255) / 256);
    k_radii_events<false><<<wgrid, 256, 0, st>>>(m, dim, lv->I, lv->J, lv->parent, lv->PJ, d_x,
                                                 cnt.get(), nullptr, nullptr, nullptr);
    k_rscan_blocks<<<nb, kScanBlock, 0, st>>>(cnt.get(), m, off.get(), sums.get());
    k_rscan_sums<<<1, kScanBlock, 0, st>>>(sums.get(), nb);
    k_rscan_add<<<nb, kScanBlock, 0, st>>>(off.get(), m, sums.get(), nb);
    GE_CUDA(cudaGetLastError());
    ctx->launches += 4;
    int total = 0;  // the number of events sizes the event arrays: one 4-byte round trip per level
    GE_CUDA(cudaMemcpyAsync(&total, off.get() + m, sizeof(int), cudaMemcpyDeviceToHost, st));
    GE_CUDA(cudaStreamSynchronize(st));
    ev_t.alloc(ctx, (size_t)std::max(total, 1));
    ev_ij.alloc(ctx, (size_t)std::max(total, 1));
    if (total > 0) {
      k_radii_events<true><<<wgrid, 256, 0, st>>>(m, dim, lv->I, lv->J, lv->parent, lv->PJ, d_x,
                                                  nullptr, off.get(), ev_t.get(), ev_ij.get());
      GE_CUDA(cudaGetLastError());
      ctx->launches++;
    }
    g.general = 1;
    g.nfam = mc;
    g.PI = lv->PI;
    g.PJ = lv->PJ;
    g.ev_off = off.get();
    g.ev_t = ev_t.get();
    g.ev_ij = ev_ij.get();
    g.xc = d_xc;
    g.rc = d_rc;
    if (mc > 0) {
      k_radii_grow<64, false><<<mc, 64, 0, st>>>(g);
      k_radii_grow<1024, true><<<mc, 1024, wide_smem, st>>>(g);
      GE_CUDA(cudaGetLastError());
      ctx->launches += 2;
    }
  }
  GE_CUDA(cudaEventRecord(ev1, st));
  GE_CUDA(cudaEventSynchronize(ev1));
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ev0, ev1) == cudaSuccess) ctx->radii_ms += ms;
  cudaEventDestroy(ev0);
  cudaEventDestroy(ev1);
}

}  // namespace ge
