// graph-embed_b200 :: flat ForceAtlas iteration for large n (kernel family K1), sm_100a.
//
// Replaces one iteration of partition::forceAtlas, /root/reference/include/forceatlas.hpp:146-270:
//   k_repulsion     <- :151-167  all-pairs repulsion, row-block x column-tile, the column tile
//                                staged in shared memory by 1-D TMA bulk copies (cp.async.bulk +
//                                mbarrier), FP64/FP32 pipes, one MUFU reciprocal square root per
//                                ordered pair.  Not a contraction: no tensor cores.
//   k_attract_step  <- :169-211, 214-217, 244-261  CSR attraction (sub-warp per row, coalesced
//                                index/weight loads, gathered coordinates), gravity, swing, speed
//                                cap and the Jacobi position update, fused; HBM-bound.
//   k_degree_mass   <- :127-140  weighted degree -> repulsion mass c = deg + 1.
// Device layout: coordinates, masses and forces are SoA [dim][ld] (ld = n padded to the column
// tile) so every access is coalesced and each array of a column tile is one contiguous bulk copy.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>
#include <type_traits>
#include <vector>

#include "ge_flat.cuh"
#include "ge_tma.cuh"

namespace ge {

// K1a.  Each thread owns IPT rows (positions and force accumulators in registers); the CTA walks
// column tiles.  Thread 0 is the TMA producer; everybody consumes through broadcast shared-memory
// loads (all lanes read the same column -> conflict free).  Work = the flat sequence of
// (row block, column tile) units, cut into gridDim.x equal contiguous shares: a CTA finishes the
// tail of one row block, sweeps whole row blocks, and starts the head of another.  Whole blocks
// are written straight to F; shared blocks go to the CTA's two partial slots and are summed in a
// fixed order by k_repulsion_fixup (deterministic, no atomics).
template <typename T, int D, int IPT, int JU, int TJ>
__global__ void __launch_bounds__(kRepMaxThreads) k_repulsion(const RepArgs<T> a) {
  constexpr int NM = Real<T>::kMassArrays;
  constexpr int NA = D + NM;
  constexpr int VEC = 16 / (int)sizeof(T);
  constexpr uint32_t kStageBytes = NA * TJ * sizeof(T);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* tiles = reinterpret_cast<T*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kRepStages * kStageBytes);

  const int tid = threadIdx.x;
  const long long W = a.total_units, G = gridDim.x, c = blockIdx.x;
  const long long u0 = W * c / G, u1 = W * (c + 1) / G;
  if (u0 >= u1) return;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kRepStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // row block holding unit u0: last block with unit0 <= u0
  int b = 0;
  {
    int lo = 0, hi = a.nblocks - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (a.blocks[mid].unit0 <= u0) lo = mid;
      else hi = mid - 1;
    }
    b = lo;
  }

  long long u = u0;
  unsigned g = 0;  // tiles this CTA has pushed through the pipeline (stage / parity bookkeeping)
  while (u < u1) {
    const BlockDesc bd = a.blocks[b];
    const int t_begin = (int)(u - bd.unit0);
    const int nt = (int)min((long long)(bd.ntiles - t_begin), u1 - u);

    T xi[IPT][D], fi[IPT][D];
#pragma unroll
    for (int t = 0; t < IPT; ++t) {
      const int i = bd.row0 + tid + t * blockDim.x;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        xi[t][k] = (i < bd.row1) ? a.pos[(int64_t)k * a.ld + i] : (T)0;
        fi[t][k] = (T)0;
      }
    }

    auto issue = [&](int l) {  // l-th tile of this segment
      const unsigned s = (g + (unsigned)l) % kRepStages;
      T* dst = tiles + (size_t)s * NA * TJ;
      const int64_t j = (int64_t)bd.j0 + (int64_t)(t_begin + l) * TJ;
      mbar_expect_tx(&full[s], kStageBytes);
#pragma unroll
      for (int k = 0; k < D; ++k)
        tma_load_1d(dst + k * TJ, a.pos + (int64_t)k * a.ld + j, TJ * sizeof(T), &full[s]);
#pragma unroll
      for (int k = 0; k < NM; ++k)
        tma_load_1d(dst + (D + k) * TJ, a.mass + (int64_t)k * a.ld + j, TJ * sizeof(T),
                    &full[s]);
    };
    // The stages the prologue refills held tiles nt-3 / nt-2 of the previous segment, which every
    // thread finished before the barrier of its last iteration.
    if (tid == 0) {
      for (int l = 0; l < kRepStages - 1 && l < nt; ++l) issue(l);
    }

    for (int l = 0; l < nt; ++l) {
      __syncthreads();  // everyone is done with tile l-1: its stage may be refilled
      if (tid == 0 && l + kRepStages - 1 < nt) issue(l + kRepStages - 1);
      const unsigned gl = g + (unsigned)l;
      const unsigned s = gl % kRepStages;
      mbar_wait(&full[s], (gl / kRepStages) & 1u);
      const T* st = tiles + (size_t)s * NA * TJ;

#pragma unroll JU
      for (int jj = 0; jj < TJ; jj += VEC) {
        T xj[D][VEC], mj[3][VEC];
#pragma unroll
        for (int k = 0; k < D; ++k) VecLoad<T, VEC>::ld(st + k * TJ + jj, xj[k]);
#pragma unroll
        for (int k = 0; k < NM; ++k) VecLoad<T, VEC>::ld(st + (D + k) * TJ + jj, mj[k]);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          // The IPT pairs of one column are independent: written stage by stage so that the
          // dependent FP chain of one pair is interleaved with the other pairs' (ILP = IPT).
          T d[IPT][D], r2[IPT], s3[IPT];
#pragma unroll
          for (int t = 0; t < IPT; ++t) {
            r2[t] = (T)0;
#pragma unroll
            for (int k = 0; k < D; ++k) {
              d[t][k] = xi[t][k] - xj[k][v];
              r2[t] = fma(d[t][k], d[t][k], r2[t]);
            }
          }
          Real<T>::template clamp_lo_n<IPT>(r2, a.eps2);
          Real<T>::template inv_cube_mass_n<IPT>(r2, mj[0][v], mj[NM > 1 ? 1 : 0][v],
                                                 mj[NM > 2 ? 2 : 0][v], s3);
#pragma unroll
          for (int t = 0; t < IPT; ++t) {
#pragma unroll
            for (int k = 0; k < D; ++k) fi[t][k] = fma(d[t][k], s3[t], fi[t][k]);
          }
        }
      }
    }
    g += (unsigned)nt;

    const bool whole = (t_begin == 0 && nt == bd.ntiles);
    const int slot = (u == u0) ? 0 : 1;
#pragma unroll
    for (int t = 0; t < IPT; ++t) {
      const int r = tid + t * blockDim.x;
      const int i = bd.row0 + r;
      if (i < bd.row1) {
        if (whole) {
          const T ci = a.mass[i] * a.repel;  // (deg_i + 1) * repel hoisted out of the pair loop
#pragma unroll
          for (int k = 0; k < D; ++k) a.F[(int64_t)k * a.ldf + (i - a.f_row_base)] = fi[t][k] * ci;
        } else {
#pragma unroll
          for (int k = 0; k < D; ++k)
            a.partial[(((size_t)c * 2 + slot) * D + k) * a.rows_per_block + r] = fi[t][k];
        }
      }
    }
    u += nt;
    ++b;
  }
}

// Sums, in CTA order, the partial results of the row blocks that k_repulsion split between CTAs.
template <typename T, int D>
__global__ void __launch_bounds__(256) k_repulsion_fixup(const RepArgs<T> a, int grid) {
  const BlockDesc bd = a.blocks[blockIdx.x];
  const long long W = a.total_units, G = grid;
  const long long U0 = bd.unit0, U1 = bd.unit0 + bd.ntiles;
  long long c = U0 * G / W;
  while (c + 1 < G && W * (c + 1) / G <= U0) ++c;
  while (c > 0 && W * c / G > U0) --c;
  if (W * c / G <= U0 && W * (c + 1) / G >= U1) return;  // swept whole by one CTA
  for (int r = threadIdx.x; bd.row0 + r < bd.row1; r += blockDim.x) {
    T acc[D];
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] = (T)0;
    for (long long cc = c; cc < G && W * cc / G < U1; ++cc) {
      const long long v0 = W * cc / G, v1 = W * (cc + 1) / G;
      if (v1 <= U0 || v0 >= v1) continue;
      const int slot = (max(v0, U0) == v0) ? 0 : 1;
#pragma unroll
      for (int k = 0; k < D; ++k)
        acc[k] += a.partial[(((size_t)cc * 2 + slot) * D + k) * a.rows_per_block + r];
    }
    const int i = bd.row0 + r;
    const T ci = a.mass[i] * a.repel;
#pragma unroll
    for (int k = 0; k < D; ++k) a.F[(int64_t)k * a.ldf + (i - a.f_row_base)] = acc[k] * ci;
  }
}

// One neighbour's coordinates from the interleaved copy (vector loads; DP = 2 or 4 reals).
template <typename T, int D>
struct Gather;
template <>
struct Gather<double, 2> {
  static constexpr int DP = 2;
  __device__ __forceinline__ static void ld(const double* b, int j, double (&x)[2]) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(b) + j);
    x[0] = v.x;
    x[1] = v.y;
  }
};
template <>
struct Gather<double, 3> {
  static constexpr int DP = 4;
  __device__ __forceinline__ static void ld(const double* b, int j, double (&x)[3]) {
    // one 256-bit load of the padded 32-byte record (sm_100: LDG.E.ENL2.256): a scattered gather
    // pays one L1 wavefront per lane and instruction, so one instruction instead of LDG.128 + LDG.64
    double pad __attribute__((unused));
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
        : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(pad)
        : "l"(b + 4 * (int64_t)j));
    (void)pad;
  }
};
template <>
struct Gather<float, 2> {
  static constexpr int DP = 2;
  __device__ __forceinline__ static void ld(const float* b, int j, float (&x)[2]) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(b) + j);
    x[0] = v.x;
    x[1] = v.y;
  }
};
template <>
struct Gather<float, 3> {
  static constexpr int DP = 4;
  __device__ __forceinline__ static void ld(const float* b, int j, float (&x)[3]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(b) + j);
    x[0] = v.x;
    x[1] = v.y;
    x[2] = v.z;
  }
};

template <typename T, int D>
__global__ void k_soa_to_gather(const T* __restrict__ soa, int64_t ld, T* __restrict__ aos) {
  constexpr int DP = Gather<T, D>::DP;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld) return;
#pragma unroll
  for (int k = 0; k < DP; ++k) aos[i * DP + k] = k < D ? soa[(int64_t)k * ld + i] : (T)0;
}

// K1b+c.  G lanes cooperate on one row (G = 4..32 chosen from the average degree); lane 0 of the
// group finishes the row: adds the repulsion sum, gravity, derives the per-vertex speed and
// writes the moved position into the NEXT coordinate buffer (Jacobi: everybody still reads the
// current one).
template <typename T, int D, int G, bool ML, bool GA>
__global__ void __launch_bounds__(256, sizeof(T) == 8 ? 4 : 6) k_attract_step(const StepArgs<T> a) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = gtid / G;
  const int lane = gtid % G;
  const bool active = r < a.nrows;
  const int i = a.row0 + (active ? r : 0);
  // Issue every load whose address is known up front before walking the row, so the row's
  // index -> coordinate gather chain overlaps them (the kernel is latency-, not issue-bound).
  const int e0 = active ? a.e_begin[r] : 0;
  int e1 = active ? a.e_end[r] : 0;
  const bool is_long = a.long_threshold > 0 && e1 - e0 > a.long_threshold;  // k_attract_step_long's
  if (is_long) e1 = e0;
  T x[D], f[D], frep[D], fprev[D], E[D];
  const bool finisher = active && lane == 0 && !is_long;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    x[k] = a.pos_cur[(int64_t)k * a.ld + i];
    f[k] = (T)0;
    frep[k] = finisher ? a.Frep[(int64_t)k * (a.ldr ? a.ldr : a.ldf) + r] : (T)0;
    fprev[k] = (finisher && a.update) ? a.Fprev[(int64_t)k * a.ldf + r] : (T)0;
    E[k] = (ML && finisher && a.Eext != nullptr) ? a.Eext[(int64_t)k * a.ldf + r] : (T)0;
  }
  const T ci = a.mass[i];
  if (a.frep_scale != (T)0) {
#pragma unroll
    for (int k = 0; k < D; ++k) frep[k] *= ci * a.frep_scale;
  }
  const bool weighted = a.W != nullptr && a.ph.use_weights;
  // EU entries per lane and trip: all index loads, then all gathers, are in flight together
  // (memory-level parallelism; the kernel is latency-bound long before it is issue-bound)
  constexpr int EU = G <= 2 ? 4 : 2;
  for (int e = e0 + lane; e < e1; e += EU * G) {
    int jn[EU];
    T wn[EU];
    bool on[EU];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      const int eu = e + u * G;
      on[u] = eu < e1;
      jn[u] = on[u] ? a.J[eu] : i;
      wn[u] = (on[u] && weighted) ? a.W[eu] : (T)1;
    }
    T dn[EU][D];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      if (a.aos_cur != nullptr) {
        Gather<T, D>::ld(a.aos_cur, jn[u], dn[u]);
      } else {
#pragma unroll
        for (int k = 0; k < D; ++k) dn[u][k] = a.pos_cur[(int64_t)k * a.ld + jn[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      T r2 = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        dn[u][k] -= x[k];
        r2 = fma(dn[u][k], dn[u][k], r2);
      }
      const T g = on[u] ? attraction_factor<T, GA>(r2, wn[u], ci, a.ph) : (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) f[k] = fma(dn[u][k], g, f[k]);
    }
  }
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < D; ++k) f[k] += __shfl_xor_sync(0xffffffffu, f[k], off, G);
  }
  if (finisher) {
#pragma unroll
    for (int k = 0; k < D; ++k) f[k] += frep[k];
    vertex_step<T, D, ML>(x, f, fprev, E, ci, a.ph);
#pragma unroll
    for (int k = 0; k < D; ++k) {
      a.Fprev[(int64_t)k * a.ldf + r] = fprev[k];
      if (a.update) a.pos_next[(int64_t)k * a.ld + i] = x[k];
    }
    if (a.update && a.aos_next != nullptr) {
      constexpr int DP = Gather<T, D>::DP;
#pragma unroll
      for (int k = 0; k < DP; ++k) a.aos_next[(int64_t)i * DP + k] = k < D ? x[k] : (T)0;
    }
  }
}

// K1b+c, staged variant for contiguous CSR rows (the flat solver).  The chain of dependent loads
// per row -- row pointer -> column index / weight -> neighbour coordinates -- is what bounds
// k_attract_step (ncu: long-scoreboard stalls on the index and weight loads).  Here a persistent CTA
// walks chunks of 256/G rows; thread 0 streams each chunk's slice of the index and weight arrays
// into shared memory with 1-D TMA bulk copies ONE CHUNK AHEAD (two stages, mbarrier completion),
// and the chunk bounds / row pointers are fetched one and two chunks ahead, so that while a chunk
// is processed the only exposed latency is the coordinate gather itself.  A chunk with more than
// kStepCap entries (very long rows) reads its indices from global memory as before.
constexpr int kStepCap = 2560;  // staged entries per chunk and stage (multiple of 4: 16-byte copies)

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

template <typename T, int D, int G, bool GA, int CAP, int MINB>
__global__ void __launch_bounds__(256, MINB) k_attract_step_staged(const StepArgs<T> a, const int nchunks) {
  constexpr int kStepCap = CAP;
  constexpr int RPC = 256 / G;
  constexpr int EU = G <= 2 ? 4 : 2;  // entries in flight per lane (5 or 6 with one lane per row: no gain)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* Ws = reinterpret_cast<T*>(smem_raw);
  int* Js = reinterpret_cast<int*>(smem_raw + 2 * kStepCap * sizeof(T));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + 2 * kStepCap * (sizeof(T) + sizeof(int)));
  const int tid = threadIdx.x;
  const int lane = tid % G, rl = tid / G;
  const bool weighted = a.W != nullptr && a.ph.use_weights;
  const int* __restrict__ rowptr = a.e_begin;  // contiguous CSR: e_end == e_begin + 1
  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int c = blockIdx.x;
  if (c >= nchunks) return;
  const int stride = gridDim.x;

  auto bounds = [&](int cc, int& b0, int& b1) {
    const int r0 = cc * RPC;
    b0 = rowptr[r0];
    b1 = rowptr[min(r0 + RPC, a.nrows)];
  };
  auto row_range = [&](int cc, int& e0, int& e1) {
    const int r = cc * RPC + rl;
    const bool ok = r < a.nrows;
    e0 = ok ? rowptr[r] : 0;
    e1 = ok ? rowptr[r + 1] : 0;
  };
  auto staged_count = [](int b0, int b1) { return ((b1 - (b0 & ~3)) + 3) & ~3; };
  auto issue = [&](int stage, int b0, int b1) {  // thread 0
    const int base = b0 & ~3;
    const int cnt = staged_count(b0, b1);
    if (cnt <= 0 || cnt > kStepCap) return;
    mbar_expect_tx(&full[stage], (uint32_t)cnt * (uint32_t)(sizeof(int) + (weighted ? sizeof(T) : 0)));
    tma_load_1d(Js + stage * kStepCap, a.J + base, (uint32_t)cnt * sizeof(int), &full[stage]);
    if (weighted) tma_load_1d(Ws + stage * kStepCap, a.W + base, (uint32_t)cnt * sizeof(T), &full[stage]);
  };

  int b0, b1, nb0 = 0, nb1 = 0, e0, e1;
  bounds(c, b0, b1);
  if (c + stride < nchunks) bounds(c + stride, nb0, nb1);
  row_range(c, e0, e1);
  if (tid == 0) issue(0, b0, b1);
  unsigned phases = 0u;  // bit s: parity of the next completion of stage s
  const T* posk[D];
#pragma unroll
  for (int kk = 0; kk < D; ++kk) posk[kk] = a.pos_cur + (int64_t)kk * a.ld;

  for (unsigned k = 0;; ++k) {
    const int stage = (int)(k & 1u);
    const int cn = c + stride, cnn = c + 2 * stride;
    const bool has_next = cn < nchunks;
    // one chunk ahead: its copies; two ahead: its bounds; the next chunk's row pointers
    if (tid == 0 && has_next) issue(stage ^ 1, nb0, nb1);
    int nnb0 = 0, nnb1 = 0, ne0 = 0, ne1 = 0;
    if (cnn < nchunks) bounds(cnn, nnb0, nnb1);
    if (has_next) row_range(cn, ne0, ne1);
    // the next chunk's per-row arrays (own position, pair sums, previous force, mass) are pulled
    // into L2 now: the first use of x in the walk otherwise waits a full DRAM round trip per chunk
    // (ncu: the top stall was the DADD that subtracts x from the first gathered neighbour)
    if (has_next && lane == 0) {
      const int rn = cn * RPC + rl;
      if (rn < a.nrows) {
        const int in = a.row0 + rn;
        const int64_t ldr_ = a.ldr ? a.ldr : a.ldf;
#pragma unroll
        for (int kk = 0; kk < D; ++kk) {
          prefetch_l2(a.pos_cur + (int64_t)kk * a.ld + in);
          prefetch_l2(a.Frep + (int64_t)kk * ldr_ + rn);
          if (a.update) prefetch_l2(a.Fprev + (int64_t)kk * a.ldf + rn);
        }
        prefetch_l2(a.mass + in);
      }
    }

    const int r = c * RPC + rl;
    const bool active = r < a.nrows;
    const int i = a.row0 + (active ? r : 0);
    const bool is_long = a.long_threshold > 0 && e1 - e0 > a.long_threshold;  // k_attract_step_long's
    if (is_long) e1 = e0;
    const bool finisher = active && lane == 0 && !is_long;
    T x[D], f[D], frep[D], fprev[D];
    const int64_t ldr = a.ldr ? a.ldr : a.ldf;
#pragma unroll
    for (int kk = 0; kk < D; ++kk) {
      x[kk] = a.pos_cur[(int64_t)kk * a.ld + i];
      f[kk] = (T)0;
      frep[kk] = finisher ? a.Frep[(int64_t)kk * ldr + r] : (T)0;
      fprev[kk] = (finisher && a.update) ? a.Fprev[(int64_t)kk * a.ldf + r] : (T)0;
    }
    const T ci = a.mass[i];

    const int base = b0 & ~3;
    const int cnt = staged_count(b0, b1);
    const bool staged = cnt > 0 && cnt <= kStepCap;
    if (staged) {
      mbar_wait(&full[stage], (phases >> stage) & 1u);
      phases ^= 1u << stage;
    }
    // The gather layout (interleaved copy or SoA) and the presence of weights are compile-time
    // inside the walk: predicating both variants costs ~2x the issue slots of this loop.
    auto walk = [&](auto aos_c, auto wt_c, const int* __restrict__ Jp, const T* __restrict__ Wp) {
      constexpr bool AOS = decltype(aos_c)::value;
      constexpr bool WT = decltype(wt_c)::value;
      for (int e = e0 + lane; e < e1; e += EU * G) {
        int jn[EU];
        T wn[EU];  // 0 on the slots past the end of the row: their term vanishes
        bool on[EU];
#pragma unroll
        for (int u = 0; u < EU; ++u) {
          const int eu = e + u * G;
          on[u] = eu < e1;
          jn[u] = on[u] ? Jp[eu] : i;
          if (WT) wn[u] = on[u] ? Wp[eu] : (T)0;
          else wn[u] = on[u] ? (T)1 : (T)0;
        }
        T dn[EU][D];
#pragma unroll
        for (int u = 0; u < EU; ++u) {
          if (AOS) {
            Gather<T, D>::ld(a.aos_cur, jn[u], dn[u]);
          } else {
#pragma unroll
            for (int kk = 0; kk < D; ++kk) dn[u][kk] = posk[kk][jn[u]];
          }
        }
#pragma unroll
        for (int u = 0; u < EU; ++u) {
          T r2 = (T)0;
#pragma unroll
          for (int kk = 0; kk < D; ++kk) {
            dn[u][kk] -= x[kk];
            if (GA) r2 = fma(dn[u][kk], dn[u][kk], r2);
          }
          T g;
          if (GA) g = on[u] ? attraction_factor<T, true>(r2, WT ? wn[u] : (T)1, ci, a.ph) : (T)0;
          else g = a.ph.attract * wn[u];
#pragma unroll
          for (int kk = 0; kk < D; ++kk) f[kk] = fma(dn[u][kk], g, f[kk]);
        }
      }
    };
    using std::false_type;
    using std::true_type;
    if (staged) {
      const int* Jp = Js + stage * kStepCap - base;
      const T* Wp = Ws + stage * kStepCap - base;
      if (a.aos_cur != nullptr) {
        if (weighted) walk(true_type{}, true_type{}, Jp, Wp);
        else walk(true_type{}, false_type{}, Jp, Wp);
      } else {
        if (weighted) walk(false_type{}, true_type{}, Jp, Wp);
        else walk(false_type{}, false_type{}, Jp, Wp);
      }
    } else {  // long rows: indices and weights straight from global memory
      if (a.aos_cur != nullptr) {
        if (weighted) walk(true_type{}, true_type{}, a.J, a.W);
        else walk(true_type{}, false_type{}, a.J, a.W);
      } else {
        if (weighted) walk(false_type{}, true_type{}, a.J, a.W);
        else walk(false_type{}, false_type{}, a.J, a.W);
      }
    }

#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
#pragma unroll
      for (int kk = 0; kk < D; ++kk) f[kk] += __shfl_xor_sync(0xffffffffu, f[kk], off, G);
    }
    if (finisher) {
      if (a.frep_scale != (T)0) {
#pragma unroll
        for (int kk = 0; kk < D; ++kk) frep[kk] *= ci * a.frep_scale;
      }
      const T E[D] = {};
#pragma unroll
      for (int kk = 0; kk < D; ++kk) f[kk] += frep[kk];
      vertex_step<T, D, false>(x, f, fprev, E, ci, a.ph);
#pragma unroll
      for (int kk = 0; kk < D; ++kk) {
        a.Fprev[(int64_t)kk * a.ldf + r] = fprev[kk];
        if (a.update) a.pos_next[(int64_t)kk * a.ld + i] = x[kk];
      }
      if (a.update && a.aos_next != nullptr) {
        constexpr int DP = Gather<T, D>::DP;
#pragma unroll
        for (int kk = 0; kk < DP; ++kk) a.aos_next[(int64_t)i * DP + kk] = kk < D ? x[kk] : (T)0;
      }
    }
    if (!has_next) break;
    __syncthreads();  // everyone has left this chunk's stage: the copy issued next trip may refill it
    c = cn;
    b0 = nb0, b1 = nb1, nb0 = nnb0, nb1 = nnb1, e0 = ne0, e1 = ne1;
  }
}

// K1b+c for the long rows of power-law graphs: one CTA per row.  All threads walk the row's entries
// (coalesced index / weight loads, 4 gathers in flight per thread), the partial sums are combined in
// a fixed order (lane tree, then warps in order: bit-reproducible) and thread 0 finishes the row
// exactly like the row kernels do.
template <typename T, int D, bool GA, int kLongThreads, bool ML = false>
__global__ void __launch_bounds__(kLongThreads) k_attract_step_long(const StepArgs<T> a,
                                                                    const int* __restrict__ rows) {
  __shared__ T red[D][kLongThreads / 32];
  const int r = rows[blockIdx.x];
  const int i = a.row0 + r;
  const int tid = threadIdx.x;
  const int e0 = a.e_begin[r], e1 = a.e_end[r];
  const bool weighted = a.W != nullptr && a.ph.use_weights;
  T x[D], f[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    x[k] = a.pos_cur[(int64_t)k * a.ld + i];
    f[k] = (T)0;
  }
  const T ci = a.mass[i];
  constexpr int EU = 4;
  for (int e = e0 + tid; e < e1; e += EU * kLongThreads) {
    int jn[EU];
    T wn[EU];
    bool on[EU];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      const int eu = e + u * kLongThreads;
      on[u] = eu < e1;
      jn[u] = on[u] ? a.J[eu] : i;
      wn[u] = (on[u] && weighted) ? a.W[eu] : (T)1;
    }
    T dn[EU][D];
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      if (a.aos_cur != nullptr) {
        Gather<T, D>::ld(a.aos_cur, jn[u], dn[u]);
      } else {
#pragma unroll
        for (int k = 0; k < D; ++k) dn[u][k] = a.pos_cur[(int64_t)k * a.ld + jn[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < EU; ++u) {
      T r2 = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        dn[u][k] -= x[k];
        r2 = fma(dn[u][k], dn[u][k], r2);
      }
      const T g = on[u] ? attraction_factor<T, GA>(r2, wn[u], ci, a.ph) : (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) f[k] = fma(dn[u][k], g, f[k]);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < D; ++k) f[k] += __shfl_xor_sync(0xffffffffu, f[k], off);
  }
  if ((tid & 31) == 0) {
#pragma unroll
    for (int k = 0; k < D; ++k) red[k][tid >> 5] = f[k];
  }
  __syncthreads();
  if (tid != 0) return;
  T frep[D], fprev[D];
  const int64_t ldr = a.ldr ? a.ldr : a.ldf;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    T acc = (T)0;
    for (int w = 0; w < kLongThreads / 32; ++w) acc += red[k][w];
    frep[k] = a.Frep[(int64_t)k * ldr + r];
    if (a.frep_scale != (T)0) frep[k] *= ci * a.frep_scale;
    fprev[k] = a.update ? a.Fprev[(int64_t)k * a.ldf + r] : (T)0;
    f[k] = acc + frep[k];
  }
  T E[D];
#pragma unroll
  for (int k = 0; k < D; ++k) E[k] = (ML && a.Eext != nullptr) ? a.Eext[(int64_t)k * a.ldf + r] : (T)0;
  vertex_step<T, D, ML>(x, f, fprev, E, ci, a.ph);
#pragma unroll
  for (int k = 0; k < D; ++k) {
    a.Fprev[(int64_t)k * a.ldf + r] = fprev[k];
    if (a.update) a.pos_next[(int64_t)k * a.ld + i] = x[k];
  }
  if (a.update && a.aos_next != nullptr) {
    constexpr int DP = Gather<T, D>::DP;
#pragma unroll
    for (int k = 0; k < DP; ++k) a.aos_next[(int64_t)i * DP + k] = k < D ? x[k] : (T)0;
  }
}

// include/forceatlas.hpp:127-140: c_i = 1 + sum of row weights (or row length).  Also emits the
// two scaled copies the FP64 pair kernel consumes (1.5 c, 1.875 c).  One thread per row.
template <typename T>
__global__ void k_mass_from_degree(const double* __restrict__ deg, int n, int64_t ld, int nm,
                                   T* __restrict__ mass) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ld) return;
  const double c = (i < n) ? deg[i] + 1.0 : 0.0;  // padding columns carry zero mass
  mass[i] = (T)c;
  if (nm > 1) mass[ld + i] = (T)(1.5 * c);
  if (nm > 2) mass[2 * ld + i] = (T)(1.875 * c);
}

// AoS double (host image) <-> SoA T (device layout).
// perm (optional): internal row i holds the caller's vertex perm[i]
template <typename T>
__global__ void k_aos_to_soa(const double* __restrict__ aos, int n, int dim, int64_t ld,
                             T* __restrict__ soa, const int* __restrict__ perm = nullptr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)ld * dim) return;
  const int k = (int)(t / ld);
  const int64_t i = t % ld;
  soa[t] = (i < n) ? (T)aos[(int64_t)(perm ? perm[i] : i) * dim + k] : (T)0;
}
template <typename T>
__global__ void k_soa_to_aos(const T* __restrict__ soa, int n, int dim, int64_t ld,
                             double* __restrict__ aos, const int* __restrict__ perm = nullptr) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * dim) return;
  const int64_t i = t / dim;
  const int k = (int)(t % dim);
  aos[(int64_t)(perm ? perm[i] : i) * dim + k] = (double)soa[(int64_t)k * ld + i];
}

// include/forceatlas.hpp:272-303 (normalize=true): centre on the mean, divide by the largest
// norm (unclamped, like the reference).  Single CTA, grid-stride; only used when the caller asks.
template <typename T, int D>
__global__ void __launch_bounds__(1024) k_normalize(T* pos, int n, int64_t ld) {
  __shared__ double red[32];
  __shared__ double bc[D + 1];
  const int tid = threadIdx.x;
  auto block_reduce = [&](double v, bool is_max) -> double {
    for (int off = 16; off > 0; off >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, v, off);
      v = is_max ? fmax(v, o) : v + o;
    }
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    if (tid < 32) {
      double w = (tid < (blockDim.x >> 5)) ? red[tid] : (is_max ? 0.0 : 0.0);
      for (int off = 16; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, w, off);
        w = is_max ? fmax(w, o) : w + o;
      }
      if (tid == 0) red[0] = w;
    }
    __syncthreads();
    const double out = red[0];
    __syncthreads();
    return out;
  };
  for (int k = 0; k < D; ++k) {
    double s = 0.0;
    for (int i = tid; i < n; i += blockDim.x) s += (double)pos[(int64_t)k * ld + i];
    const double tot = block_reduce(s, false);
    if (tid == 0) bc[k] = tot / n;
  }
  __syncthreads();
  double mx = 0.0;
  for (int i = tid; i < n; i += blockDim.x) {
    double m2 = 0.0;
    for (int k = 0; k < D; ++k) {
      const double v = (double)pos[(int64_t)k * ld + i] - bc[k];
      m2 += v * v;
    }
    mx = fmax(mx, sqrt(m2));
  }
  const double maxlen = block_reduce(mx, true);
  for (int i = tid; i < n; i += blockDim.x)
    for (int k = 0; k < D; ++k)
      pos[(int64_t)k * ld + i] = (T)(((double)pos[(int64_t)k * ld + i] - bc[k]) / maxlen);
}

// ---------------------------------------------------------------------------------------------
// launchers (also used by ge_multilevel.cu for aggregates too large for one CTA)
// ---------------------------------------------------------------------------------------------
namespace {
int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

template <typename T>
size_t repulsion_smem(int dim, int tj) {
  return (size_t)kRepStages * (dim + Real<T>::kMassArrays) * tj * sizeof(T) +
         kRepStages * sizeof(uint64_t);
}

template <typename T, int D, int JU, int TJ>
const void* repulsion_kernel_d(int ipt) {
  return ipt == 1 ? (const void*)k_repulsion<T, D, 1, JU, TJ>
                  : ipt == 2 ? (const void*)k_repulsion<T, D, 2, JU, TJ> : (const void*)k_repulsion<T, D, 4, JU, TJ>;
}
template <typename T, int TJ>
const void* repulsion_kernel_t(int dim, int ipt, int ju) {
  if (dim == 2) return ju == 1 ? repulsion_kernel_d<T, 2, 1, TJ>(ipt) : repulsion_kernel_d<T, 2, 2, TJ>(ipt);
  return ju == 1 ? repulsion_kernel_d<T, 3, 1, TJ>(ipt) : repulsion_kernel_d<T, 3, 2, TJ>(ipt);
}
template <typename T>
const void* repulsion_kernel(int dim, int ipt, int ju, int tj) {
  return tj == kTileJSmall ? repulsion_kernel_t<T, kTileJSmall>(dim, ipt, ju)
                           : repulsion_kernel_t<T, kTileJ>(dim, ipt, ju);
}
}  // namespace

template <typename T>
RepulsionPlan<T>::RepulsionPlan(ge_context* ctx, int dim, const std::vector<RowSegment>& segments)
    : ctx_(ctx), dim_(dim) {
  // Launch shape measured on B200 (tools/sweep_rep.py; every shape lands within 10 % because the
  // kernel is issue-bound): 512 threads, 2 rows per thread (4 for FP64 d = 3), the
  // column loop unrolled twice.  Small sweeps use narrower CTAs so that there are enough
  // (row block, tile) units to share out.
  long long total_rows = 0, total_pairs = 0;
  for (const auto& sg : segments) {
    total_rows += sg.row1 - sg.row0;
    total_pairs += (long long)(sg.row1 - sg.row0) * (sg.j1 - sg.j0);
  }
  ipt_ = env_int("GE_REP_IPT", (sizeof(T) == 8 && dim == 3) ? 4 : 2);
  threads_ = env_int("GE_REP_THREADS", 512);
  tile_ = kTileJ;
  // Small sweeps (up to a few million pairs: large aggregates, graphs of 1-2 thousand vertices)
  // are latency-bound per launch: fine-grained units (128 rows x 64 columns, one row per thread)
  // put every SM to work instead of a handful of CTAs (measured n = 1000: 20 vs 30 us per
  // iteration; from n = 4000 the wide configuration wins again).
  if (total_pairs < 5000000LL && !std::getenv("GE_REP_THREADS")) {
    tile_ = kTileJSmall;
    threads_ = 128;
    ipt_ = 1;
  }
  // narrower CTAs only when there would otherwise be too few (row block, tile) units to share out
  // (a 1/8 row block of a 500k-vertex graph still has 120k units: keep the wide shape)
  while (threads_ > 128 &&
         (total_pairs / ((long long)threads_ * ipt_ * tile_)) < 8LL * ctx->sm_count)
    threads_ /= 2;
  ju_ = env_int("GE_REP_JU", 2);
  const void* fn = repulsion_kernel<T>(dim_, ipt_, ju_, tile_);
  const size_t smem = repulsion_smem<T>(dim_, tile_);
  if (smem > 48 * 1024)
    GE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  GE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads_, smem));
  GE_REQUIRE(occ > 0, "repulsion kernel does not fit on an SM");
  const int rows_per_block = threads_ * ipt_;
  std::vector<BlockDesc> blocks;
  long long units = 0;
  for (const auto& sg : segments) {
    const int ntiles = (sg.j1 - sg.j0) / tile_;
    for (int r = sg.row0; r < sg.row1; r += rows_per_block) {
      blocks.push_back(BlockDesc{r, std::min(sg.row1, r + rows_per_block), sg.j0, ntiles, units});
      units += ntiles;
    }
  }
  nblocks_ = (int)blocks.size();
  total_units_ = units;
  grid_ = (int)std::min<long long>((long long)ctx->sm_count * occ, std::max<long long>(units, 1));
  blocks_.alloc(ctx, std::max<size_t>(blocks.size(), 1));
  blocks_.upload(ctx, blocks.data(), blocks.size());
  partial_.alloc(ctx, (size_t)grid_ * 2 * dim_ * rows_per_block);
  GE_CUDA(cudaStreamSynchronize(ctx->stream));
  if (std::getenv("GE_VERBOSE"))
    std::fprintf(stderr, "[ge] repulsion plan: threads=%d ipt=%d tile=%d grid=%d (occ %d) blocks=%d units=%lld\n",
                 threads_, ipt_, tile_, grid_, occ, nblocks_, total_units_);
}

template <typename T>
void RepulsionPlan<T>::launch(const T* pos, const T* mass, int64_t ld, T* F, int64_t ldf,
                              int f_row_base, T repel, T eps2) {
  if (nblocks_ == 0 || total_units_ == 0) return;
  RepArgs<T> a;
  a.pos = pos;
  a.mass = mass;
  a.F = F;
  a.partial = partial_.get();
  a.blocks = blocks_.get();
  a.ld = ld;
  a.ldf = ldf;
  a.total_units = total_units_;
  a.nblocks = nblocks_;
  a.rows_per_block = threads_ * ipt_;
  a.f_row_base = f_row_base;
  a.repel = repel;
  a.eps2 = eps2;
  void* args[] = {(void*)&a};
  GE_CUDA(cudaLaunchKernel(repulsion_kernel<T>(dim_, ipt_, ju_, tile_), dim3(grid_), dim3(threads_), args,
                           repulsion_smem<T>(dim_, tile_), ctx_->stream));
  ctx_->launches++;
  if (dim_ == 2) k_repulsion_fixup<T, 2><<<nblocks_, 256, 0, ctx_->stream>>>(a, grid_);
  else k_repulsion_fixup<T, 3><<<nblocks_, 256, 0, ctx_->stream>>>(a, grid_);
  GE_CUDA(cudaGetLastError());
  ctx_->launches++;
}

template class RepulsionPlan<double>;
template class RepulsionPlan<float>;

namespace {
template <typename T, int D, bool ML, bool GA>
void launch_step_g(ge_context* ctx, const StepArgs<T>& a, int group) {
  const int g = group <= 1 ? 1 : group <= 2 ? 2 : group <= 4 ? 4 : group <= 8 ? 8 : group <= 16 ? 16 : 32;
  const int64_t threads = (int64_t)a.nrows * g;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  if (g == 1) k_attract_step<T, D, 1, ML, GA><<<grid, 256, 0, ctx->stream>>>(a);
  else if (g == 2) k_attract_step<T, D, 2, ML, GA><<<grid, 256, 0, ctx->stream>>>(a);
  else if (g == 4) k_attract_step<T, D, 4, ML, GA><<<grid, 256, 0, ctx->stream>>>(a);
  else if (g == 8) k_attract_step<T, D, 8, ML, GA><<<grid, 256, 0, ctx->stream>>>(a);
  else if (g == 16) k_attract_step<T, D, 16, ML, GA><<<grid, 256, 0, ctx->stream>>>(a);
  else k_attract_step<T, D, 32, ML, GA><<<grid, 256, 0, ctx->stream>>>(a);
}
template <typename T, int D, bool ML>
void launch_step_ga(ge_context* ctx, const StepArgs<T>& a, int group) {
  if (a.ph.general_attraction) launch_step_g<T, D, ML, true>(ctx, a, group);
  else launch_step_g<T, D, ML, false>(ctx, a, group);
}
}  // namespace

namespace {
template <typename T, int D, int G, bool GA, int CAP, int MINB>
void launch_staged_c(ge_context* ctx, const StepArgs<T>& a) {
  const size_t smem = 2 * CAP * (sizeof(T) + sizeof(int)) + 2 * sizeof(uint64_t);
  auto fn = k_attract_step_staged<T, D, G, GA, CAP, MINB>;
  static std::atomic<int> occ_dev[64];  // per instantiation and device (function attributes are per device)
  const int di = ctx->device & 63;
  int occ = occ_dev[di].load();
  if (occ == 0) {
    GE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 256, smem));
    GE_REQUIRE(occ > 0, "staged attraction kernel does not fit on an SM");
    occ_dev[di].store(occ);
  }
  const int rpc = 256 / G;
  const int nchunks = (a.nrows + rpc - 1) / rpc;
  const int grid = std::min(nchunks, ctx->sm_count * occ);
  fn<<<grid, 256, smem, ctx->stream>>>(a, nchunks);
}
template <typename T, int D, int G, bool GA>
void launch_staged_g(ge_context* ctx, const StepArgs<T>& a) {
  static const int cap = env_int("GE_STEP_CAP", kStepCap);
  if (cap <= 2048) launch_staged_c<T, D, G, GA, 2048, 4>(ctx, a);
  else launch_staged_c<T, D, G, GA, kStepCap, 3>(ctx, a);
}
template <typename T, int D, bool GA>
void launch_staged_d(ge_context* ctx, const StepArgs<T>& a, int group) {
  // (a 3072-entry / 3-CTA shape of the one-lane variant measured 47 % vs 58 %; cause not isolated)
  if (group <= 1) launch_staged_c<T, D, 1, GA, 4096, 2>(ctx, a);
  else if (group <= 2) launch_staged_g<T, D, 2, GA>(ctx, a);
  else if (group <= 4) launch_staged_g<T, D, 4, GA>(ctx, a);
  else launch_staged_g<T, D, 8, GA>(ctx, a);
}
}  // namespace

// Contiguous CSR rows (a.e_end == a.e_begin + 1), J / W padded by >= 4 entries, flat physics.
template <typename T>
void launch_attract_step_staged(ge_context* ctx, const StepArgs<T>& a, int dim, int group) {
  if (a.nrows == 0) return;
  if (dim == 2) {
    if (a.ph.general_attraction) launch_staged_d<T, 2, true>(ctx, a, group);
    else launch_staged_d<T, 2, false>(ctx, a, group);
  } else {
    if (a.ph.general_attraction) launch_staged_d<T, 3, true>(ctx, a, group);
    else launch_staged_d<T, 3, false>(ctx, a, group);
  }
  GE_CUDA(cudaGetLastError());
  ctx->launches++;
}

template <typename T>
void launch_attract_step(ge_context* ctx, const StepArgs<T>& a, int dim, int group, bool ml) {
  if (a.nrows == 0) return;
  if (dim == 2) {
    if (ml) launch_step_ga<T, 2, true>(ctx, a, group);
    else launch_step_ga<T, 2, false>(ctx, a, group);
  } else {
    if (ml) launch_step_ga<T, 3, true>(ctx, a, group);
    else launch_step_ga<T, 3, false>(ctx, a, group);
  }
  GE_CUDA(cudaGetLastError());
  ctx->launches++;
}

template <typename T>
void launch_attract_step_long(ge_context* ctx, const StepArgs<T>& a, int dim, const int* rows, int nlong,
                              int threads, bool ml) {
  if (nlong == 0) return;
  const bool ga = a.ph.general_attraction != 0;
  auto go = [&](auto d_c, auto ga_c) {
    constexpr int D = decltype(d_c)::value;
    constexpr bool GA = decltype(ga_c)::value;
    if (ml) k_attract_step_long<T, D, GA, 128, true><<<nlong, 128, 0, ctx->stream>>>(a, rows);  // the multilevel tier
    else if (threads >= 512) k_attract_step_long<T, D, GA, 512><<<nlong, 512, 0, ctx->stream>>>(a, rows);
    else k_attract_step_long<T, D, GA, 32><<<nlong, 32, 0, ctx->stream>>>(a, rows);
  };
  using std::integral_constant;
  if (dim == 2) {
    if (ga) go(integral_constant<int, 2>{}, std::true_type{});
    else go(integral_constant<int, 2>{}, std::false_type{});
  } else {
    if (ga) go(integral_constant<int, 3>{}, std::true_type{});
    else go(integral_constant<int, 3>{}, std::false_type{});
  }
  GE_CUDA(cudaGetLastError());
  ctx->launches++;
}
template void launch_attract_step_long<double>(ge_context*, const StepArgs<double>&, int, const int*, int, int, bool);
template void launch_attract_step_long<float>(ge_context*, const StepArgs<float>&, int, const int*, int, int, bool);

template void launch_attract_step<double>(ge_context*, const StepArgs<double>&, int, int, bool);
template void launch_attract_step<float>(ge_context*, const StepArgs<float>&, int, int, bool);

int group_for_degree(double avg_deg) {
  // measured on B200 (n = 2M, avg degree 10, d = 3): 2 lanes per row 0.244 ms, 1: 0.269, 4: 0.283,
  // 8: 0.40 -- short rows want few lanes (the per-row epilogue runs on one lane of the group)
  return avg_deg <= 16 ? 2 : avg_deg <= 32 ? 4 : avg_deg <= 64 ? 8 : avg_deg <= 128 ? 16 : 32;
}

// ---------------------------------------------------------------------------------------------
// device-resident flat solver
// ---------------------------------------------------------------------------------------------
namespace {

template <typename T>
class FlatSolverT final : public FlatSolver {
 public:
  FlatSolverT(ge_context* c, const ge_csr& A, int dim, const ge_params& p, int rb, int re,
              int part, int parts, const double* shared_deg)
      : n_(A.rows), dim_(dim), rb_(rb), re_(re), parts_(parts) {
    ctx = c;
    const bool verbose_ctor = std::getenv("GE_VERBOSE_PLAN") != nullptr;
    double t_ctor = now_ms();
    auto lap = [&](const char* what) {
      if (!verbose_ctor) return;
      cudaStreamSynchronize(ctx->stream);
      const double t = now_ms();
      std::fprintf(stderr, "[ge] flat plan ctor dev %d %-18s %8.3f ms\n", ctx->device, what, t - t_ctor);
      t_ctor = t;
    };
    ph_ = make_physics<T>(p);
    ld_ = round_up(std::max(n_, 1), kTileJ);
    nrows_ = re_ - rb_;
    ldf_ = round_up(std::max(nrows_, 1), 32);
    constexpr int NM = Real<T>::kMassArrays;

    // Internal renumbering (single-rank plans on large graphs that will iterate many times):
    // breadth-first order makes a row's neighbours close in memory, so the gathers of the
    // attraction kernel hit L1/L2 lines instead of one HBM sector each (ncu at n = 2M, random
    // numbering: 1.1 GB of DRAM traffic for 0.5 GB of algorithmic bytes).  The all-pairs sweep is
    // order-agnostic.  The caller's numbering is restored at the upload / download transposes.
    const bool reorder = rb_ == 0 && re_ == n_ && n_ >= 65536 && p.iterations >= 16 &&
                         std::getenv("GE_NO_REORDER") == nullptr;
    std::vector<int> perm, inv, rI, rJ;
    std::vector<double> rD;
    const int32_t* I = A.indptr;
    const int32_t* J = A.indices;
    const double* Dw = A.data;
    if (reorder) {
      perm.reserve(n_);
      inv.assign(n_, -1);
      for (int root = 0; root < n_; ++root) {
        if (inv[root] >= 0) continue;
        inv[root] = (int)perm.size();
        perm.push_back(root);
        for (size_t head = perm.size() - 1; head < perm.size(); ++head) {
          const int u = perm[head];
          for (int e = A.indptr[u]; e < A.indptr[u + 1]; ++e) {
            const int v = A.indices[e];
            if (inv[v] < 0) {
              inv[v] = (int)perm.size();
              perm.push_back(v);
            }
          }
        }
      }
      rI.resize(n_ + 1);
      rJ.resize(A.indptr[n_]);
      if (A.data) rD.resize(A.indptr[n_]);
      rI[0] = 0;
      for (int r = 0; r < n_; ++r) {
        const int o = perm[r];
        int w = rI[r];
        for (int e = A.indptr[o]; e < A.indptr[o + 1]; ++e, ++w) {
          rJ[w] = inv[A.indices[e]];
          if (A.data) rD[w] = A.data[e];
        }
        rI[r + 1] = w;
      }
      I = rI.data();
      J = rJ.data();
      Dw = A.data ? rD.data() : nullptr;
      perm_.alloc(ctx, n_);
      perm_.upload(ctx, perm.data(), n_);
    }

    lap("renumbering");
    // Vertex masses need every row's degree (include/forceatlas.hpp:127-140); rows outside the
    // owned block contribute nothing else.
    std::vector<double> deg_own;
    const bool weighted = p.use_weights && A.data != nullptr;
    if (reorder) shared_deg = nullptr;  // (the caller's row sums are in the caller's numbering)
    if (shared_deg == nullptr) deg_own.resize(n_);
    double* deg_w = deg_own.data();
    const double* deg_r = shared_deg ? shared_deg : deg_own.data();
    if (shared_deg == nullptr) {
      // row sums in CSR order (bit-identical whatever the thread count: rows are independent); a
      // few host threads on large graphs, where this pass is milliseconds of every plan creation
      auto rows = [&](int i0, int i1) {
        for (int i = i0; i < i1; ++i) {
          double s = 0.0;
          if (weighted) {
            for (int e = I[i]; e < I[i + 1]; ++e) s += Dw[e];
          } else {
            s = 1.0 * (I[i + 1] - I[i]);
          }
          deg_w[i] = s;
        }
      };
      const int nt = (weighted && I[n_] > (1 << 20)) ? 8 : 1;
      if (nt == 1) {
        rows(0, n_);
      } else {
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t)
          pool.emplace_back(rows, (int)((int64_t)n_ * t / nt), (int)((int64_t)n_ * (t + 1) / nt));
        for (auto& th : pool) th.join();
      }
    }
    DevBuf<double> d_deg(ctx, std::max(n_, 1));
    d_deg.upload(ctx, deg_r, n_);
    mass_.alloc(ctx, (size_t)NM * ld_);
    k_mass_from_degree<T><<<(unsigned)((ld_ + 255) / 256), 256, 0, ctx->stream>>>(
        d_deg.get(), n_, ld_, NM, mass_.get());
    ctx->launches++;

    lap("masses");
    // owned CSR rows, re-based to local entry offsets
    const int e0 = I[rb_], e1 = I[re_];
    const int lnnz = e1 - e0;
    std::vector<int> rowptr(nrows_ + 1);
    for (int r = 0; r <= nrows_; ++r) rowptr[r] = I[rb_ + r] - e0;
    rowptr_.alloc(ctx, nrows_ + 1);
    rowptr_.upload(ctx, rowptr.data(), nrows_ + 1);
    J_.alloc(ctx, lnnz + 8);  // + padding: the staged kernel copies 16-byte-aligned slices
    GE_CUDA(cudaMemsetAsync(J_.get() + lnnz, 0, 8 * sizeof(int), ctx->stream));
    J_.upload(ctx, J + e0, lnnz);
    std::vector<T> w;
    if (weighted) {
      W_.alloc(ctx, lnnz + 8);
      GE_CUDA(cudaMemsetAsync(W_.get() + lnnz, 0, 8 * sizeof(T), ctx->stream));
      if (std::is_same<T, double>::value) {  // no conversion: straight from the caller's array
        W_.upload(ctx, reinterpret_cast<const T*>(Dw + e0), lnnz);
      } else {
        w.resize(lnnz);
        for (int e = 0; e < lnnz; ++e) w[e] = (T)Dw[e0 + e];
        W_.upload(ctx, w.data(), lnnz);
      }
    }
    lap("graph upload");
    // Rows far longer than the rest (power-law graphs) get a CTA each; the row kernels skip them.
    {
      // three tiers by row length: the row kernels (G lanes per row, chunked) up to long_threshold_
      // entries, one warp per row up to 8x that, one 512-thread CTA per row beyond
      long_threshold_ = env_int("GE_LONG_ROW", 96);
      const int cta_threshold = 8 * long_threshold_;
      std::vector<int> lrows, mrows;
      int64_t long_entries = 0;
      if (long_threshold_ > 0)
        for (int r = 0; r < nrows_; ++r) {
          const int len = rowptr[r + 1] - rowptr[r];
          if (len > long_threshold_) {
            (len > cta_threshold ? lrows : mrows).push_back(r);
            long_entries += len;
          }
        }
      nlong_ = (int)lrows.size();
      nmid_ = (int)mrows.size();
      if (nlong_ > 0) {
        long_rows_.alloc(ctx, lrows.size());
        long_rows_.upload(ctx, lrows.data(), lrows.size());
      }
      if (nmid_ > 0) {
        mid_rows_.alloc(ctx, mrows.size());
        mid_rows_.upload(ctx, mrows.data(), mrows.size());
      }
      if (nlong_ + nmid_ == 0) long_threshold_ = 0;
      const int nshort = nrows_ - nlong_ - nmid_;
      avg_deg_ = nshort > 0 ? double(lnnz - long_entries) / nshort : 0.0;
      // Gather locality after the renumbering: the mean distance |i - j| over the entries.  2-D
      // geometric graphs stay within a few hundred positions (the neighbours' coordinates are L1
      // hits and the SoA gathers are cheapest); 3-D meshes and power-law graphs do not (Delaunay:
      // ~9000, R-MAT: ~26000), every gathered coordinate then costs a 32-byte L2 sector per
      // dimension, and the interleaved copy (one sector per neighbour) pays for its extra write.
      double span = 0.0;
      int64_t sampled = 0;
      const int stride = nrows_ > (1 << 16) ? 16 : 1;  // a sample of the rows is enough for a threshold
      for (int r = 0; r < nrows_; r += stride)
        for (int e = rowptr[r]; e < rowptr[r + 1]; ++e, ++sampled) span += std::abs((rb_ + r) - J[e0 + e]);
      mean_span_ = sampled > 0 ? span / sampled : 0.0;
      if (std::getenv("GE_GATHER_COPY_REORDERED") == nullptr)
        gather_copy_reordered_ = mean_span_ > env_int("GE_GATHER_SPAN", 2048);
      if (std::getenv("GE_VERBOSE"))
        std::fprintf(stderr, "[ge] flat plan rows=%d entries=%d long rows=%d+%d (%.1f%% of the entries) "
                             "avg degree of the rest %.1f mean |i-j| %.0f gather copy %d\n",
                     nrows_, lnnz, nmid_, nlong_, lnnz ? 100.0 * long_entries / lnnz : 0.0, avg_deg_, mean_span_,
                     (int)(use_gather_copy_ && (perm_.size() == 0 || gather_copy_reordered_)));
    }
    GE_CUDA(cudaStreamSynchronize(ctx->stream));  // the renumbered host arrays die with this scope

    lap("row tiers + span");
    aos_[0].alloc(ctx, (size_t)gather_dp() * ld_);
    aos_[1].alloc(ctx, (size_t)gather_dp() * ld_);
    own0_.alloc(ctx, (size_t)dim_ * ld_);
    own1_.alloc(ctx, (size_t)dim_ * ld_);
    own0_.zero(ctx->stream);
    own1_.zero(ctx->stream);
    buf_[0] = own0_.get();
    buf_[1] = own1_.get();
    Frep_.alloc(ctx, (size_t)dim_ * ldf_);
    Fprev_.alloc(ctx, (size_t)dim_ * ldf_);
    Frep_.zero(ctx->stream);
    Fprev_.zero(ctx->stream);
    stage_.alloc(ctx, (size_t)std::max(n_, 1) * dim_);

    lap("buffers");
    // Whole-graph plans on large graphs evaluate every unordered pair once (ge_flat_sym.cu); the
    // column-side scratch grows with n^2 / 2048, so very large graphs (and row-block plans, whose
    // pairs are not closed under transposition) keep the ordered sweep.
    {
      const char* e = std::getenv("GE_REP_SYM");
      // parts > 1: this plan is rank `part` of a symmetric multi-rank solve -- it evaluates its
      // share of the unordered pairs over the FULL length and the ranks add their sums
      // (reduce-scatter) before the step; row-block plans (parts == 1, rows != all) cannot.
      bool want = (e ? std::atoi(e) != 0 : true) && (parts_ > 1 || (rb_ == 0 && re_ == n_)) &&
                  n_ >= env_int("GE_SYM_MIN_ROWS", 32768);
      if (want) {
        size_t free_b = 0, total_b = 0;
        GE_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const double need = RepulsionSymPlan<T>::scratch_bytes(dim_, ld_, parts_);
        want = need < 0.5 * double(free_b) && need < 1e9 * env_int("GE_SYM_MAX_SCRATCH_GB", 16);
      }
      GE_REQUIRE(want || parts_ == 1, "symmetric multi-rank plan refused (graph too small / too "
                                      "large for the column scratch, or GE_REP_SYM=0)");
      if (want) {
        sym_.reset(new RepulsionSymPlan<T>(ctx, dim_, ld_, part, parts_));
        S_.alloc(ctx, (size_t)dim_ * ld_);
        sums_ = S_.get();
      } else {
        rep_.reset(new RepulsionPlan<T>(ctx, dim_, {RowSegment{rb_, re_, 0, (int)ld_}}));
      }
    }
    for (auto& e : ev_) GE_CUDA(cudaEventCreate(&e));
    GE_CUDA(cudaStreamSynchronize(ctx->stream));
    lap("repulsion plan");
  }
  ~FlatSolverT() override {
    for (auto& e : ev_)
      if (e) cudaEventDestroy(e);
  }

  int64_t ld() const override { return ld_; }
  int elem_size() const override { return (int)sizeof(T); }
  void bind_coords(void* b0, void* b1) override {
    buf_[0] = (T*)b0;
    buf_[1] = (T*)b1;
    GE_CUDA(cudaMemsetAsync(b0, 0, (size_t)dim_ * ld_ * sizeof(T), ctx->stream));
    GE_CUDA(cudaMemsetAsync(b1, 0, (size_t)dim_ * ld_ * sizeof(T), ctx->stream));
    cur_ = 0;
    aos_valid_ = stepped_ = false;
  }
  void upload_coords(const double* aos) override {
    stage_.upload(ctx, aos, (size_t)n_ * dim_);
    const int64_t total = ld_ * dim_;
    k_aos_to_soa<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        stage_.get(), n_, dim_, ld_, buf_[cur_], perm_.size() ? perm_.get() : nullptr);
    ctx->launches++;
    // the other buffer must hold valid data outside the owned rows when ranks exchange slices
    GE_CUDA(cudaMemcpyAsync(buf_[cur_ ^ 1], buf_[cur_], (size_t)total * sizeof(T),
                            cudaMemcpyDeviceToDevice, ctx->stream));
    Fprev_.zero(ctx->stream);
    aos_valid_ = stepped_ = false;
  }
  void download_coords(double* aos) override {
    const int64_t total = (int64_t)n_ * dim_;
    if (total == 0) return;
    k_soa_to_aos<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        buf_[cur_], n_, dim_, ld_, stage_.get(), perm_.size() ? perm_.get() : nullptr);
    ctx->launches++;
    stage_.download(ctx, aos, (size_t)total);
    GE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  void download_forces(double* aos) override {
    const int64_t total = (int64_t)nrows_ * dim_;
    if (total == 0) return;
    k_soa_to_aos<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        Fprev_.get(), nrows_, dim_, ldf_, stage_.get(), perm_.size() ? perm_.get() : nullptr);
    ctx->launches++;
    stage_.download(ctx, aos, (size_t)total);
    GE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  int gather_dp() const { return dim_ == 2 ? 2 : 4; }
  void refresh_gather_copy() {
    const unsigned grid = (unsigned)((ld_ + 255) / 256);
    if (dim_ == 2) k_soa_to_gather<T, 2><<<grid, 256, 0, ctx->stream>>>(buf_[cur_], ld_, aos_[cur_].get());
    else k_soa_to_gather<T, 3><<<grid, 256, 0, ctx->stream>>>(buf_[cur_], ld_, aos_[cur_].get());
    ctx->launches++;
  }
  void* cur_coords() override { return buf_[cur_]; }
  void* next_coords() override { return buf_[cur_ ^ 1]; }
  void swap() override {
    cur_ ^= 1;
    aos_valid_ = stepped_;  // the step kernel just wrote the owned rows of the new current copy
    stepped_ = false;
  }

  bool symmetric() const override { return sym_ != nullptr; }
  void bind_pair_sums(void* full) override {
    GE_REQUIRE(sym_ != nullptr, "not a symmetric plan");
    sums_ = (T*)full;
  }
  void* pair_sums() override { return sums_; }
  void launch_iteration(bool update) override {
    GE_REQUIRE(!(sym_ && parts_ > 1),
               "multi-rank symmetric plan: call launch_repulsion, add the ranks' sums, then launch_step");
    launch_repulsion();
    launch_step(update);
  }
  void launch_repulsion() override {
    if (nrows_ == 0 && !sym_) return;
    if (prof_) GE_CUDA(cudaEventRecord(ev_[0], ctx->stream));
    if (kernel_mask_ & 1) {
      if (sym_) sym_->launch(buf_[cur_], mass_.get(), sums_, ph_.eps2);
      else rep_->launch(buf_[cur_], mass_.get(), ld_, Frep_.get(), ldf_, rb_, ph_.repel, ph_.eps2);
    }
    if (prof_) {  // read back after the step, so that the step queues behind the sweep as usual
      GE_CUDA(cudaEventRecord(ev_[1], ctx->stream));
      rep_pending_ = true;
    }
  }
  void launch_step(bool update) override {
    if (nrows_ == 0) return;
    if (prof_) GE_CUDA(cudaEventRecord(ev_[2], ctx->stream));
    StepArgs<T> sa;
    sa.e_begin = rowptr_.get();
    sa.e_end = rowptr_.get() + 1;
    sa.J = J_.get();
    sa.W = W_.size() ? W_.get() : nullptr;
    sa.pos_cur = buf_[cur_];
    sa.pos_next = buf_[cur_ ^ 1];
    if (use_gather_copy_ && (perm_.size() == 0 || gather_copy_reordered_)) {
      // (with the breadth-first renumbering the SoA gathers are already local and the copy only
      // costs its own write traffic: 0.158 ms without vs 0.182 ms with, n = 2M)
      // The copy of the current buffer is exact when this plan wrote every row of it in the
      // previous step; otherwise (first step, or rows updated by other ranks) it is rebuilt.
      if (!(aos_valid_ && nrows_ == n_)) refresh_gather_copy();
      sa.aos_cur = aos_[cur_].get();
      sa.aos_next = aos_[cur_ ^ 1].get();
    }
    sa.Frep = Frep_.get();
    if (sym_) {  // raw pair sums over the full length; the step kernel applies c_i * repel
      sa.Frep = sums_ + rb_;
      sa.ldr = ld_;
      sa.frep_scale = ph_.repel;
    }
    sa.Fprev = Fprev_.get();
    sa.mass = mass_.get();
    sa.Eext = nullptr;
    sa.ld = ld_;
    sa.ldf = ldf_;
    sa.row0 = rb_;
    sa.nrows = nrows_;
    sa.update = update ? 1 : 0;
    sa.ph = ph_;
    sa.long_threshold = long_threshold_;
    if ((kernel_mask_ & 2) && nlong_ > 0) launch_attract_step_long<T>(ctx, sa, dim_, long_rows_.get(), nlong_, 512);
    if ((kernel_mask_ & 2) && nmid_ > 0) launch_attract_step_long<T>(ctx, sa, dim_, mid_rows_.get(), nmid_, 32);
    if (kernel_mask_ & 2) {
      // one lane per row while a 256-row chunk fits the staging buffer (fewest instructions per
      // row: measured 58-63 % of the HBM peak against 57-61 % with two lanes), else 2-8 lanes
      const int group = env_int("GE_STEP_GROUP", avg_deg_ * 256 <= 0.8 * 4096 ? 1 : group_for_degree(avg_deg_));
      // staged (TMA) variant: groups up to 8 lanes; chunks of 256/G rows must mostly fit kStepCap
      if (staged_step_ && group <= 8 &&
          avg_deg_ * (256 / std::max(group, 1)) <= 0.8 * (group <= 1 ? 4096 : kStepCap))
        launch_attract_step_staged<T>(ctx, sa, dim_, group);
      else
        launch_attract_step<T>(ctx, sa, dim_, group, false);
      stepped_ = update;
    }
    if (prof_) {
      GE_CUDA(cudaEventRecord(ev_[3], ctx->stream));
      GE_CUDA(cudaEventSynchronize(ev_[3]));
      float t_step = 0, t_rep = 0;
      GE_CUDA(cudaEventElapsedTime(&t_step, ev_[2], ev_[3]));
      step_ms_ += t_step;
      step_n_++;
      if (rep_pending_) {
        GE_CUDA(cudaEventElapsedTime(&t_rep, ev_[0], ev_[1]));
        rep_ms_ += t_rep;
        rep_n_++;
        rep_pending_ = false;
      }
    }
  }

  void normalize() override {
    if (dim_ == 2)
      k_normalize<T, 2><<<1, 1024, 0, ctx->stream>>>(buf_[cur_], n_, ld_);
    else
      k_normalize<T, 3><<<1, 1024, 0, ctx->stream>>>(buf_[cur_], n_, ld_);
    ctx->launches++;
    GE_CUDA(cudaGetLastError());
  }
  void select_kernels(int mask) override { kernel_mask_ = mask; }
  void profile(bool enable) override {
    prof_ = enable;
    rep_ms_ = step_ms_ = 0;
    rep_n_ = step_n_ = 0;
  }
  void profile_get(double* rep_ms, int64_t* rep_n, double* step_ms, int64_t* step_n) override {
    if (rep_ms) *rep_ms = rep_ms_;
    if (rep_n) *rep_n = rep_n_;
    if (step_ms) *step_ms = step_ms_;
    if (step_n) *step_n = step_n_;
  }

 private:
  int n_, dim_, rb_, re_, parts_ = 1, nrows_ = 0;
  T* sums_ = nullptr;  // [dim][ld] raw pair sums of the symmetric sweep (own S_ or caller-bound)
  int64_t ld_ = 0, ldf_ = 0;
  Physics<T> ph_;
  double avg_deg_ = 0;
  DevBuf<T> mass_, own0_, own1_, Frep_, Fprev_, W_, aos_[2];
  bool use_gather_copy_ = std::getenv("GE_NO_GATHER_COPY") == nullptr;
  bool staged_step_ = env_int("GE_STEP_STAGED", 1) != 0;
  bool gather_copy_reordered_ = env_int("GE_GATHER_COPY_REORDERED", 0) != 0;
  bool aos_valid_ = false, stepped_ = false;
  DevBuf<int> rowptr_, J_, perm_, long_rows_, mid_rows_;
  int long_threshold_ = 0, nlong_ = 0, nmid_ = 0;
  double mean_span_ = 0.0;
  std::unique_ptr<RepulsionPlan<T>> rep_;
  std::unique_ptr<RepulsionSymPlan<T>> sym_;
  DevBuf<T> S_;
  DevBuf<double> stage_;
  T* buf_[2] = {nullptr, nullptr};
  int cur_ = 0;
  bool prof_ = false;
  int kernel_mask_ = 3;
  double rep_ms_ = 0, step_ms_ = 0;
  int64_t rep_n_ = 0, step_n_ = 0;
  cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
  bool rep_pending_ = false;
};

}  // namespace

FlatSolver* make_flat_solver(ge_context* ctx, const ge_csr& A, int dim, const ge_params& p,
                             int row_begin, int row_end, int part, int parts, const double* shared_deg) {
  GE_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
  GE_REQUIRE(A.rows == A.cols, "A must be square");
  GE_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= A.rows, "bad row block");
  GE_REQUIRE(parts >= 1 && part >= 0 && part < parts, "bad rank / world size");
  if (p.precision == GE_F32)
    return new FlatSolverT<float>(ctx, A, dim, p, row_begin, row_end, part, parts, shared_deg);
  return new FlatSolverT<double>(ctx, A, dim, p, row_begin, row_end, part, parts, shared_deg);
}

// include/forceatlas.hpp:127-140: weighted row sums (or row lengths) of every row, in CSR order; the
// multi-device solve computes them once and hands them to every device's plan.
void flat_degrees(const ge_csr& A, const ge_params& p, std::vector<double>& deg) {
  const int n = A.rows;
  deg.resize(std::max(n, 1));
  const bool weighted = p.use_weights && A.data != nullptr;
  auto rows = [&](int i0, int i1) {
    for (int i = i0; i < i1; ++i) {
      double s = 0.0;
      if (weighted) {
        for (int e = A.indptr[i]; e < A.indptr[i + 1]; ++e) s += A.data[e];
      } else {
        s = 1.0 * (A.indptr[i + 1] - A.indptr[i]);
      }
      deg[i] = s;
    }
  };
  const int nt = (weighted && A.indptr[n] > (1 << 20)) ? 8 : 1;
  if (nt == 1) {
    rows(0, n);
    return;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t) pool.emplace_back(rows, (int)((int64_t)n * t / nt), (int)((int64_t)n * (t + 1) / nt));
  for (auto& th : pool) th.join();
}

}  // namespace ge
