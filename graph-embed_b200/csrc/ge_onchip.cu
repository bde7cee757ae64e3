// graph-embed_b200 :: on-chip persistent ForceAtlas solvers (kernel families K3 and K2), sm_100a.
//
// k_onchip_cta   All iterations of one small all-pairs problem inside ONE CTA: positions and
//                masses live in shared memory (double-buffered: one __syncthreads per iteration),
//                L lanes share one vertex's pair loop and combine with warp shuffles, the per-vertex
//                state (x, previous force, mass) stays in registers.  Zero launches and zero global
//                synchronisation per iteration.  Used for
//                  * the coarsest level: partition::forceAtlas with its default 100 000 iterations on
//                    n ~ 30-100 vertices (/root/reference/include/forceatlas.hpp:146-270, called
//                    from src/embed.cpp:586), which dominates embed() wall time, and
//                  * aggregates of 33..1024 members in forceAtlasMultilevel (:390-538) followed by
//                    the centre / max-normalise / prolongation epilogue (:539-570).
// k_onchip_warp  The same for aggregates of 2..32 members: one lane per member, several
//                equal-size aggregates packed into one warp, __syncwarp instead of __syncthreads.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cooperative_groups.h>

#include "ge_onchip.cuh"
#include "ge_tma.cuh"

namespace ge {

namespace {

// ---- position exchange between the CTAs of a cluster without a cluster barrier -------------------
// st.async writes one value into a peer's shared memory and completes its bytes on the PEER's
// mbarrier; the receiver waits on its own mbarrier for the bytes of one iteration.  Measured
// (tools/micro/latency.cu, 8 CTAs): 327 cycles per exchange against 684 for plain DSMEM stores +
// barrier.cluster (whose arrive has to drain the stores first: ERRBAR was 20 % of the stall samples
// of the cluster kernels).
__device__ __forceinline__ uint32_t cluster_addr(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async(uint32_t dst, double v, uint32_t peer_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(dst),
               "l"(__double_as_longlong(v)), "r"(peer_bar)
               : "memory");
}
__device__ __forceinline__ void st_async(uint32_t dst, float v, uint32_t peer_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst),
               "r"(__float_as_uint(v)), "r"(peer_bar)
               : "memory");
}

// Shared-memory column records: positions are AoS with DP = 2 (d = 2) or 4 (d = 3, one pad) reals
// per column, masses AoS (c, 1.5c, 1.875c, pad) in FP64 / (c) in FP32, so that one column is
// fetched with 16-byte shared loads; both arrays are padded with zero-mass columns up to a
// multiple of the pair-loop trip width, which removes every bounds predicate from the loop.
template <typename T, int D>
struct Col;
template <>
struct Col<double, 2> {
  static constexpr int DP = 2, MP = 4;
  __device__ __forceinline__ static void pos(const double* b, int j, double (&x)[2]) {
    const double2 v = reinterpret_cast<const double2*>(b)[j];
    x[0] = v.x;
    x[1] = v.y;
  }
  __device__ __forceinline__ static void mass(const double* b, int j, double& m0, double& m1, double& m2) {
    const double2 v = reinterpret_cast<const double2*>(b)[2 * j];
    m0 = v.x;
    m1 = v.y;
    m2 = b[4 * j + 2];
  }
};
template <>
struct Col<double, 3> {
  static constexpr int DP = 4, MP = 4;
  __device__ __forceinline__ static void pos(const double* b, int j, double (&x)[3]) {
    const double2 v = reinterpret_cast<const double2*>(b)[2 * j];
    x[0] = v.x;
    x[1] = v.y;
    x[2] = b[4 * j + 2];
  }
  __device__ __forceinline__ static void mass(const double* b, int j, double& m0, double& m1, double& m2) {
    Col<double, 2>::mass(b, j, m0, m1, m2);
  }
};
template <>
struct Col<float, 2> {
  static constexpr int DP = 2, MP = 1;
  __device__ __forceinline__ static void pos(const float* b, int j, float (&x)[2]) {
    const float2 v = reinterpret_cast<const float2*>(b)[j];
    x[0] = v.x;
    x[1] = v.y;
  }
  __device__ __forceinline__ static void mass(const float* b, int j, float& m0, float& m1, float& m2) {
    m0 = m1 = m2 = b[j];
  }
};
template <>
struct Col<float, 3> {
  static constexpr int DP = 4, MP = 1;
  __device__ __forceinline__ static void pos(const float* b, int j, float (&x)[3]) {
    const float4 v = reinterpret_cast<const float4*>(b)[j];
    x[0] = v.x;
    x[1] = v.y;
    x[2] = v.z;
  }
  __device__ __forceinline__ static void mass(const float* b, int j, float& m0, float& m1, float& m2) {
    m0 = m1 = m2 = b[j];
  }
};

__host__ __device__ inline int onchip_spad(int s, int L, int U) { return (s + L * U - 1) / (L * U) * (L * U); }

// BIG = false: up to 512 threads (128 registers each), U = 4 independent pairs per trip of the pair
// loop (instruction-level parallelism); BIG = true: up to 1024 threads (64 registers), U = 2 and
// twice the resident warps (thread-level parallelism) for the larger problems.
// GA: general attraction (linlog / delta / nohubs) compiled in.
template <typename T, int D, bool ML, int L, bool BIG, bool GA>
__global__ void __launch_bounds__(BIG ? 1024 : 512) k_onchip_cta(const OnchipArgs<T> a) {
  constexpr int NM = Real<T>::kMassArrays;
  constexpr int U = BIG ? 2 : 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ double bc[D + 1];

  const int4 task = a.tasks[blockIdx.x];
  const int slot0 = task.x, s = task.y, agg = task.z;
  using C = Col<T, D>;
  constexpr int DP = C::DP, MP = C::MP;
  const int S = onchip_spad(s, L, U);
  T* pos = reinterpret_cast<T*>(smem_raw);  // [2][S][DP]
  T* ms = pos + 2 * S * DP;                 // [S][MP]

  const int tid = threadIdx.x;
  const int lv = tid / L;
  const int part = tid % L;
  const bool owner = lv < s;
  const int slot = slot0 + (owner ? lv : 0);
  const int v = a.vtx ? a.vtx[slot] : slot;
  const Physics<T> ph = a.ph;  // registers, not constant-bank reloads inside the loop

  T x[D], fprev[D], E[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    x[k] = (T)a.init_aos[(int64_t)v * D + k];
    fprev[k] = (T)0;
    E[k] = (ML && a.Eext) ? a.Eext[(int64_t)k * a.ld + slot] : (T)0;
  }
  const T ci = a.mass[slot];
  const T ci_repel = ci * ph.repel;
  const int eb = owner ? a.e_begin[slot] : 0;
  const int ee = owner ? a.e_end[slot] : 0;
  // The first two attraction entries of this lane live in registers (coarse graphs have ~1-2
  // entries per lane); longer rows continue from the compact list through L1.
  int ej[2] = {0, 0};
  T ew[2] = {(T)0, (T)0};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int e = eb + part + q * L;
    if (e < ee) {
      ej[q] = a.e_idx[e] - slot0;
      ew[q] = (a.e_w != nullptr && ph.use_weights) ? a.e_w[e] : (T)1;
    }
  }
  for (int i = tid; i < 2 * S * DP + S * MP; i += blockDim.x) pos[i] = (T)0;  // zero-mass padding
  __syncthreads();
  if (owner && part == 0) {
#pragma unroll
    for (int k = 0; k < D; ++k) pos[lv * DP + k] = x[k];
    ms[lv * MP] = ci;
    if (NM > 1) ms[lv * MP + (NM > 1 ? 1 : 0)] = (T)1.5 * ci;
    if (NM > 2) ms[lv * MP + (NM > 2 ? 2 : 0)] = (T)1.875 * ci;
  }
  __syncthreads();

  const T* pc = pos;
  T* pn = pos + S * DP;
  const int iters = a.forces_only ? 1 : a.iters;
  for (int it = 0; it < iters; ++it) {
    T f[D];
#pragma unroll
    for (int k = 0; k < D; ++k) f[k] = (T)0;
    {  // every lane runs the pair loop (lanes without a vertex compute a discarded row), so the
       // warp-wide clamp vote inside it is executed convergently
      for (int j0 = part; j0 < ((a.debug_skip & 1) ? 0 : S); j0 += L * U) {
        T d[U][D], r2[U], s3[U], m0[U], m1[U], m2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * L;
          T xj[D];
          C::pos(pc, j, xj);
          C::mass(ms, j, m0[u], m1[u], m2[u]);
          r2[u] = (T)0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            d[u][k] = x[k] - xj[k];
            r2[u] = fma(d[u][k], d[u][k], r2[u]);
          }
        }
        // one warp-uniform clamp test for the U pairs; it also starts a basic block in which the
        // U reciprocal-square-root chains below stay interleaved (ILP = U)
        Real<T>::template clamp_lo_n<U>(r2, ph.eps2);
        Real<T>::template inv_cube_mass_v<U>(r2, m0, m1, m2, s3);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int k = 0; k < D; ++k) f[k] = fma(d[u][k], s3[u], f[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < D; ++k) f[k] *= ci_repel;
    }
    if (owner) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (eb + part + q * L < ee) {
          T xj[D], d[D];
          C::pos(pc, ej[q], xj);
          T r2 = (T)0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            d[k] = xj[k] - x[k];
            r2 = fma(d[k], d[k], r2);
          }
          const T g = attraction_factor<T, GA>(r2, ew[q], ci, ph);
#pragma unroll
          for (int k = 0; k < D; ++k) f[k] = fma(d[k], g, f[k]);
        }
      }
      for (int e = eb + part + 2 * L; e < ee; e += L) {
        const int j = a.e_idx[e] - slot0;
        const T w = (a.e_w != nullptr && ph.use_weights) ? a.e_w[e] : (T)1;
        T xj[D], d[D];
        C::pos(pc, j, xj);
        T r2 = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          d[k] = xj[k] - x[k];
          r2 = fma(d[k], d[k], r2);
        }
        const T g = attraction_factor<T, GA>(r2, w, ci, ph);
#pragma unroll
        for (int k = 0; k < D; ++k) f[k] = fma(d[k], g, f[k]);
      }
    }
#pragma unroll
    for (int off = L >> 1; off > 0; off >>= 1) {
#pragma unroll
      for (int k = 0; k < D; ++k) f[k] += __shfl_xor_sync(0xffffffffu, f[k], off);
    }
    if (!(a.debug_skip & 2)) vertex_step<T, D, ML>(x, f, fprev, E, ci, ph);
    if (a.forces_only) break;
    if (owner && part == 0) {
#pragma unroll
      for (int k = 0; k < D; ++k) pn[lv * DP + k] = x[k];
    }
    const T* tmp = pc;
    pc = pn;
    pn = const_cast<T*>(tmp);
    if (!(a.debug_skip & 4)) __syncthreads();
  }

  if (a.forces_only) {
    if (owner && part == 0) {
#pragma unroll
      for (int k = 0; k < D; ++k) a.out_aos[(int64_t)v * D + k] = (double)fprev[k];
    }
    return;
  }
  if (!ML && !a.normalize) {
    if (owner && part == 0) {
#pragma unroll
      for (int k = 0; k < D; ++k) a.out_aos[(int64_t)v * D + k] = (double)x[k];
    }
    return;
  }

  // Epilogue: subtract the mean, divide by the largest norm (:272-303 flat / :539-564 multilevel),
  // and for the multilevel kernel map into the parent ball (:565-569).
  auto block_reduce = [&](double val, bool is_max) -> double {
    for (int off = 16; off > 0; off >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, val, off);
      val = is_max ? fmax(val, o) : val + o;
    }
    if ((tid & 31) == 0) red[tid >> 5] = val;
    __syncthreads();
    if (tid < 32) {
      double w = (tid < (int)((blockDim.x + 31) >> 5)) ? red[tid] : 0.0;
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
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double part_sum = 0.0;
    for (int i = tid; i < s; i += blockDim.x) part_sum += (double)pc[i * DP + k];
    const double tot = block_reduce(part_sum, false);
    if (tid == 0) bc[k] = tot / s;
  }
  __syncthreads();
  double mx = 0.0;
  for (int i = tid; i < s; i += blockDim.x) {
    double m2 = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double c = (double)pc[i * DP + k] - bc[k];
      m2 += c * c;
    }
    mx = fmax(mx, sqrt(m2));
  }
  double maxlen = block_reduce(mx, true);
  if (ML && maxlen < kEpsilon) maxlen = kEpsilon;
  for (int i = tid; i < s; i += blockDim.x) {
    const int vi = a.vtx ? a.vtx[slot0 + i] : slot0 + i;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double c = ((double)pc[i * DP + k] - bc[k]) / maxlen;
      a.out_aos[(int64_t)vi * D + k] =
          ML ? a.cA_aos[(int64_t)agg * D + k] + a.rA[agg] * c : c;
    }
  }
}

// K3 across a thread-block cluster: the coarsest-level flat solve with the vertices split over the
// C CTAs of one cluster (C SMs instead of one).  Every CTA keeps the full, double-buffered position
// array in its own shared memory; after the per-vertex epilogue the owning lanes store the new
// position into the NEXT buffer of every CTA of the cluster through distributed shared memory, and
// one cluster barrier per iteration (arrive.release / wait.acquire) publishes them.  Still zero
// launches and zero global-memory traffic per iteration.
template <typename T, int D, int L, bool GA>
__global__ void __launch_bounds__(512) k_onchip_cluster(const OnchipArgs<T> a, int per_cta) {
  namespace cg = cooperative_groups;
  constexpr int NM = Real<T>::kMassArrays;
  constexpr int U = 4;
  using C = Col<T, D>;
  constexpr int DP = C::DP, MP = C::MP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ double bc[D + 1];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();

  const int4 task = a.tasks[0];
  const int s = task.y;
  const int S = onchip_spad(s, L, U);
  T* pos = reinterpret_cast<T*>(smem_raw);  // [2][S][DP]
  T* ms = pos + 2 * S * DP;                 // [S][MP]

  const int tid = threadIdx.x;
  const int lv = tid / L, part = tid % L;
  const int gv = rank * per_cta + lv;
  const bool owner = lv < per_cta && gv < s;
  const int slot = owner ? gv : 0;
  const Physics<T> ph = a.ph;

  T x[D], fprev[D], E[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    x[k] = (T)a.init_aos[(int64_t)slot * D + k];
    fprev[k] = (T)0;
    E[k] = (T)0;
  }
  const T ci = a.mass[slot];
  const T ci_repel = ci * ph.repel;
  const int eb = owner ? a.e_begin[slot] : 0;
  const int ee = owner ? a.e_end[slot] : 0;
  int ej[2] = {0, 0};
  T ew[2] = {(T)0, (T)0};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int e = eb + part + q * L;
    if (e < ee) {
      ej[q] = a.e_idx[e];
      ew[q] = (a.e_w != nullptr && ph.use_weights) ? a.e_w[e] : (T)1;
    }
  }
  for (int i = tid; i < 2 * S * DP + S * MP; i += blockDim.x) pos[i] = (T)0;
  __syncthreads();
  for (int i = tid; i < s; i += blockDim.x) {  // every CTA loads the whole graph's state
#pragma unroll
    for (int k = 0; k < D; ++k) pos[i * DP + k] = (T)a.init_aos[(int64_t)i * D + k];
    const T c = a.mass[i];
    ms[i * MP] = c;
    if (NM > 1) ms[i * MP + (NM > 1 ? 1 : 0)] = (T)1.5 * c;
    if (NM > 2) ms[i * MP + (NM > 2 ? 2 : 0)] = (T)1.875 * c;
  }
  __shared__ __align__(8) uint64_t xbar[2];  // one mbarrier per position buffer (exchange mode 1)
  const bool async_x = a.exchange == 1;
  if (tid == 0) {
    mbar_init(&xbar[0], 1);
    mbar_init(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();  // also: nobody writes into a peer's shared memory before it is initialised
  const uint32_t pos_addr = smem_addr(pos);
  const uint32_t bar_addr[2] = {smem_addr(&xbar[0]), smem_addr(&xbar[1])};
  const uint32_t xbytes = (uint32_t)s * D * (uint32_t)sizeof(T);
  unsigned xphase = 0u;  // bit b: parity of the next completion of xbar[b]

  const T* pc = pos;
  T* pn = pos + S * DP;
  int nxt = 1;
  // own-position part of the step (:205-211): the gravity factor gravity * c_i / |x_i| needs only
  // the vertex's own position, so it is formed while the new positions travel to the peers
  T grav;
  {
    T m2 = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) m2 = fma(x[k], x[k], m2);
    grav = ph.gravity * ci * Real<T>::rsqrt_acc(m2);
  }
  for (int it = 0; it < a.iters; ++it) {
    if (async_x && tid == 0) mbar_expect_tx(&xbar[nxt], xbytes);  // this iteration's arrivals
    T f[D];
#pragma unroll
    for (int k = 0; k < D; ++k) f[k] = (T)0;
    {  // every lane runs the pair loop (lanes without a vertex compute a discarded row), so the
       // warp-wide clamp vote inside it is executed convergently
      for (int j0 = part; j0 < ((a.debug_skip & 1) ? 0 : S); j0 += L * U) {
        T d[U][D], r2[U], s3[U], m0[U], m1[U], m2[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = j0 + u * L;
          T xj[D];
          C::pos(pc, j, xj);
          C::mass(ms, j, m0[u], m1[u], m2[u]);
          r2[u] = (T)0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            d[u][k] = x[k] - xj[k];
            r2[u] = fma(d[u][k], d[u][k], r2[u]);
          }
        }
        // one warp-uniform clamp test for the U pairs; it also starts a basic block in which the
        // U reciprocal-square-root chains below stay interleaved (ILP = U)
        Real<T>::template clamp_lo_n<U>(r2, ph.eps2);
        Real<T>::template inv_cube_mass_v<U>(r2, m0, m1, m2, s3);
#pragma unroll
        for (int u = 0; u < U; ++u) {
#pragma unroll
          for (int k = 0; k < D; ++k) f[k] = fma(d[u][k], s3[u], f[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < D; ++k) f[k] *= ci_repel;
    }
    if (owner) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (eb + part + q * L < ee) {
          T xj[D], d[D];
          C::pos(pc, ej[q], xj);
          T r2 = (T)0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            d[k] = xj[k] - x[k];
            r2 = fma(d[k], d[k], r2);
          }
          const T g = attraction_factor<T, GA>(r2, ew[q], ci, ph);
#pragma unroll
          for (int k = 0; k < D; ++k) f[k] = fma(d[k], g, f[k]);
        }
      }
      for (int e = eb + part + 2 * L; e < ee; e += L) {
        const int j = a.e_idx[e];
        const T w = (a.e_w != nullptr && ph.use_weights) ? a.e_w[e] : (T)1;
        T xj[D], d[D];
        C::pos(pc, j, xj);
        T r2 = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          d[k] = xj[k] - x[k];
          r2 = fma(d[k], d[k], r2);
        }
        const T g = attraction_factor<T, GA>(r2, w, ci, ph);
#pragma unroll
        for (int k = 0; k < D; ++k) f[k] = fma(d[k], g, f[k]);
      }
    }
#pragma unroll
    for (int off = L >> 1; off > 0; off >>= 1) {
#pragma unroll
      for (int k = 0; k < D; ++k) f[k] += __shfl_xor_sync(0xffffffffu, f[k], off);
    }
    if (!(a.debug_skip & 2)) {  // vertex_step<T, D, false> with the gravity factor prepared ahead
      T sw2 = (T)0, f2 = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const T fk = fma(-x[k], grav, f[k]);
        const T dk = fk - fprev[k];
        sw2 = fma(dk, dk, sw2);
        f2 = fma(fk, fk, f2);
        f[k] = fk;
      }
      const T swing = sw2 > (T)0 ? sw2 * Real<T>::rsqrt_acc(sw2) : sw2;
      const T ssw = swing > (T)0 ? swing * Real<T>::rsqrt_acc(swing) : swing;
      T speed = ph.ks * ph.gspeed * Real<T>::rcp_acc((T)1 + ph.gspeed * ssw);
      const T cap = ph.ksmax * Real<T>::rsqrt_acc(f2);
      if (speed > cap) speed = cap;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        x[k] = fma(f[k], speed, x[k]);
        fprev[k] = f[k];
      }
    }
    if (owner) {  // the L lanes of the group share out the csize peer stores
      for (int rr = part; rr < csize; rr += L) {
        if (async_x) {
          const uint32_t dst = cluster_addr(pos_addr + (uint32_t)(((size_t)nxt * S + gv) * DP * sizeof(T)), rr);
          const uint32_t pbar = cluster_addr(bar_addr[nxt], rr);
#pragma unroll
          for (int k = 0; k < D; ++k) st_async(dst + k * (uint32_t)sizeof(T), x[k], pbar);
        } else {
          T* dst = cluster.map_shared_rank(pos, rr) + (size_t)nxt * S * DP + (size_t)gv * DP;
#pragma unroll
          for (int k = 0; k < D; ++k) dst[k] = x[k];
        }
      }
    }
    {  // while the positions travel: next iteration's gravity factor
      T m2 = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) m2 = fma(x[k], x[k], m2);
      grav = ph.gravity * ci * Real<T>::rsqrt_acc(m2);
    }
    if (async_x) {
      // every vertex's new position arrives through st.async: wait for this buffer's bytes.  No
      // CTA can be more than one iteration ahead (it needs everybody's positions to go on), so the
      // two buffers / barriers never see traffic of two different iterations.
      mbar_wait(&xbar[nxt], (xphase >> nxt) & 1u);
      xphase ^= 1u << nxt;
    } else {
      cluster.sync();
    }
    const T* tmp = pc;
    pc = pn;
    pn = const_cast<T*>(tmp);
    nxt ^= 1;
  }
  if (async_x) cluster.sync();  // nobody leaves while a peer may still be storing into it

  if (!a.normalize) {
    if (owner && part == 0) {
#pragma unroll
      for (int k = 0; k < D; ++k) a.out_aos[(int64_t)gv * D + k] = (double)x[k];
    }
    return;
  }
  if (rank != 0) return;  // include/forceatlas.hpp:272-303 by CTA 0 (every CTA holds all positions)
  auto block_reduce = [&](double val, bool is_max) -> double {
    for (int off = 16; off > 0; off >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, val, off);
      val = is_max ? fmax(val, o) : val + o;
    }
    if ((tid & 31) == 0) red[tid >> 5] = val;
    __syncthreads();
    if (tid < 32) {
      double w = (tid < (int)((blockDim.x + 31) >> 5)) ? red[tid] : 0.0;
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
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double part_sum = 0.0;
    for (int i = tid; i < s; i += blockDim.x) part_sum += (double)pc[i * DP + k];
    const double tot = block_reduce(part_sum, false);
    if (tid == 0) bc[k] = tot / s;
  }
  __syncthreads();
  double mx = 0.0;
  for (int i = tid; i < s; i += blockDim.x) {
    double m2 = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double c = (double)pc[i * DP + k] - bc[k];
      m2 += c * c;
    }
    mx = fmax(mx, sqrt(m2));
  }
  const double maxlen = block_reduce(mx, true);
  for (int i = tid; i < s; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < D; ++k)
      a.out_aos[(int64_t)i * D + k] = ((double)pc[i * DP + k] - bc[k]) / maxlen;
  }
}

// Second-generation cluster solve for the coarsest level (what bounds embed() wall time: 100 000
// dependent iterations, /root/reference/include/forceatlas.hpp:146-270 called from
// src/embed.cpp:586).  Same decomposition as k_onchip_cluster; what changed is the length of the
// dependent chain inside one iteration:
//   * the pair loop handles U = 8 (or 4) columns per lane and trip with a branch-free clamp and two
//     accumulator sets, so one warp keeps 8 independent reciprocal-square-root chains in flight
//     (one warp per scheduler: the instruction-level parallelism has to come from inside it);
//   * the part of the per-vertex step that depends only on the vertex's own position (|x|, the
//     gravity factor) is computed between the arrive and the wait of the cluster barrier, where the
//     warp would otherwise idle for the ~400 cycles the barrier takes;
//   * DENSE: coarsest graphs of power-law hierarchies are (nearly) complete (R-MAT-20: 54 vertices,
//     2916 entries), which made the CSR attraction loop as long as the pair loop; the owned rows of
//     the weight matrix are kept in shared memory and folded into the pair loop (one extra FMA per
//     pair: d * (c_i c_j repel / r^3 - attract a_ij)).
template <typename T>
__device__ __forceinline__ void clamp_branchless(T& r2, T lo);
template <>
__device__ __forceinline__ void clamp_branchless<double>(double& r2, double lo) {
  r2 = Real<double>::clamp_lo(r2, lo);
}
template <>
__device__ __forceinline__ void clamp_branchless<float>(float& r2, float lo) {
  r2 = fmaxf(r2, lo);
}

template <typename T, int D, int L, int U, bool GA, bool DENSE>
__global__ void __launch_bounds__(256) k_onchip_cluster2(const OnchipArgs<T> a, int per_cta) {
  namespace cg = cooperative_groups;
  constexpr int NM = Real<T>::kMassArrays;
  using C = Col<T, D>;
  constexpr int DP = C::DP, MP = C::MP;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[32];
  __shared__ double bc[D + 1];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();

  const int4 task = a.tasks[0];
  const int s = task.y;
  const int S = onchip_spad(s, L, U);
  const int WS = S + 8;                      // row stride of the weight rows (bank spread)
  T* pos = reinterpret_cast<T*>(smem_raw);   // [2][S][DP]
  T* ms = pos + 2 * S * DP;                  // [S][MP]
  T* Wm = ms + S * MP;                       // DENSE: [per_cta][WS]  attract * a_ij of the owned rows

  const int tid = threadIdx.x;
  const int lv = tid / L, part = tid % L;
  const int gv = rank * per_cta + lv;
  const bool owner = lv < per_cta && gv < s;
  const int slot = owner ? gv : 0;
  const Physics<T> ph = a.ph;

  T x[D], fprev[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    x[k] = (T)a.init_aos[(int64_t)slot * D + k];
    fprev[k] = (T)0;
  }
  const T ci = a.mass[slot];
  const T ci_repel = ci * ph.repel;
  const int eb = owner ? a.e_begin[slot] : 0;
  const int ee = owner ? a.e_end[slot] : 0;
  int ej[2] = {0, 0};
  T ew[2] = {(T)0, (T)0};
  if (!DENSE) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int e = eb + part + q * L;
      if (e < ee) {
        ej[q] = a.e_idx[e];
        ew[q] = (a.e_w != nullptr && ph.use_weights) ? a.e_w[e] : (T)1;
      }
    }
  }
  const int smem_elems = 2 * S * DP + S * MP + (DENSE ? per_cta * WS : 0);
  for (int i = tid; i < smem_elems; i += blockDim.x) pos[i] = (T)0;
  __syncthreads();
  for (int i = tid; i < s; i += blockDim.x) {  // every CTA loads the whole graph's state
#pragma unroll
    for (int k = 0; k < D; ++k) pos[i * DP + k] = (T)a.init_aos[(int64_t)i * D + k];
    const T c = a.mass[i];
    ms[i * MP] = c;
    if (NM > 1) ms[i * MP + (NM > 1 ? 1 : 0)] = (T)1.5 * c;
    if (NM > 2) ms[i * MP + (NM > 2 ? 2 : 0)] = (T)1.875 * c;
  }
  if (DENSE && owner) {
    for (int e = eb + part; e < ee; e += L) {
      const T w = (a.e_w != nullptr && ph.use_weights) ? a.e_w[e] : (T)1;
      atomicAdd(&Wm[lv * WS + a.e_idx[e]], ph.attract * w);
    }
  }
  __shared__ __align__(8) uint64_t xbar[2];  // one mbarrier per position buffer (exchange mode 1)
  const bool async_x = a.exchange == 1;
  if (tid == 0) {
    mbar_init(&xbar[0], 1);
    mbar_init(&xbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();  // also: nobody writes into a peer's shared memory before it is initialised
  const uint32_t pos_addr = smem_addr(pos);
  const uint32_t bar_addr[2] = {smem_addr(&xbar[0]), smem_addr(&xbar[1])};
  const uint32_t xbytes = (uint32_t)s * D * (uint32_t)sizeof(T);
  unsigned xphase = 0u;
  const T* wrow = Wm + (owner ? lv : 0) * WS;

  const T* pc = pos;
  T* pn = pos + S * DP;
  int nxt = 1;
  // own-position part of the step (:205-211): 1/|x| and the gravity factor
  T m2 = (T)0;
#pragma unroll
  for (int k = 0; k < D; ++k) m2 = fma(x[k], x[k], m2);
  T grav = ph.gravity * ci * Real<T>::rsqrt_acc(m2);
  for (int it = 0; it < a.iters; ++it) {
    if (async_x && tid == 0) mbar_expect_tx(&xbar[nxt], xbytes);  // this iteration's arrivals
    T f0[D], f1[D], fa[D];
#pragma unroll
    for (int k = 0; k < D; ++k) f0[k] = f1[k] = fa[k] = (T)0;
    if (!DENSE && owner) {  // attraction entries held in registers: independent of the pair loop
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (eb + part + q * L < ee) {
          T xj[D], d[D];
          C::pos(pc, ej[q], xj);
          T r2 = (T)0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            d[k] = xj[k] - x[k];
            r2 = fma(d[k], d[k], r2);
          }
          const T g = attraction_factor<T, GA>(r2, ew[q], ci, ph);
#pragma unroll
          for (int k = 0; k < D; ++k) fa[k] = fma(d[k], g, fa[k]);
        }
      }
    }
    // every lane runs the pair loop (lanes without a vertex compute a discarded row)
    for (int j0 = part; j0 < S; j0 += L * U) {
      T d[U][D], r2[U], s3[U], m0[U], m1[U], m2c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int j = j0 + u * L;
        T xj[D];
        C::pos(pc, j, xj);
        C::mass(ms, j, m0[u], m1[u], m2c[u]);
        r2[u] = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          d[u][k] = x[k] - xj[k];
          r2[u] = fma(d[u][k], d[u][k], r2[u]);
        }
        clamp_branchless<T>(r2[u], ph.eps2);
      }
      Real<T>::template inv_cube_mass_v<U>(r2, m0, m1, m2c, s3);
      if (DENSE) {
#pragma unroll
        for (int u = 0; u < U; ++u) s3[u] = fma(s3[u], ci_repel, -wrow[j0 + u * L]);
      }
#pragma unroll
      for (int u = 0; u < U; u += 2) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
          f0[k] = fma(d[u][k], s3[u], f0[k]);
          f1[k] = fma(d[u + 1][k], s3[u + 1], f1[k]);
        }
      }
    }
    T f[D];
#pragma unroll
    for (int k = 0; k < D; ++k) f[k] = DENSE ? f0[k] + f1[k] : fma(f0[k] + f1[k], ci_repel, fa[k]);
    if (!DENSE && owner) {
      // (entries beyond two per lane continue from the compact list through L1)
      for (int e = eb + part + 2 * L; e < ee; e += L) {
        const int j = a.e_idx[e];
        const T w = (a.e_w != nullptr && ph.use_weights) ? a.e_w[e] : (T)1;
        T xj[D], d[D];
        C::pos(pc, j, xj);
        T r2 = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          d[k] = xj[k] - x[k];
          r2 = fma(d[k], d[k], r2);
        }
        const T g = attraction_factor<T, GA>(r2, w, ci, ph);
#pragma unroll
        for (int k = 0; k < D; ++k) f[k] = fma(d[k], g, f[k]);
      }
    }
#pragma unroll
    for (int off = L >> 1; off > 0; off >>= 1) {
#pragma unroll
      for (int k = 0; k < D; ++k) f[k] += __shfl_xor_sync(0xffffffffu, f[k], off);
    }
    {  // :205-261 with the gravity factor prepared ahead (flat kernel: |x| and swing unclamped)
      T sw2 = (T)0, f2 = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        const T fk = fma(-x[k], grav, f[k]);
        const T dk = fk - fprev[k];
        sw2 = fma(dk, dk, sw2);
        f2 = fma(fk, fk, f2);
        f[k] = fk;
      }
      const T swing = sw2 > (T)0 ? sw2 * Real<T>::rsqrt_acc(sw2) : sw2;
      const T ssw = swing > (T)0 ? swing * Real<T>::rsqrt_acc(swing) : swing;
      T speed = ph.ks * ph.gspeed * Real<T>::rcp_acc((T)1 + ph.gspeed * ssw);
      const T cap = ph.ksmax * Real<T>::rsqrt_acc(f2);
      if (speed > cap) speed = cap;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        x[k] = fma(f[k], speed, x[k]);
        fprev[k] = f[k];
      }
    }
    if (owner) {  // the L lanes of the group share out the csize peer stores
      for (int rr = part; rr < csize; rr += L) {
        if (async_x) {
          const uint32_t dst = cluster_addr(pos_addr + (uint32_t)(((size_t)nxt * S + gv) * DP * sizeof(T)), rr);
          const uint32_t pbar = cluster_addr(bar_addr[nxt], rr);
#pragma unroll
          for (int k = 0; k < D; ++k) st_async(dst + k * (uint32_t)sizeof(T), x[k], pbar);
        } else {
          T* dst = cluster.map_shared_rank(pos, rr) + (size_t)nxt * S * DP + (size_t)gv * DP;
#pragma unroll
          for (int k = 0; k < D; ++k) dst[k] = x[k];
        }
      }
    }
    if (!async_x) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    // while the positions travel: what the next iteration needs from the vertex's own new position
    m2 = (T)0;
#pragma unroll
    for (int k = 0; k < D; ++k) m2 = fma(x[k], x[k], m2);
    grav = ph.gravity * ci * Real<T>::rsqrt_acc(m2);
    if (async_x) {
      mbar_wait(&xbar[nxt], (xphase >> nxt) & 1u);
      xphase ^= 1u << nxt;
    } else {
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    const T* tmp = pc;
    pc = pn;
    pn = const_cast<T*>(tmp);
    nxt ^= 1;
  }
  if (async_x) cluster.sync();  // nobody leaves while a peer may still be storing into it

  if (!a.normalize) {
    if (owner && part == 0) {
#pragma unroll
      for (int k = 0; k < D; ++k) a.out_aos[(int64_t)gv * D + k] = (double)x[k];
    }
    return;
  }
  if (rank != 0) return;  // include/forceatlas.hpp:272-303 by CTA 0 (every CTA holds all positions)
  auto block_reduce = [&](double val, bool is_max) -> double {
    for (int off = 16; off > 0; off >>= 1) {
      const double o = __shfl_xor_sync(0xffffffffu, val, off);
      val = is_max ? fmax(val, o) : val + o;
    }
    if ((tid & 31) == 0) red[tid >> 5] = val;
    __syncthreads();
    if (tid < 32) {
      double w = (tid < (int)((blockDim.x + 31) >> 5)) ? red[tid] : 0.0;
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
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double part_sum = 0.0;
    for (int i = tid; i < s; i += blockDim.x) part_sum += (double)pc[i * DP + k];
    const double tot = block_reduce(part_sum, false);
    if (tid == 0) bc[k] = tot / s;
  }
  __syncthreads();
  double mx = 0.0;
  for (int i = tid; i < s; i += blockDim.x) {
    double mm = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double c = (double)pc[i * DP + k] - bc[k];
      mm += c * c;
    }
    mx = fmax(mx, sqrt(mm));
  }
  const double maxlen = block_reduce(mx, true);
  for (int i = tid; i < s; i += blockDim.x) {
#pragma unroll
    for (int k = 0; k < D; ++k)
      a.out_aos[(int64_t)i * D + k] = ((double)pc[i * DP + k] - bc[k]) / maxlen;
  }
}

constexpr int kWarpsPerCta = 8;

template <typename T, int D>
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_onchip_warp(const OnchipArgs<T> a,
                                                                   int npacks) {
  constexpr int NM = Real<T>::kMassArrays;
  __shared__ T pos_s[kWarpsPerCta][2][D][32];
  __shared__ T ms_s[kWarpsPerCta][NM][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wid = blockIdx.x * kWarpsPerCta + warp;
  if (wid >= npacks) return;
  const int4 pack = a.tasks[wid];
  const int slot0 = pack.x, s = pack.y, count = pack.z;
  const bool active = lane < s * count;
  const int g = active ? lane / s : 0;
  const int base = g * s;
  const int slot = slot0 + (active ? lane : 0);
  const int v = a.vtx ? a.vtx[slot] : slot;

  T x[D], fprev[D], E[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    x[k] = (T)a.init_aos[(int64_t)v * D + k];
    fprev[k] = (T)0;
    E[k] = a.Eext ? a.Eext[(int64_t)k * a.ld + slot] : (T)0;
  }
  const T ci = a.mass[slot];
  const T ci_repel = ci * a.ph.repel;
  const int eb = active ? a.e_begin[slot] : 0;
  const int ee = active ? a.e_end[slot] : 0;
#pragma unroll
  for (int k = 0; k < D; ++k) pos_s[warp][0][k][lane] = x[k];
  ms_s[warp][0][lane] = ci;
  if (NM > 1) ms_s[warp][NM > 1 ? 1 : 0][lane] = (T)1.5 * ci;
  if (NM > 2) ms_s[warp][NM > 2 ? 2 : 0][lane] = (T)1.875 * ci;
  __syncwarp();

  int cur = 0;
  const int iters = a.forces_only ? 1 : a.iters;
  for (int it = 0; it < iters; ++it) {
    T f[D];
#pragma unroll
    for (int k = 0; k < D; ++k) f[k] = (T)0;
    {
      // U = 4 members per trip (all lanes, so that the clamp vote is convergent) (the self pair contributes exactly 0, Q4; lanes past the end of
      // the aggregate re-read its first member with mass 0)
      for (int j0 = base; j0 < base + s; j0 += 4) {
        T d[4][D], r2[4], s3[4], m0[4], m1[4], m2[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = j0 + u < base + s;
          const int j = ok ? j0 + u : base;
          r2[u] = (T)0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            d[u][k] = x[k] - pos_s[warp][cur][k][j];
            r2[u] = fma(d[u][k], d[u][k], r2[u]);
          }
          m0[u] = ok ? ms_s[warp][0][j] : (T)0;
          m1[u] = ok ? ms_s[warp][NM > 1 ? 1 : 0][j] : (T)0;
          m2[u] = ok ? ms_s[warp][NM > 2 ? 2 : 0][j] : (T)0;
        }
        Real<T>::template clamp_lo_n<4>(r2, a.ph.eps2);
        Real<T>::template inv_cube_mass_v<4>(r2, m0, m1, m2, s3);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int k = 0; k < D; ++k) f[k] = fma(d[u][k], s3[u], f[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < D; ++k) f[k] *= ci_repel;
    }
    if (active) {
      for (int e = eb; e < ee; ++e) {
        const int j = a.e_idx[e] - slot0;
        const T w = (a.e_w != nullptr && a.ph.use_weights) ? a.e_w[e] : (T)1;
        T d[D];
        T r2 = (T)0;
#pragma unroll
        for (int k = 0; k < D; ++k) {
          d[k] = pos_s[warp][cur][k][j] - x[k];
          r2 = fma(d[k], d[k], r2);
        }
        const T gf = attraction_factor<T, true>(r2, w, ci, a.ph);
#pragma unroll
        for (int k = 0; k < D; ++k) f[k] = fma(d[k], gf, f[k]);
      }
    }
    vertex_step<T, D, true>(x, f, fprev, E, ci, a.ph);
    if (a.forces_only) break;
#pragma unroll
    for (int k = 0; k < D; ++k) pos_s[warp][cur ^ 1][k][lane] = x[k];
    cur ^= 1;
    __syncwarp();
  }

  if (!active) return;
  if (a.forces_only) {
#pragma unroll
    for (int k = 0; k < D; ++k) a.out_aos[(int64_t)v * D + k] = (double)fprev[k];
    return;
  }
  // :539-570, member order ascending exactly like the reference's serial loops.
  double avg[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double sum = 0.0;
    for (int j = base; j < base + s; ++j) sum += (double)pos_s[warp][cur][k][j];
    avg[k] = sum / s;
  }
  double maxlen = 0.0;
  for (int j = base; j < base + s; ++j) {
    double m2 = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const double c = (double)pos_s[warp][cur][k][j] - avg[k];
      m2 += c * c;
    }
    maxlen = fmax(maxlen, sqrt(m2));
  }
  if (maxlen < kEpsilon) maxlen = kEpsilon;
  const int agg = a.agg_of_slot[slot];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const double c = ((double)x[k] - avg[k]) / maxlen;
    a.out_aos[(int64_t)v * D + k] = a.cA_aos[(int64_t)agg * D + k] + a.rA[agg] * c;
  }
}

template <typename T>
size_t cta_smem(int dim, int max_size, int lanes, bool big) {
  const int S = onchip_spad(max_size, lanes, big ? 2 : 4);
  const int DP = dim == 2 ? 2 : 4, MP = sizeof(T) == 8 ? 4 : 1;
  return (size_t)(2 * DP + MP) * S * sizeof(T);
}

}  // namespace

namespace {
template <typename T, int D, bool ML, bool BIG, bool GA>
const void* cta_kernel_l(int L) {
  switch (L) {
    case 1: return (const void*)k_onchip_cta<T, D, ML, 1, BIG, GA>;
    case 2: return (const void*)k_onchip_cta<T, D, ML, 2, BIG, GA>;
    case 4: return (const void*)k_onchip_cta<T, D, ML, 4, BIG, GA>;
    default: return (const void*)k_onchip_cta<T, D, ML, 8, BIG, GA>;
  }
}
template <typename T, int D, bool ML>
const void* cta_kernel(int L, bool big, bool ga) {
  if (big) return ga ? cta_kernel_l<T, D, ML, true, true>(L) : cta_kernel_l<T, D, ML, true, false>(L);
  return ga ? cta_kernel_l<T, D, ML, false, true>(L) : cta_kernel_l<T, D, ML, false, false>(L);
}
}  // namespace

template <typename T>
void launch_onchip_cta(ge_context* ctx, const OnchipArgs<T>& a, int ntasks, int dim, bool ml,
                       int lanes, int threads, int max_size) {
  if (ntasks == 0) return;
  const bool big = threads > 512, ga = a.ph.general_attraction != 0;
  const size_t smem = cta_smem<T>(dim, max_size, lanes, big);
  const void* fn = dim == 2 ? (ml ? cta_kernel<T, 2, true>(lanes, big, ga) : cta_kernel<T, 2, false>(lanes, big, ga))
                            : (ml ? cta_kernel<T, 3, true>(lanes, big, ga) : cta_kernel<T, 3, false>(lanes, big, ga));
  if (smem > 40 * 1024)  // (static shared memory counts against the 48 KB default limit too)
    GE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&a};
  GE_CUDA(cudaLaunchKernel(fn, dim3(ntasks), dim3(threads), args, smem, ctx->stream));
  ctx->launches++;
}

template <typename T>
void launch_onchip_warp(ge_context* ctx, const OnchipArgs<T>& a, int npacks, int dim) {
  if (npacks == 0) return;
  const int ctas = (npacks + kWarpsPerCta - 1) / kWarpsPerCta;
  if (dim == 2)
    k_onchip_warp<T, 2><<<ctas, kWarpsPerCta * 32, 0, ctx->stream>>>(a, npacks);
  else
    k_onchip_warp<T, 3><<<ctas, kWarpsPerCta * 32, 0, ctx->stream>>>(a, npacks);
  GE_CUDA(cudaGetLastError());
  ctx->launches++;
}

namespace {
template <typename T, int D, bool GA>
const void* cluster_kernel(int L) {
  switch (L) {
    case 1: return (const void*)k_onchip_cluster<T, D, 1, GA>;
    case 2: return (const void*)k_onchip_cluster<T, D, 2, GA>;
    case 4: return (const void*)k_onchip_cluster<T, D, 4, GA>;
    case 16: return (const void*)k_onchip_cluster<T, D, 16, GA>;
    case 32: return (const void*)k_onchip_cluster<T, D, 32, GA>;
    default: return (const void*)k_onchip_cluster<T, D, 8, GA>;
  }
}
}  // namespace

namespace {
template <typename T, int D, int L>
const void* cluster2_kernel_l(int U, bool ga, bool dense) {
  if (ga) return (const void*)k_onchip_cluster2<T, D, L, 4, true, false>;
  if (dense) return U == 8 ? (const void*)k_onchip_cluster2<T, D, L, 8, false, true>
                           : (const void*)k_onchip_cluster2<T, D, L, 4, false, true>;
  return U == 8 ? (const void*)k_onchip_cluster2<T, D, L, 8, false, false>
                : (const void*)k_onchip_cluster2<T, D, L, 4, false, false>;
}
template <typename T, int D>
const void* cluster2_kernel(int L, int U, bool ga, bool dense) {
  return L == 16 ? cluster2_kernel_l<T, D, 16>(U, ga, dense) : cluster2_kernel_l<T, D, 8>(U, ga, dense);
}
int env_or(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}
}  // namespace

// The whole flat problem on one cluster of `csize` CTAs (slot == vertex id).
template <typename T>
void launch_onchip_cluster(ge_context* ctx, const OnchipArgs<T>& a, int n, int dim, int csize, int64_t nnz) {
  const int per_cta = (n + csize - 1) / csize;
  const bool ga = a.ph.general_attraction != 0;
  // ---- second-generation kernel (dense coarse graphs): 8 or 16 lanes per vertex, <= 256 threads ----
  // Measured (tools/profile_small.py k3v2, profiles/r02_k3v2_d2.txt): on the (nearly) complete
  // 54-vertex coarsest graph of a power-law hierarchy the dense variant takes 1.00 us/iteration
  // (8 CTAs x 16 lanes, 4 columns per trip) against 1.72 for the first-generation kernel, whose CSR
  // attraction loop is as long as its pair loop there.  On sparse coarse graphs (~6 entries per
  // row) the first-generation kernel stays ahead in every shape (n = 100: 1.51 vs 1.57-2.0 us), so
  // it keeps serving them: GE_K3_V2=1 forces the new kernel, GE_K3_V1=1 the old one.
  bool dense = !ga && nnz > (int64_t)2 * 8 * n && n <= 256;
  if (const char* v = std::getenv("GE_K3_DENSE")) dense = !ga && std::atoi(v) != 0;
  const bool want_v2 = env_or("GE_K3_V1", 0) == 0 && (dense || env_or("GE_K3_V2", 0) != 0 ||
                                                       std::getenv("GE_K3_DENSE") != nullptr);
  if (want_v2 && per_cta * 8 <= 256) {
    int L = per_cta * 16 <= 256 ? 16 : 8;
    if (const char* v = std::getenv("GE_ONCHIP_LANES")) L = std::atoi(v) >= 16 ? 16 : 8;
    if (per_cta * L > 256) L = 8;
    int U = env_or("GE_K3_U", 4) >= 8 ? 8 : 4;
    if (ga) U = 4;
    const int threads = (int)round_up((int64_t)per_cta * L, 32);
    const void* fn = dim == 2 ? cluster2_kernel<T, 2>(L, U, ga, dense) : cluster2_kernel<T, 3>(L, U, ga, dense);
    const int S = onchip_spad(n, L, U);
    const int DP = dim == 2 ? 2 : 4, MP = sizeof(T) == 8 ? 4 : 1;
    const size_t smem = ((size_t)(2 * DP + MP) * S + (dense ? (size_t)per_cta * (S + 8) : 0)) * sizeof(T);
    if (smem > 40 * 1024)
      GE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (csize > 8) GE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    void* args[] = {(void*)&a, (void*)&per_cta};
    GE_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
    ctx->launches++;
    return;
  }
  // measured (tools/profile_small.py k3sweep): 8 lanes per vertex, as in the single-CTA kernel
  // (n = 73: 8 CTAs x 8 lanes 1.3 us/iteration vs 2.2 us on one CTA; n = 157: 1.8 us)
  // (16 lanes: one trip of the pair loop instead of two up to n = 64 -- 1.11 vs 1.17 us at n = 64,
  // d = 2; no gain at n = 97, a loss at n = 200; 32 lanes lose everywhere)
  int L = 1;
  while (L < 8 && per_cta * (L * 2) <= 512) L *= 2;
  if (L == 8 && per_cta * 16 <= 128) L = 16;
  // a full warp per vertex when that saves a trip of the pair loop (64 < n <= 128 on 16 CTAs:
  // one trip of 4 columns per lane instead of two; n = 100: 1.34 -> 1.27 us, d = 3: 1.62 -> 1.51)
  if (L == 16 && n > 64 && n <= 128 && per_cta * 32 <= 256) L = 32;
  if (const char* v = std::getenv("GE_ONCHIP_LANES")) {
    L = 1;
    while (L < 32 && L * 2 <= std::atoi(v)) L *= 2;  // 1, 2, 4, 8, 16 or 32
  }
  while (L > 1 && per_cta * L > 512) L /= 2;
  const int threads = (int)round_up((int64_t)per_cta * L, 32);
  GE_REQUIRE(threads <= 512, "cluster solve: too many vertices per CTA");
  const void* fn = dim == 2 ? (ga ? cluster_kernel<T, 2, true>(L) : cluster_kernel<T, 2, false>(L))
                            : (ga ? cluster_kernel<T, 3, true>(L) : cluster_kernel<T, 3, false>(L));
  const size_t smem = cta_smem<T>(dim, n, L, false);
  if (smem > 40 * 1024)
    GE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (csize > 8) GE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  void* args[] = {(void*)&a, (void*)&per_cta};
  GE_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  ctx->launches++;
}

template void launch_onchip_cta<double>(ge_context*, const OnchipArgs<double>&, int, int, bool, int, int, int);
template void launch_onchip_cta<float>(ge_context*, const OnchipArgs<float>&, int, int, bool, int, int, int);
template void launch_onchip_warp<double>(ge_context*, const OnchipArgs<double>&, int, int);
template void launch_onchip_warp<float>(ge_context*, const OnchipArgs<float>&, int, int);

// ---------------------------------------------------------------------------------------------
// Coarsest-level flat solve: the whole graph is one task of k_onchip_cta.
// ---------------------------------------------------------------------------------------------
namespace {

// Lanes per vertex: the largest power of two (<= 32) such that n * L threads fit in one CTA.
int lanes_for(int n, int max_threads) {
  if (const char* v = std::getenv("GE_ONCHIP_LANES")) return std::atoi(v);
  // measured on B200 (tools/profile_small.py k3sweep): 8 lanes per vertex is the sweet spot between
  // pair-loop length and the per-warp cost of the (redundantly executed) per-vertex epilogue
  int L = 1;
  while (L < 8 && (int64_t)n * (L * 2) <= max_threads) L *= 2;
  return L;
}
int onchip_thread_budget(int n) {
  if (const char* v = std::getenv("GE_ONCHIP_THREADS")) return std::atoi(v);
  return n > 64 ? 1024 : 512;
}

template <typename T>
void onchip_flat_t(ge_context* ctx, const ge_csr& A, int dim, const ge_params& p, double* coords,
                   double* forces_out, bool forces_only, DevBuf<double>* keep) {
  const int n = A.rows;
  const int nnz = A.indptr[n];
  const bool weighted = p.use_weights && A.data != nullptr;
  std::vector<T> mass(n), w;
  for (int i = 0; i < n; ++i) {  // include/forceatlas.hpp:127-140
    double s = 0.0;
    if (weighted) {
      for (int e = A.indptr[i]; e < A.indptr[i + 1]; ++e) s += A.data[e];
    } else {
      s = 1.0 * (A.indptr[i + 1] - A.indptr[i]);
    }
    mass[i] = (T)(s + 1.0);
  }
  DevBuf<T> d_mass(ctx, n), d_w;
  DevBuf<int> d_I(ctx, n + 1), d_J(ctx, std::max(nnz, 1));
  DevBuf<double> d_init(ctx, (size_t)n * dim), d_out(ctx, (size_t)n * dim);
  DevBuf<int4> d_task(ctx, 1);
  d_mass.upload(ctx, mass.data(), n);
  d_I.upload(ctx, A.indptr, n + 1);
  d_J.upload(ctx, A.indices, nnz);
  if (weighted) {
    w.resize(nnz);
    for (int e = 0; e < nnz; ++e) w[e] = (T)A.data[e];
    d_w.alloc(ctx, std::max(nnz, 1));
    d_w.upload(ctx, w.data(), nnz);
  }
  d_init.upload(ctx, coords, (size_t)n * dim);
  const int L = lanes_for(n, onchip_thread_budget(n));
  const int threads = (int)round_up((int64_t)n * L, 32);
  const int4 task = make_int4(0, n, 0, L);
  d_task.upload(ctx, &task, 1);

  OnchipArgs<T> a;
  a.init_aos = d_init.get();
  a.mass = d_mass.get();
  a.e_begin = d_I.get();
  a.e_end = d_I.get() + 1;
  a.e_idx = d_J.get();
  a.e_w = weighted ? d_w.get() : nullptr;
  a.ld = n;
  a.tasks = d_task.get();
  a.out_aos = d_out.get();
  a.iters = p.iterations;
  a.forces_only = forces_only ? 1 : 0;
  a.normalize = p.normalize;
  a.ph = make_physics<T>(p);
  if (const char* v = std::getenv("GE_ONCHIP_SKIP")) a.debug_skip = std::atoi(v);
  a.exchange = 1;  // st.async + mbarrier (0: DSMEM stores + cluster barrier)
  if (const char* v = std::getenv("GE_K3_EXCHANGE")) a.exchange = std::atoi(v);
  // Larger coarsest levels are spread over a thread-block cluster (measured crossover n ~ 40).
  int csize = 1;
  if (!forces_only) {
    // measured with the st.async exchange (tools/k3_cluster_sweep.py, profiles/r02_k3_cluster_sweep.txt):
    // one CTA up to ~30 vertices (n = 27: 0.68 us vs 0.85 on a cluster), 8 CTAs up to ~95
    // (n = 34: 0.88 vs 1.09 on one CTA; n = 64: 0.94), the non-portable 16-CTA cluster beyond
    // (n = 100: 1.36 vs 1.39; n = 200: 2.17 vs 2.54; d = 3 gains more)
    csize = n >= 96 ? 16 : n >= 32 ? 8 : 1;
    if ((int64_t)nnz > (int64_t)16 * n && !a.ph.general_attraction && n <= 256)  // dense: see launch_onchip_cluster
      csize = n >= 32 ? 8 : n >= 24 ? 4 : 1;
    if (const char* v = std::getenv("GE_CLUSTER")) csize = std::atoi(v);
    csize = std::max(1, std::min(csize, 16));
  }
  if (csize > 1)
    launch_onchip_cluster<T>(ctx, a, n, dim, csize, nnz);
  else
    launch_onchip_cta<T>(ctx, a, 1, dim, false, L, threads, n);
  if (keep == nullptr) d_out.download(ctx, forces_only ? forces_out : coords, (size_t)n * dim);
  GE_CUDA(cudaStreamSynchronize(ctx->stream));
  if (keep != nullptr) *keep = std::move(d_out);
}

}  // namespace

void onchip_flat_solve(ge_context* ctx, const ge_csr& A, int dim, const ge_params& p,
                       double* coords, double* forces_out, bool forces_only, DevBuf<double>* keep) {
  GE_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
  GE_REQUIRE(A.rows >= 1 && A.rows <= kOnchipMaxVertices, "on-chip flat solve needs 1..1024 vertices");
  if (p.precision == GE_F32)
    onchip_flat_t<float>(ctx, A, dim, p, coords, forces_out, forces_only, keep);
  else
    onchip_flat_t<double>(ctx, A, dim, p, coords, forces_out, forces_only, keep);
}

}  // namespace ge
