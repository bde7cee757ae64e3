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
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ge_flat.cuh"

namespace ge {

// ---- mbarrier / TMA bulk-copy primitives ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_addr(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

template <typename T, int VEC>
struct VecLoad;
template <>
struct VecLoad<double, 2> {
  __device__ __forceinline__ static void ld(const double* p, double (&v)[2]) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x;
    v[1] = t.y;
  }
};
template <>
struct VecLoad<float, 4> {
  __device__ __forceinline__ static void ld(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x;
    v[1] = t.y;
    v[2] = t.z;
    v[3] = t.w;
  }
};

// K1a.  Each thread owns IPT rows (positions and force accumulators in registers); the CTA walks
// its column range tile by tile.  Thread 0 is the TMA producer; everybody consumes through
// broadcast shared-memory loads (all lanes read the same column -> conflict free).
template <typename T, int D, int IPT>
__global__ void __launch_bounds__(kRepMaxThreads) k_repulsion(const RepArgs<T> a) {
  constexpr int NM = Real<T>::kMassArrays;
  constexpr int NA = D + NM;
  constexpr int VEC = 16 / (int)sizeof(T);
  constexpr uint32_t kStageBytes = NA * kTileJ * sizeof(T);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* tiles = reinterpret_cast<T*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kRepStages * kStageBytes);

  const BlockDesc bd = a.blocks[blockIdx.x];
  const int ntiles = (bd.j1 - bd.j0) / kTileJ;
  const int tid = threadIdx.x;

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

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kRepStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int tile) {
    const int s = tile % kRepStages;
    T* dst = tiles + (size_t)s * NA * kTileJ;
    const int64_t j = (int64_t)bd.j0 + (int64_t)tile * kTileJ;
    mbar_expect_tx(&full[s], kStageBytes);
#pragma unroll
    for (int k = 0; k < D; ++k)
      tma_load_1d(dst + k * kTileJ, a.pos + (int64_t)k * a.ld + j, kTileJ * sizeof(T), &full[s]);
#pragma unroll
    for (int k = 0; k < NM; ++k)
      tma_load_1d(dst + (D + k) * kTileJ, a.mass + (int64_t)k * a.ld + j, kTileJ * sizeof(T),
                  &full[s]);
  };

  if (tid == 0) {
    for (int t = 0; t < kRepStages - 1 && t < ntiles; ++t) issue(t);
  }

  for (int tile = 0; tile < ntiles; ++tile) {
    __syncthreads();  // everyone is done with tile-1: its stage may be refilled
    if (tid == 0 && tile + kRepStages - 1 < ntiles) issue(tile + kRepStages - 1);
    const int s = tile % kRepStages;
    mbar_wait(&full[s], (uint32_t)((tile / kRepStages) & 1));
    const T* st = tiles + (size_t)s * NA * kTileJ;

#pragma unroll 1
    for (int jj = 0; jj < kTileJ; jj += VEC) {
      T xj[D][VEC], mj[3][VEC];
#pragma unroll
      for (int k = 0; k < D; ++k) VecLoad<T, VEC>::ld(st + k * kTileJ + jj, xj[k]);
#pragma unroll
      for (int k = 0; k < NM; ++k) VecLoad<T, VEC>::ld(st + (D + k) * kTileJ + jj, mj[k]);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
          T d[D];
          T r2 = (T)0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            d[k] = xi[t][k] - xj[k][v];
            r2 = fma(d[k], d[k], r2);
          }
          r2 = Real<T>::clamp_lo(r2, a.eps2);
          const T s3 = Real<T>::inv_cube_mass(r2, mj[0][v], mj[NM > 1 ? 1 : 0][v],
                                              mj[NM > 2 ? 2 : 0][v]);
#pragma unroll
          for (int k = 0; k < D; ++k) fi[t][k] = fma(d[k], s3, fi[t][k]);
        }
      }
    }
  }

#pragma unroll
  for (int t = 0; t < IPT; ++t) {
    const int i = bd.row0 + tid + t * blockDim.x;
    if (i < bd.row1) {
      const T ci = a.mass[i] * a.repel;  // (deg_i + 1) * repel hoisted out of the pair loop
#pragma unroll
      for (int k = 0; k < D; ++k) a.F[(int64_t)k * a.ldf + (i - a.f_row_base)] = fi[t][k] * ci;
    }
  }
}

// K1b+c.  G lanes cooperate on one row (G = 4..32 chosen from the average degree); lane 0 of the
// group finishes the row: adds the repulsion sum, gravity, derives the per-vertex speed and
// writes the moved position into the NEXT coordinate buffer (Jacobi: everybody still reads the
// current one).
template <typename T, int D, int G, bool ML>
__global__ void __launch_bounds__(256) k_attract_step(const StepArgs<T> a) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = gtid / G;
  const int lane = gtid % G;
  const bool active = r < a.nrows;
  const int i = a.row0 + (active ? r : 0);
  T x[D], f[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    x[k] = a.pos_cur[(int64_t)k * a.ld + i];
    f[k] = (T)0;
  }
  const T ci = a.mass[i];
  if (active) {
    const int e1 = a.e_end[r];
    for (int e = a.e_begin[r] + lane; e < e1; e += G) {
      const int j = a.J[e];
      const T w = (a.W != nullptr && a.ph.use_weights) ? a.W[e] : (T)1;
      T d[D];
      T r2 = (T)0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        d[k] = a.pos_cur[(int64_t)k * a.ld + j] - x[k];
        r2 = fma(d[k], d[k], r2);
      }
      const T g = attraction_factor<T>(r2, w, ci, a.ph);
#pragma unroll
      for (int k = 0; k < D; ++k) f[k] = fma(d[k], g, f[k]);
    }
  }
#pragma unroll
  for (int off = G / 2; off > 0; off >>= 1) {
#pragma unroll
    for (int k = 0; k < D; ++k) f[k] += __shfl_xor_sync(0xffffffffu, f[k], off, G);
  }
  if (active && lane == 0) {
    T fprev[D], E[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      f[k] += a.Frep[(int64_t)k * a.ldf + r];
      fprev[k] = a.update ? a.Fprev[(int64_t)k * a.ldf + r] : (T)0;
      E[k] = (ML && a.Eext != nullptr) ? a.Eext[(int64_t)k * a.ldf + r] : (T)0;
    }
    vertex_step<T, D, ML>(x, f, fprev, E, ci, a.ph);
#pragma unroll
    for (int k = 0; k < D; ++k) {
      a.Fprev[(int64_t)k * a.ldf + r] = fprev[k];
      if (a.update) a.pos_next[(int64_t)k * a.ld + i] = x[k];
    }
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
template <typename T>
__global__ void k_aos_to_soa(const double* __restrict__ aos, int n, int dim, int64_t ld,
                             T* __restrict__ soa) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)ld * dim) return;
  const int k = (int)(t / ld);
  const int64_t i = t % ld;
  soa[t] = (i < n) ? (T)aos[i * dim + k] : (T)0;
}
template <typename T>
__global__ void k_soa_to_aos(const T* __restrict__ soa, int n, int dim, int64_t ld,
                             double* __restrict__ aos) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n * dim) return;
  const int64_t i = t / dim;
  const int k = (int)(t % dim);
  aos[t] = (double)soa[(int64_t)k * ld + i];
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
template <typename T>
size_t repulsion_smem(int dim) {
  return (size_t)kRepStages * (dim + Real<T>::kMassArrays) * kTileJ * sizeof(T) +
         kRepStages * sizeof(uint64_t);
}

template <typename T>
const void* repulsion_kernel(int dim, int ipt) {
  if (dim == 2)
    return ipt == 1 ? (const void*)k_repulsion<T, 2, 1>
                    : ipt == 2 ? (const void*)k_repulsion<T, 2, 2> : (const void*)k_repulsion<T, 2, 4>;
  return ipt == 1 ? (const void*)k_repulsion<T, 3, 1>
                  : ipt == 2 ? (const void*)k_repulsion<T, 3, 2> : (const void*)k_repulsion<T, 3, 4>;
}

template <typename T>
void launch_repulsion(ge_context* ctx, const RepArgs<T>& a, int nblocks, int threads, int ipt,
                      int dim) {
  if (nblocks == 0) return;
  const void* fn = repulsion_kernel<T>(dim, ipt);
  const size_t smem = repulsion_smem<T>(dim);
  if (smem > 48 * 1024)
    GE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* args[] = {(void*)&a};
  GE_CUDA(cudaLaunchKernel(fn, dim3(nblocks), dim3(threads), args, smem, ctx->stream));
  ctx->launches++;
}

namespace {
template <typename T, int D, bool ML>
void launch_step_g(ge_context* ctx, const StepArgs<T>& a, int group) {
  const int g = group <= 4 ? 4 : group <= 8 ? 8 : group <= 16 ? 16 : 32;
  const int64_t threads = (int64_t)a.nrows * g;
  const unsigned grid = (unsigned)((threads + 255) / 256);
  if (g == 4) k_attract_step<T, D, 4, ML><<<grid, 256, 0, ctx->stream>>>(a);
  else if (g == 8) k_attract_step<T, D, 8, ML><<<grid, 256, 0, ctx->stream>>>(a);
  else if (g == 16) k_attract_step<T, D, 16, ML><<<grid, 256, 0, ctx->stream>>>(a);
  else k_attract_step<T, D, 32, ML><<<grid, 256, 0, ctx->stream>>>(a);
}
}  // namespace

template <typename T>
void launch_attract_step(ge_context* ctx, const StepArgs<T>& a, int dim, int group, bool ml) {
  if (a.nrows == 0) return;
  if (dim == 2) {
    if (ml) launch_step_g<T, 2, true>(ctx, a, group);
    else launch_step_g<T, 2, false>(ctx, a, group);
  } else {
    if (ml) launch_step_g<T, 3, true>(ctx, a, group);
    else launch_step_g<T, 3, false>(ctx, a, group);
  }
  GE_CUDA(cudaGetLastError());
  ctx->launches++;
}

template size_t repulsion_smem<double>(int);
template size_t repulsion_smem<float>(int);
template const void* repulsion_kernel<double>(int, int);
template const void* repulsion_kernel<float>(int, int);
template void launch_repulsion<double>(ge_context*, const RepArgs<double>&, int, int, int, int);
template void launch_repulsion<float>(ge_context*, const RepArgs<float>&, int, int, int, int);
template void launch_attract_step<double>(ge_context*, const StepArgs<double>&, int, int, bool);
template void launch_attract_step<float>(ge_context*, const StepArgs<float>&, int, int, bool);

int group_for_degree(double avg_deg) {
  return avg_deg <= 6 ? 4 : avg_deg <= 12 ? 8 : avg_deg <= 24 ? 16 : 32;
}

// ---------------------------------------------------------------------------------------------
// device-resident flat solver
// ---------------------------------------------------------------------------------------------
namespace {

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

template <typename T>
class FlatSolverT final : public FlatSolver {
 public:
  FlatSolverT(ge_context* c, const ge_csr& A, int dim, const ge_params& p, int rb, int re)
      : n_(A.rows), dim_(dim), rb_(rb), re_(re) {
    ctx = c;
    ph_ = make_physics<T>(p);
    ld_ = round_up(std::max(n_, 1), kTileJ);
    nrows_ = re_ - rb_;
    ldf_ = round_up(std::max(nrows_, 1), 32);
    constexpr int NM = Real<T>::kMassArrays;

    // Vertex masses need every row's degree (include/forceatlas.hpp:127-140); rows outside the
    // owned block contribute nothing else.
    std::vector<double> deg(n_);
    const bool weighted = p.use_weights && A.data != nullptr;
    for (int i = 0; i < n_; ++i) {
      double s = 0.0;
      if (weighted) {
        for (int e = A.indptr[i]; e < A.indptr[i + 1]; ++e) s += A.data[e];
      } else {
        s = 1.0 * (A.indptr[i + 1] - A.indptr[i]);
      }
      deg[i] = s;
    }
    DevBuf<double> d_deg(std::max(n_, 1));
    d_deg.upload(ctx, deg.data(), n_);
    mass_.alloc((size_t)NM * ld_);
    k_mass_from_degree<T><<<(unsigned)((ld_ + 255) / 256), 256, 0, ctx->stream>>>(
        d_deg.get(), n_, ld_, NM, mass_.get());
    ctx->launches++;

    // owned CSR rows, re-based to local entry offsets
    const int e0 = A.indptr[rb_], e1 = A.indptr[re_];
    const int lnnz = e1 - e0;
    std::vector<int> rowptr(nrows_ + 1);
    for (int r = 0; r <= nrows_; ++r) rowptr[r] = A.indptr[rb_ + r] - e0;
    rowptr_.alloc(nrows_ + 1);
    rowptr_.upload(ctx, rowptr.data(), nrows_ + 1);
    J_.alloc(std::max(lnnz, 1));
    J_.upload(ctx, A.indices + e0, lnnz);
    std::vector<T> w;
    if (weighted) {
      w.resize(lnnz);
      for (int e = 0; e < lnnz; ++e) w[e] = (T)A.data[e0 + e];
      W_.alloc(std::max(lnnz, 1));
      W_.upload(ctx, w.data(), lnnz);
    }
    avg_deg_ = nrows_ > 0 ? double(lnnz) / nrows_ : 0.0;

    own0_.alloc((size_t)dim_ * ld_);
    own1_.alloc((size_t)dim_ * ld_);
    own0_.zero(ctx->stream);
    own1_.zero(ctx->stream);
    buf_[0] = own0_.get();
    buf_[1] = own1_.get();
    Frep_.alloc((size_t)dim_ * ldf_);
    Fprev_.alloc((size_t)dim_ * ldf_);
    Frep_.zero(ctx->stream);
    Fprev_.zero(ctx->stream);
    stage_.alloc((size_t)std::max(n_, 1) * dim_);

    plan_repulsion();
    for (auto& e : ev_) GE_CUDA(cudaEventCreate(&e));
    GE_CUDA(cudaStreamSynchronize(ctx->stream));
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
  }
  void upload_coords(const double* aos) override {
    stage_.upload(ctx, aos, (size_t)n_ * dim_);
    const int64_t total = ld_ * dim_;
    k_aos_to_soa<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        stage_.get(), n_, dim_, ld_, buf_[cur_]);
    ctx->launches++;
    // the other buffer must hold valid data outside the owned rows when ranks exchange slices
    GE_CUDA(cudaMemcpyAsync(buf_[cur_ ^ 1], buf_[cur_], (size_t)total * sizeof(T),
                            cudaMemcpyDeviceToDevice, ctx->stream));
    Fprev_.zero(ctx->stream);
  }
  void download_coords(double* aos) override {
    const int64_t total = (int64_t)n_ * dim_;
    if (total == 0) return;
    k_soa_to_aos<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        buf_[cur_], n_, dim_, ld_, stage_.get());
    ctx->launches++;
    stage_.download(ctx, aos, (size_t)total);
    GE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  void download_forces(double* aos) override {
    const int64_t total = (int64_t)nrows_ * dim_;
    if (total == 0) return;
    k_soa_to_aos<T><<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        Fprev_.get(), nrows_, dim_, ldf_, stage_.get());
    ctx->launches++;
    stage_.download(ctx, aos, (size_t)total);
    GE_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  void* cur_coords() override { return buf_[cur_]; }
  void* next_coords() override { return buf_[cur_ ^ 1]; }
  void swap() override { cur_ ^= 1; }

  void launch_iteration(bool update) override {
    if (nrows_ == 0) return;
    if (prof_) GE_CUDA(cudaEventRecord(ev_[0], ctx->stream));
    RepArgs<T> ra;
    ra.pos = buf_[cur_];
    ra.mass = mass_.get();
    ra.F = Frep_.get();
    ra.blocks = blocks_.get();
    ra.ld = ld_;
    ra.ldf = ldf_;
    ra.f_row_base = rb_;
    ra.repel = ph_.repel;
    ra.eps2 = ph_.eps2;
    launch_repulsion<T>(ctx, ra, rep_blocks_n_, rep_threads_, rep_ipt_, dim_);
    if (prof_) GE_CUDA(cudaEventRecord(ev_[1], ctx->stream));
    StepArgs<T> sa;
    sa.e_begin = rowptr_.get();
    sa.e_end = rowptr_.get() + 1;
    sa.J = J_.get();
    sa.W = W_.size() ? W_.get() : nullptr;
    sa.pos_cur = buf_[cur_];
    sa.pos_next = buf_[cur_ ^ 1];
    sa.Frep = Frep_.get();
    sa.Fprev = Fprev_.get();
    sa.mass = mass_.get();
    sa.Eext = nullptr;
    sa.ld = ld_;
    sa.ldf = ldf_;
    sa.row0 = rb_;
    sa.nrows = nrows_;
    sa.update = update ? 1 : 0;
    sa.ph = ph_;
    launch_attract_step<T>(ctx, sa, dim_, env_int("GE_STEP_GROUP", group_for_degree(avg_deg_)), false);
    if (prof_) {
      GE_CUDA(cudaEventRecord(ev_[2], ctx->stream));
      GE_CUDA(cudaEventSynchronize(ev_[2]));
      float t_rep = 0, t_step = 0;
      GE_CUDA(cudaEventElapsedTime(&t_rep, ev_[0], ev_[1]));
      GE_CUDA(cudaEventElapsedTime(&t_step, ev_[1], ev_[2]));
      rep_ms_ += t_rep;
      step_ms_ += t_step;
      rep_n_++;
      step_n_++;
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
  // Choose threads-per-CTA and rows-per-thread so that the grid fills whole waves of the
  // 148 SMs x resident CTAs; overridable with GE_REP_THREADS / GE_REP_IPT for sweeps.
  void plan_repulsion() {
    const int sms = ctx->sm_count;
    double best_score = -1.0;
    const int force_thr = env_int("GE_REP_THREADS", 0), force_ipt = env_int("GE_REP_IPT", 0);
    const size_t smem = repulsion_smem<T>(dim_);
    for (int ipt : {4, 2, 1}) {
      if (force_ipt && ipt != force_ipt) continue;
      const void* fn = repulsion_kernel<T>(dim_, ipt);
      for (int thr = 128; thr <= kRepMaxThreads; thr += 32) {
        if (force_thr && thr != force_thr) continue;
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, thr, smem) != cudaSuccess ||
            occ == 0) {
          cudaGetLastError();
          continue;
        }
        const int64_t rows_per_cta = (int64_t)thr * ipt;
        const int64_t ctas = (nrows_ + rows_per_cta - 1) / rows_per_cta;
        const int64_t slots = (int64_t)sms * occ;
        const int64_t waves = (ctas + slots - 1) / slots;
        // useful row-slots / row-slots paid for ...
        double score;
        if (waves == 1) {
          const int64_t per_sm = (ctas + sms - 1) / sms;  // CTAs on the busiest SM
          score = double(nrows_) / double(per_sm * sms * rows_per_cta);
        } else {
          score = double(nrows_) / double(waves * slots * rows_per_cta);
        }
        // ... times a mild preference for register blocking (fewer shared-memory reads per
        // pair) and for at least 8 resident warps per SM (latency hiding on the FP64 pipe).
        const int res_ctas = (int)std::min<int64_t>(occ, (ctas + sms - 1) / sms);
        const double warps = double(res_ctas) * thr / 32.0;
        score *= (ipt == 4 ? 1.0 : ipt == 2 ? 0.97 : 0.90);
        score *= std::min(1.0, 0.6 + 0.05 * warps);
        if (score > best_score) {
          best_score = score;
          rep_threads_ = thr;
          rep_ipt_ = ipt;
        }
      }
    }
    GE_REQUIRE(best_score > 0, "no launchable repulsion configuration");
    const int rows_per_cta = rep_threads_ * rep_ipt_;
    std::vector<BlockDesc> blocks;
    for (int r = rb_; r < re_; r += rows_per_cta)
      blocks.push_back(BlockDesc{r, std::min(re_, r + rows_per_cta), 0, (int)ld_});
    rep_blocks_n_ = (int)blocks.size();
    blocks_.alloc(std::max<size_t>(blocks.size(), 1));
    blocks_.upload(ctx, blocks.data(), blocks.size());
    GE_CUDA(cudaStreamSynchronize(ctx->stream));
    if (std::getenv("GE_VERBOSE"))
      std::fprintf(stderr, "[ge] repulsion plan: threads=%d ipt=%d ctas=%d smem=%zu\n", rep_threads_,
                   rep_ipt_, rep_blocks_n_, smem);
  }

  int n_, dim_, rb_, re_, nrows_ = 0;
  int64_t ld_ = 0, ldf_ = 0;
  Physics<T> ph_;
  double avg_deg_ = 0;
  DevBuf<T> mass_, own0_, own1_, Frep_, Fprev_, W_;
  DevBuf<int> rowptr_, J_;
  DevBuf<BlockDesc> blocks_;
  DevBuf<double> stage_;
  T* buf_[2] = {nullptr, nullptr};
  int cur_ = 0;
  int rep_threads_ = 256, rep_ipt_ = 4, rep_blocks_n_ = 0;
  bool prof_ = false;
  double rep_ms_ = 0, step_ms_ = 0;
  int64_t rep_n_ = 0, step_n_ = 0;
  cudaEvent_t ev_[3] = {nullptr, nullptr, nullptr};
};

}  // namespace

FlatSolver* make_flat_solver(ge_context* ctx, const ge_csr& A, int dim, const ge_params& p,
                             int row_begin, int row_end) {
  GE_REQUIRE(dim == 2 || dim == 3, "dim must be 2 or 3");
  GE_REQUIRE(A.rows == A.cols, "A must be square");
  GE_REQUIRE(0 <= row_begin && row_begin <= row_end && row_end <= A.rows, "bad row block");
  if (p.precision == GE_F32) return new FlatSolverT<float>(ctx, A, dim, p, row_begin, row_end);
  return new FlatSolverT<double>(ctx, A, dim, p, row_begin, row_end);
}

}  // namespace ge
