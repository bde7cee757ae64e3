// graph-embed_b200 :: symmetric all-pairs repulsion sweep (kernel K1s), sm_100a.
//
// Replaces the O(n^2) loop of partition::forceAtlas, /root/reference/include/forceatlas.hpp:151-167.
// The reference adds, for every ORDERED pair,  -(xj-xi)/dis * ci*cj*repel/dis^2  to row i.  The
// term is exactly antisymmetric in floating point ((xj-xi) == -(xi-xj), the squared distance and
// the clamp are the same for (i,j) and (j,i)), so this kernel evaluates every UNORDERED pair once:
//     s   = max(|xi-xj|^2, eps^2)^(-3/2)                (one MUFU reciprocal square root)
//     S_i += (xi-xj) * (s * c_j)          row side, accumulated in registers
//     S_j -= (xi-xj) * (s * c_i)          column side
// and the caller multiplies S by c_i * repel.  16 FP-pipe instructions per unordered pair
// (d = 2, FP64) instead of 2 x 12.  Not a contraction: no tensor cores.
//
// Work = the upper triangle of (row block, column tile) units, cut into equal contiguous shares
// (one per resident CTA, as in k_repulsion).  Each thread owns IPT rows; all lanes of a warp meet
// the same CG columns of a tile together, so the column side needs a sum over the 32 lanes: a
// butterfly that halves the number of live values at each exchange (lane l evaluates the columns
// in the order  group ^ (top bits of l), so the first exchanges need no selects), ending with one
// lane holding one column's sum.  Warps deposit these into per-warp shared-memory rows; at the end
// of a tile the CTA adds the warps in a fixed order and writes the tile's column sums to its own
// slab of `colpartial` (every (block, tile) unit is visited exactly once per launch: no atomics,
// bit-reproducible).  k_sym_reduce then forms  S_i = rows(i) - sum over blocks of columns(i).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ge_flat.cuh"
#include "ge_tma.cuh"

namespace ge {

void sym_pass_bounds(int ntile, int npass, int q, int& pt0, int& pt1);

namespace {
constexpr int ilog2(int v) { return v <= 1 ? 0 : 1 + ilog2(v >> 1); }
}  // namespace

// One column tile against the thread's IPT rows.  SYM: also accumulate and deposit column sums.
template <typename T, int D, int IPT, int CG, bool SYM>
__device__ __forceinline__ void sym_tile(const T* __restrict__ st, const T (&xi)[IPT][D],
                                         const T (&ci)[IPT], T (&fi)[IPT][D],
                                         T* __restrict__ colacc_w, const T eps2, const int lane) {
  constexpr int TJ = kTileJ;
  constexpr int VEC = 16 / (int)sizeof(T);
  constexpr int NG = CG / VEC;      // column groups of one 16-byte load each
  constexpr int LG = ilog2(NG);     // exchanges resolved by the lane-dependent column order
  constexpr int LV = ilog2(VEC);    // exchanges that pick a half with selects
  constexpr int LC = ilog2(CG);
  static_assert(NG >= 1 && (1 << LG) == NG && LC <= 5, "bad column group");
  const int m = (SYM && LG > 0) ? (lane >> (5 - LG)) : 0;
  constexpr unsigned kFull = 0xffffffffu;

#pragma unroll 1
  for (int jj = 0; jj < TJ; jj += CG) {
    T g[CG][D];
    if (SYM) {
#pragma unroll
      for (int c = 0; c < CG; ++c)
#pragma unroll
        for (int k = 0; k < D; ++k) g[c][k] = (T)0;
    }
#pragma unroll
    for (int q = 0; q < NG; ++q) {
      const int cp = jj + VEC * (q ^ m);
      T xj[D][VEC], cj[VEC];
#pragma unroll
      for (int k = 0; k < D; ++k) VecLoad<T, VEC>::ld(st + k * TJ + cp, xj[k]);
      VecLoad<T, VEC>::ld(st + D * TJ + cp, cj);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
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
        Real<T>::template clamp_lo_n<IPT>(r2, eps2);
        Real<T>::template inv_cube_n<IPT>(r2, s3);
#pragma unroll
        for (int t = 0; t < IPT; ++t) {
          const T u = s3[t] * cj[v];
#pragma unroll
          for (int k = 0; k < D; ++k) fi[t][k] = fma(d[t][k], u, fi[t][k]);
        }
        if (SYM) {
#pragma unroll
          for (int t = 0; t < IPT; ++t) {
            const T w = s3[t] * ci[t];
#pragma unroll
            for (int k = 0; k < D; ++k) g[q * VEC + v][k] = fma(d[t][k], w, g[q * VEC + v][k]);
          }
        }
      }
    }
    if (SYM) {
      // slot group q of lane l holds column group q ^ m: partners across lane bit (4 - s) hold
      // the same column groups in slots that differ in bit (LG-1-s) -- keep the low half
#pragma unroll
      for (int s = 0; s < LG; ++s) {
        const int half = NG >> (s + 1);
        const int lx = 16 >> s;
#pragma unroll
        for (int q = 0; q < half; ++q)
#pragma unroll
          for (int v = 0; v < VEC; ++v)
#pragma unroll
            for (int k = 0; k < D; ++k)
              g[q * VEC + v][k] += __shfl_xor_sync(kFull, g[(q + half) * VEC + v][k], lx);
      }
      // within the remaining group: the lane bit picks which half it keeps
#pragma unroll
      for (int s = 0; s < LV; ++s) {
        const int half = VEC >> (s + 1);
        const int lx = 16 >> (LG + s);
        const bool up = (lane & lx) != 0;
#pragma unroll
        for (int v = 0; v < half; ++v)
#pragma unroll
          for (int k = 0; k < D; ++k) {
            const T send = up ? g[v][k] : g[v + half][k];
            const T mine = up ? g[v + half][k] : g[v][k];
            g[v][k] = mine + __shfl_xor_sync(kFull, send, lx);
          }
      }
      // lane l now holds column (l >> (5 - LC)) of the group, spread over 2^(5-LC) lanes
#pragma unroll
      for (int lx = (16 >> LC); lx > 0; lx >>= 1)
#pragma unroll
        for (int k = 0; k < D; ++k) g[0][k] += __shfl_xor_sync(kFull, g[0][k], lx);
      if ((lane & ((32 >> LC) - 1)) == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) colacc_w[k * TJ + jj + (lane >> (5 - LC))] = g[0][k];
      }
    }
  }
}

// Launch shapes: 1024 rows per block (512 threads x 2 rows, or 256 x 4) for long row ranges.  Plans
// over many short segments shrink the block: whole 1024-row blocks waste lanes on padding rows, give
// too few units to share out evenly, and the part of a block on its segment's diagonal is swept in
// full (every ordered pair), which costs rb/s extra for a segment of s rows.  SHRINK 1: 512 rows
// (256 threads x 2), SHRINK 2: 256 rows (128 threads x 2, four CTAs per SM) -- R-MAT-20's large
// aggregates hold 1100 - 3800 rows each.
template <int IPT, int SHRINK>
struct SymShape {
  static constexpr int kThreads = SHRINK == 2 ? 128 : (IPT >= 4 || SHRINK == 1) ? 256 : 512;
  static constexpr int kMinBlocks = SHRINK == 2 ? 4 : SHRINK == 1 ? 2 : 0;  // 0: no constraint
};

template <typename T, int D, int IPT, int CG, int SHRINK = 0>
__global__ void __launch_bounds__(SymShape<IPT, SHRINK>::kThreads, SymShape<IPT, SHRINK>::kMinBlocks)
    k_repulsion_sym(const RepSymArgs<T> a) {
  constexpr int TJ = kTileJ;
  constexpr int NA = D + 1;
  constexpr int NW = SymShape<IPT, SHRINK>::kThreads / 32;
  constexpr uint32_t kStageBytes = NA * TJ * sizeof(T);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* tiles = reinterpret_cast<T*>(smem_raw);
  T* colacc = tiles + (size_t)kRepStages * NA * TJ;  // [NW][D][TJ]
  uint64_t* full = reinterpret_cast<uint64_t*>(colacc + (size_t)NW * D * TJ);

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int nthreads = blockDim.x;
  const int nwarps = nthreads >> 5;
  T* colacc_w = colacc + (size_t)(tid >> 5) * D * TJ;
  const long long W = a.total_units, G = gridDim.x, c = blockIdx.x;
  const long long u0 = W * c / G, u1 = W * (c + 1) / G;
  if (u0 >= u1) return;

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kRepStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  int b = 0;  // block holding unit u0: last block with unit0 <= u0
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
  unsigned g = 0;  // tiles pushed through the pipeline so far (stage / parity bookkeeping)
  while (u < u1) {
    const SymBlockDesc bd = a.blocks[b];
    const int t_begin = (int)(u - bd.unit0);
    const int nt = (int)min((long long)(bd.ntiles - t_begin), u1 - u);
    const int gt0 = bd.t_first + t_begin;  // global index of this segment's first tile

    T xi[IPT][D], fi[IPT][D], ci[IPT];
#pragma unroll
    for (int t = 0; t < IPT; ++t) {
      const int i = bd.row0 + tid + t * nthreads;
      const bool ok = i < bd.row1;
      ci[t] = ok ? a.mass[i] : (T)0;  // rows outside the block must not push on the columns
#pragma unroll
      for (int k = 0; k < D; ++k) {
        xi[t][k] = ok ? a.pos[(int64_t)k * a.ld + i] : (T)0;
        fi[t][k] = (T)0;
      }
    }

    auto issue = [&](int l) {
      const unsigned s = (g + (unsigned)l) % kRepStages;
      T* dst = tiles + (size_t)s * NA * TJ;
      const int64_t j = (int64_t)(gt0 + l) * TJ;
      mbar_expect_tx(&full[s], kStageBytes);
#pragma unroll
      for (int k = 0; k < D; ++k)
        tma_load_1d(dst + k * TJ, a.pos + (int64_t)k * a.ld + j, TJ * sizeof(T), &full[s]);
      tma_load_1d(dst + D * TJ, a.mass + j, TJ * sizeof(T), &full[s]);
    };
    if (tid == 0) {
      for (int l = 0; l < kRepStages - 1 && l < nt; ++l) issue(l);
    }

    for (int l = 0; l < nt; ++l) {
      __syncthreads();  // everyone is done with tile l-1 (its stage and colacc may be reused)
      if (tid == 0 && l + kRepStages - 1 < nt) issue(l + kRepStages - 1);
      const unsigned gl = g + (unsigned)l;
      const unsigned s = gl % kRepStages;
      mbar_wait(&full[s], (gl / kRepStages) & 1u);
      const T* st = tiles + (size_t)s * NA * TJ;
      const int gt = gt0 + l;
      if (gt >= bd.tile_sym0) {
        sym_tile<T, D, IPT, CG, true>(st, xi, ci, fi, colacc_w, a.eps2, lane);
        __syncthreads();
        T* dst = a.colpartial + bd.col_off + (int64_t)(gt - bd.col_t0) * TJ;
        for (int col = tid; col < TJ; col += nthreads) {
#pragma unroll
          for (int k = 0; k < D; ++k) {
            T acc = (T)0;
            for (int w = 0; w < nwarps; ++w) acc += colacc[((size_t)w * D + k) * TJ + col];
            dst[(int64_t)k * bd.ncols + col] = acc;
          }
        }
      } else {
        sym_tile<T, D, IPT, CG, false>(st, xi, ci, fi, colacc_w, a.eps2, lane);
      }
    }
    g += (unsigned)nt;

    const bool whole = (t_begin == 0 && nt == bd.ntiles);
    const int slot = (u == u0) ? 0 : 1;
#pragma unroll
    for (int t = 0; t < IPT; ++t) {
      const int r = tid + t * nthreads;
      const int i = bd.row0 + r;
      if (i < bd.row1) {
        if (whole) {
#pragma unroll
          for (int k = 0; k < D; ++k) a.Srow[(int64_t)k * a.ld + i] = fi[t][k];
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

// S_i = (row sums of i, from S itself or from the CTA partial slots in CTA order)
//       - (column sums of i, one term per block of the same segment above it, in block order).
template <typename T, int D>
__global__ void __launch_bounds__(256) k_sym_reduce(const RepSymArgs<T> a, int grid, int64_t len) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len) return;
  const SymTileRef tr = a.tiles[i / kTileJ];
  T acc[D];
#pragma unroll
  for (int k = 0; k < D; ++k) acc[k] = (T)0;
  bool live = true;
  if (tr.row_block >= 0) {
    const SymBlockDesc bd = a.blocks[tr.row_block];
    live = i < bd.row1;  // padding rows behind the end of a segment
    if (live) {
      const int r = (int)(i - bd.row0);
      const long long W = a.total_units, G = grid;
      const long long U0 = bd.unit0, U1 = bd.unit0 + bd.ntiles;
      long long c = U0 * G / W;
      while (c + 1 < G && W * (c + 1) / G <= U0) ++c;
      while (c > 0 && W * c / G > U0) --c;
      if (W * c / G <= U0 && W * (c + 1) / G >= U1) {  // swept whole by one CTA
#pragma unroll
        for (int k = 0; k < D; ++k) acc[k] = a.Srow[(int64_t)k * a.ld + i];
      } else {
        for (long long cc = c; cc < G && W * cc / G < U1; ++cc) {
          const long long v0 = W * cc / G, v1 = W * (cc + 1) / G;
          if (v1 <= U0 || v0 >= v1) continue;
          const int slot = (max(v0, U0) == v0) ? 0 : 1;
#pragma unroll
          for (int k = 0; k < D; ++k)
            acc[k] += a.partial[(((size_t)cc * 2 + slot) * D + k) * a.rows_per_block + r];
        }
      }
    }
  }
  if (live) {
#pragma unroll 4
    for (int bb = tr.col_b0; bb < tr.col_b0 + tr.col_n; ++bb) {
      const int c0 = __ldg(&a.blocks[bb].col_t0) * kTileJ;
      const int nc = __ldg(&a.blocks[bb].ncols);
      const long long off = __ldg(&a.blocks[bb].col_off);
      if (i >= c0 && i < (int64_t)c0 + nc) {
#pragma unroll
        for (int k = 0; k < D; ++k) acc[k] -= a.colpartial[off + (int64_t)k * nc + (i - c0)];
      }
    }
  }
  if (a.accumulate) {  // a later pass over another column panel: add to what the earlier ones left
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] += a.S[(int64_t)k * a.ld + i];
  }
  if (a.out_scale != (T)0) {
    const T sc = a.mass[i] * a.out_scale;
#pragma unroll
    for (int k = 0; k < D; ++k) acc[k] *= sc;
  }
#pragma unroll
  for (int k = 0; k < D; ++k) a.S[(int64_t)k * a.ld + i] = acc[k];
}

// ---------------------------------------------------------------------------------------------
namespace {
int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

template <typename T, int D>
const void* sym_kernel_d(int ipt, int cg, int shrink) {
  if (shrink == 2) return cg >= 8 ? (const void*)k_repulsion_sym<T, D, 2, 8, 2> : (const void*)k_repulsion_sym<T, D, 2, 4, 2>;
  if (shrink == 1) return cg >= 8 ? (const void*)k_repulsion_sym<T, D, 2, 8, 1> : (const void*)k_repulsion_sym<T, D, 2, 4, 1>;
  if (ipt >= 4) return cg >= 8 ? (const void*)k_repulsion_sym<T, D, 4, 8> : (const void*)k_repulsion_sym<T, D, 4, 4>;
  return cg >= 8 ? (const void*)k_repulsion_sym<T, D, 2, 8> : (const void*)k_repulsion_sym<T, D, 2, 4>;
}
template <typename T>
const void* sym_kernel(int dim, int ipt, int cg, int shrink) {
  return dim == 2 ? sym_kernel_d<T, 2>(ipt, cg, shrink) : sym_kernel_d<T, 3>(ipt, cg, shrink);
}
template <typename T>
size_t sym_smem(int dim, int threads) {
  return (size_t)kRepStages * (dim + 1) * kTileJ * sizeof(T) +
         (size_t)(threads / 32) * dim * kTileJ * sizeof(T) + kRepStages * sizeof(uint64_t);
}

// The triangular unit list of a set of segments: inside segment [s0, s1), block g covers rows
// [s0 + g*RB, min(s1, s0 + (g+1)*RB)) and the column tiles from its own first row to the end of the
// segment.  `blocks` are the blocks clipped to units [U0, U1) of the concatenated list (a rank's
// share); every block of the triangle, in the plan or not, has a global ordinal.
struct SymBlockFull {
  SymBlockDesc d;   // unit0 / col_off filled per pass
  int seg, gord;    // segment id, global ordinal of the block
};
struct SymLayout {
  std::vector<SymBlockFull> blocks;
  std::vector<int> tile_seg, tile_gord;  // per 256-row tile of [0, ld): its segment and block ordinal (-1: none)
  long long units = 0, colpartial_elems = 0, pairs = 0;
  int64_t reduce_len = 0;
};
SymLayout sym_layout(int dim, int64_t ld, const std::vector<SymSegment>& segs, int rb, int part, int parts) {
  auto seg_tiles = [&](const SymSegment& sg) { return (int)((sg.row1 - (int64_t)sg.row0 + kTileJ - 1) / kTileJ); };
  long long total = 0;
  for (const auto& sg : segs) {
    const int nt = seg_tiles(sg);
    const int nblk = (sg.row1 - sg.row0 + rb - 1) / rb;
    for (int g = 0; g < nblk; ++g) total += nt - (int)((int64_t)g * rb / kTileJ);
  }
  const long long U0 = total * part / parts, U1 = total * (part + 1) / parts;
  SymLayout L;
  L.tile_seg.assign((size_t)(ld / kTileJ), -1);
  L.tile_gord.assign((size_t)(ld / kTileJ), -1);
  long long prefix = 0;
  int gord = 0, seg_id = 0;
  for (const auto& sg : segs) {
    const int nt = seg_tiles(sg);
    const int t_seg = sg.row0 / kTileJ;  // global index of the segment's first tile
    const int nblk = (sg.row1 - sg.row0 + rb - 1) / rb;
    L.reduce_len = std::max<int64_t>(L.reduce_len, (int64_t)(t_seg + nt) * kTileJ);
    for (int g = 0; g < nblk; ++g, ++gord) {
      const int tf = (int)((int64_t)g * rb / kTileJ);  // first tile of the block, within the segment
      const long long b0 = prefix, b1 = prefix + (nt - tf);
      prefix = b1;
      const int row0 = sg.row0 + g * rb;
      const int row1 = (int)std::min<int64_t>(sg.row1, (int64_t)row0 + rb);
      const int bt1 = std::min(nt, tf + rb / kTileJ);  // the block's own row tiles [tf, bt1)
      for (int t = tf; t < bt1; ++t) {
        L.tile_seg[(size_t)t_seg + t] = seg_id;
        L.tile_gord[(size_t)t_seg + t] = gord;
      }
      const long long lo = std::max(b0, U0), hi = std::min(b1, U1);
      if (lo >= hi) continue;
      SymBlockFull f;
      f.seg = seg_id;
      f.gord = gord;
      SymBlockDesc& d = f.d;
      d.row0 = row0;
      d.row1 = row1;
      d.t_first = t_seg + tf + (int)(lo - b0);
      d.ntiles = (int)(hi - lo);
      d.tile_sym0 = t_seg + (int)((row1 - sg.row0 + kTileJ - 1) / kTileJ);
      d.col_t0 = std::max(d.t_first, d.tile_sym0);
      d.ncols = std::max(0, d.t_first + d.ntiles - d.col_t0) * kTileJ;
      d.unit0 = 0;
      d.col_off = 0;
      L.units += d.ntiles;
      L.colpartial_elems += (long long)dim * d.ncols;
      const long long rows = d.row1 - d.row0;
      L.pairs += rows * (long long)(d.col_t0 - d.t_first) * kTileJ + 2 * rows * (long long)d.ncols;
      L.blocks.push_back(f);
    }
    ++seg_id;
  }
  return L;
}

// One pass = the plan's blocks restricted to the column tiles [pt0, pt1): its own block list, unit
// numbering, column-slab offsets and per-tile references.
struct SymPass {
  std::vector<SymBlockDesc> blocks;
  std::vector<SymTileRef> tiles;
  long long units = 0, colpartial_elems = 0;
};
SymPass sym_pass(int dim, const SymLayout& L, int pt0, int pt1) {
  SymPass P;
  std::vector<int> seg_of, gord_of;
  for (const auto& f : L.blocks) {
    const int t0 = std::max(f.d.t_first, pt0), t1 = std::min(f.d.t_first + f.d.ntiles, pt1);
    if (t0 >= t1) continue;
    SymBlockDesc d = f.d;
    d.t_first = t0;
    d.ntiles = t1 - t0;
    d.col_t0 = std::max(d.t_first, d.tile_sym0);
    d.ncols = std::max(0, d.t_first + d.ntiles - d.col_t0) * kTileJ;
    d.unit0 = P.units;
    d.col_off = P.colpartial_elems;
    P.units += d.ntiles;
    P.colpartial_elems += (long long)dim * d.ncols;
    P.blocks.push_back(d);
    seg_of.push_back(f.seg);
    gord_of.push_back(f.gord);
  }
  // per row tile: the pass block that sweeps its rows, and the pass blocks of its segment above it
  P.tiles.assign(L.tile_seg.size(), SymTileRef{-1, 0, 0, 0});
  const int nb = (int)P.blocks.size();
  std::vector<int> seg_first;  // first pass block of each segment (pass blocks are segment-major)
  for (int b = 0; b < nb; ++b) {
    if ((int)seg_first.size() <= seg_of[b]) seg_first.resize(seg_of[b] + 1, -1);
    if (seg_first[seg_of[b]] < 0) seg_first[seg_of[b]] = b;
  }
  int cursor = 0;  // tiles ascend with (segment, block ordinal): one forward scan over the blocks
  for (size_t t = 0; t < P.tiles.size(); ++t) {
    const int sg = L.tile_seg[t], go = L.tile_gord[t];
    if (sg < 0) continue;
    while (cursor < nb && (seg_of[cursor] < sg || (seg_of[cursor] == sg && gord_of[cursor] < go))) ++cursor;
    const int first = (sg < (int)seg_first.size() && seg_first[sg] >= 0) ? seg_first[sg] : cursor;
    SymTileRef r;
    r.row_block = (cursor < nb && seg_of[cursor] == sg && gord_of[cursor] == go) ? cursor : -1;
    r.col_b0 = first;
    r.col_n = std::max(0, cursor - first);
    r.pad = 0;
    P.tiles[t] = r;
  }
  return P;
}
}  // namespace

// Column-tile panel [pt0, pt1) of pass q of npass: equal numbers of column-side entries per pass (the
// triangle's columns fill linearly, so the boundaries go with the square root).
void sym_pass_bounds(int ntile, int npass, int q, int& pt0, int& pt1) {
  pt0 = (int)std::floor(ntile * std::sqrt(double(q) / npass));
  pt1 = q == npass - 1 ? ntile : (int)std::floor(ntile * std::sqrt(double(q + 1) / npass));
}

void sym_pass_share(int64_t ld, int part, int parts, int npass, int q, std::vector<int>& out) {
  const SymLayout L = sym_layout(2, ld, {SymSegment{0, (int)ld}}, 1024, part, parts);
  int pt0, pt1;
  sym_pass_bounds((int)(ld / kTileJ), npass, q, pt0, pt1);
  out.clear();
  if (pt1 <= pt0) return;
  const SymPass P = sym_pass(2, L, pt0, pt1);
  for (const auto& d : P.blocks) {
    const int v[5] = {d.row0, d.row1, d.t_first, d.ntiles, d.tile_sym0};
    out.insert(out.end(), v, v + 5);
  }
}

void sym_share(int64_t ld, int part, int parts, std::vector<int>& out) {
  const SymLayout L = sym_layout(2, ld, {SymSegment{0, (int)ld}}, 1024, part, parts);
  out.clear();
  for (const auto& f : L.blocks) {
    const SymBlockDesc& d = f.d;
    const int v[5] = {d.row0, d.row1, d.t_first, d.ntiles, d.tile_sym0};
    out.insert(out.end(), v, v + 5);
  }
}

template <typename T>
double RepulsionSymPlan<T>::scratch_bytes(int dim, int64_t ld, int parts) {
  const int rb = 1024;  // both launch shapes (512 x 2, 256 x 4) cover 1024 rows per block
  // the triangle holds ~ ld^2 / (2 rb) column entries per dimension, shared out over the parts; the
  // plan cuts the sweep into passes over column panels so that one pass stays under the cap
  const double all = 1.05 * double(dim) * sizeof(T) * double(ld) * double(ld) / (2.0 * rb) / parts + 1e6;
  const double cap = 1048576.0 * env_int("GE_SYM_SCRATCH_MB", 2048);
  return std::min(all, 1.3 * cap + 1e6) + double(dim) * sizeof(T) * double(ld);
}

template <typename T>
RepulsionSymPlan<T>::RepulsionSymPlan(ge_context* ctx, int dim, int64_t ld, int part, int parts)
    : ctx_(ctx), dim_(dim), ld_(ld) {
  init({SymSegment{0, (int)ld}}, part, parts);
}

template <typename T>
RepulsionSymPlan<T>::RepulsionSymPlan(ge_context* ctx, int dim, int64_t ld,
                                      const std::vector<SymSegment>& segments)
    : ctx_(ctx), dim_(dim), ld_(ld) {
  init(segments, 0, 1);
}

template <typename T>
void RepulsionSymPlan<T>::init(const std::vector<SymSegment>& segments, int part, int parts) {
  ge_context* ctx = ctx_;
  // measured on B200 (tools/sweep_sym.py, n = 300k): FP64 is best with 2 rows per thread and 512
  // threads (d = 2: 53.8 ms vs 54.3; d = 3: 68.6 vs 73.8), FP32 with 4 rows and 256 threads
  // (25.8 vs 27.5; 32.1 vs 38.0); 8-column groups beat 4-column groups everywhere by 5-8 %
  ipt_ = env_int("GE_SYM_IPT", sizeof(T) == 8 ? 2 : 4) >= 4 ? 4 : 2;
  cg_ = env_int("GE_SYM_CG", 8) >= 8 ? 8 : 4;
  {  // smaller blocks when the segments are short on average (never for the flat sweep's one segment)
    long long rows = 0;
    for (const auto& sg : segments) rows += sg.row1 - sg.row0;
    const long long nseg = (long long)segments.size();
    shrink_ = 0;
    if (parts == 1 && nseg > 1) shrink_ = rows < 2048 * nseg ? 2 : rows < 4096 * nseg ? 1 : 0;
    if (const char* e = std::getenv("GE_SYM_HALF")) shrink_ = std::atoi(e) != 0 ? 1 : 0;
    if (const char* e = std::getenv("GE_SYM_SHRINK")) shrink_ = std::max(0, std::min(2, std::atoi(e)));
    if (shrink_) ipt_ = 2;
  }
  threads_ = shrink_ == 2 ? 128 : (ipt_ >= 4 || shrink_ == 1) ? 256 : 512;
  const int rb = threads_ * ipt_;
  const void* fn = sym_kernel<T>(dim_, ipt_, cg_, shrink_);
  const size_t smem = sym_smem<T>(dim_, threads_);
  if (smem > 48 * 1024)
    GE_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  GE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, threads_, smem));
  GE_REQUIRE(occ > 0, "symmetric repulsion kernel does not fit on an SM");
  for (const auto& sg : segments)
    GE_REQUIRE(sg.row0 % kTileJ == 0 && sg.row0 <= sg.row1 && sg.row1 <= ld_, "bad segment");
  SymLayout L = sym_layout(dim_, ld_, segments, rb, part, parts);
  pairs_ = L.pairs;
  reduce_len_ = parts > 1 ? ld_ : L.reduce_len;  // multi-rank: every row's sums enter the reduce-scatter
  rb_ = rb;
  // The column-side slabs hold n^2 / (2 rb) entries per dimension.  When that exceeds the cap the
  // sweep is cut into passes over column panels that reuse one scratch buffer (same kernels, same
  // pairs; k_sym_reduce adds each pass's sums to the previous ones), so memory stays bounded for
  // any n instead of falling back to the ordered sweep.
  const double cap_elems = 1048576.0 * env_int("GE_SYM_SCRATCH_MB", 2048) / sizeof(T);
  int npass = (int)std::min<double>(64.0, std::ceil(std::max(1.0, double(L.colpartial_elems) / cap_elems)));
  npass = std::max(1, env_int("GE_SYM_PASSES", npass));
  const int ntile = (int)(ld_ / kTileJ);
  long long max_units = 1, max_col = 1;
  size_t max_blocks = 1;
  std::vector<SymPass> passes;
  // equal numbers of column-side entries per pass: the triangle's columns fill linearly, so the
  // panel boundaries go with the square root
  for (int q = 0; q < npass; ++q) {
    int pt0, pt1;
    sym_pass_bounds(ntile, npass, q, pt0, pt1);
    if (pt1 <= pt0) continue;
    passes.push_back(sym_pass(dim_, L, pt0, pt1));
    if (passes.back().units == 0) {
      passes.pop_back();
      continue;
    }
    max_units = std::max(max_units, passes.back().units);
    max_col = std::max(max_col, passes.back().colpartial_elems);
    max_blocks = std::max(max_blocks, passes.back().blocks.size());
  }
  grid_ = (int)std::min<long long>((long long)ctx->sm_count * occ, max_units);
  pass_.clear();
  for (auto& P : passes) {
    PassDev pd;
    pd.nblocks = (int)P.blocks.size();
    pd.units = P.units;
    pd.grid = (int)std::min<long long>(grid_, std::max<long long>(P.units, 1));
    pd.blocks.alloc(ctx, std::max<size_t>(P.blocks.size(), 1));
    pd.blocks.upload(ctx, P.blocks.data(), P.blocks.size());
    pd.tiles.alloc(ctx, std::max<size_t>(P.tiles.size(), 1));
    pd.tiles.upload(ctx, P.tiles.data(), P.tiles.size());
    pass_.push_back(std::move(pd));
  }
  total_units_ = L.units;
  nblocks_ = (int)L.blocks.size();
  partial_.alloc(ctx, (size_t)grid_ * 2 * dim_ * rb);
  srow_.alloc(ctx, (size_t)dim_ * ld_);
  colpartial_elems_ = (size_t)max_col;
  colpartial_.alloc(ctx, colpartial_elems_);
  GE_CUDA(cudaStreamSynchronize(ctx->stream));
  if (std::getenv("GE_VERBOSE"))
    std::fprintf(stderr,
                 "[ge] symmetric repulsion plan: segments=%zu threads=%d ipt=%d cg=%d grid=%d (occ %d) "
                 "blocks=%d units=%lld passes=%zu column scratch %.1f MB (one pass would need %.1f MB)\n",
                 segments.size(), threads_, ipt_, cg_, grid_, occ, nblocks_, total_units_, pass_.size(),
                 double(max_col) * sizeof(T) / 1e6, double(L.colpartial_elems) * sizeof(T) / 1e6);
}

template <typename T>
void RepulsionSymPlan<T>::launch(const T* pos, const T* mass, T* S, T eps2, T out_scale) {
  if (colpartial_.size() == 0) colpartial_.alloc(ctx_, colpartial_elems_);
  bool first = true;
  for (size_t q = 0; q < pass_.size(); ++q) {
    const PassDev& pd = pass_[q];
    const bool last = q + 1 == pass_.size();
    RepSymArgs<T> a;
    a.pos = pos;
    a.mass = mass;
    a.S = S;
    a.Srow = srow_.get();
    a.partial = partial_.get();
    a.colpartial = colpartial_.get();
    a.blocks = pd.blocks.get();
    a.tiles = pd.tiles.get();
    a.ld = ld_;
    a.total_units = std::max<long long>(pd.units, 1);
    a.nblocks = pd.nblocks;
    a.rows_per_block = rb_;
    a.eps2 = eps2;
    a.out_scale = last ? out_scale : (T)0;
    a.accumulate = first ? 0 : 1;
    if (pd.nblocks > 0 && pd.units > 0) {
      void* args[] = {(void*)&a};
      GE_CUDA(cudaLaunchKernel(sym_kernel<T>(dim_, ipt_, cg_, shrink_), dim3(pd.grid), dim3(threads_), args,
                               sym_smem<T>(dim_, threads_), ctx_->stream));
      ctx_->launches++;
    }
    if (reduce_len_ > 0) {
      const unsigned rgrid = (unsigned)((reduce_len_ + 255) / 256);
      if (dim_ == 2) k_sym_reduce<T, 2><<<rgrid, 256, 0, ctx_->stream>>>(a, pd.grid, reduce_len_);
      else k_sym_reduce<T, 3><<<rgrid, 256, 0, ctx_->stream>>>(a, pd.grid, reduce_len_);
      GE_CUDA(cudaGetLastError());
      ctx_->launches++;
    }
    first = false;
  }
  if (pass_.empty() && reduce_len_ > 0) {  // a rank without any pair: its sums are zero
    GE_CUDA(cudaMemsetAsync(S, 0, sizeof(T) * (size_t)dim_ * ld_, ctx_->stream));
  }
}

template class RepulsionSymPlan<double>;
template class RepulsionSymPlan<float>;

}  // namespace ge
