// graph-embed_b200 :: argument blocks and launchers of the tiled multi-CTA kernels (K1).
#ifndef GE_FLAT_CUH
#define GE_FLAT_CUH

#include <vector>

#include "ge_context.h"

namespace ge {

constexpr int kTileJ = 256;      // column-tile entries (large sweeps); arrays are padded to this
constexpr int kTileJSmall = 64;  // column-tile entries for small, latency-bound sweeps
constexpr int kRepStages = 3;    // TMA pipeline depth
constexpr int kRepMaxThreads = 512;

// One row block of the all-pairs sweep: rows [row0,row1) against ntiles column tiles starting at
// column j0.  unit0 is the exclusive prefix sum of ntiles over the block list: the sweep is a flat
// sequence of (row block, column tile) units that is cut into equal contiguous shares, one per
// resident CTA ("stream-K"), so the 148 SMs finish together whatever n is.
struct BlockDesc {
  int row0, row1;
  int j0, ntiles;
  long long unit0;
};

struct RowSegment {  // rows [row0,row1) interact with columns [j0,j1) (j0, j1 multiples of kTileJ)
  int row0, row1, j0, j1;
};

template <typename T>
struct RepArgs {
  const T* pos;    // [D][ld]
  const T* mass;   // [kMassArrays][ld]
  T* F;            // [D][ldf]  rows indexed (i - f_row_base)
  T* partial;      // [grid][2][D][rows_per_block] raw sums of row blocks shared between CTAs
  const BlockDesc* blocks;
  int64_t ld, ldf;
  long long total_units;
  int nblocks, rows_per_block;
  int f_row_base;
  T repel, eps2;
};

// Launch plan of the repulsion sweep over a set of row segments (flat solve: one segment covering
// every column; multilevel: one segment per large aggregate).
template <typename T>
class RepulsionPlan {
 public:
  RepulsionPlan(ge_context* ctx, int dim, const std::vector<RowSegment>& segments);
  // F[k][i - f_row_base] = (deg_i + 1) * repel * sum_j ...   for every row of every segment
  void launch(const T* pos, const T* mass, int64_t ld, T* F, int64_t ldf, int f_row_base, T repel,
              T eps2);
  int threads() const { return threads_; }
  int ipt() const { return ipt_; }
  int grid() const { return grid_; }

 private:
  ge_context* ctx_;
  int dim_, threads_ = 512, ipt_ = 2, ju_ = 1, tile_ = kTileJ, grid_ = 0, nblocks_ = 0;
  long long total_units_ = 0;
  DevBuf<BlockDesc> blocks_;
  DevBuf<T> partial_;
};

// ---- symmetric all-pairs sweep (ge_flat_sym.cu) ------------------------------------------------
// The pair term is antisymmetric ((xi-xj) ci cj repel / dis^3, include/forceatlas.hpp:154-165), so
// every unordered pair is evaluated once and applied to both endpoints.  The sweep covers the
// upper triangle of (row block, column tile) units: tiles inside the row block's own rows are
// evaluated in full (row side only), tiles to the right of it feed the row AND the column.
struct SymBlockDesc {
  int row0, row1;     // rows of the block
  int t_first;        // first column tile (global tile index) of this plan's units in the block
  int ntiles;         // number of units (consecutive tiles)
  int tile_sym0;      // tiles >= tile_sym0 lie to the right of the block: symmetric
  int col_t0, ncols;  // column slab of this block in `colpartial`: tiles from col_t0, ncols columns
  long long unit0;    // exclusive prefix sum of ntiles
  long long col_off;  // element offset of the slab, laid out [D][ncols]
};

// Where the rows of one 256-entry tile get their sums from (k_sym_reduce): the plan block that
// sweeps them (row side; -1: none in this plan), and the plan blocks of the same segment that lie
// above them and therefore pushed on them from the column side.
struct SymTileRef {
  int row_block;  // index into the plan's block list, or -1
  int col_b0;     // first plan block of the same segment
  int col_n;      // number of plan blocks of that segment strictly above this tile's block
  int pad;
};

template <typename T>
struct RepSymArgs {
  const T* pos;      // [D][ld]
  const T* mass;     // [ld]  c = deg + 1 (0 on padding)
  T* S;              // [D][ld] raw sums: sum_j c_j (xi-xj)/dis^3 over every pair this plan owns
  T* Srow;           // [D][ld] row sums of blocks swept whole by one CTA (input of k_sym_reduce)
  T* partial;        // [grid][2][D][rows_per_block] row sums of blocks shared between CTAs
  T* colpartial;     // column-side sums, one slab per block (written exactly once per launch)
  const SymBlockDesc* blocks;
  const SymTileRef* tiles;  // [ld / kTileJ]
  int64_t ld;
  long long total_units;
  int nblocks, rows_per_block;
  T eps2;
  T out_scale;       // != 0: S = sums * c_i * out_scale (the force itself); 0: raw sums
  int accumulate;    // k_sym_reduce: add to S (a later pass over another column panel) instead of assigning
};

struct SymSegment {  // rows == columns [row0, row1); row0 a multiple of kTileJ
  int row0, row1;
};

// Launch plan of the symmetric sweep over a set of disjoint segments (flat solve: one segment
// [0, ld); multilevel: one segment per large aggregate, all pairs inside each): share `part` of
// `parts` equal contiguous cuts of the triangular unit list (parts > 1: every rank produces partial
// sums over the full length, which the ranks add with a reduce-scatter).
template <typename T>
class RepulsionSymPlan {
 public:
  RepulsionSymPlan(ge_context* ctx, int dim, int64_t ld, int part, int parts);
  RepulsionSymPlan(ge_context* ctx, int dim, int64_t ld, const std::vector<SymSegment>& segments);
  // bytes of column-side scratch the plan would need (worst share)
  static double scratch_bytes(int dim, int64_t ld, int parts);
  // S[k][i] = sum over this plan's pairs; multiply by c_i * repel to obtain the force
  // (out_scale != 0: the kernel does it, S = sums * c_i * out_scale)
  void launch(const T* pos, const T* mass, T* S, T eps2, T out_scale = (T)0);
  long long pairs() const { return pairs_; }  // ordered pairs covered by one launch
  int grid() const { return grid_; }

 private:
  void init(const std::vector<SymSegment>& segments, int part, int parts);
  struct PassDev {  // one pass over a column panel: its block list and tile references
    DevBuf<SymBlockDesc> blocks;
    DevBuf<SymTileRef> tiles;
    int nblocks = 0, grid = 0;
    long long units = 0;
  };
  ge_context* ctx_;
  int dim_, threads_ = 256, ipt_ = 4, cg_ = 8, grid_ = 0, nblocks_ = 0, rb_ = 1024;
  int shrink_ = 0;  // 1: 512-row blocks (256 threads x 2 rows), 2: 256-row blocks: plans over many short segments
  int64_t ld_ = 0, reduce_len_ = 0;
  long long total_units_ = 0, pairs_ = 0;
  std::vector<PassDev> pass_;
  DevBuf<T> partial_, colpartial_, srow_;
  size_t colpartial_elems_ = 0;
};

// The (row0, row1, tile_first, ntiles, tile_sym0) quintuples of share `part` of `parts` (host).
void sym_share(int64_t ld, int part, int parts, std::vector<int>& out);
// The same for pass q of npass column-panel passes of that share (RepulsionSymPlan's cut).
void sym_pass_share(int64_t ld, int part, int parts, int npass, int q, std::vector<int>& out);

template <typename T>
struct StepArgs {
  const int* e_begin;  // per owned row: first / one-past-last entry in J, W
  const int* e_end;
  const int* J;
  const T* W;          // nullptr: unit weights
  const T* pos_cur;    // [D][ld]
  T* pos_next;         // [D][ld]
  // Interleaved copies of the coordinates for the gather side: one 16-byte (d = 2) or 32-byte
  // (d = 3, padded to 4 reals; FP32: 8 / 16 bytes) record per vertex, so a neighbour costs one
  // L2 sector instead of one per dimension.  nullptr: gather from the SoA arrays.
  const T* aos_cur = nullptr;  // [ld][DP]
  T* aos_next = nullptr;       // [ld][DP]
  const T* Frep;       // [D][ldr] (owned rows)
  int64_t ldr = 0;     // leading dimension of Frep (0: ldf)
  T frep_scale = 0;    // != 0: Frep holds raw pair sums, force = Frep * c_i * frep_scale
  T* Fprev;            // [D][ldf]
  const T* mass;       // [ld]
  const T* Eext;       // [D][ldf] multilevel external-pull numerators, or nullptr
  int64_t ld, ldf;
  int row0, nrows;
  int update;          // 0: only write the total force into Fprev (parity hook)
  // > 0: rows with more entries than this are left to k_attract_step_long (one CTA per row); the
  // row kernels skip them entirely.  Power-law graphs: half of the entries of R-MAT-18 sit in the
  // 2 % of rows longer than 512 entries, the longest has 25 000.
  int long_threshold = 0;
  Physics<T> ph;
};

// CSR attraction + gravity + step; `group` lanes per row; ml selects the multilevel clamps.
template <typename T>
void launch_attract_step(ge_context* ctx, const StepArgs<T>& a, int dim, int group, bool ml);
// The rows listed in `rows` (local row indices, nlong of them), one CTA each; ml: the multilevel
// clamps and external pull, 128 threads per row.
template <typename T>
void launch_attract_step_long(ge_context* ctx, const StepArgs<T>& a, int dim, const int* rows, int nlong,
                              int threads, bool ml = false);

int group_for_degree(double avg_deg);

}  // namespace ge
#endif
