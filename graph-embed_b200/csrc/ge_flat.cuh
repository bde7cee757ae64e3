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
  const T* Frep;       // [D][ldf] (owned rows)
  T* Fprev;            // [D][ldf]
  const T* mass;       // [ld]
  const T* Eext;       // [D][ldf] multilevel external-pull numerators, or nullptr
  int64_t ld, ldf;
  int row0, nrows;
  int update;          // 0: only write the total force into Fprev (parity hook)
  Physics<T> ph;
};

// CSR attraction + gravity + step; `group` lanes per row; ml selects the multilevel clamps.
template <typename T>
void launch_attract_step(ge_context* ctx, const StepArgs<T>& a, int dim, int group, bool ml);

int group_for_degree(double avg_deg);

}  // namespace ge
#endif
