// graph-embed_b200 :: argument blocks and launchers of the tiled multi-CTA kernels (K1).
#ifndef GE_FLAT_CUH
#define GE_FLAT_CUH

#include "ge_context.h"

namespace ge {

constexpr int kTileJ = 256;      // column-tile entries
constexpr int kRepStages = 3;    // TMA pipeline depth
constexpr int kRepMaxThreads = 512;

struct BlockDesc {
  int row0, row1;  // rows of this CTA: [row0, row1)
  int j0, j1;      // column range (multiples of kTileJ)
};

template <typename T>
struct RepArgs {
  const T* pos;    // [D][ld]
  const T* mass;   // [kMassArrays][ld]
  T* F;            // [D][ldf]  rows indexed (i - f_row_base)
  const BlockDesc* blocks;
  int64_t ld, ldf;
  int f_row_base;
  T repel, eps2;
};

template <typename T>
struct StepArgs {
  const int* e_begin;  // per owned row: first / one-past-last entry in J, W
  const int* e_end;
  const int* J;
  const T* W;          // nullptr: unit weights
  const T* pos_cur;    // [D][ld]
  T* pos_next;         // [D][ld]
  const T* Frep;       // [D][ldf] (owned rows)
  T* Fprev;            // [D][ldf]
  const T* mass;       // [ld]
  const T* Eext;       // [D][ldf] multilevel external-pull numerators, or nullptr
  int64_t ld, ldf;
  int row0, nrows;
  int update;          // 0: only write the total force into Fprev (parity hook)
  Physics<T> ph;
};

template <typename T>
size_t repulsion_smem(int dim);
template <typename T>
const void* repulsion_kernel(int dim, int ipt);
// Tiled all-pairs repulsion over the row blocks in a.blocks.
template <typename T>
void launch_repulsion(ge_context* ctx, const RepArgs<T>& a, int nblocks, int threads, int ipt, int dim);
// CSR attraction + gravity + step; `group` lanes per row; ml selects the multilevel clamps.
template <typename T>
void launch_attract_step(ge_context* ctx, const StepArgs<T>& a, int dim, int group, bool ml);

int group_for_degree(double avg_deg);

}  // namespace ge
#endif
