// graph-embed_b200 :: argument block and launchers of the on-chip (persistent) solvers.
#ifndef GE_ONCHIP_CUH
#define GE_ONCHIP_CUH

#include "ge_context.h"

namespace ge {

// All per-vertex arrays are indexed by SLOT: an internal ordering of the vertices in which the
// members of one aggregate are contiguous and keep the member order of their P_T row (for the
// flat solve slot == vertex id).  Host-facing arrays (init_aos, out_aos, cA_aos) are the row-major
// double images the C ABI carries, indexed by global vertex / aggregate id.
template <typename T>
struct OnchipArgs {
  const double* init_aos = nullptr;  // [n][D] positions at iteration 0, by global vertex id
  const int* vtx = nullptr;          // slot -> global vertex id (nullptr: identity)
  const int* agg_of_slot = nullptr;  // slot -> aggregate id (multilevel)
  const T* mass = nullptr;           // [slots] c = deg + 1
  const int* e_begin = nullptr;      // per slot: attraction entries [e_begin, e_end) in e_idx/e_w
  const int* e_end = nullptr;
  const int* e_idx = nullptr;        // neighbour, as an absolute slot
  const T* e_w = nullptr;            // weight (nullptr: 1)
  const T* Eext = nullptr;           // [D][ld] external-pull numerators (multilevel)
  int64_t ld = 0;
  const double* cA_aos = nullptr;    // [m][D] parent centres
  const double* rA = nullptr;        // [m]    parent radii
  const int4* tasks = nullptr;       // CTA tier: {slot0, size, aggregate, -}
                                     // warp tier: {slot0, size, aggregates in the pack, -}
  double* out_aos = nullptr;         // [n][D] final coordinates (or forces when forces_only)
  int iters = 0;
  int forces_only = 0;
  int debug_skip = 0;                // measurement only (GE_ONCHIP_SKIP): 1 = no pair loop, 2 = no epilogue, 4 = no barrier
  int normalize = 0;                 // flat epilogue of include/forceatlas.hpp:272-303
  int exchange = 0;                  // cluster kernels: 1 = st.async + mbarrier, 0 = DSMEM stores + cluster barrier
  Physics<T> ph;
};

// One CTA per task (aggregate of 33..1024 members, or the whole coarsest-level graph).
template <typename T>
void launch_onchip_cta(ge_context* ctx, const OnchipArgs<T>& a, int ntasks, int dim, bool ml,
                       int lanes, int threads, int max_size);
// One warp per pack of equal-size aggregates (2..32 members, also singletons in forces mode).
template <typename T>
void launch_onchip_warp(ge_context* ctx, const OnchipArgs<T>& a, int npacks, int dim);

}  // namespace ge
#endif
