// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// extern "C" entry points around the reference's own functions, linked against
// the reference's unmodified src/embed.cpp, src/partitioner.cpp,
// src/matrixutils.cpp and src/export.cpp (compiled where they lie under /root/reference by
// oracle/Makefile; outputs only in oracle/_ref/).  Loaded with ctypes by
// tests/ and by bench.py's CPU-baseline / --impl reference legs.
#include <omp.h>

#include <cstring>
#include <vector>

#include "embed.hpp"
#include "export.hpp"

static unsigned g_seed = 0;
extern "C" unsigned ge_ref_current_seed(void) { return g_seed; }

namespace partition {
// Strong symbols defined by the reference's header-only forceatlas.hpp inside
// embed.o (include/forceatlas.hpp:89-103, 314-331); re-declared here because
// the header cannot be included in a second translation unit.
void forceAtlas(const SparseMatrix& A, const int dim, std::vector<std::vector<double>>& coords,
                const int iterations, const double ks, const double ksmax, const double repel,
                const double attract, const double gravity, const bool useWeights,
                const bool linlog, const bool nohubs, const double delta, const double tolerate,
                const bool normalize);
void forceAtlasMultilevel(const SparseMatrix& A, const SparseMatrix& P, const std::vector<int>& v_A,
                          const std::vector<std::vector<double>>& coords_A,
                          const std::vector<double>& r_A, std::vector<std::vector<double>>& coords,
                          int dim, int iterations, double ks, double ksmax, bool useWeights,
                          bool linlog, bool nohubs, double repel, double attract, double gravity,
                          double delta, double tolerate);
}  // namespace partition

namespace {

struct Params {
  int iterations;
  double ks, ksmax, repel, attract, gravity, delta, tolerate;
  int useWeights, linlog, nohubs, normalize;
};

SparseMatrix make_csr(int rows, int cols, const int* I, const int* J, const double* D) {
  const int nnz = I[rows];
  std::vector<double> data(nnz, 1.0);
  if (D) data.assign(D, D + nnz);
  return SparseMatrix(std::vector<int>(I, I + rows + 1), std::vector<int>(J, J + nnz),
                      std::move(data), rows, cols);
}

std::vector<std::vector<double>> unflatten(const double* x, int n, int d) {
  std::vector<std::vector<double>> out(n, std::vector<double>(d));
  for (int i = 0; i < n; i++)
    for (int k = 0; k < d; k++) out[i][k] = x[(size_t)i * d + k];
  return out;
}

void flatten(const std::vector<std::vector<double>>& c, int d, double* x) {
  for (size_t i = 0; i < c.size(); i++)
    for (int k = 0; k < d; k++) x[i * d + k] = c[i][k];
}

std::vector<SparseMatrix> g_hierarchy;  // last result of ref_partition

}  // namespace

extern "C" {

void ref_set_seed(unsigned seed) { g_seed = seed; }
int ref_max_threads(void) { return omp_get_max_threads(); }

// partition::forceAtlas, 15-argument form.  random_init != 0 passes empty coords
// (the reference then draws them itself from the seeded generator).
void ref_flat_forceatlas(int n, const int* I, const int* J, const double* D, int dim,
                         double* coords, int random_init, const Params* p, int nthreads) {
  omp_set_num_threads(nthreads);
  SparseMatrix A = make_csr(n, n, I, J, D);
  std::vector<std::vector<double>> c;
  if (!random_init) c = unflatten(coords, n, dim);
  partition::forceAtlas(A, dim, c, p->iterations, p->ks, p->ksmax, p->repel, p->attract,
                        p->gravity, p->useWeights != 0, p->linlog != 0, p->nohubs != 0, p->delta,
                        p->tolerate, p->normalize != 0);
  flatten(c, dim, coords);
}

// partition::forceAtlasMultilevel, 18-argument form.
void ref_multilevel_forceatlas(int n, const int* I, const int* J, const double* D, int m,
                               const int* PI, const int* PJ, const int* v_A,
                               const double* coords_A, const double* r_A, int dim,
                               const Params* p, double* coords_out, int nthreads) {
  omp_set_num_threads(nthreads);
  SparseMatrix A = make_csr(n, n, I, J, D);
  SparseMatrix P = make_csr(m, n, PI, PJ, nullptr);
  std::vector<int> vA(v_A, v_A + n);
  std::vector<std::vector<double>> cA = unflatten(coords_A, m, dim);
  std::vector<double> rA(r_A, r_A + m);
  std::vector<std::vector<double>> c(n, std::vector<double>(dim));
  partition::forceAtlasMultilevel(A, P, vA, cA, rA, c, dim, p->iterations, p->ks, p->ksmax,
                                  p->useWeights != 0, p->linlog != 0, p->nohubs != 0, p->repel,
                                  p->attract, p->gravity, p->delta, p->tolerate);
  flatten(c, dim, coords_out);
}

static void build_levels(int L, const int* An, const int* const* AI, const int* const* AJ,
                         const double* const* AD, const int* Pm, const int* const* PI,
                         const int* const* PJ, std::vector<SparseMatrix>& As,
                         std::vector<SparseMatrix>& Ps) {
  for (int l = 0; l <= L; l++) As.push_back(make_csr(An[l], An[l], AI[l], AJ[l], AD[l]));
  for (int l = 0; l < L; l++) Ps.push_back(make_csr(Pm[l], An[l], PI[l], PJ[l], nullptr));
}

// partition::embed (src/embed.cpp:561).  Returns wall seconds of the embed call,
// bracketed like examples/embedder.cpp:219-222.
double ref_embed(int L, const int* An, const int* const* AI, const int* const* AJ,
                 const double* const* AD, const int* Pm, const int* const* PI,
                 const int* const* PJ, int dim, double* coords_out, int nthreads) {
  omp_set_num_threads(nthreads);
  std::vector<SparseMatrix> As, Ps;
  build_levels(L, An, AI, AJ, AD, Pm, PI, PJ, As, Ps);
  linalgcpp::Timer timer(linalgcpp::Timer::Start::True);
  std::vector<std::vector<double>> c = partition::embed(As, Ps, dim);
  timer.Click();
  flatten(c, dim, coords_out);
  return timer[0];
}

// Radii pin: runs embedMultilevel(level 1) to obtain the coarse inputs, then
// embedViaMultilevel(level 0) with a capturing functor, which receives exactly
// the r_A / rescaled coords_A that src/embed.cpp:166-329 (== :615-778) computes.
// Outputs: coords_A_in (m x dim, before rescale), r_Ac (mc or 0), coords_Ac,
// coords_A_out / r_A_out (m) as handed to the level-0 embedder.
// Returns the number of valid entries in r_Ac (0 => base case).
int ref_radii_case(int L, const int* An, const int* const* AI, const int* const* AJ,
                   const double* const* AD, const int* Pm, const int* const* PI,
                   const int* const* PJ, int dim, double* coords_A_in, double* r_Ac_out,
                   double* coords_Ac_out, double* coords_A_out, double* r_A_out) {
  omp_set_num_threads(1);
  std::vector<SparseMatrix> As, Ps;
  build_levels(L, An, AI, AJ, AD, Pm, PI, PJ, As, Ps);
  std::vector<double> r_Ac;
  std::vector<std::vector<double>> coords_Ac;
  std::vector<std::vector<double>> cA = partition::embedMultilevel(As, Ps, dim, 1, r_Ac, coords_Ac);
  flatten(cA, dim, coords_A_in);
  for (size_t i = 0; i < r_Ac.size(); i++) r_Ac_out[i] = r_Ac[i];
  flatten(coords_Ac, dim, coords_Ac_out);
  std::vector<double> r_A;
  std::vector<std::vector<double>> coords_A;
  auto capture = [&](const SparseMatrix&, const SparseMatrix&, const std::vector<int>&,
                     const std::vector<std::vector<double>>& cAr, const std::vector<double>& rAr,
                     std::vector<std::vector<double>>& coords, const int d) {
    flatten(cAr, d, coords_A_out);
    for (size_t i = 0; i < rAr.size(); i++) r_A_out[i] = rAr[i];
    for (auto& row : coords) row.assign(d, 0.0);
  };
  partition::embedViaMultilevel(As, Ps, dim, 0, r_A, coords_A, capture);
  return (int)r_Ac.size();
}

// partition::partition(A, coarseningFactor, false, true, 1.0, matchingIterations, false)
// (call shape of examples/embedder.cpp:187).  Returns the number of levels; the
// P_T matrices are fetched with ref_hierarchy_rows / ref_hierarchy_get.
// The caller-side line examples/embedder.cpp:215, `P.Mult(A).Mult(P.Transpose())`, evaluated with
// the stand-in linalgcpp container (host/compat/sparsematrix.hpp: row-wise Gustavson product,
// ascending columns).  Returns nnz; out_idx / out_val need room for nnz(A) entries.
long ref_galerkin(int n, const int* I, const int* J, const double* D, int m, const int* PI,
                  const int* PJ, int* out_ptr, int* out_idx, double* out_val) {
  const int nnz = I[n];
  SparseMatrix A(std::vector<int>(I, I + n + 1), std::vector<int>(J, J + nnz),
                 std::vector<double>(D, D + nnz), n, n);
  SparseMatrix P(std::vector<int>(PI, PI + m + 1), std::vector<int>(PJ, PJ + n),
                 std::vector<double>(n, 1.0), m, n);
  const SparseMatrix C = P.Mult(A).Mult(P.Transpose());
  std::memcpy(out_ptr, C.GetIndptr().data(), sizeof(int) * (m + 1));
  std::memcpy(out_idx, C.GetIndices().data(), sizeof(int) * C.GetIndices().size());
  std::memcpy(out_val, C.GetData().data(), sizeof(double) * C.GetData().size());
  return (long)C.GetIndices().size();
}

// partition::writeCoords / writePartition (src/export.cpp:16-39) on caller-supplied arrays.
void ref_write_coords(const double* x, int n, int d, const char* path) {
  std::vector<std::vector<double>> coords(n, std::vector<double>(d));
  for (int i = 0; i < n; i++)
    for (int k = 0; k < d; k++) coords[i][k] = x[(size_t)i * d + k];
  partition::writeCoords(coords, path);
}
void ref_write_partition(const int* part, int n, const char* path) {
  partition::writePartition(std::vector<int>(part, part + n), path);
}

int ref_partition(int n, const int* I, const int* J, const double* D, double coarseningFactor,
                  int matchingIterations, int nthreads) {
  omp_set_num_threads(nthreads);
  SparseMatrix A = make_csr(n, n, I, J, D);
  g_hierarchy = partition::partition(A, coarseningFactor, false, true, 1.0, matchingIterations, false);
  return (int)g_hierarchy.size();
}
int ref_hierarchy_rows(int level) { return g_hierarchy[level].Rows(); }
int ref_hierarchy_cols(int level) { return g_hierarchy[level].Cols(); }
void ref_hierarchy_get(int level, int* PI, int* PJ) {
  const auto& P = g_hierarchy[level];
  std::memcpy(PI, P.GetIndptr().data(), sizeof(int) * (P.Rows() + 1));
  std::memcpy(PJ, P.GetIndices().data(), sizeof(int) * P.GetIndices().size());
}

}  // extern "C"
