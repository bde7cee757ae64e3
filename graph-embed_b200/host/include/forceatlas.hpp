// graph-embed_b200 drop-in :: same signatures and default arguments as
// /root/reference/include/forceatlas.hpp:89-103, :307-308, :314-331, running on a B200 through the
// C ABI (include/graph_embed_b200.h).  Unlike the reference these are `inline`, so the header may
// be included from several translation units.
#ifndef GE_B200_FORCEATLAS_HPP
#define GE_B200_FORCEATLAS_HPP

#include <cmath>
#include <random>
#include <vector>

#include "ge_b200_runtime.hpp"

namespace partition {

// include/forceatlas.hpp:66-87
inline double abs(double val) { return (val < 0) ? -val : val; }
inline double distance(const std::vector<double>& v1, const std::vector<double>& v2) {
  double sum = 0.0;
  for (size_t i = 0; i < v1.size(); i++) {
    double d = v2[i] - v1[i];
    sum += d * d;
  }
  return std::sqrt(sum);
}
inline double magnitude(const std::vector<double>& v) {
  double sum = 0.0;
  for (size_t i = 0; i < v.size(); i++) sum += v[i] * v[i];
  return std::sqrt(sum);
}

// include/forceatlas.hpp:89-305
inline void forceAtlas(const SparseMatrix& A, const int dim, std::vector<std::vector<double>>& coords,
                       const int iterations = 100000, const double ks = 0.1, const double ksmax = 1.0,
                       const double repel = 1.0, const double attract = 1.0, const double gravity = 1.0,
                       const bool useWeights = true, const bool linlog = false,
                       const bool nohubs = false, const double delta = 1.0,
                       const double tolerate = 1.0, const bool normalize = false) {
  const int n = A.Rows();
  std::vector<double> x;
  if (coords.size() == 0) {  // :118-125, same generator and draw order
    x.resize(static_cast<size_t>(n) * dim);
    const unsigned seed = ge_b200::options().seed ? ge_b200::options().seed : std::random_device()();
    ge_reference_uniform(seed, static_cast<int64_t>(x.size()), x.data());
  } else {
    x = ge_b200::flatten(coords, dim);
  }
  ge_params p;
  ge_params_default_flat(&p);
  p.iterations = iterations;
  p.ks = ks;
  p.ksmax = ksmax;
  p.repel = repel;
  p.attract = attract;
  p.gravity = gravity;
  p.use_weights = useWeights;
  p.linlog = linlog;
  p.nohubs = nohubs;
  p.delta = delta;
  p.tolerate = tolerate;
  p.normalize = normalize;
  p.precision = ge_b200::options().precision;
  const ge_csr a = ge_b200::view(A);
  ge_b200::check(ge_flat_forceatlas(ge_b200::default_context(), &a, dim, x.data(), &p));
  coords = ge_b200::unflatten(x, n, dim);
}

// include/forceatlas.hpp:307-312
inline std::vector<std::vector<double>> forceAtlas(const SparseMatrix& A, const int dim = 2) {
  std::vector<std::vector<double>> coords(0);
  forceAtlas(A, dim, coords);
  return coords;
}

// include/forceatlas.hpp:314-574 (note the reference's parameter order differs from forceAtlas)
inline void forceAtlasMultilevel(const SparseMatrix& A, const SparseMatrix& P,
                                 const std::vector<int>& v_A,
                                 const std::vector<std::vector<double>>& coords_A,
                                 const std::vector<double>& r_A,
                                 std::vector<std::vector<double>>& coords, int dim = 2,
                                 int iterations = 10, double ks = 0.1, double ksmax = 1.0,
                                 bool useWeights = true, bool linlog = false, bool nohubs = false,
                                 double repel = 1.0, double attract = 1.0, double gravity = 1.0,
                                 double delta = 1.0, double tolerate = 1.0) {
  ge_params p;
  ge_params_default_multilevel(&p);
  p.iterations = iterations;
  p.ks = ks;
  p.ksmax = ksmax;
  p.repel = repel;
  p.attract = attract;
  p.gravity = gravity;
  p.use_weights = useWeights;
  p.linlog = linlog;
  p.nohubs = nohubs;
  p.delta = delta;
  p.tolerate = tolerate;
  p.precision = ge_b200::options().precision;
  p.seed = ge_b200::options().seed;
  const ge_csr a = ge_b200::view(A), pt = ge_b200::view(P);
  const std::vector<double> cA = ge_b200::flatten(coords_A, dim);
  std::vector<double> out(static_cast<size_t>(A.Rows()) * dim);
  ge_b200::check(ge_multilevel_forceatlas(ge_b200::default_context(), &a, &pt, v_A.data(), cA.data(),
                                          r_A.data(), nullptr, out.data(), dim, &p));
  coords = ge_b200::unflatten(out, A.Rows(), dim);  // the caller pre-sizes it (src/embed.cpp:786)
}

}  // namespace partition

#endif
