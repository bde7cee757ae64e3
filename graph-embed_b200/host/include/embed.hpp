// graph-embed_b200 drop-in :: /root/reference/include/embed.hpp:70-78 (embed, embedMultilevel) and
// :28-65 (the pluggable-embedder interface embedVia / embedViaMultilevel), same signatures, running
// on a B200 through the C ABI.  Swap the include path and link libgraphembed_b200.so instead of
// libpartitioner.a; see INTEGRATION.md.
//
// Not provided (out of the hot path): anyToMultilevel and
// embedViaMinimization.
#ifndef GE_B200_EMBED_HPP
#define GE_B200_EMBED_HPP

#include <functional>
#include <iostream>
#include <vector>

#include "forceatlas.hpp"

namespace partition {

typedef std::function<void(const SparseMatrix&, const SparseMatrix&, const std::vector<int>&,
                           const std::vector<std::vector<double>>&, const std::vector<double>&,
                           std::vector<std::vector<double>>&, const int)>
    MultilevelEmbedder;  // include/embed.hpp:28-35

namespace detail {
inline std::vector<std::vector<double>> run_embed(const std::vector<SparseMatrix>& As,
                                                  const std::vector<SparseMatrix>& ps, const int d,
                                                  const int index, std::vector<double>* r_A,
                                                  std::vector<std::vector<double>>* coords_A) {
  const int L = static_cast<int>(ps.size()) - index;
  std::vector<ge_csr> a, p;
  for (size_t l = index; l < As.size(); ++l) a.push_back(ge_b200::view(As[l]));
  for (size_t l = index; l < ps.size(); ++l) p.push_back(ge_b200::view(ps[l]));
  ge_embed_options opt;
  ge_embed_options_default(&opt);
  opt.precision = ge_b200::options().precision;
  opt.seed = ge_b200::options().seed;
  opt.verbose = ge_b200::options().verbose;
  opt.first_layer = index + 1;  // the reference numbers its progress lines from the finest level
  const int n = As[index].Rows();
  const int m = L > 0 ? As[index + 1].Rows() : 0;
  // embedMultilevel's out-parameters (level index+1's radii and rescaled coordinates) are copied
  // back from the device only when the caller asked for them: partition::embed downloads the
  // finest coordinates alone
  std::vector<double> x(static_cast<size_t>(n) * d), rA(r_A ? m : 0), cA(coords_A ? static_cast<size_t>(m) * d : 0);
  ge_b200::check(ge_embed(ge_b200::default_context(), L, a.data(), p.data(), d, &opt, x.data(),
                          (r_A && m) ? rA.data() : nullptr, (coords_A && m) ? cA.data() : nullptr, nullptr));
  if (r_A) *r_A = rA;
  if (coords_A) *coords_A = ge_b200::unflatten(cA, m, d);
  return ge_b200::unflatten(x, n, d);
}
}  // namespace detail

// src/embed.cpp:576-796.  r_A / coords_A receive the radii and rescaled coordinates of level
// index+1 (empty at the coarsest level), as the reference's out-parameters do.
inline std::vector<std::vector<double>> embedMultilevel(const std::vector<SparseMatrix>& As,
                                                        const std::vector<SparseMatrix>& ps, const int d,
                                                        const int index, std::vector<double>& r_A,
                                                        std::vector<std::vector<double>>& coords_A) {
  return detail::run_embed(As, ps, d, index, &r_A, &coords_A);
}

// src/embed.cpp:561-574
inline std::vector<std::vector<double>> embed(const std::vector<SparseMatrix>& As,
                                              const std::vector<SparseMatrix>& ps, const int d) {
  if (As.size() != ps.size() + 1) throw std::invalid_argument("embed: As.size() != P_Ts.size() + 1");
  return detail::run_embed(As, ps, d, 0, nullptr, nullptr);
}

// src/embed.cpp:108-335.  Like the reference, only the level at `levelIndex` goes through
// `embedder`; the coarser levels are produced by embedMultilevel (src/embed.cpp:144).
inline std::vector<std::vector<double>> embedViaMultilevel(
    const std::vector<SparseMatrix>& As, const std::vector<SparseMatrix>& P_Ts, const int d,
    const int levelIndex, std::vector<double>& r_A, std::vector<std::vector<double>>& coords_A,
    MultilevelEmbedder embedder) {
  if (levelIndex == static_cast<int>(P_Ts.size())) {  // :121-138
    std::cout << "embedding layer " << levelIndex + 1 << ": getting base coords" << std::endl;
    r_A.clear();
    coords_A.clear();
    const int n = As[levelIndex].Rows();
    // One aggregate holding every vertex.  (The reference's construction at :126-132 pairs an
    // (n+1)-entry indptr with Rows() == 1, so its row 0 lists vertex 0 only and the remaining
    // coordinate rows are never sized; the drop-in builds the 1 x n aggregation evidently meant.)
    std::vector<int> members(n);
    for (int i = 0; i < n; i++) members[i] = i;
    SparseMatrix P_T(std::vector<int>{0, n}, members, std::vector<double>(n, 1.0), 1, n);
    std::vector<std::vector<double>> coords(n, std::vector<double>(d));
    std::vector<int> v_A(n, 0);
    std::vector<std::vector<double>> origin = {std::vector<double>(d, 0.0)};
    std::vector<double> one = {1.0};
    embedder(As[levelIndex], P_T, v_A, origin, one, coords, d);
    return coords;
  }
  std::vector<double> r_Ac;
  std::vector<std::vector<double>> coords_Ac;
  coords_A = embedMultilevel(As, P_Ts, d, levelIndex + 1, r_Ac, coords_Ac);  // :144
  const SparseMatrix& P_T = P_Ts[levelIndex];
  const int n = As[levelIndex].Rows();
  const int m = static_cast<int>(coords_A.size());
  std::cout << "embeding layer " << levelIndex + 1 << std::endl;
  std::vector<double> cA = ge_b200::flatten(coords_A, d);
  r_A.assign(m, 0.0);
  if (r_Ac.empty()) {  // :167-230
    ge_b200::check(ge_level_radii(m, d, cA.data(), r_A.data(), nullptr, nullptr, nullptr, nullptr));
  } else {  // :231-329
    const ge_csr Ac = ge_b200::view(As[levelIndex + 1]), Pc = ge_b200::view(P_Ts[levelIndex + 1]);
    const std::vector<double> cAc = ge_b200::flatten(coords_Ac, d);
    ge_b200::check(ge_level_radii(m, d, cA.data(), r_A.data(), &Ac, &Pc, cAc.data(), r_Ac.data()));
  }
  coords_A = ge_b200::unflatten(cA, m, d);
  std::vector<int> vertex_A(n);  // P_T.Transpose().GetIndices(), :156
  for (int a = 0; a < P_T.Rows(); ++a)
    for (int c = P_T.GetIndptr()[a]; c < P_T.GetIndptr()[a + 1]; ++c) vertex_A[P_T.GetIndices()[c]] = a;
  std::vector<std::vector<double>> coords(n, std::vector<double>(d));
  embedder(As[levelIndex], P_T, vertex_A, coords_A, r_A, coords, d);  // :332
  return coords;
}

// src/embed.cpp:85-106
inline std::vector<std::vector<double>> embedVia(const std::vector<SparseMatrix>& As,
                                                 const std::vector<SparseMatrix>& P_Ts, const int d,
                                                 MultilevelEmbedder embedder) {
  if (As.size() != P_Ts.size() + 1) throw std::invalid_argument("embedVia: As.size() != P_Ts.size() + 1");
  std::vector<double> none;
  std::vector<std::vector<double>> none2;
  return embedViaMultilevel(As, P_Ts, d, 0, none, none2, embedder);
}

// The B200 per-aggregate solver as an embedVia-compatible functor (100 iterations, as
// src/embed.cpp:793 passes).
inline MultilevelEmbedder forceAtlasMultilevelEmbedder(int iterations = 100) {
  return [iterations](const SparseMatrix& A, const SparseMatrix& P_T, const std::vector<int>& v_A,
                      const std::vector<std::vector<double>>& coords_A, const std::vector<double>& r_A,
                      std::vector<std::vector<double>>& coords, const int d) {
    forceAtlasMultilevel(A, P_T, v_A, coords_A, r_A, coords, d, iterations);
  };
}

}  // namespace partition

#endif
