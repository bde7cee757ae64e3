// graph-embed_b200 drop-in :: glue between the reference's C++ types and the C ABI.
#ifndef GE_B200_RUNTIME_HPP
#define GE_B200_RUNTIME_HPP

#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "graph_embed_b200.h"
#include "matrixutils.hpp"

namespace ge_b200 {

// Knobs the reference does not have (precision, seed, GPU count); change before the first call
// into the library (the GPU count is read once, when the process-wide context is created).
struct Options {
  int precision = GE_F64;
  unsigned seed = 0;   // 0: std::random_device like the reference
  bool verbose = true; // print the reference's "embedding layer N" progress lines
  int gpus = 0;        // 0: the environment variable GE_GPUS, else 1.  > 1: partition::forceAtlas on
                       // graphs of >= 32768 vertices is sharded over that many GPUs of the box
};
inline Options& options() {
  static Options o;
  return o;
}

// Process-wide context (the reference keeps no state either; the context only owns streams, the
// device memory pools and, with several GPUs, the NCCL communicator).
inline ge_context* default_context() {
  struct Holder {
    ge_context* ctx = nullptr;
    Holder() {
      int gpus = options().gpus;
      if (gpus <= 0) {
        const char* e = std::getenv("GE_GPUS");
        gpus = e ? std::atoi(e) : 1;
      }
      const ge_status st = gpus > 1 ? ge_context_create_multi(gpus, nullptr, &ctx)
                                    : ge_context_create(-1, nullptr, &ctx);
      if (st != GE_OK) throw std::runtime_error(std::string("graph-embed_b200: ") + ge_last_error());
    }
    ~Holder() { ge_context_destroy(ctx); }
  };
  static Holder h;
  return h.ctx;
}

// The reference reports problems through assert() only (src/embed.cpp:564-570); the drop-in
// throws instead of continuing with garbage, because there is no CPU path to fall back to.
inline void check(ge_status st) {
  if (st != GE_OK) throw std::runtime_error(std::string("graph-embed_b200: ") + ge_last_error());
}

inline ge_csr view(const SparseMatrix& A) {
  ge_csr c;
  c.rows = A.Rows();
  c.cols = A.Cols();
  c.nnz = static_cast<int64_t>(A.GetIndices().size());
  c.indptr = A.GetIndptr().data();
  c.indices = A.GetIndices().data();
  c.data = A.GetData().data();
  return c;
}

inline std::vector<double> flatten(const std::vector<std::vector<double>>& c, int d) {
  std::vector<double> x(c.size() * static_cast<size_t>(d));
  for (size_t i = 0; i < c.size(); ++i)
    for (int k = 0; k < d; ++k) x[i * d + k] = c[i][k];
  return x;
}

inline std::vector<std::vector<double>> unflatten(const std::vector<double>& x, int n, int d) {
  std::vector<std::vector<double>> c(n, std::vector<double>(d));
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < d; ++k) c[i][k] = x[static_cast<size_t>(i) * d + k];
  return c;
}

// A_{l+1} = P_T * A_l * P_T^T on the device: the line every caller of partition::embed writes as
//   As.push_back(P.Mult(As.back()).Mult(P.Transpose()));   (examples/embedder.cpp:213-216)
// becomes  As.push_back(ge_b200::galerkin(As.back(), P));
inline SparseMatrix galerkin(const SparseMatrix& A, const SparseMatrix& P_T) {
  const ge_csr a = view(A), p = view(P_T);
  std::vector<int> indptr(static_cast<size_t>(P_T.Rows()) + 1), indices(A.GetIndices().size());
  std::vector<double> data(A.GetIndices().size());
  int64_t nnz = 0;
  check(ge_galerkin(default_context(), &a, &p, indptr.data(), indices.data(), data.data(),
                    static_cast<int64_t>(indices.size()), &nnz, nullptr));
  indices.resize(static_cast<size_t>(nnz));
  data.resize(static_cast<size_t>(nnz));
  return SparseMatrix(std::move(indptr), std::move(indices), std::move(data), P_T.Rows(), P_T.Rows());
}

}  // namespace ge_b200

#endif
