// graph-embed_b200 drop-in :: the result writers of /root/reference/include/export.hpp:21-23
// (src/export.cpp:16-39), same names and file formats, header-only:
//   writePartition  one aggregate id per line
//   writeCoords     one vertex per line, every coordinate followed by one blank (default ostream
//                   formatting, i.e. 6 significant digits, as the reference writes them)
// and the three text files examples/embedder.cpp:230-289 hands to scripts/plot-graph.py
// (ge_b200::writePlotInputs).  Host-only; nothing here touches the GPU.
#ifndef GE_B200_EXPORT_HPP
#define GE_B200_EXPORT_HPP

#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "matrixutils.hpp"

namespace partition {

inline void writePartition(const std::vector<int>& partition, const std::string& outputpath) {
  std::ofstream out(outputpath);
  for (const int part : partition) out << part << "\n";
}

inline void writeCoords(const std::vector<std::vector<double>>& coords, const std::string& outputpath) {
  std::ofstream out(outputpath);
  for (const auto& row : coords) {
    for (const double x : row) out << x << " ";
    out << "\n";
  }
}

}  // namespace partition

namespace ge_b200 {

// examples/embedder.cpp:230-289: `partpath` holds "n k", the k level sizes, then for every level
// the member lists of its aggregates (one aggregate per line); `coordspath` holds x y z per vertex
// (z = 0 for a 2-D layout); `matpath` holds one "i j" line per stored entry of A.
inline void writePlotInputs(const SparseMatrix& A, const std::vector<SparseMatrix>& hierarchy,
                            const std::vector<std::vector<double>>& coords, const int dimension,
                            const std::string& partpath, const std::string& coordspath,
                            const std::string& matpath) {
  if (dimension != 2 && dimension != 3) throw std::invalid_argument("writePlotInputs: dimension must be 2 or 3");
  const int n = A.Rows();
  {
    std::ofstream part(partpath);
    const int k = static_cast<int>(hierarchy.size());
    if (k == 0) {  // :240-247: no hierarchy -> one level of singletons
      part << n << " " << 1 << "\n" << n << " \n";
      for (int i = 0; i < n; ++i) part << i << " \n";
    } else {
      part << n << " " << k << "\n";
      for (const auto& P : hierarchy) part << P.Rows() << " ";
      part << "\n";
      for (const auto& P : hierarchy) {
        const std::vector<int>& I = P.GetIndptr();
        const std::vector<int>& J = P.GetIndices();
        for (int a = 0; a < P.Rows(); ++a) {
          for (int c = I[a]; c < I[a + 1]; ++c) part << J[c] << " ";
          part << "\n";
        }
      }
    }
  }
  {
    std::ofstream out(coordspath);
    for (int i = 0; i < n; ++i) {
      out << coords[i][0] << " " << coords[i][1] << " ";
      if (dimension == 3) out << coords[i][2];
      else out << 0.0;
      out << "\n";
    }
  }
  {
    std::ofstream out(matpath);
    const std::vector<int>& I = A.GetIndptr();
    const std::vector<int>& J = A.GetIndices();
    for (int i = 0; i < n; ++i)
      for (int c = I[i]; c < I[i + 1]; ++c) out << i << " " << J[c] << "\n";
  }
}

}  // namespace ge_b200

#endif
