// graph-embed_b200 drop-in :: type aliases of /root/reference/include/matrixutils.hpp:17-19.
// (identity / toLaplacian / fromLaplacian of that header are unused by the embed path and are not
// part of the hot path)
#ifndef GE_B200_MATRIXUTILS_HPP
#define GE_B200_MATRIXUTILS_HPP

#include "sparsematrix.hpp"  // linalgcpp's, or graph-embed_b200/host/compat/sparsematrix.hpp

using SparseMatrix = linalgcpp::SparseMatrix<double>;
using coord = std::vector<double>;
using coordinates = std::vector<coord>;

#endif
