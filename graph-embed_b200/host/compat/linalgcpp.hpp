// graph-embed_b200 :: umbrella header matching `#include "linalgcpp.hpp"`
// (/root/reference/include/partitioner.hpp:16).  See sparsematrix.hpp.
#ifndef GE_B200_COMPAT_LINALGCPP_HPP
#define GE_B200_COMPAT_LINALGCPP_HPP
#include "sparsematrix.hpp"
#endif
