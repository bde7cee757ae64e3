// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// Force-included (-include) ahead of the reference's UNMODIFIED sources when they
// are compiled into oracle/_ref/.  The reference seeds every std::mt19937 from
// std::random_device (/root/reference/include/forceatlas.hpp:104-105, 332-333;
// /root/reference/src/embed.cpp:351-352, 780-781), which makes it
// non-reproducible.  Rather than patching the sources, the token
// `random_device` is redirected to a stand-in whose operator() returns a seed
// the test driver sets: every generator the reference constructs then starts
// from mt19937(seed).  With one OpenMP thread the multilevel draws are
// aggregate-major, member-major, k-minor (forceatlas.hpp:341, 356-358).
#ifndef GE_ORACLE_REF_PRELUDE_HPP
#define GE_ORACLE_REF_PRELUDE_HPP
#include <random>
extern "C" unsigned ge_ref_current_seed(void);
namespace std {
struct ge_seeded_random_device {
  typedef unsigned int result_type;
  result_type operator()() { return ge_ref_current_seed(); }
};
}  // namespace std
#define random_device ge_seeded_random_device
#endif
