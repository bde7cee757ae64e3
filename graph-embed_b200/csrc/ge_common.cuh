// graph-embed_b200 :: shared device helpers (sm_100a).
#ifndef GE_COMMON_CUH
#define GE_COMMON_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/graph_embed_b200.h"

namespace ge {

// ---------------------------------------------------------------------------------------------
// error plumbing (host)
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
struct Fail {
  ge_status st;
};
#define GE_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (call);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::ge::set_error(std::string(#call) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                      std::to_string(__LINE__) + ")");                                         \
      throw ::ge::Fail{_e == cudaErrorMemoryAllocation ? GE_ERR_OOM : GE_ERR_CUDA};            \
    }                                                                                          \
  } while (0)
#define GE_REQUIRE(cond, msg)                      \
  do {                                             \
    if (!(cond)) {                                 \
      ::ge::set_error(std::string("invalid: ") + msg); \
      throw ::ge::Fail{GE_ERR_INVALID};            \
    }                                              \
  } while (0)

constexpr double kEpsilon = 0.00001;  // include/forceatlas.hpp:110, :337

// ---------------------------------------------------------------------------------------------
// per-type math.  The reference computes  direction * Fr = -(xj - xi)/dis * ci*cj*repel/dis^2
// with dis = max(|xj - xi|, eps)  (include/forceatlas.hpp:154-165).  On the device this is
// (xi - xj) * [ci*repel] * cj * max(r2, eps^2)^(-3/2): one reciprocal square root per ordered
// pair, no division, no sqrt.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct Real;

template <>
struct Real<double> {
  static constexpr int kMassArrays = 3;  // c, 1.5 c, 1.875 c
  // MUFU.RSQ64H seed (low word zero, ~2^-20 relative accuracy for normal inputs).
  __device__ __forceinline__ static double rsqrt_seed(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
  }
  // max(r2, eps2) for r2 >= 0: ordering of non-negative doubles == ordering of their bit
  // patterns as integers, which keeps the compare off the FP64 pipe.
  __device__ __forceinline__ static double clamp_lo(double r2, double lo) {
    return (__double_as_longlong(r2) < __double_as_longlong(lo)) ? lo : r2;
  }
  // The same clamp for N values at once.  Pairs closer than eps are rare, so the common path is a
  // warp-uniform test on the smallest high word (N-1 integer min + 1 compare + 1 branch) and the
  // exact 64-bit clamp runs only when some lane might need it: 1.5 instead of 4 issue slots per
  // pair in the FP64 repulsion loop, which is issue-bound.
  template <int N>
  __device__ __forceinline__ static void clamp_lo_n(double (&r2)[N], double lo) {
    int hmin = __double2hiint(r2[0]);
#pragma unroll
    for (int t = 1; t < N; ++t) hmin = min(hmin, __double2hiint(r2[t]));
    if (__any_sync(0xffffffffu, hmin <= __double2hiint(lo))) {
#pragma unroll
      for (int t = 0; t < N; ++t) r2[t] = clamp_lo(r2[t], lo);
    }
  }
  // s = cj * r2^(-3/2), with q = seed, e = 1 - r2 q^2:
  //   r2^(-3/2) = q^3 (1-e)^(-3/2) = q^3 (1 + 3/2 e + 15/8 e^2 + O(e^3)),   |e| <~ 4e-6
  // so the truncation error is ~2.2 e^3 < 2e-16 relative.  6 FP64-pipe instructions per pair.
  // N independent pairs against one column, stage by stage (instruction-level parallelism across
  // the pairs hides the FP64 pipe latency).
  template <int N>
  __device__ __forceinline__ static void inv_cube_mass_n(const double (&r2)[N], double c, double c15,
                                                         double c1875, double (&s)[N]) {
    double q[N], q2[N], e[N], w[N];
#pragma unroll
    for (int t = 0; t < N; ++t) q[t] = rsqrt_seed(r2[t]);
#pragma unroll
    for (int t = 0; t < N; ++t) q2[t] = q[t] * q[t];
#pragma unroll
    for (int t = 0; t < N; ++t) e[t] = fma(-r2[t], q2[t], 1.0);
#pragma unroll
    for (int t = 0; t < N; ++t) q2[t] = q2[t] * q[t];
#pragma unroll
    for (int t = 0; t < N; ++t) w[t] = fma(e[t], c1875, c15);
#pragma unroll
    for (int t = 0; t < N; ++t) w[t] = fma(e[t], w[t], c);
#pragma unroll
    for (int t = 0; t < N; ++t) s[t] = q2[t] * w[t];
  }
  // s = r2^(-3/2) without a mass factor (the symmetric sweep applies c_j on the row side and c_i
  // on the column side): same series with the constants 1, 3/2, 15/8.
  template <int N>
  __device__ __forceinline__ static void inv_cube_n(const double (&r2)[N], double (&s)[N]) {
    double q[N], q2[N], e[N], w[N];
#pragma unroll
    for (int t = 0; t < N; ++t) q[t] = rsqrt_seed(r2[t]);
#pragma unroll
    for (int t = 0; t < N; ++t) q2[t] = q[t] * q[t];
#pragma unroll
    for (int t = 0; t < N; ++t) e[t] = fma(-r2[t], q2[t], 1.0);
#pragma unroll
    for (int t = 0; t < N; ++t) q2[t] = q2[t] * q[t];
#pragma unroll
    for (int t = 0; t < N; ++t) w[t] = fma(e[t], 1.875, 1.5);
#pragma unroll
    for (int t = 0; t < N; ++t) w[t] = fma(e[t], w[t], 1.0);
#pragma unroll
    for (int t = 0; t < N; ++t) s[t] = q2[t] * w[t];
  }
  // N independent pairs against N different columns.
  template <int N>
  __device__ __forceinline__ static void inv_cube_mass_v(const double (&r2)[N], const double (&c)[N],
                                                         const double (&c15)[N],
                                                         const double (&c1875)[N], double (&s)[N]) {
    double q[N], q2[N], e[N], w[N];
#pragma unroll
    for (int t = 0; t < N; ++t) q[t] = rsqrt_seed(r2[t]);
#pragma unroll
    for (int t = 0; t < N; ++t) q2[t] = q[t] * q[t];
#pragma unroll
    for (int t = 0; t < N; ++t) e[t] = fma(-r2[t], q2[t], 1.0);
#pragma unroll
    for (int t = 0; t < N; ++t) q2[t] = q2[t] * q[t];
#pragma unroll
    for (int t = 0; t < N; ++t) w[t] = fma(e[t], c1875[t], c15[t]);
#pragma unroll
    for (int t = 0; t < N; ++t) w[t] = fma(e[t], w[t], c[t]);
#pragma unroll
    for (int t = 0; t < N; ++t) s[t] = q2[t] * w[t];
  }
  // 1/sqrt(x) to ~1e-16: q (1-e)^(-1/2) = q (1 + e/2 + 3/8 e^2 + O(e^3)).  x = 0 -> NaN (callers
  // that can see zero guard it); 5 FP64-pipe instructions instead of CUDA's sqrt/div sequences.
  __device__ __forceinline__ static double rsqrt_acc(double x) {
    const double q = rsqrt_seed(x);
    const double e = fma(-x, q * q, 1.0);
    return fma(q * e, fma(e, 0.375, 0.5), q);
  }
  // 1/x to ~1e-16 from the MUFU.RCP64H seed: q (1 + e + e^2), e = 1 - x q.
  __device__ __forceinline__ static double rcp_acc(double x) {
    double q;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(x));
    const double e = fma(-x, q, 1.0);
    return fma(fma(e, e, e), q, q);
  }
  __device__ __forceinline__ static double sqrt_(double x) { return sqrt(x); }
  __device__ __forceinline__ static double log1p_(double x) { return log(1.0 + x); }
  __device__ __forceinline__ static double pow_(double x, double y) { return pow(x, y); }
};

template <>
struct Real<float> {
  static constexpr int kMassArrays = 1;  // c
  __device__ __forceinline__ static float rsqrt_seed(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
  }
  __device__ __forceinline__ static float clamp_lo(float r2, float lo) { return fmaxf(r2, lo); }
  template <int N>
  __device__ __forceinline__ static void clamp_lo_n(float (&r2)[N], float lo) {
#pragma unroll
    for (int t = 0; t < N; ++t) r2[t] = fmaxf(r2[t], lo);
  }
  template <int N>
  __device__ __forceinline__ static void inv_cube_mass_n(const float (&r2)[N], float c, float, float,
                                                         float (&s)[N]) {
    float q[N];
#pragma unroll
    for (int t = 0; t < N; ++t) q[t] = rsqrt_seed(r2[t]);
#pragma unroll
    for (int t = 0; t < N; ++t) s[t] = (q[t] * q[t]) * (q[t] * c);
  }
  template <int N>
  __device__ __forceinline__ static void inv_cube_n(const float (&r2)[N], float (&s)[N]) {
    float q[N];
#pragma unroll
    for (int t = 0; t < N; ++t) q[t] = rsqrt_seed(r2[t]);
#pragma unroll
    for (int t = 0; t < N; ++t) s[t] = (q[t] * q[t]) * q[t];
  }
  template <int N>
  __device__ __forceinline__ static void inv_cube_mass_v(const float (&r2)[N], const float (&c)[N],
                                                         const float (&)[N], const float (&)[N],
                                                         float (&s)[N]) {
#pragma unroll
    for (int t = 0; t < N; ++t) {
      const float q = rsqrt_seed(r2[t]);
      s[t] = (q * q) * (q * c[t]);
    }
  }
  __device__ __forceinline__ static float rsqrt_acc(float x) {
    const float q = rsqrt_seed(x);
    return q * fmaf(fmaf(-x, q * q, 1.0f), 0.5f, 1.0f);
  }
  __device__ __forceinline__ static float rcp_acc(float x) {
    float q;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(x));
    return fmaf(fmaf(-x, q, 1.0f), q, q);
  }
  __device__ __forceinline__ static float sqrt_(float x) { return sqrtf(x); }
  __device__ __forceinline__ static float log1p_(float x) { return logf(1.0f + x); }
  __device__ __forceinline__ static float pow_(float x, float y) { return powf(x, y); }
};

// Physical constants of one solve, already converted to the compute type.
template <typename T>
struct Physics {
  T ks, ksmax, repel, attract, gravity, delta, gspeed /* tolerate * 1.0 / 1.0, :244 */;
  T eps, eps2, inv_eps;
  int use_weights, linlog, nohubs;
  int general_attraction;  // linlog || nohubs || delta != 1
};

template <typename T>
inline Physics<T> make_physics(const ge_params& p) {
  Physics<T> ph;
  ph.ks = (T)p.ks;
  ph.ksmax = (T)p.ksmax;
  ph.repel = (T)p.repel;
  ph.attract = (T)p.attract;
  ph.gravity = (T)p.gravity;
  ph.delta = (T)p.delta;
  ph.gspeed = (T)(p.tolerate * 1.0 / 1.0);
  ph.eps = (T)kEpsilon;
  ph.eps2 = (T)(kEpsilon * kEpsilon);
  ph.inv_eps = (T)(1.0 / kEpsilon);
  ph.use_weights = p.use_weights;
  ph.linlog = p.linlog;
  ph.nohubs = p.nohubs;
  ph.general_attraction = (p.linlog || p.nohubs || p.delta != 1.0) ? 1 : 0;
  return ph;
}

// Attraction term factor g such that  force += (xj - xi) * g   (include/forceatlas.hpp:171-202).
// With the defaults (linlog=false, delta=1, nohubs=false) the clamped distance cancels:
//   direction * Fa = (xj-xi)/dis * attract * dis * a  =  (xj-xi) * attract * a.
// GA = false compiles only the default path (no log / pow code in the hot kernels' loop bodies).
template <typename T, bool GA = true>
__device__ __forceinline__ T attraction_factor(T r2, T a, T deg_ip1, const Physics<T>& ph) {
  if (!GA || !ph.general_attraction) return ph.attract * a;
  T dis = Real<T>::sqrt_(r2);
  if (dis < ph.eps) dis = ph.eps;
  T fa = dis;
  if (ph.linlog) fa = Real<T>::log1p_(fa);
  if (ph.delta == (T)1) {
    fa = fa * a;
  } else if (ph.delta != (T)0) {
    fa = (a < (T)0 ? (T)-1 : (T)1) * Real<T>::pow_(a < (T)0 ? -a : a, ph.delta) * fa;
  }
  if (ph.nohubs) fa = fa / deg_ip1;
  return ph.attract * fa / dis;
}

// Per-vertex epilogue of one iteration: gravity (:205-211 / :411-414,469-474), swing (:214-217 /
// :477-487), speed and cap (:248-255), displacement (:257-260).  `f` enters holding repulsion +
// attraction and leaves as the total force; E holds the multilevel external-pull numerators
// sum(100 * direction) which the reference divides by the clamped |x_i| (:453-465).
// Square roots and divisions are rewritten on reciprocal square roots (MUFU seed + series):
//   mag = sqrt(m2)            ->  1/mag   = rsqrt(m2)
//   swing = sqrt(sw2)         ->  sw2 * rsqrt(sw2);   sqrt(swing) likewise
//   ksmax / sqrt(f2)          ->  ksmax * rsqrt(f2)   (f2 = 0 -> NaN -> "no cap", as with +inf)
// Flat kernel (ML=false): |x| and swing unclamped, x = 0 gives NaN like the reference (Q3).
template <typename T, int D, bool ML>
__device__ __forceinline__ void vertex_step(T (&x)[D], T (&f)[D], T (&fprev)[D], const T (&E)[D],
                                            T deg_ip1, const Physics<T>& ph) {
  T m2 = (T)0;
#pragma unroll
  for (int k = 0; k < D; ++k) m2 = fma(x[k], x[k], m2);
  T inv_mag;
  if (ML) {
    inv_mag = (m2 < ph.eps2) ? ph.inv_eps : Real<T>::rsqrt_acc(m2);
  } else {
    inv_mag = Real<T>::rsqrt_acc(m2);
  }
  const T g = ph.gravity * deg_ip1 * inv_mag;
  T sw2 = (T)0, f2 = (T)0;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    T fk = f[k];
    if (ML) fk = fma(E[k], inv_mag, fk);
    fk = fma(-x[k], g, fk);
    const T dk = fk - fprev[k];
    sw2 = fma(dk, dk, sw2);
    f2 = fma(fk, fk, f2);
    f[k] = fk;
  }
  T swing;
  if (ML) {
    const T c = sw2 < ph.eps2 ? ph.eps2 : sw2;
    swing = c * Real<T>::rsqrt_acc(c);
  } else {
    swing = sw2 > (T)0 ? sw2 * Real<T>::rsqrt_acc(sw2) : sw2;
  }
  const T ssw = swing > (T)0 ? swing * Real<T>::rsqrt_acc(swing) : swing;
  T speed = ph.ks * ph.gspeed * Real<T>::rcp_acc((T)1 + ph.gspeed * ssw);
  const T cap = ph.ksmax * Real<T>::rsqrt_acc(f2);
  if (speed > cap) speed = cap;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    x[k] = fma(f[k], speed, x[k]);
    fprev[k] = f[k];
  }
}

__host__ __device__ inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

}  // namespace ge

#endif  // GE_COMMON_CUH
