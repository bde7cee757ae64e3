// graph-embed_b200 :: mbarrier + 1-D TMA bulk-copy primitives and 16-byte shared-memory loads
// shared by the tiled all-pairs kernels (ge_flat.cu, ge_flat_sym.cu), sm_100a.
#ifndef GE_TMA_CUH
#define GE_TMA_CUH

#include <stdint.h>

namespace ge {

// ---- mbarrier / TMA bulk-copy primitives ------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}
// Waits for the phase with the given parity.  try_wait suspends the thread in hardware for a
// bounded time per call; the poll count is bounded too, so a bulk copy that never lands (a bug)
// traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_addr(bar);
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();
}
// 1-D TMA: global -> shared, completion counted in bytes on the mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_addr(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

template <typename T, int VEC>
struct VecLoad;
template <>
struct VecLoad<double, 2> {
  __device__ __forceinline__ static void ld(const double* p, double (&v)[2]) {
    const double2 t = *reinterpret_cast<const double2*>(p);
    v[0] = t.x;
    v[1] = t.y;
  }
};
template <>
struct VecLoad<float, 4> {
  __device__ __forceinline__ static void ld(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x;
    v[1] = t.y;
    v[2] = t.z;
    v[3] = t.w;
  }
};

}  // namespace ge

#endif  // GE_TMA_CUH
