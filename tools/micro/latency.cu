// Dependent-chain latencies on the SM (cycles per op, one warp unless stated): the numbers the
// on-chip coarsest-level solve (csrc/ge_onchip.cu) is bounded by.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o latency tools/micro/latency.cu && ./latency
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ double rsq_seed(double x) {
  double y;
  asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}

template <int OP>
__global__ void k_chain(double* out, long long* cyc, int iters, double a, double b) {
  __shared__ double sm[256];
  sm[threadIdx.x] = a + threadIdx.x;
  __syncthreads();
  double x = a + threadIdx.x * 1e-3, y = b;
  int idx = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (OP == 0) x = fma(x, y, b);                                   // DFMA
      if (OP == 1) x = rsq_seed(x) + 1.5;                               // MUFU.RSQ64H + DADD
      if (OP == 2) x = __shfl_xor_sync(0xffffffffu, x, 1) + 1.0;        // SHFL(64-bit = 2) + DADD
      if (OP == 3) { idx = (int)sm[idx & 255] & 255; }                  // LDS.64 -> F2I -> LDS
      if (OP == 4) x = x + y;                                           // DADD
      if (OP == 5) x = x * y;                                           // DMUL
      if (OP == 6) { float f = (float)x; f = fmaf(f, 1.0001f, 0.5f); x = (double)f; }  // cvt + FFMA + cvt
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x + idx;
}

__global__ void __cluster_dims__(8, 1, 1) k_cluster_bar(long long* cyc, int iters) {
  cg::cluster_group cl = cg::this_cluster();
  cl.sync();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) cl.sync();
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void __cluster_dims__(8, 1, 1) k_cluster_store_bar(long long* cyc, int iters) {
  __shared__ double buf[2][256];
  cg::cluster_group cl = cg::this_cluster();
  const int rank = cl.block_rank();
  cl.sync();
  long long t0 = clock64();
  double x = threadIdx.x;
  for (int i = 0; i < iters; ++i) {
    if (threadIdx.x < 8) {
      double* dst = cl.map_shared_rank(&buf[i & 1][0], threadIdx.x);
      dst[rank] = x;
    }
    cl.sync();
    x += buf[i & 1][(rank + 1) & 7];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (x == -1.0) cyc[1] = 0;
}

// The same exchange without a cluster barrier: every CTA sends its value to each peer with
// st.async (the store itself completes bytes on the receiver's mbarrier), the receiver waits on its
// own mbarrier for the 8 x 8 bytes of the iteration (two phases alternate with the two buffers).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __cluster_dims__(8, 1, 1) k_cluster_st_async(long long* cyc, int iters) {
  __shared__ double buf[2][256];
  __shared__ __align__(8) unsigned long long bar[2];
  cg::cluster_group cl = cg::this_cluster();
  const int rank = cl.block_rank();
  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[b])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cl.sync();
  long long t0 = clock64();
  double x = threadIdx.x;
  unsigned phase[2] = {0u, 0u};
  for (int i = 0; i < iters; ++i) {
    const int b = i & 1;
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 64;" ::"r"(smem_u32(&bar[b])) : "memory");
    if (threadIdx.x < 8) {
      unsigned dst, rbar;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(smem_u32(&buf[b][rank])), "r"(threadIdx.x));
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&bar[b])), "r"(threadIdx.x));
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                   ::"r"(dst), "l"(__double_as_longlong(x)), "r"(rbar) : "memory");
    }
    unsigned done = 0;
    while (!done)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                   : "=r"(done) : "r"(smem_u32(&bar[b])), "r"(phase[b]) : "memory");
    phase[b] ^= 1u;
    x += buf[b][(rank + 1) & 7];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  if (x == -1.0) cyc[1] = 0;
  cl.sync();
}

__global__ void k_syncthreads(long long* cyc, int iters) {
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, 4096 * sizeof(double));
  cudaMallocManaged(&cyc, 4 * sizeof(long long));
  const int iters = 2000;
  const char* names[] = {"DFMA dependent", "MUFU.RSQ64H + DADD", "SHFL.64 + DADD", "LDS.64 -> cvt -> LDS",
                         "DADD dependent", "DMUL dependent", "F2F + FFMA + F2F"};
  for (int nw = 1; nw <= 8; nw *= 2) {
    for (int op = 0; op < 7; ++op) {
      switch (op) {
        case 0: k_chain<0><<<1, 32 * nw>>>(out, cyc, iters, 1.0, 0.999); break;
        case 1: k_chain<1><<<1, 32 * nw>>>(out, cyc, iters, 1.0, 0.999); break;
        case 2: k_chain<2><<<1, 32 * nw>>>(out, cyc, iters, 1.0, 0.999); break;
        case 3: k_chain<3><<<1, 32 * nw>>>(out, cyc, iters, 1.0, 0.999); break;
        case 4: k_chain<4><<<1, 32 * nw>>>(out, cyc, iters, 1.0, 0.999); break;
        case 5: k_chain<5><<<1, 32 * nw>>>(out, cyc, iters, 1.0, 0.999); break;
        case 6: k_chain<6><<<1, 32 * nw>>>(out, cyc, iters, 1.0, 0.999); break;
      }
      cudaDeviceSynchronize();
      printf("warps=%d  %-24s %7.2f cycles/op\n", nw, names[op], double(cyc[0]) / (iters * 16.0));
    }
  }
  for (int threads = 32; threads <= 512; threads *= 4) {
    k_syncthreads<<<1, threads>>>(cyc, 10000);
    cudaDeviceSynchronize();
    printf("__syncthreads, %3d threads: %7.1f cycles\n", threads, double(cyc[0]) / 10000);
    k_cluster_bar<<<8, threads>>>(cyc, 10000);
    cudaDeviceSynchronize();
    printf("cluster.sync (8 CTAs), %3d threads: %7.1f cycles\n", threads, double(cyc[0]) / 10000);
    k_cluster_store_bar<<<8, threads>>>(cyc, 10000);
    cudaDeviceSynchronize();
    printf("8 DSMEM stores + cluster.sync + LDS, %3d threads: %7.1f cycles\n", threads, double(cyc[0]) / 10000);
    k_cluster_st_async<<<8, threads>>>(cyc, 10000);
    cudaDeviceSynchronize();
    printf("8 st.async stores + mbarrier wait + LDS, %3d threads: %7.1f cycles (%s)\n", threads,
           double(cyc[0]) / 10000, cudaGetErrorString(cudaGetLastError()));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
