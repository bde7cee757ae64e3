// graph-embed_b200 :: one process, several B200s behind the C ABI.
//
// ge_context_create_multi builds a context that owns one sub-context (device, stream, memory pool)
// per GPU and an NCCL communicator over them (ncclCommInitAll: single process, one rank per device).
// ge_flat_forceatlas on such a context runs partition::forceAtlas
// (/root/reference/include/forceatlas.hpp:89-305) as the symmetric multi-rank plan of ge_flat_sym.cu:
//   every device evaluates 1/N of the unordered pairs over the full length      (k_repulsion_sym)
//   in-place reduce-scatter of the pair sums, per dimension                      (NCCL, NVLink)
//   attraction + gravity + step on the device's row block                        (k_attract_step_staged)
//   in-place all-gather of the new positions, per dimension                      (NCCL)
// There is no all-reduce: the reference's global swing / traction sums are dead code.
// The two exchanges move n*d*w bytes each per iteration (8 MB at n = 500k, d = 2) against >= 19 ms
// of kernels per device, which is why they are plain NCCL calls and not a fused peer-memory kernel.
//
// NCCL is bound at run time (dlopen of libnccl.so.2) so that the library keeps loading on boxes
// without it and a host process that already carries an NCCL (PyTorch) shares that copy.
#include <dlfcn.h>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "ge_context.h"

namespace ge {

namespace {
// the slice of nccl.h this file needs (values are part of NCCL's stable ABI)
typedef void* nccl_comm_t;
enum { kNcclSuccess = 0, kNcclFloat32 = 7, kNcclFloat64 = 8, kNcclSum = 0 };
struct NcclApi {
  void* handle = nullptr;
  int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(nccl_comm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& why) {
    const char* names[] = {std::getenv("GE_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (n == nullptr) continue;
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) {
      why = "libnccl.so.2 not found (set GE_NCCL_LIB)";
      return false;
    }
    auto sym = [&](const char* name) { return dlsym(handle, name); };
    CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
    GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
    ReduceScatter = (decltype(ReduceScatter))sym("ncclReduceScatter");
    AllGather = (decltype(AllGather))sym("ncclAllGather");
    AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
    Broadcast = (decltype(Broadcast))sym("ncclBroadcast");
    GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !GroupStart || !GroupEnd || !ReduceScatter || !AllGather ||
        !AllReduce || !Broadcast) {
      why = "libnccl lacks a required symbol";
      return false;
    }
    return true;
  }
};
}  // namespace

struct MultiState {
  NcclApi nccl;
  std::vector<ge_context*> dev;  // dev[0] is the owning context itself
  std::vector<nccl_comm_t> comm;
  ~MultiState() {
    for (size_t r = 0; r < comm.size(); ++r)
      if (comm[r]) {
        cudaSetDevice(dev[r]->device);
        nccl.CommDestroy(comm[r]);
      }
    for (size_t r = 1; r < dev.size(); ++r) ge_context_destroy(dev[r]);
  }
};

void multi_destroy(MultiState* m) { delete m; }
int multi_size(const ge_context* ctx) { return ctx->multi ? (int)ctx->multi->dev.size() : 1; }

#define GE_NCCL(api, call)                                                                      \
  do {                                                                                          \
    const int _r = (call);                                                                      \
    if (_r != kNcclSuccess) {                                                                   \
      set_error(std::string(#call) + ": " +                                                     \
                ((api).GetErrorString ? (api).GetErrorString(_r) : "NCCL error") + " (" __FILE__ \
                ":" + std::to_string(__LINE__) + ")");                                          \
      throw Fail{GE_ERR_CUDA};                                                                  \
    }                                                                                           \
  } while (0)

void multi_attach(ge_context* ctx, int ndev, const int* devices) {
  std::unique_ptr<MultiState> m(new MultiState);
  std::string why;
  if (!m->nccl.load(why)) {
    set_error("multi-GPU context: " + why);
    throw Fail{GE_ERR_UNSUPPORTED};
  }
  m->dev.push_back(ctx);
  std::vector<int> ids(ndev);
  ids[0] = ctx->device;
  for (int r = 1; r < ndev; ++r) {
    ids[r] = devices ? devices[r] : r;
    ge_context* c = nullptr;
    const ge_status st = ge_context_create(ids[r], nullptr, &c);
    if (st != GE_OK) throw Fail{st};
    m->dev.push_back(c);
  }
  m->comm.assign(ndev, nullptr);
  GE_NCCL(m->nccl, m->nccl.CommInitAll(m->comm.data(), ndev, ids.data()));
  GE_CUDA(cudaSetDevice(ctx->device));
  ctx->multi = m.release();
}

ge_context* multi_device(ge_context* ctx, int r) { return ctx->multi ? ctx->multi->dev[r] : ctx; }

void multi_broadcast_f64(ge_context* ctx, const std::vector<double*>& bufs, size_t count, int root) {
  MultiState& M = *ctx->multi;
  if (count == 0) return;
  GE_NCCL(M.nccl, M.nccl.GroupStart());
  for (size_t r = 0; r < M.dev.size(); ++r)
    GE_NCCL(M.nccl, M.nccl.Broadcast(bufs[root], bufs[r], count, kNcclFloat64, root, M.comm[r], M.dev[r]->stream));
  GE_NCCL(M.nccl, M.nccl.GroupEnd());
}

void multi_allreduce_sum_f64(ge_context* ctx, const std::vector<double*>& bufs, size_t count) {
  MultiState& M = *ctx->multi;
  if (count == 0) return;
  GE_NCCL(M.nccl, M.nccl.GroupStart());
  for (size_t r = 0; r < M.dev.size(); ++r)
    GE_NCCL(M.nccl, M.nccl.AllReduce(bufs[r], bufs[r], count, kNcclFloat64, kNcclSum, M.comm[r], M.dev[r]->stream));
  GE_NCCL(M.nccl, M.nccl.GroupEnd());
}

void multi_sync(ge_context* ctx) {
  if (!ctx->multi) {
    GE_CUDA(cudaStreamSynchronize(ctx->stream));
    return;
  }
  for (ge_context* c : ctx->multi->dev) {
    GE_CUDA(cudaSetDevice(c->device));
    GE_CUDA(cudaStreamSynchronize(c->stream));
  }
  GE_CUDA(cudaSetDevice(ctx->device));
}

void multi_collect_counters(ge_context* ctx) {
  if (!ctx->multi) return;
  for (size_t r = 1; r < ctx->multi->dev.size(); ++r) {
    ge_context* c = ctx->multi->dev[r];
    ctx->launches += c->launches;
    ctx->h2d_bytes += c->h2d_bytes;
    ctx->d2h_bytes += c->d2h_bytes;
    ctx->grid_tier_ms = std::max(ctx->grid_tier_ms, c->grid_tier_ms);
    c->launches = 0;
    c->h2d_bytes = c->d2h_bytes = c->grid_tier_ms = 0;
  }
}

// partition::forceAtlas on every device of the context (see the header comment).
void multi_flat_solve(ge_context* ctx, const ge_csr& A, int dim, double* coords, const ge_params& p) {
  MultiState& M = *ctx->multi;
  const int N = (int)M.dev.size();
  const int n = A.rows;
  const int64_t ld = round_up(std::max(n, 1), 256);
  GE_REQUIRE(ld % N == 0, "device count does not divide the padded row count");
  const int64_t R = ld / N;
  const bool verbose = std::getenv("GE_VERBOSE") != nullptr;
  double t_mark = now_ms();
  auto lap = [&](const char* what) {
    if (!verbose) return;
    multi_sync(ctx);
    const double t = now_ms();
    std::fprintf(stderr, "[ge] multi flat n=%d N=%d %-14s %8.3f ms\n", n, N, what, t - t_mark);
    t_mark = t;
  };
  std::vector<std::unique_ptr<FlatSolver>> solver(N);
  std::vector<double> deg;  // every plan needs every row's mass: computed once, not once per device
  flat_degrees(A, p, deg);
  std::vector<ge_status> status(N, GE_OK);
  std::vector<std::string> errors(N);
  {  // plan creation (graph upload, masses, scratch) and the coordinate upload: one host thread per device
    std::vector<std::thread> pool;
    for (int r = 0; r < N; ++r)
      pool.emplace_back([&, r] {
        try {
          GE_CUDA(cudaSetDevice(M.dev[r]->device));
          const int rb = (int)std::min<int64_t>(n, r * R), re = (int)std::min<int64_t>(n, (r + 1) * R);
          solver[r].reset(make_flat_solver(M.dev[r], A, dim, p, rb, re, r, N, deg.data()));
          GE_REQUIRE(solver[r]->symmetric(), "symmetric plan refused");
          solver[r]->upload_coords(coords);
          GE_CUDA(cudaStreamSynchronize(M.dev[r]->stream));
        } catch (const Fail& f) {
          status[r] = f.st;
          errors[r] = ge_last_error();
        }
      });
    for (auto& t : pool) t.join();
    for (int r = 0; r < N; ++r)
      if (status[r] != GE_OK) {
        set_error("device " + std::to_string(M.dev[r]->device) + ": " + errors[r]);
        throw Fail{status[r]};
      }
  }
  lap("plans + upload");
  const int dt = solver[0]->elem_size() == 8 ? kNcclFloat64 : kNcclFloat32;
  const size_t w = (size_t)solver[0]->elem_size();
  for (int it = 0; it < p.iterations; ++it) {
    for (int r = 0; r < N; ++r) {
      GE_CUDA(cudaSetDevice(M.dev[r]->device));
      solver[r]->launch_repulsion();
    }
    GE_NCCL(M.nccl, M.nccl.GroupStart());
    for (int r = 0; r < N; ++r)
      for (int k = 0; k < dim; ++k) {
        char* s = (char*)solver[r]->pair_sums() + (size_t)k * ld * w;
        GE_NCCL(M.nccl, M.nccl.ReduceScatter(s, s + (size_t)r * R * w, (size_t)R, dt, kNcclSum,
                                             M.comm[r], M.dev[r]->stream));
      }
    GE_NCCL(M.nccl, M.nccl.GroupEnd());
    for (int r = 0; r < N; ++r) {
      GE_CUDA(cudaSetDevice(M.dev[r]->device));
      solver[r]->launch_step(true);
    }
    GE_NCCL(M.nccl, M.nccl.GroupStart());
    for (int r = 0; r < N; ++r)
      for (int k = 0; k < dim; ++k) {
        char* x = (char*)solver[r]->next_coords() + (size_t)k * ld * w;
        GE_NCCL(M.nccl, M.nccl.AllGather(x + (size_t)r * R * w, x, (size_t)R, dt, M.comm[r],
                                         M.dev[r]->stream));
      }
    GE_NCCL(M.nccl, M.nccl.GroupEnd());
    for (int r = 0; r < N; ++r) solver[r]->swap();
  }
  lap("iterations");
  for (int r = 1; r < N; ++r) {
    GE_CUDA(cudaSetDevice(M.dev[r]->device));
    GE_CUDA(cudaStreamSynchronize(M.dev[r]->stream));
    ctx->launches += M.dev[r]->launches;  // the owning context reports the launches of all devices
    M.dev[r]->launches = 0;
    ctx->h2d_bytes += M.dev[r]->h2d_bytes;
    M.dev[r]->h2d_bytes = 0;
  }
  GE_CUDA(cudaSetDevice(ctx->device));
  if (p.normalize) solver[0]->normalize();  // every device holds all positions after the all-gather
  solver[0]->download_coords(coords);
  lap("download");
  for (int r = N - 1; r >= 0; --r) {  // free each plan's memory on its own device
    GE_CUDA(cudaSetDevice(M.dev[r]->device));
    solver[r].reset();
  }
  GE_CUDA(cudaSetDevice(ctx->device));
  lap("release");
}

}  // namespace ge
