"""ncu driver: the on-chip coarsest-level solve (K3) and the attraction+step kernel alone.
usage: python tools/profile_small.py k3 [n] [iters] | attr [n] [f64|f32]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

mode = sys.argv[1]
ctx = capi.Context(0)
if mode == "k3":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 91
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
    A = graphs.rgg(3000, 10.0, seed=1)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=n)
    Ac = As[-1]
    print("coarse n", Ac.shape[0], "nnz", Ac.nnz)
    x0 = capi.reference_uniform(1, Ac.shape[0] * 2).reshape(-1, 2)
    import time
    for rep in range(2):
        t = time.time()
        ctx.flat_forceatlas(Ac, 2, x0, capi.flat_params(iterations=iters))
        print("K3 %d iters: %.3f ms -> %.3f us/iter" % (iters, 1e3 * (time.time() - t), 1e6 * (time.time() - t) / iters))
elif mode == "k3dense":
    # the (nearly) complete coarsest graph of a power-law hierarchy: dense variant of the cluster kernel
    import time
    import scipy.sparse as sp
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    rng = np.random.default_rng(3)
    M = rng.random((54, 54))
    Ac = graphs.canonical(sp.csr_matrix(M + M.T))
    x0 = capi.reference_uniform(1, 54 * 3).reshape(-1, 3)
    for rep in range(2):
        t = time.time()
        ctx.flat_forceatlas(Ac, 3, x0, capi.flat_params(iterations=iters))
        print("K3 dense n=54 d=3 %d iters: %.3f us/iter" % (iters, 1e6 * (time.time() - t) / iters))
elif mode == "k3sweep":
    import time
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    prec = capi.GE_F32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else capi.GE_F64
    for target in (34, 54, 64, 97, 200):
        A = graphs.rgg(40 * target, 10.0, seed=1)
        As, Ps = graphs.coarsen(A, 0.25, min_coarse=target)
        Ac = As[-1]
        n = Ac.shape[0]
        x0 = capi.reference_uniform(1, n * dim).reshape(-1, dim)
        shapes = ((1, 4), (1, 8), (2, 8), (4, 8), (8, 8), (4, 4), (8, 4))
        if os.environ.get("K3_WIDE"):
            shapes = ((8, 8), (8, 16), (4, 16))
        for cs, L in shapes:
            if (n + cs - 1) // cs * L > (1024 if cs == 1 else 512):
                continue
            os.environ["GE_ONCHIP_LANES"] = str(L)
            os.environ["GE_CLUSTER"] = str(cs)
            ts = []
            for iters in (1, 20001):
                ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters, precision=prec))
                t = time.time()
                ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters, precision=prec))
                ts.append(time.time() - t)
            print("n=%d nnz=%d cluster=%d L=%d: %.3f us/iter (overhead %.2f ms)" % (n, Ac.nnz, cs, L, 1e6 * (ts[1] - ts[0]) / 20000, 1e3 * ts[0]), flush=True)
elif mode == "k3v2":
    # us/iteration of the coarsest-level solve: first- vs second-generation cluster kernel over
    # cluster size x lanes x columns per trip, at n ~ 34, 64, 100, 157 and on the dense 54-vertex
    # coarsest graph of a power-law hierarchy
    import time
    import scipy.sparse as sp
    dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cases = []
    for target in (34, 64, 100, 157):
        A = graphs.rgg(40 * target, 10.0, seed=1)
        As, Ps = graphs.coarsen(A, 0.25, min_coarse=target)
        cases.append(("rgg-coarse", As[-1]))
    rng = np.random.default_rng(3)
    M = rng.random((54, 54))
    cases.append(("dense54", graphs.canonical(sp.csr_matrix(M + M.T))))

    def measure(Ac, env):
        for k in ("GE_K3_V1", "GE_CLUSTER", "GE_ONCHIP_LANES", "GE_K3_U", "GE_K3_DENSE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        n = Ac.shape[0]
        x0 = capi.reference_uniform(1, n * dim).reshape(-1, dim)
        ts = []
        for iters in (1, 20001):
            ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
            t = time.time()
            ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
            ts.append(time.time() - t)
        return 1e6 * (ts[1] - ts[0]) / 20000

    for name, Ac in cases:
        n = Ac.shape[0]
        print("%s n=%d nnz=%d d=%d" % (name, n, Ac.nnz, dim), flush=True)
        print("   default (auto)            : %.3f us/iter" % measure(Ac, {}), flush=True)
        print("   v1 default                : %.3f us/iter" % measure(Ac, {"GE_K3_V1": "1"}), flush=True)
        for cs in (4, 8, 16):
            for L in (8, 16):
                if (n + cs - 1) // cs * L > 256:
                    continue
                for U in (4, 8):
                    for dense in ((0, 1) if Ac.nnz > 8 * n else (0,)):
                        t = measure(Ac, {"GE_CLUSTER": str(cs), "GE_ONCHIP_LANES": str(L), "GE_K3_U": str(U),
                                         "GE_K3_DENSE": str(dense)})
                        print("   v2 cluster=%2d L=%2d U=%d dense=%d: %.3f us/iter" % (cs, L, U, dense, t), flush=True)
elif mode == "k3parts":
    import time
    for target, dim in ((42, 3), (34, 2), (91, 2)):
        A = graphs.rgg(40 * target, 10.0, seed=1)
        As, Ps = graphs.coarsen(A, 0.25, min_coarse=target)
        Ac = As[-1]
        n = Ac.shape[0]
        x0 = capi.reference_uniform(1, n * dim).reshape(-1, dim)
        for skip in (0, 1, 2, 3):
            os.environ["GE_ONCHIP_SKIP"] = str(skip)
            ts = []
            for iters in (1, 20001):
                ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
                t = time.time()
                ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
                ts.append(time.time() - t)
            print("n=%d d=%d skip=%d (1=pairs 2=epilogue 4=barrier): %.3f us/iter" % (n, dim, skip, 1e6 * (ts[1] - ts[0]) / 20000), flush=True)
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    prec = capi.GE_F32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else capi.GE_F64
    A = graphs.rgg(n, 10.0, seed=11)
    n = A.shape[0]
    plan = ctx.flat_plan(A, 3, capi.flat_params(precision=prec))
    plan.upload(capi.reference_uniform(5, n * 3).reshape(n, 3))
    plan.select_kernels(2)
    plan.iterate(2)
    plan.sync()
    plan.profile(True)
    plan.iterate(3)
    print(plan.profile_get(), n, A.nnz)
