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
