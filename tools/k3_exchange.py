"""us/iteration of the coarsest-level solve with the two position exchanges of the cluster kernels:
GE_K3_EXCHANGE=0 (DSMEM stores + cluster barrier) vs 1 (st.async + mbarrier)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs
import scipy.sparse as sp

ctx = capi.Context(0)
cases = []
for target in (34, 42, 64, 100, 157, 300):
    A = graphs.rgg(40 * target, 10.0, seed=1)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=target)
    cases.append(("rgg-coarse", As[-1]))
rng = np.random.default_rng(3)
M = rng.random((54, 54))
cases.append(("dense54", graphs.canonical(sp.csr_matrix(M + M.T))))
for dim in (2, 3):
    for name, Ac in cases:
        n = Ac.shape[0]
        x0 = capi.reference_uniform(1, n * dim).reshape(-1, dim)
        res, out = [], []
        for mode in ("0", "1"):
            os.environ["GE_K3_EXCHANGE"] = mode
            ts = []
            for iters in (1, 20001):
                ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
                t = time.time()
                x = ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
                ts.append(time.time() - t)
            res.append(1e6 * (ts[1] - ts[0]) / 20000)
            out.append(ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=50)))
        same = bool(np.array_equal(out[0], out[1]))
        print("%s n=%d nnz=%d d=%d: barrier %.3f us/iter, st.async %.3f us/iter, identical positions after 50 iterations: %s"
              % (name, n, Ac.nnz, dim, res[0], res[1], same), flush=True)
