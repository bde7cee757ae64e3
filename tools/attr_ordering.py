"""Is the attraction kernel's shortfall on 3-D meshes an ORDERING problem?  The same Delaunay graph
with its vertices numbered (a) at random + the plan's breadth-first renumbering (what a caller
gets), (b) along a Morton curve of the generating points with the internal renumbering off."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.load_package()
from graph_embed_b200 import capi, graphs
from scipy.spatial import Delaunay
import scipy.sparse as sp

n = 1_000_000
rng = np.random.default_rng(3)
pts = rng.random((n, 3))
def morton(p, bits=10):
    q = np.minimum((p * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    code = np.zeros(len(p), dtype=np.uint64)
    for b in range(bits):
        for k in range(3):
            code |= ((q[:, k] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + k)
    return code
order = np.argsort(morton(pts), kind="stable")
tets = Delaunay(pts).simplices
pr = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
rows = np.concatenate([tets[:, i] for i, _ in pr]); cols = np.concatenate([tets[:, j] for _, j in pr])
A = graphs._finish(rows, cols, n)
inv = np.empty(n, dtype=np.int64); inv[order] = np.arange(n)
coo = A.tocoo()
B = graphs.canonical(sp.csr_matrix((coo.data, (inv[coo.row], inv[coo.col])), shape=A.shape))
peak = 6544.0
ctx = capi.Context(0)
for name, M, env in (("random numbering + internal BFS", A, {}), ("Morton numbering, no internal renumbering", B, {"GE_NO_REORDER": "1"}),
                     ("Morton numbering + internal BFS", B, {})):
    for dim in (2, 3):
        os.environ.pop("GE_NO_REORDER", None)
        os.environ.update(env)
        b = float(M.nnz) * 12 + float(n) * (4 + 8 + 5 * dim * 8)
        plan = ctx.flat_plan(M, dim, capi.flat_params())
        plan.upload(capi.reference_uniform(5, n * dim).reshape(n, dim))
        plan.select_kernels(2)
        plan.iterate(2); plan.sync(); plan.profile(True); plan.iterate(5)
        prof = plan.profile_get(); plan.close()
        ms = prof["attract_step_ms"] / prof["attract_step_launches"]
        print("%-45s d=%d %.4f ms %.3f of HBM" % (name, dim, ms, b / (ms * 1e-3) / 1e9 / peak), flush=True)
