"""Driver for the per-aggregate level solve (K2): one forceAtlasMultilevel call on the finest level
of a synthetic hierarchy.  usage: python tools/profile_levels.py [delaunay|rgg|rmat] [n|scale] [dim] [f64|f32]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

kind = sys.argv[1] if len(sys.argv) > 1 else "delaunay"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 500_000
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 3
prec = capi.GE_F32 if (len(sys.argv) > 4 and sys.argv[4] == "f32") else capi.GE_F64
if kind == "delaunay":
    A, cf = graphs.delaunay3d(size, seed=1), 0.125
elif kind == "rgg":
    A, cf = graphs.rgg(size, 10.0, seed=1), 0.25
else:
    A, cf = graphs.rmat(size, 16, seed=1), 0.25
As, Ps = graphs.coarsen(A, cf, min_coarse=64, max_levels=1)
P = Ps[0]
s = np.diff(P.indptr)
print("n", A.shape[0], "nnz", A.nnz, "aggregates", P.shape[0], "size max", s.max(), "pairs/iter", int((s * (s - 1)).sum()),
      "hist", np.bincount(np.minimum(s, 40))[:41].tolist())
m = P.shape[0]
rng = np.random.default_rng(0)
cA, rA = rng.normal(size=(m, dim)), rng.random(m) + 0.1
ctx = capi.Context(0)
p = capi.multilevel_params(precision=prec, seed=3)
for rep in range(3):
    l0 = ctx.launches
    t = time.time()
    x = ctx.multilevel_forceatlas(A, P, cA, rA, dim, p)
    print("level solve %.2f ms, launches %d" % (1e3 * (time.time() - t), ctx.launches - l0))
assert np.isfinite(x).all()
