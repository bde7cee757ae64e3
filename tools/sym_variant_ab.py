"""One build of the library (GE_LIB selects it) on the symmetric repulsion sweep: ms per iteration at
d = 2 and 3 (FP64) and the worst sampled force error against the oracle.  Run once per build:
  python tools/sym_variant_ab.py [n];  GE_LIB=graph-embed_b200/lib/alt/libgraphembed_b200.so python tools/sym_variant_ab.py [n]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
A = graphs.rgg(n, 10.0, seed=7)
n = A.shape[0]
O = entry.load_oracle()
ctx = capi.Context(0)
for dim in (2, 3):
    x0 = capi.reference_uniform(23, n * dim).reshape(n, dim)
    plan = ctx.flat_plan(A, dim, capi.flat_params())
    plan.upload(x0)
    plan.iterate(1)
    plan.sync()
    plan.upload(x0)
    plan.profile(True)
    plan.iterate(1)
    F = plan.download_forces()
    plan.iterate(3)
    p = plan.profile_get()
    plan.close()
    ms = p["repulsion_ms"] / p["repulsion_launches"]
    worst = 0.0
    for r in np.random.default_rng(0).choice(n, 12, replace=False):
        Fr, S = O.flat_forces(A, dim, x0, rows=(int(r), int(r) + 1))
        worst = max(worst, float(np.linalg.norm(F[r] - Fr[r]) / S[r]))
    print("lib=%s n=%d d=%d: %.2f ms/iteration, %.4e pair-interactions/s, force error %.2e" % (
        os.path.basename(os.path.dirname(capi.LIB_PATH)), n, dim, ms, float(n) * (n - 1) / (ms * 1e-3), worst), flush=True)
