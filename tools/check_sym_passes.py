"""Symmetric sweep with column-panel passes: timing at n = 500k (1 vs 2 vs 4 passes) and at a size
whose one-pass scratch would not fit (n = 2M: 32 GB).  python tools/check_sym_passes.py [n_big]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

ctx = capi.Context(0)
A = graphs.rgg(500_000, 10.0, seed=7)
n = A.shape[0]
x0 = capi.reference_uniform(23, n * 2).reshape(n, 2)
ref = None
for passes in ("1", "2", "4"):
    os.environ["GE_SYM_PASSES"] = passes
    plan = ctx.flat_plan(A, 2, capi.flat_params())
    plan.upload(x0)
    plan.iterate(1)
    plan.sync()
    plan.profile(True)
    plan.iterate(3)
    prof = plan.profile_get()
    x = plan.download()
    plan.close()
    if ref is None:
        ref = x
    print("n=%d passes=%s: repulsion %.2f ms/iteration, max |dx| vs one pass %.2e" % (
        n, passes, prof["repulsion_ms"] / prof["repulsion_launches"], np.abs(x - ref).max()), flush=True)
del os.environ["GE_SYM_PASSES"]
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
B = graphs.rgg(nb, 10.0, seed=9)
nb = B.shape[0]
y0 = capi.reference_uniform(5, nb * 2).reshape(nb, 2)
plan = ctx.flat_plan(B, 2, capi.flat_params())
print("n=%d symmetric=%s" % (nb, plan.symmetric), flush=True)
plan.upload(y0)
plan.profile(True)
plan.iterate(1)
prof = plan.profile_get()
F = plan.download_forces()
plan.close()
ms = prof["repulsion_ms"] / prof["repulsion_launches"]
print("n=%d: repulsion %.1f ms/iteration = %.3e pair-interactions/s" % (nb, ms, float(nb) * (nb - 1) / (ms * 1e-3)))
O = entry.load_oracle()
worst = 0.0
for r in np.random.default_rng(0).choice(nb, 6, replace=False):
    Fr, S = O.flat_forces(B, 2, y0, rows=(int(r), int(r) + 1))
    worst = max(worst, float(np.linalg.norm(F[r] - Fr[r]) / S[r]))
print("sampled force error vs oracle: %.2e" % worst)
