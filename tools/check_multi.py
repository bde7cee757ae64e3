"""Diagnostic: one process driving N GPUs (ge_context_create_multi) against the single-GPU plan and
the oracle after ONE flat iteration; prints where the largest differences are.
  python tools/check_multi.py [ndev] [n] [dim]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

O = entry.load_oracle()
ndev = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
dim = int(sys.argv[3]) if len(sys.argv) > 3 else 2
A = graphs.rgg(n, 10.0, seed=3)
n = A.shape[0]
x0 = capi.reference_uniform(5, n * dim).reshape(n, dim)
single = capi.Context(0)
multi = capi.Context(devices=list(range(ndev)))
p1 = capi.flat_params(iterations=1)
one = single.flat_forceatlas(A, dim, x0, p1)
got = multi.flat_forceatlas(A, dim, x0, p1)
d = np.abs(got - one).max(axis=1)
worst = np.argsort(d)[-8:][::-1]
print("max |multi - single| = %.3e at rows %s" % (d.max(), worst.tolist()))
p = O.Params()
for r in list(worst) + list(np.random.default_rng(0).choice(n, 8, replace=False)):
    F, S = O.flat_forces(A, dim, x0, p, rows=(int(r), int(r) + 1))
    f = F[r]
    fn = float(np.sqrt((f * f).sum()))
    speed = min(p.ks / (1.0 + np.sqrt(fn)), p.ksmax / fn)
    xr = x0[r] + f * speed
    print("row %6d |F| %.3e scale %.3e speed %.3e | single-oracle %.3e multi-oracle %.3e multi-single %.3e" % (
        r, fn, S[r], speed, np.abs(one[r] - xr).max(), np.abs(got[r] - xr).max(), d[r]))
for it in (1, 1, 1):
    t = time.time()
    multi.flat_forceatlas(A, dim, x0, p1)
    print("multi call %.1f ms" % (1e3 * (time.time() - t)))
for it in (1, 1):
    t = time.time()
    single.flat_forceatlas(A, dim, x0, p1)
    print("single call %.1f ms" % (1e3 * (time.time() - t)))
