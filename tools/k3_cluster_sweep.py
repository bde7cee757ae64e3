"""us/iteration of the coarsest-level solve over the cluster size (st.async exchange), d = 2 and 3."""
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
for target in (20, 27, 34, 42, 54, 64, 80, 100, 130, 157, 200):
    A = graphs.rgg(40 * target, 10.0, seed=1)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=target)
    Ac = As[-1]
    n = Ac.shape[0]
    for dim in (2, 3):
        x0 = capi.reference_uniform(1, n * dim).reshape(-1, dim)
        row = []
        for cs in ("auto", "1", "2", "4", "8", "16"):
            if cs == "auto":
                os.environ.pop("GE_CLUSTER", None)
            else:
                os.environ["GE_CLUSTER"] = cs
            try:
                ts = []
                for iters in (1, 20001):
                    ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
                    t = time.time()
                    ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
                    ts.append(time.time() - t)
                row.append("%s=%.3f" % (cs, 1e6 * (ts[1] - ts[0]) / 20000))
            except Exception as e:
                row.append("%s=err" % cs)
        print("n=%d d=%d us/iter by cluster size: %s" % (n, dim, "  ".join(row)), flush=True)
