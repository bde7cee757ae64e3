"""us/iteration of the coarsest-level solve over lanes per vertex x cluster size (v1 cluster kernel)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.load_package()
from graph_embed_b200 import capi, graphs
ctx = capi.Context(0)
for target in (42, 64, 100, 157, 200):
    A = graphs.rgg(40 * target, 10.0, seed=1)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=target)
    Ac = As[-1]
    n = Ac.shape[0]
    for dim in (2, 3):
        x0 = capi.reference_uniform(1, n * dim).reshape(-1, dim)
        row = []
        for cs in ("8", "16"):
            for L in ("8", "16", "32"):
                os.environ["GE_CLUSTER"] = cs
                os.environ["GE_ONCHIP_LANES"] = L
                ts = []
                for iters in (1, 20001):
                    ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
                    t = time.time()
                    ctx.flat_forceatlas(Ac, dim, x0, capi.flat_params(iterations=iters))
                    ts.append(time.time() - t)
                row.append("c%s/L%s=%.3f" % (cs, L, 1e6 * (ts[1] - ts[0]) / 20000))
        print("n=%d d=%d: %s" % (n, dim, "  ".join(row)), flush=True)
