"""ncu driver on the reference partitioner's hierarchies: one level solve (the large-aggregate tier:
k_repulsion_sym over segments) and one radii step (k_radii_grow) on the device.
usage: python tools/profile_refhier.py [rmat20|delaunay1000000] [level]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs
from helpers import load_ref_hierarchy

name = sys.argv[1] if len(sys.argv) > 1 else "delaunay1000000"
l = int(sys.argv[2]) if len(sys.argv) > 2 else 1
As, Ps, _ = load_ref_hierarchy(graphs, name)
A, P = As[l], Ps[l]
m, dim = P.shape[0], 3
rng = np.random.default_rng(0)
cA, rA = rng.normal(size=(m, dim)), rng.random(m) + 0.1
ctx = capi.Context(0)
p = capi.multilevel_params(seed=3, iterations=10)
t = time.time()
x = ctx.multilevel_forceatlas(A, P, cA, rA, dim, p)
print("level %d solve (10 iterations) %.1f ms" % (l, 1e3 * (time.time() - t)))
# radii of this level's vertices inside the families of the next level
if l + 1 < len(Ps):
    mc = Ps[l + 1].shape[0]
    cAc, rAc = rng.normal(size=(mc, dim)), rng.random(mc) + 0.1
    xs = rng.normal(size=(As[l + 1].shape[0], dim))
    t = time.time()
    ctx.level_radii(xs, dim, As[l + 1], Ps[l + 1], cAc, rAc)
    print("radii of level %d: %.1f ms" % (l + 1, 1e3 * (time.time() - t)))
