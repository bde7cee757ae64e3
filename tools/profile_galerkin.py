"""Small driver for ncu: one level of the Galerkin product on the device.
usage: python tools/profile_galerkin.py [n] [kind rgg|rmat]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "rgg"
A = graphs.rgg(n, 10.0, seed=13) if kind == "rgg" else graphs.rmat(n, 16, seed=3)
As, Ps = graphs.coarsen(A, 0.25, min_coarse=1000, max_levels=1)
ctx = capi.Context(0)
ctx.galerkin(A, Ps[0])
C, st = ctx.galerkin(A, Ps[0], with_stats=True)
print(A.shape, Ps[0].shape, st)
