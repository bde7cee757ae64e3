"""Small driver for ncu: the attraction + step kernel alone on a graph larger than L2.
usage: python tools/profile_attr.py [n] [dim] [f64|f32] [iters] [rgg|delaunay|rmat]
(rmat: n is the scale)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
prec = capi.GE_F32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else capi.GE_F64
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
kind = sys.argv[5] if len(sys.argv) > 5 else "rgg"
A = (graphs.rgg(n, 10.0, seed=11) if kind == "rgg" else graphs.delaunay3d(n, seed=3) if kind == "delaunay"
     else graphs.rmat(n, 16, seed=5))
n = A.shape[0]
os.environ["GE_REP_SYM"] = "0"
ctx = capi.Context(0)
plan = ctx.flat_plan(A, dim, capi.flat_params(precision=prec))
plan.upload(capi.reference_uniform(5, n * dim).reshape(n, dim))
plan.select_kernels(2)
plan.iterate(iters)
plan.sync()
print("done", n, ctx.launches)
