"""Launch-shape sweep of the symmetric repulsion kernel against the ordered sweep on one GPU.
usage: python tools/sweep_sym.py [n] [dim] [f64|f32]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prec = capi.GE_F32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else capi.GE_F64
A = graphs.rgg(n, 10.0, seed=7)
n = A.shape[0]
x0 = capi.reference_uniform(23, n * dim).reshape(n, dim)
ctx = capi.Context(0)
peak = ctx.fma_peak_tflops(prec)
print("n", n, "dim", dim, "peak TF", peak)
flops = float(n) * (n - 1) * (5 * dim + 4)
for sym, ipt, cg in [(0, 0, 0), (1, 4, 8), (1, 4, 4), (1, 2, 8), (1, 2, 4)]:
    os.environ["GE_REP_SYM"] = str(sym)
    os.environ["GE_SYM_IPT"], os.environ["GE_SYM_CG"] = str(ipt), str(cg)
    plan = ctx.flat_plan(A, dim, capi.flat_params(precision=prec))
    plan.upload(x0)
    plan.iterate(1)
    plan.sync()
    plan.profile(True)
    plan.iterate(2)
    p = plan.profile_get()
    ms = p["repulsion_ms"] / p["repulsion_launches"]
    print("sym=%d ipt=%d cg=%d  %.2f ms  %.2f TF algorithmic  %.1f%% of peak" %
          (sym, ipt, cg, ms, flops / ms / 1e9, 100 * flops / ms / 1e9 / peak), flush=True)
    plan.close()
