"""Sweep of the repulsion kernel's launch shape (threads per CTA x rows per thread) on one GPU.
usage: python tools/sweep_rep.py [n] [dim] [f64|f32]"""
import os
import sys

import numpy as np

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
print("peak TF", ctx.fma_peak_tflops(prec))
flops = float(n) * (n - 1) * (5 * dim + 4)
import itertools
for ju, ipt, thr in itertools.product((1, 2), (1, 2, 4), (128, 256, 512)):
    if True:
        os.environ["GE_REP_THREADS"], os.environ["GE_REP_IPT"], os.environ["GE_REP_JU"] = str(thr), str(ipt), str(ju)
        try:
            plan = ctx.flat_plan(A, dim, capi.flat_params(precision=prec))
        except capi.GeError as e:
            print(ipt, thr, "n/a", e)
            continue
        plan.upload(x0)
        plan.iterate(1)
        plan.sync()
        plan.profile(True)
        plan.iterate(2)
        p = plan.profile_get()
        ms = p["repulsion_ms"] / p["repulsion_launches"]
        print("ju=%d ipt=%d thr=%3d  %.2f ms  %.2f TF" % (ju, ipt, thr, ms, flops / ms / 1e9), flush=True)
        plan.close()
