"""The CSR attraction + step kernel alone (no repulsion) on a graph larger than L2: staged (TMA)
variant against the direct one, lanes per row.
usage: python tools/sweep_attr.py [n] [dim] [avg_degree]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 3
deg = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
A = graphs.rgg(n, deg, seed=11)
n, nnz = A.shape[0], A.nnz
x0 = capi.reference_uniform(5, n * dim).reshape(n, dim)
ctx = capi.Context(0)
os.environ["GE_REP_SYM"] = "0"
for prec, w, name in ((capi.GE_F64, 8, "f64"), (capi.GE_F32, 4, "f32")):
    b = nnz * (4 + w) + n * (4 + w + 5 * dim * w)
    for staged, group, cap, aos in ((0, 2, 2560, 0), (1, 2, 2560, 0), (1, 2, 2048, 0), (1, 1, 4096, 0)):
        os.environ["GE_STEP_STAGED"], os.environ["GE_STEP_GROUP"] = str(staged), str(group)
        os.environ["GE_STEP_CAP"], os.environ["GE_GATHER_COPY_REORDERED"] = str(cap), str(aos)
        plan = ctx.flat_plan(A, dim, capi.flat_params(precision=prec))
        plan.upload(x0)
        plan.select_kernels(2)
        plan.iterate(2)
        plan.sync()
        plan.profile(True)
        plan.iterate(10)
        p = plan.profile_get()
        x = plan.download()
        plan.close()
        ms = p["attract_step_ms"] / p["attract_step_launches"]
        print("%s n=%d nnz=%d d=%d staged=%d lanes=%d cap=%d aos=%d  %.4f ms  %.0f GB/s  (%.1f%% of 6544)  checksum %.12e"
              % (name, n, nnz, dim, staged, group, cap, aos, ms, b / ms / 1e6, 100 * b / ms / 1e6 / 6544, float(abs(x).sum())),
              flush=True)
