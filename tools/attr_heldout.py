"""Attraction + step kernel on the held-out graphs (Delaunay 1M points, optionally R-MAT): fraction of
the HBM roofline per dimension and precision.  python tools/attr_heldout.py [delaunay|rmat] [n or scale]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.load_package()
from graph_embed_b200 import capi, graphs
kind = sys.argv[1] if len(sys.argv) > 1 else "delaunay"
size = int(sys.argv[2]) if len(sys.argv) > 2 else (1_000_000 if kind == "delaunay" else 18)
import scipy.sparse as sp
cache = "/tmp/attr_heldout_%s_%d.npz" % (kind, size)  # several builds (GE_LIB) on one box: generate once
if os.path.exists(cache):
    A = sp.load_npz(cache)
else:
    A = graphs.delaunay3d(size, seed=3) if kind == "delaunay" else graphs.largest_component(graphs.rmat(size, 16, seed=5))
    sp.save_npz(cache, A, compressed=False)
n, nnz = A.shape[0], A.nnz
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6544.0
ctx = capi.Context(0)
for dim in (2, 3):
    for prec, w in ((capi.GE_F64, 8), (capi.GE_F32, 4)):
        b = float(nnz) * (4 + w) + float(n) * (4 + w + 5 * dim * w)
        plan = ctx.flat_plan(A, dim, capi.flat_params(precision=prec))
        plan.upload(capi.reference_uniform(5, n * dim).reshape(n, dim))
        plan.select_kernels(2)
        plan.iterate(2)
        plan.sync()
        plan.profile(True)
        plan.iterate(8)
        prof = plan.profile_get()
        plan.close()
        ms = prof["attract_step_ms"] / prof["attract_step_launches"]
        print("%s n=%d nnz=%d d=%d w=%d: %.4f ms  %.3f of HBM peak %.0f GB/s" % (
            kind, n, nnz, dim, w, ms, b / (ms * 1e-3) / 1e9 / peak, peak), flush=True)
