"""Attraction + step kernel on the held-out Delaunay graph under the launch knobs
(GE_STEP_GROUP lanes per row, GE_GATHER_COPY_REORDERED, GE_STEP_STAGED): fraction of the HBM roofline."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
entry.load_package()
from graph_embed_b200 import capi, graphs
kind = sys.argv[1] if len(sys.argv) > 1 else "delaunay"
A = graphs.delaunay3d(1_000_000, seed=3) if kind == "delaunay" else graphs.rgg(2_000_000, 10.0, seed=11)
n, nnz = A.shape[0], A.nnz
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6544.0
ctx = capi.Context(0)
for dim in (2, 3):
    b = float(nnz) * 12 + float(n) * (4 + 8 + 5 * dim * 8)
    for env in ({}, {"GE_STEP_GROUP": "1"}, {"GE_STEP_GROUP": "2"}, {"GE_STEP_GROUP": "4"},
                {"GE_GATHER_COPY_REORDERED": "0"}, {"GE_GATHER_COPY_REORDERED": "1"},
                {"GE_STEP_GROUP": "1", "GE_GATHER_COPY_REORDERED": "0"}, {"GE_STEP_STAGED": "0"}):
        for k in ("GE_STEP_GROUP", "GE_GATHER_COPY_REORDERED", "GE_STEP_STAGED"):
            os.environ.pop(k, None)
        os.environ.update(env)
        plan = ctx.flat_plan(A, dim, capi.flat_params())
        plan.upload(capi.reference_uniform(5, n * dim).reshape(n, dim))
        plan.select_kernels(2)
        plan.iterate(2)
        plan.sync()
        plan.profile(True)
        plan.iterate(5)
        prof = plan.profile_get()
        plan.close()
        ms = prof["attract_step_ms"] / prof["attract_step_launches"]
        print("%s d=%d %-55s %.4f ms  %.3f of HBM" % (kind, dim, env, ms, b / (ms * 1e-3) / 1e9 / peak), flush=True)
