"""Every kernel family once, at sizes compute-sanitizer gets through in a minute or two:
  compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py
  compute-sanitizer --tool synccheck python tools/sanitize_smoke.py
(compute-sanitizer is closed on the round-2 GPU pool: the script was run plain there.)
Results are checked for finiteness only (parity is the test suite's job)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

entry.load_package()
from graph_embed_b200 import capi, graphs

only = set(sys.argv[1:])


def want(name):
    return not only or name in only


def finite(name, *arrays):
    for a in arrays:
        assert np.isfinite(a).all(), name
    print("ok", name, flush=True)


ctx = capi.Context(0)

if want("flat"):
    A = graphs.rgg(3000, 10.0, seed=3)
    n = A.shape[0]
    for dim in (2, 3):
        x0 = capi.reference_uniform(7, n * dim).reshape(n, dim)
        for prec in (capi.GE_F64, capi.GE_F32):
            for sym in ("0", "1"):
                os.environ["GE_REP_SYM"] = sym
                F = ctx.flat_forces(A, dim, x0, capi.flat_params(precision=prec))
                x = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(precision=prec, iterations=3))
                finite("flat tiled d=%d prec=%d sym=%s" % (dim, prec, sym), F, x)
    del os.environ["GE_REP_SYM"]
    # power-law rows: the long-row attraction tiers
    B = graphs.largest_component(graphs.rmat(12, 16, seed=5))
    nb = B.shape[0]
    y0 = capi.reference_uniform(9, nb * 3).reshape(nb, 3)
    finite("flat tiled rmat-12 d=3", ctx.flat_forceatlas(B, 3, y0, capi.flat_params(iterations=2)))
    os.environ["GE_LONG_ROW"] = "8"
    finite("flat tiled rmat-12 long rows from 8", ctx.flat_forceatlas(B, 3, y0, capi.flat_params(iterations=2)))
    del os.environ["GE_LONG_ROW"]
    # column-panel passes of the symmetric sweep
    os.environ["GE_SYM_PASSES"] = "3"
    x0 = capi.reference_uniform(7, n * 2).reshape(n, 2)
    finite("flat symmetric, 3 passes", ctx.flat_forceatlas(A, 2, x0, capi.flat_params(iterations=2)))
    del os.environ["GE_SYM_PASSES"]

if want("onchip"):
    for n, deg in ((20, 6.0), (50, 6.0), (120, 6.0), (54, 40.0), (300, 8.0)):
        A = graphs.largest_component(graphs.rgg(n, deg, seed=11))
        k = A.shape[0]
        for dim in (2, 3):
            x0 = capi.reference_uniform(3, k * dim).reshape(k, dim)
            for prec in (capi.GE_F64, capi.GE_F32):
                x = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(precision=prec, iterations=12))
                finite("on-chip n=%d entries/row=%.0f d=%d prec=%d" % (k, A.nnz / k, dim, prec), x)

if want("multilevel"):
    A = graphs.rgg(9000, 10.0, seed=5)
    n = A.shape[0]
    sizes = [3000, 1500, 1100, 700, 300, 100, 40, 33, 20, 5, 2, 1, 1]
    agg = np.empty(n, dtype=np.int64)
    pos, a = 0, 0
    for s in sizes:
        agg[pos:pos + s] = a
        pos, a = pos + s, a + 1
    while pos < n:  # the rest in aggregates of 7
        agg[pos:pos + 7] = a
        pos, a = pos + 7, a + 1
    m = a
    P_T = graphs.aggregation_matrix(agg, m)
    for dim in (2, 3):
        cA = capi.reference_uniform(13, m * dim).reshape(m, dim) * 10
        rA = np.full(m, 0.3)
        for prec in (capi.GE_F64, capi.GE_F32):
            for knob in ({}, {"GE_ML_SYM_MIN_MPAIRS": "0"}, {"GE_ML_SYM": "0"}):
                os.environ.update(knob)
                p = capi.multilevel_params(precision=prec, iterations=10, seed=5)
                x = ctx.multilevel_forceatlas(A, P_T, cA, rA, dim, p)
                F = ctx.multilevel_forces(A, P_T, cA, x, dim, p)
                for key in knob:
                    del os.environ[key]
                finite("multilevel d=%d prec=%d %s" % (dim, prec, knob), x, F)
        x = ctx.multilevel_forceatlas(A, P_T, cA, rA, dim, capi.multilevel_params(iterations=4, seed=5),
                                      aggregates=(1, 40))
        finite("multilevel shard d=%d" % dim, x)

if want("radii"):
    A = graphs.rgg(6000, 10.0, seed=8)
    As, P_Ts = graphs.coarsen(A, 0.25, min_coarse=40, seed=1)
    for dim in (2, 3):
        k = len(P_Ts)
        xc = capi.reference_uniform(3, As[k].shape[0] * dim).reshape(-1, dim)
        xc, rc = ctx.level_radii(xc, dim)
        finite("radii base m=%d d=%d" % (As[k].shape[0], dim), xc, rc)
        for l in range(k - 1, 0, -1):
            x = capi.reference_uniform(17 + l, As[l].shape[0] * dim).reshape(-1, dim)
            for batch in ("0", "1"):
                os.environ["GE_RADII_BATCH"] = batch
                xl, rl = ctx.level_radii(x, dim, As[l], P_Ts[l], xc, rc)
                finite("radii level %d m=%d d=%d batch=%s" % (l, As[l].shape[0], dim, batch), xl, rl)
            del os.environ["GE_RADII_BATCH"]
            xc, rc = xl, rl
    # one family holding many members (wide kernel, events staged in shared memory / global)
    B = graphs.largest_component(graphs.rmat(11, 16, seed=2))
    nb = B.shape[0]
    aggb = np.minimum(np.arange(nb) // 600, 2).astype(np.int64)
    P = graphs.aggregation_matrix(aggb, 3)
    xb = capi.reference_uniform(4, nb * 3).reshape(nb, 3)
    xcb, rcb = ctx.level_radii(capi.reference_uniform(6, 9).reshape(3, 3), 3)
    finite("radii hub families", *ctx.level_radii(xb, 3, B, P, xcb, rcb))

if want("galerkin"):
    A = graphs.rgg(8000, 10.0, seed=2)
    As, P_Ts = graphs.coarsen(A, 0.25, min_coarse=60, seed=3)
    for l, P in enumerate(P_Ts):
        Ac = ctx.galerkin(As[l], P)
        finite("galerkin level %d -> %d rows" % (l, Ac.shape[0]), Ac.data)
    B = graphs.largest_component(graphs.rmat(12, 16, seed=5))
    Bs, Q_Ts = graphs.coarsen(B, 0.25, min_coarse=60, seed=3)
    for l, P in enumerate(Q_Ts[:2]):
        finite("galerkin rmat level %d" % l, ctx.galerkin(Bs[l], P).data)

if want("embed"):
    A = graphs.rgg(12000, 10.0, seed=4)
    As, P_Ts = graphs.coarsen(A, 0.25, min_coarse=60, seed=2)
    for dim in (2, 3):
        for seed in (0, 5):
            x, st = ctx.embed(As, P_Ts, dim, seed=seed, coarse_iterations=40, level_iterations=12)
            finite("embed d=%d seed=%d levels=%d" % (dim, seed, len(P_Ts)), x)
print("sanitize_smoke: done, launches", ctx.launches)
