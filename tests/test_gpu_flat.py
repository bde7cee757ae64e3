"""GPU parity of the flat ForceAtlas kernels (K1 tiled multi-CTA path, K3 on-chip path) against
the oracle, through the C ABI.  Force tolerances are the ones BASELINE.json states: 1e-10 (FP64)
and 1e-4 (FP32), relative to the conditioning scale of each vertex's force sum."""
import numpy as np
import pytest

from helpers import TOL_F32, TOL_F64, force_error, load_flat_golden

pytestmark = pytest.mark.gpu


def _graph(graphs, name):
    if name == "grid12":
        return load_flat_golden()[0]
    if name == "rgg2000":
        return graphs.rgg(2000, 10.0, seed=1)
    if name == "galerkin":  # weighted, with self-loops (quirk Q5)
        A = graphs.rgg(2400, 10.0, seed=2)
        As, _ = graphs.coarsen(A, 0.25, min_coarse=200, max_levels=1)
        return As[1]
    raise KeyError(name)


@pytest.mark.parametrize("name", ["grid12", "rgg2000", "galerkin"])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("path", [1, 2])
def test_forces_fp64(ctx, capi, oracle, graphs, name, dim, path):
    A = _graph(graphs, name)
    n = A.shape[0]
    if path == 2 and n > 1024:
        pytest.skip("on-chip path holds at most 1024 vertices")
    x0 = capi.reference_uniform(17, n * dim).reshape(n, dim)
    F_ref, S = oracle.flat_forces(A, dim, x0)
    F = ctx.flat_forces(A, dim, x0, capi.flat_params(), path=path)
    err = force_error(F, F_ref, S)
    assert err.max() < TOL_F64, err.max()
    # plain relative error of the whole force field as well
    assert np.linalg.norm(F - F_ref) / np.linalg.norm(F_ref) < TOL_F64


@pytest.mark.parametrize("name", ["grid12", "rgg2000"])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("path", [1, 2])
def test_forces_fp32(ctx, capi, oracle, graphs, name, dim, path):
    A = _graph(graphs, name)
    n = A.shape[0]
    if path == 2 and n > 1024:
        pytest.skip("on-chip path holds at most 1024 vertices")
    x0 = capi.reference_uniform(17, n * dim).reshape(n, dim)
    F_ref, S = oracle.flat_forces(A, dim, x0)
    F = ctx.flat_forces(A, dim, x0, capi.flat_params(precision=capi.GE_F32), path=path)
    assert force_error(F, F_ref, S).max() < TOL_F32


@pytest.mark.parametrize("kw", [dict(use_weights=0), dict(linlog=1), dict(nohubs=1), dict(delta=0.5),
                                dict(delta=0.0), dict(ks=0.3, ksmax=2.0, repel=2.0, attract=0.7, gravity=1.5, tolerate=0.8)])
@pytest.mark.parametrize("path", [1, 2])
def test_forces_options(ctx, capi, oracle, graphs, kw, path):
    A = _graph(graphs, "galerkin")
    n = A.shape[0]
    x0 = capi.reference_uniform(5, n * 2).reshape(n, 2)
    okw = {"useWeights" if k == "use_weights" else k: v for k, v in kw.items()}
    F_ref, S = oracle.flat_forces(A, 2, x0, oracle.Params(**okw))
    F = ctx.flat_forces(A, 2, x0, capi.flat_params(**kw), path=path)
    assert force_error(F, F_ref, S).max() < TOL_F64


def test_coincident_points_and_padding(ctx, capi, oracle, graphs):
    """eps clamp (include/forceatlas.hpp:155-157) and tile padding: n not a multiple of 256."""
    A = graphs.grid2d(9, 29)  # n = 261
    n = A.shape[0]
    x0 = capi.reference_uniform(1, n * 2).reshape(n, 2)
    x0[10] = x0[11]          # distance 0 -> clamped, direction 0
    x0[20] = x0[21] + 1e-7   # below eps
    F_ref, S = oracle.flat_forces(A, 2, x0)
    for path in (1, 2):
        F = ctx.flat_forces(A, 2, x0, capi.flat_params(), path=path)
        assert np.isfinite(F).all()
        assert force_error(F, F_ref, S).max() < TOL_F64


@pytest.mark.parametrize("k", [1, 5, 25])
@pytest.mark.parametrize("path", ["tiled", "onchip"])
def test_positions_after_k_iterations(ctx, capi, oracle, k, path, monkeypatch):
    """Golden positions of the compiled reference after k iterations.  The map is chaotic (a 1-ulp
    change of the initial coordinates moves the ORACLE's own result by 1e-5 after 25 iterations in
    2-D), so the bound is the oracle's measured sensitivity to a 1-ulp input perturbation times 50,
    floored at 1e-12."""
    A, z = load_flat_golden()
    monkeypatch.setenv("GE_ONCHIP_MAX", "0" if path == "tiled" else "1024")
    for dim in (2, 3):
        x0 = z["x0_d%d" % dim]
        ref = z["x_d%d_k%d" % (dim, k)]
        sign = np.random.default_rng(0).choice([-1.0, 1.0], size=x0.shape)
        pert, _ = oracle.flat_run(A, dim, x0 * (1 + sign * 2.2e-16), oracle.Params(iterations=k))
        tol = max(1e-12, 50 * np.abs(pert - ref).max())
        x = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=k))
        assert np.abs(x - ref).max() < tol, (np.abs(x - ref).max(), tol)


def test_normalize_option(ctx, capi, oracle, monkeypatch):
    """normalize=true epilogue (include/forceatlas.hpp:272-303) on every flat path (tiled, single
    CTA, cluster); 7 iterations, bound = the oracle's own sensitivity to a 1-ulp input change."""
    A, z = load_flat_golden()
    x0, ref = z["x0_d2"], z["x_d2_k7_normalize"]
    sign = np.random.default_rng(0).choice([-1.0, 1.0], size=x0.shape)
    pert, _ = oracle.flat_run(A, 2, x0 * (1 + sign * 2.2e-16), oracle.Params(iterations=7, normalize=True))
    tol = max(1e-10, 50 * np.abs(pert - ref).max())
    for onchip_max, cluster in (("0", "1"), ("1024", "1"), ("1024", "8")):
        monkeypatch.setenv("GE_ONCHIP_MAX", onchip_max)
        monkeypatch.setenv("GE_CLUSTER", cluster)
        x = ctx.flat_forceatlas(A, 2, x0, capi.flat_params(iterations=7, normalize=1))
        assert np.abs(x - ref).max() < tol, (onchip_max, cluster, np.abs(x - ref).max(), tol)
        assert abs(np.linalg.norm(x, axis=1).max() - 1.0) < 1e-12


def test_origin_vertex_gives_nan_like_reference(ctx, capi, graphs):
    """include/forceatlas.hpp:205 divides by |x_i| unclamped: a vertex at the origin goes NaN."""
    A = graphs.grid2d(3, 3)
    x0 = capi.reference_uniform(2, 18).reshape(9, 2)
    x0[4] = 0.0
    for path in (1, 2):
        F = ctx.flat_forces(A, 2, x0, capi.flat_params(), path=path)
        assert np.isnan(F[4]).all() and np.isfinite(np.delete(F, 4, axis=0)).all()


def test_plan_row_blocks_equal_full(ctx, capi, oracle, graphs):
    """Row-block sharding (multi-GPU layout): two plans owning half the rows each reproduce the
    single-plan forces and positions exactly."""
    A = graphs.rgg(1500, 10.0, seed=4)
    n = A.shape[0]
    x0 = capi.reference_uniform(9, n * 2).reshape(n, 2)
    p = capi.flat_params()
    full = ctx.flat_plan(A, 2, p)
    full.upload(x0)
    full.iterate(1)
    x_full, f_full = full.download(), full.download_forces()
    half = n // 2
    xs = []
    for rows in ((0, half), (half, n)):
        pl = ctx.flat_plan(A, 2, p, rows=rows)
        pl.upload(x0)
        pl.iterate(1)
        assert np.array_equal(pl.download_forces(), f_full[rows[0]:rows[1]])
        xs.append(pl.download()[rows[0]:rows[1]])
    assert np.array_equal(np.vstack(xs), x_full)


def test_full_size_properties(ctx, capi, oracle, graphs):
    """BASELINE config 4 size (n = 500 000, all pairs, d = 2) through size-independent properties:
    (1) Newton's third law: repulsion and attraction are antisymmetric, so the forces minus the
        gravity term sum to zero; (2) 48 sampled rows agree with the oracle to 1e-10."""
    n = 500_000
    A = graphs.rgg(n, 10.0, seed=7)
    n = A.shape[0]
    x0 = capi.reference_uniform(23, n * 2).reshape(n, 2)
    F = ctx.flat_forces(A, 2, x0, capi.flat_params(), path=1)
    assert np.isfinite(F).all()
    deg = np.asarray(A.sum(axis=1)).ravel()
    grav = -x0 / np.linalg.norm(x0, axis=1, keepdims=True) * (deg + 1)[:, None]
    net = (F - grav).sum(axis=0)
    assert np.abs(net).max() < 1e-9 * np.abs(F - grav).sum()
    rows = np.random.default_rng(0).choice(n, 48, replace=False)
    for r in rows:
        F_ref, S = oracle.flat_forces(A, 2, x0, rows=(int(r), int(r) + 1))
        assert np.linalg.norm(F[r] - F_ref[r]) / S[r] < TOL_F64


@pytest.mark.parametrize("csize", [2, 4, 8])
@pytest.mark.parametrize("dim", [2, 3])
def test_cluster_solve_matches_single_cta(ctx, capi, oracle, csize, dim, monkeypatch):
    """The thread-block-cluster variant of the coarsest-level solve (DSMEM position exchange) gives
    the positions of the single-CTA kernel: same arithmetic per vertex, only the placement differs,
    so the results are bit-identical; both are within the chaos-aware bound of the golden
    reference positions."""
    A, z = load_flat_golden()
    monkeypatch.setenv("GE_ONCHIP_MAX", "1024")
    monkeypatch.setenv("GE_K3_V1", "1")         # the first-generation cluster kernel
    monkeypatch.setenv("GE_ONCHIP_LANES", "4")  # same lane count -> same summation order
    x0 = z["x0_d%d" % dim]
    for k in (1, 25):
        monkeypatch.setenv("GE_CLUSTER", "1")
        x1 = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=k))
        monkeypatch.setenv("GE_CLUSTER", str(csize))
        xc = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=k))
        assert np.array_equal(x1, xc), np.abs(x1 - xc).max()


@pytest.mark.parametrize("csize,lanes,u", [(4, 8, 4), (8, 8, 8), (8, 16, 4), (16, 16, 8), (8, 8, 4)])
@pytest.mark.parametrize("dense", [0, 1])
def test_cluster_solve_v2_against_golden(ctx, capi, oracle, csize, lanes, u, dense, monkeypatch):
    """The second-generation cluster kernel (8 columns per lane and trip, own-position work inside
    the barrier, optional dense weight rows folded into the pair loop) in every launch shape:
    the compiled reference's golden positions within the chaos-aware bound, d = 2 and 3."""
    A, z = load_flat_golden()
    monkeypatch.setenv("GE_ONCHIP_MAX", "1024")
    monkeypatch.setenv("GE_CLUSTER", str(csize))
    monkeypatch.setenv("GE_ONCHIP_LANES", str(lanes))
    monkeypatch.setenv("GE_K3_U", str(u))
    monkeypatch.setenv("GE_K3_DENSE", str(dense))
    for dim in (2, 3):
        x0 = z["x0_d%d" % dim]
        for k in (1, 5, 25):
            ref = z["x_d%d_k%d" % (dim, k)]
            sign = np.random.default_rng(0).choice([-1.0, 1.0], size=x0.shape)
            pert, _ = oracle.flat_run(A, dim, x0 * (1 + sign * 2.2e-16), oracle.Params(iterations=k))
            tol = max(1e-12, 50 * np.abs(pert - ref).max())
            x = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=k))
            assert np.abs(x - ref).max() < tol, (dim, k, np.abs(x - ref).max(), tol)


def test_dense_weighted_coarse_graph_with_self_loops(ctx, capi, oracle, graphs):
    """A coarsest level as power-law hierarchies produce it: (nearly) complete, real weights,
    diagonal entries (quirk Q5).  The default launch picks the dense variant; positions after k
    iterations against the oracle within the chaos-aware bound, plus normalize and options."""
    rng = np.random.default_rng(3)
    n = 54
    M = rng.random((n, n)) * (rng.random((n, n)) < 0.9)
    M = M + M.T + np.diag(rng.random(n) * 5)
    import scipy.sparse as sp
    A = graphs.canonical(sp.csr_matrix(M))
    for dim in (2, 3):
        x0 = capi.reference_uniform(9, n * dim).reshape(n, dim)
        for k in (1, 3, 10):
            for kw in (dict(), dict(normalize=1), dict(attract=0.5, repel=2.0, gravity=0.3)):
                okw = dict(kw)
                ref, _ = oracle.flat_run(A, dim, x0, oracle.Params(iterations=k, **okw))
                sign = np.random.default_rng(0).choice([-1.0, 1.0], size=x0.shape)
                pert, _ = oracle.flat_run(A, dim, x0 * (1 + sign * 2.2e-16), oracle.Params(iterations=k, **okw))
                tol = max(1e-12, 50 * np.abs(pert - ref).max())
                x = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=k, **kw))
                assert np.abs(x - ref).max() < tol, (dim, k, kw, np.abs(x - ref).max(), tol)


@pytest.mark.parametrize("long_row", ["512", "48"])
@pytest.mark.parametrize("dim", [2, 3])
def test_power_law_rows_one_cta_per_long_row(ctx, capi, oracle, graphs, long_row, dim, monkeypatch):
    """Power-law graphs: rows longer than GE_LONG_ROW entries are handled by one CTA each
    (k_attract_step_long), the rest by the row kernels; forces and 2-iteration positions against
    the oracle, weighted and unweighted, with and without the breadth-first renumbering."""
    monkeypatch.setenv("GE_LONG_ROW", long_row)
    monkeypatch.setenv("GE_ONCHIP_MAX", "0")
    A = graphs.rmat(13, 16, seed=4)
    n = A.shape[0]
    assert np.diff(A.indptr).max() > 512
    B = A.copy()
    B.data = np.random.default_rng(1).uniform(0.2, 3.0, B.nnz)
    B = graphs.canonical((B + B.T) * 0.5)
    x0 = capi.reference_uniform(21, n * dim).reshape(n, dim)
    for M in (A, B):
        F_ref, S = oracle.flat_forces(M, dim, x0)
        F = ctx.flat_forces(M, dim, x0, capi.flat_params(), path=1)
        assert force_error(F, F_ref, S).max() < TOL_F64
        ref, _ = oracle.flat_run(M, dim, x0, oracle.Params(iterations=2))
        got = ctx.flat_forceatlas(M, dim, x0, capi.flat_params(iterations=2))
        assert np.abs(got - ref).max() < 1e-9 * np.abs(ref).max()
    if dim == 3 and long_row == "512":   # a plan large enough to be renumbered (>= 65536 vertices)
        C = graphs.rmat(17, 16, seed=6)
        m = C.shape[0]
        assert m >= 65536
        y0 = capi.reference_uniform(22, m * dim).reshape(m, dim)
        plan = ctx.flat_plan(C, dim, capi.flat_params(iterations=100))
        plan.upload(y0)
        plan.iterate(1)
        F = plan.download_forces()
        plan.close()
        deg = np.diff(C.indptr)
        rows = list(np.argsort(deg)[-4:]) + list(np.random.default_rng(0).choice(m, 12, replace=False))
        for r in rows:
            F_ref, S = oracle.flat_forces(C, dim, y0, rows=(int(r), int(r) + 1))
            assert np.linalg.norm(F[r] - F_ref[r]) / S[r] < TOL_F64, (r, int(deg[r]))


def test_internal_renumbering_is_transparent(ctx, capi, oracle, graphs, monkeypatch):
    """Single-rank plans on large graphs renumber the vertices breadth-first for gather locality;
    forces and positions come back in the caller's numbering and agree with the un-renumbered
    plan (summation order inside a row changes, hence the tolerance) and with the oracle."""
    A = graphs.rgg(70_000, 10.0, seed=3)
    n = A.shape[0]
    x0 = capi.reference_uniform(13, n * 3).reshape(n, 3)
    p = capi.flat_params()
    out = {}
    for tag, env in (("bfs", None), ("plain", "1")):
        if env:
            monkeypatch.setenv("GE_NO_REORDER", env)
        else:
            monkeypatch.delenv("GE_NO_REORDER", raising=False)
        plan = ctx.flat_plan(A, 3, p)
        plan.upload(x0)
        plan.iterate(2)
        out[tag] = (plan.download(), plan.download_forces())
        plan.close()
    assert np.abs(out["bfs"][0] - out["plain"][0]).max() < 1e-11
    scale = np.linalg.norm(out["plain"][1], axis=1).max()
    assert np.abs(out["bfs"][1] - out["plain"][1]).max() < 1e-10 * scale
    rows = np.random.default_rng(1).choice(n, 16, replace=False)
    F = ctx.flat_forces(A, 3, x0, p, path=1)
    for r in rows:
        F_ref, S = oracle.flat_forces(A, 3, x0, rows=(int(r), int(r) + 1))
        assert np.linalg.norm(F[r] - F_ref[r]) / S[r] < TOL_F64


def test_graph_replay_matches_eager_launches(ctx, capi, graphs, monkeypatch):
    """Long flat solves on mid-size graphs replay a captured 2-iteration CUDA graph; the result is
    bit-identical to launching every kernel eagerly (same kernels, same order)."""
    A = graphs.rgg(3000, 10.0, seed=8)
    n = A.shape[0]
    x0 = capi.reference_uniform(2, n * 2).reshape(n, 2)
    monkeypatch.setenv("GE_ONCHIP_MAX", "0")
    for iters in (64, 101):
        monkeypatch.delenv("GE_NO_GRAPH", raising=False)
        l0 = ctx.launches
        xg = ctx.flat_forceatlas(A, 2, x0, capi.flat_params(iterations=iters))
        launched = ctx.launches - l0
        monkeypatch.setenv("GE_NO_GRAPH", "1")
        xe = ctx.flat_forceatlas(A, 2, x0, capi.flat_params(iterations=iters))
        assert np.array_equal(xg, xe)
        assert launched >= 3 * iters  # repulsion + fix-up + attraction/step per iteration are counted
