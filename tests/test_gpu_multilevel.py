"""GPU parity of the per-aggregate solver (K2: singleton / warp / CTA / segmented tiers), the
prolongation epilogue and the whole embed() driver, through the C ABI."""
import numpy as np
import pytest

from helpers import TOL_F32, TOL_F64, force_error, layout_stats, load_hier_golden

pytestmark = pytest.mark.gpu


def _case(graphs, name):
    if name == "grid30":            # the reference partitioner's own hierarchy (sizes <= 8)
        As, Ps, _ = load_hier_golden()
        return As, Ps
    if name == "rgg_big_aggs":      # coarsening 0.02: aggregates of 30..150 members -> CTA tier
        return graphs.coarsen(graphs.rgg(4000, 10.0, seed=3), 0.02, min_coarse=20)
    if name == "rmat":              # skewed sizes
        return graphs.coarsen(graphs.rmat(12, 8, seed=2), 0.25, min_coarse=40)
    raise KeyError(name)


@pytest.mark.parametrize("name", ["grid30", "rgg_big_aggs", "rmat"])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("cta_max", [512, 40])
def test_forces_fp64(ctx, capi, oracle, graphs, name, dim, cta_max, monkeypatch):
    """cta_max=40 pushes every aggregate above 40 members through the segmented multi-CTA tier."""
    monkeypatch.setenv("GE_CTA_MAX", str(cta_max))
    As, Ps = _case(graphs, name)
    rng = np.random.default_rng(1)
    for l in range(min(2, len(Ps))):
        A, P = As[l], Ps[l]
        n, m = A.shape[0], P.shape[0]
        cA = rng.normal(size=(m, dim))
        x = capi.reference_uniform(3 + l, n * dim).reshape(n, dim)
        _, F_ref, S = oracle.multilevel_run(A, P, cA, np.ones(m), dim, x, oracle.Params(iterations=1), forces_iter=0)
        F = ctx.multilevel_forces(A, P, cA, x, dim, capi.multilevel_params())
        err = force_error(F, F_ref, S)
        assert err.max() < TOL_F64, (l, err.max(), np.argmax(err))


def test_forces_fp32(ctx, capi, oracle, graphs):
    As, Ps = _case(graphs, "rgg_big_aggs")
    A, P = As[0], Ps[0]
    n, m = A.shape[0], P.shape[0]
    cA = np.random.default_rng(1).normal(size=(m, 2))
    x = capi.reference_uniform(3, n * 2).reshape(n, 2)
    _, F_ref, S = oracle.multilevel_run(A, P, cA, np.ones(m), 2, x, oracle.Params(iterations=1), forces_iter=0)
    F = ctx.multilevel_forces(A, P, cA, x, 2, capi.multilevel_params(precision=capi.GE_F32))
    assert force_error(F, F_ref, S).max() < TOL_F32


def test_quirk_q1_edge_dropped(ctx, capi, oracle):
    """include/forceatlas.hpp:417 compares the global neighbour id with the LOCAL member index:
    with aggregate {1,0,2} (member order 1,0,2) vertex 1 (local 0) loses its edge to global 0."""
    import scipy.sparse as sp
    A = sp.csr_matrix(np.array([[0, 1, 1, 0], [1, 0, 1, 1], [1, 1, 0, 0], [0, 1, 0, 0]], dtype=float))
    P = sp.csr_matrix((np.ones(4), [1, 0, 2, 3], [0, 3, 4]), shape=(2, 4))
    cA = np.array([[0.0, 0.0], [2.0, 1.0]])
    x = capi.reference_uniform(4, 8).reshape(4, 2)
    _, F_ref, S = oracle.multilevel_run(A, P, cA, np.ones(2), 2, x, oracle.Params(iterations=1), forces_iter=0)
    F = ctx.multilevel_forces(A, P, cA, x, 2, capi.multilevel_params())
    assert force_error(F, F_ref, S).max() < TOL_F64


@pytest.mark.parametrize("name", ["grid30", "rgg_big_aggs", "rmat"])
@pytest.mark.parametrize("k", [1, 3])
@pytest.mark.parametrize("cta_max", [512, 40])
def test_positions_after_k_iterations(ctx, capi, oracle, graphs, name, k, cta_max, monkeypatch):
    """k iterations + centre/normalise/prolongation from the SAME initial local coordinates."""
    monkeypatch.setenv("GE_CTA_MAX", str(cta_max))
    As, Ps = _case(graphs, name)
    A, P = As[0], Ps[0]
    n, m = A.shape[0], P.shape[0]
    rng = np.random.default_rng(2)
    cA, rA = rng.normal(size=(m, 2)), rng.random(m) * 0.3 + 0.05
    init = oracle.multilevel_init(P, 2, 5)
    ref = oracle.multilevel_run(A, P, cA, rA, 2, init, oracle.Params(iterations=k))
    x = ctx.multilevel_forceatlas(A, P, cA, rA, 2, capi.multilevel_params(iterations=k), init=init)
    assert np.abs(x - ref).max() < 1e-10, np.abs(x - ref).max()


@pytest.mark.parametrize("name", ["rgg_big_aggs", "rmat"])
@pytest.mark.parametrize("dim", [2, 3])
def test_long_rows_of_the_segmented_tier(ctx, capi, oracle, graphs, name, dim, monkeypatch):
    """Rows of the large-aggregate tier with many neighbours inside their aggregate (hubs of a
    power-law level) are listed by the prep kernel and walked by one CTA each
    (k_attract_step_long<.., ML>); GE_ML_LONG_ROW=4 sends every row with more than 4 internal
    entries that way.  Forces and 3-iteration positions against the oracle, and the same
    positions to rounding as without the tier (only the order of a row's partial sums changes)."""
    monkeypatch.setenv("GE_CTA_MAX", "40")
    As, Ps = _case(graphs, name)
    A, P = As[0], Ps[0]
    n, m = A.shape[0], P.shape[0]
    rng = np.random.default_rng(7)
    cA, rA = rng.normal(size=(m, dim)), rng.random(m) * 0.3 + 0.05
    x = capi.reference_uniform(4, n * dim).reshape(n, dim)
    init = oracle.multilevel_init(P, dim, 5)
    _, F_ref, S = oracle.multilevel_run(A, P, cA, np.ones(m), dim, x, oracle.Params(iterations=1), forces_iter=0)
    ref = oracle.multilevel_run(A, P, cA, rA, dim, init, oracle.Params(iterations=3))
    out = {}
    for long_row in ("4", "0"):
        monkeypatch.setenv("GE_ML_LONG_ROW", long_row)
        F = ctx.multilevel_forces(A, P, cA, x, dim, capi.multilevel_params())
        assert force_error(F, F_ref, S).max() < TOL_F64, long_row
        out[long_row] = ctx.multilevel_forceatlas(A, P, cA, rA, dim, capi.multilevel_params(iterations=3), init=init)
        assert np.abs(out[long_row] - ref).max() < 1e-10, long_row
    assert np.abs(out["4"] - out["0"]).max() < 1e-12


def test_seeded_init_matches_reference_stream(ctx, capi, oracle):
    """init=NULL draws the reference's stream (forceatlas.hpp:341,356-358) from params.seed: the
    result equals the golden output of the COMPILED REFERENCE for the same seed (k=1,3)."""
    As, Ps, z = load_hier_golden()
    for dim in (2, 3):
        for l in (0, 1):
            for k in (1, 3):
                cA, rA = z["ml_cA_l%d_d%d" % (l, dim)], z["ml_rA_l%d_d%d" % (l, dim)]
                x = ctx.multilevel_forceatlas(As[l], Ps[l], cA, rA, dim, capi.multilevel_params(iterations=k, seed=5))
                assert np.abs(x - z["ml_x_l%d_d%d_k%d" % (l, dim, k)]).max() < 1e-10


@pytest.mark.parametrize("name", ["grid30", "rgg_big_aggs", "rmat"])
def test_prolongation_properties_100_iterations(ctx, capi, oracle, graphs, name):
    """Exact properties of forceatlas.hpp:539-569 after the full 100 iterations: every member lies in
    its parent ball, the farthest member of a multi-member aggregate lies ON it, the members'
    centroid is the parent centre, singletons sit exactly on it; and the layout statistics agree
    with the oracle's run from the same initial coordinates."""
    As, Ps = _case(graphs, name)
    A, P = As[0], Ps[0]
    n, m = A.shape[0], P.shape[0]
    rng = np.random.default_rng(2)
    cA, rA = rng.normal(size=(m, 2)) * 5, rng.random(m) * 0.3 + 0.05
    init = oracle.multilevel_init(P, 2, 5)
    x = ctx.multilevel_forceatlas(A, P, cA, rA, 2, capi.multilevel_params(), init=init)
    assert np.isfinite(x).all()
    v_A = capi.vertex_to_aggregate(P)
    d = np.linalg.norm(x - cA[v_A], axis=1)
    assert (d <= rA[v_A] * (1 + 1e-12)).all()
    size = np.diff(P.indptr)
    far = np.zeros(m)
    np.maximum.at(far, v_A, d)
    multi = size > 1
    assert np.allclose(far[multi], rA[multi], rtol=1e-12)
    assert np.array_equal(x[P.indices[P.indptr[:-1][~multi]]], cA[~multi])
    cent = np.zeros((m, 2))
    np.add.at(cent, v_A, x)
    assert np.abs(cent / size[:, None] - cA).max() < 1e-9
    ref = oracle.multilevel_run(A, P, cA, rA, 2, init, oracle.Params(iterations=100))
    s1, s2 = layout_stats(A, x), layout_stats(A, ref)
    for key in s1:
        assert abs(s1[key] - s2[key]) < 0.05 * abs(s2[key]), (key, s1, s2)


def test_embed_against_reference_golden(ctx, capi, oracle):
    """partition::embed on the reference partitioner's hierarchy (grid 30x30, 3 levels, 100 000
    coarsest iterations).  Trajectories are chaotic, so the layouts are compared on statistics
    (tolerance 20 %) against the golden output of the compiled reference."""
    As, Ps, z = load_hier_golden()
    x, st = ctx.embed(As, Ps, 2, seed=21)
    assert np.isfinite(x).all() and st["kernel_launches"] > 0
    assert st["pair_interactions"] == pytest.approx(
        100000 * 16 * 15 + 100 * sum(int((np.diff(P.indptr) * (np.diff(P.indptr) - 1)).sum()) for P in Ps))
    s1, s2 = layout_stats(As[0], x), layout_stats(As[0], z["embed_d2_seed21"])
    for key in s1:
        assert abs(s1[key] - s2[key]) < 0.2 * abs(s2[key]), (key, s1, s2)


def test_embed_short_run_matches_oracle(ctx, capi, oracle):
    """With few iterations per level the whole driver (flat init stream, radii, rescale, multilevel
    init stream, prolongation) is comparable position by position with the oracle's embed()."""
    As, Ps, _ = load_hier_golden()
    for dim in (2, 3):
        ref = oracle.embed(As, Ps, dim, seed=9, coarse_iterations=20, level_iterations=3)
        x, _ = ctx.embed(As, Ps, dim, seed=9, coarse_iterations=20, level_iterations=3)
        assert np.abs(x - ref).max() < 1e-8 * np.abs(ref).max(), np.abs(x - ref).max()


def test_config2_full_size_properties(ctx, capi, graphs):
    """BASELINE config 2 (RGG, n = 100 000, avg degree 10, d = 2, coarsening 0.25): NaN-free and
    every vertex inside the ball of its aggregate at the finest level."""
    A = graphs.rgg(100_000, 10.0, seed=12345)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=100)
    x, st = ctx.embed(As, Ps, 2, seed=1)
    assert np.isfinite(x).all()
    assert x.shape == (A.shape[0], 2)
    v_A = capi.vertex_to_aggregate(Ps[0])
    # aggregates are compact relative to the layout: mean member-to-centroid distance is small
    cent = np.zeros((Ps[0].shape[0], 2))
    np.add.at(cent, v_A, x)
    cent /= np.diff(Ps[0].indptr)[:, None]
    spread = np.linalg.norm(x - cent[v_A], axis=1).mean()
    extent = np.linalg.norm(x - x.mean(0), axis=1).max()
    assert spread < 0.05 * extent


def test_sharded_aggregates_equal_unsharded(ctx, capi, oracle, graphs):
    """ge_multilevel_forceatlas_shard over two aggregate ranges, summed, is bit-identical to the
    unsharded call (aggregates are independent, include/forceatlas.hpp:340-341); and the sharded
    embed driver equals ge_embed for the same seed."""
    from graph_embed_b200 import sharding
    As, Ps = _case(graphs, "rmat")
    A, P = As[0], Ps[0]
    m = P.shape[0]
    rng = np.random.default_rng(2)
    cA, rA = rng.normal(size=(m, 2)), rng.random(m) * 0.3 + 0.05
    init = oracle.multilevel_init(P, 2, 5)
    p = capi.multilevel_params(iterations=20)
    full = ctx.multilevel_forceatlas(A, P, cA, rA, 2, p, init=init)
    blocks = sharding.aggregate_blocks(A, P, 3)
    parts = [ctx.multilevel_forceatlas(A, P, cA, rA, 2, p, init=init, aggregates=b) for b in blocks]
    assert np.array_equal(sum(parts), full)
    v_A = capi.vertex_to_aggregate(P)
    for b, part in zip(blocks, parts):
        foreign = (v_A < b[0]) | (v_A >= b[1])
        assert not part[foreign].any()

    class OneRank:  # the driver's collective is a no-op at world size 1
        @staticmethod
        def all_reduce(t):
            return t
    x1 = sharding.embed_sharded(ctx, OneRank, As, Ps, 2, seed=7, rank=0, world=1, coarse_iterations=200)
    x2, _ = ctx.embed(As, Ps, 2, seed=7, coarse_iterations=200)
    assert np.array_equal(x1, x2)


def test_config1_embed_against_reference(ctx, capi, graphs):
    """BASELINE config 1 at full size: 100 x 100 grid, the hierarchy the reference's partitioner
    builds (10000 -> 1270 -> 209 -> 43 -> 34), d = 2, 100 000 coarsest iterations.  Layout statistics
    against the compiled reference's own embed() output (golden)."""
    from helpers import load_config1_golden
    As, Ps, z = load_config1_golden(graphs)
    assert [A.shape[0] for A in As] == [10000, 1270, 209, 43, 34]
    x, st = ctx.embed(As, Ps, 2, seed=3)
    assert np.isfinite(x).all()
    s1, s2 = layout_stats(As[0], x), layout_stats(As[0], z["embed_d2_seed3"].astype(np.float64))
    for key in s1:
        assert abs(s1[key] - s2[key]) < 0.25 * abs(s2[key]), (key, s1, s2)


def _embed_properties(ctx, capi, As, Ps, dim):
    x, st = ctx.embed(As, Ps, dim, seed=1)
    assert np.isfinite(x).all() and x.shape == (As[0].shape[0], dim)
    v_A = capi.vertex_to_aggregate(Ps[0])
    cent = np.zeros((Ps[0].shape[0], dim))
    np.add.at(cent, v_A, x)
    cent /= np.diff(Ps[0].indptr)[:, None]
    spread = np.linalg.norm(x - cent[v_A], axis=1).mean()
    extent = np.linalg.norm(x - x.mean(0), axis=1).max()
    assert spread < 0.05 * extent
    pairs = 100000.0 * As[-1].shape[0] * (As[-1].shape[0] - 1) + 100.0 * sum(
        float((np.diff(P.indptr).astype(np.int64) * (np.diff(P.indptr) - 1)).sum()) for P in Ps)
    assert st["pair_interactions"] == pytest.approx(pairs)
    return x


@pytest.mark.parametrize("name", ["rmat16_d3", "delaunay200k_d3"])
def test_config3_config5_shapes_reduced(ctx, capi, graphs, name):
    """BASELINE configs 3 (R-MAT, d = 3, coarsening 0.25) and 5 (3-D Delaunay mesh, d = 3,
    coarsening 0.125) at reduced size on the stand-in generator; the reference partitioner's own
    hierarchies for both configs run in tests/test_gpu_refhier.py."""
    if name == "rmat16_d3":
        As, Ps = graphs.coarsen(graphs.rmat(16, 16, seed=1), 0.25, min_coarse=64)
    else:
        As, Ps = graphs.coarsen(graphs.delaunay3d(200_000, seed=1), 0.125, min_coarse=64)
    _embed_properties(ctx, capi, As, Ps, 3)


def test_unseeded_mode_draws_on_device(ctx, capi, graphs):
    """seed = 0 is the reference's std::random_device mode: two runs differ, both are valid
    layouts (finite, every member inside its parent ball)."""
    As, Ps = _case(graphs, "rgg_big_aggs")
    A, P = As[0], Ps[0]
    m = P.shape[0]
    rng = np.random.default_rng(2)
    cA, rA = rng.normal(size=(m, 2)) * 5, rng.random(m) * 0.3 + 0.05
    x1 = ctx.multilevel_forceatlas(A, P, cA, rA, 2, capi.multilevel_params(seed=0))
    x2 = ctx.multilevel_forceatlas(A, P, cA, rA, 2, capi.multilevel_params(seed=0))
    v_A = capi.vertex_to_aggregate(P)
    for x in (x1, x2):
        assert np.isfinite(x).all()
        assert (np.linalg.norm(x - cA[v_A], axis=1) <= rA[v_A] * (1 + 1e-12)).all()
    assert not np.array_equal(x1, x2)
    e1, _ = ctx.embed(As, Ps, 2, seed=0, coarse_iterations=500)
    assert np.isfinite(e1).all()


def test_staged_uploads_equal_direct_copies(ctx, capi, graphs, monkeypatch):
    """Large host->device copies go through a multi-threaded pinned staging ring; the results are
    bit-identical to plain cudaMemcpyAsync from pageable memory (1.2M-vertex level: the CSR arrays
    are 20-100 MB each, well past the 16 MB staging threshold)."""
    A = graphs.rgg(1_200_000, 10.0, seed=2)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=64, max_levels=1)
    m = Ps[0].shape[0]
    rng = np.random.default_rng(0)
    cA, rA = rng.normal(size=(m, 2)), rng.random(m) + 0.1
    p = capi.multilevel_params(iterations=3, seed=4)
    monkeypatch.delenv("GE_NO_STAGING", raising=False)
    x1 = ctx.multilevel_forceatlas(As[0], Ps[0], cA, rA, 2, p)
    monkeypatch.setenv("GE_NO_STAGING", "1")
    x2 = ctx.multilevel_forceatlas(As[0], Ps[0], cA, rA, 2, p)
    assert np.isfinite(x1).all() and np.array_equal(x1, x2)


def test_embed_fp32_option(ctx, capi, graphs):
    """precision = GE_F32 runs every kernel family in single precision: same exact prolongation
    properties (the epilogue is evaluated in double), statistics close to the FP64 layout."""
    As, Ps = graphs.coarsen(graphs.rgg(6000, 10.0, seed=4), 0.25, min_coarse=50)
    x32, _ = ctx.embed(As, Ps, 2, seed=3, precision=capi.GE_F32)
    x64, _ = ctx.embed(As, Ps, 2, seed=3, precision=capi.GE_F64)
    assert np.isfinite(x32).all()
    s32, s64 = layout_stats(As[0], x32), layout_stats(As[0], x64)
    for key in s32:
        assert abs(s32[key] - s64[key]) < 0.25 * abs(s64[key]), (key, s32, s64)


@pytest.mark.parametrize("kw", [dict(use_weights=0), dict(linlog=1), dict(nohubs=1), dict(delta=0.5), dict(delta=0.0),
                                dict(ks=0.3, ksmax=2.0, repel=2.0, attract=0.7, gravity=1.5, tolerate=0.8)])
@pytest.mark.parametrize("cta_max", [512, 40])
def test_multilevel_options(ctx, capi, oracle, graphs, kw, cta_max, monkeypatch):
    """Non-default forceAtlasMultilevel arguments (include/forceatlas.hpp:320-331) on a weighted
    Galerkin level with self-loops, through every tier: forces and 2-iteration positions."""
    monkeypatch.setenv("GE_CTA_MAX", str(cta_max))
    As, Ps = graphs.coarsen(graphs.rgg(6000, 10.0, seed=6), 0.1, min_coarse=10)
    A, P = As[1], Ps[1]   # level 1: weighted edges + diagonal entries
    n, m = A.shape[0], P.shape[0]
    rng = np.random.default_rng(3)
    cA, rA = rng.normal(size=(m, 2)), rng.random(m) * 0.3 + 0.05
    okw = {"useWeights" if k == "use_weights" else k: v for k, v in kw.items()}
    x = capi.reference_uniform(8, n * 2).reshape(n, 2)
    _, F_ref, S = oracle.multilevel_run(A, P, cA, np.ones(m), 2, x, oracle.Params(iterations=1, **okw), forces_iter=0)
    F = ctx.multilevel_forces(A, P, cA, x, 2, capi.multilevel_params(**kw))
    assert force_error(F, F_ref, S).max() < TOL_F64
    init = oracle.multilevel_init(P, 2, 5)
    ref = oracle.multilevel_run(A, P, cA, rA, 2, init, oracle.Params(iterations=2, **okw))
    got = ctx.multilevel_forceatlas(A, P, cA, rA, 2, capi.multilevel_params(iterations=2, **kw), init=init)
    assert np.abs(got - ref).max() < 1e-10
