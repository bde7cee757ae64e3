"""One process driving several B200s behind the C ABI (ge_context_create_multi): the sharded
partition::forceAtlas (include/forceatlas.hpp:89-305) against the single-GPU plan and the oracle.
Needs >= 2 GPUs on the box (skipped on a single-GPU box; `gpurun --gpus 2` runs it)."""
import numpy as np
import pytest

from helpers import TOL_F64, force_error

pytestmark = pytest.mark.gpu


def _multi(capi, ndev):
    try:
        return capi.Context(devices=list(range(ndev)))
    except capi.GeError as e:
        if e.status == capi.GE_ERR_INVALID and "more devices" in str(e):
            pytest.skip("box has fewer than %d GPUs" % ndev)
        raise


@pytest.mark.parametrize("ndev", [2, 4, 8])
@pytest.mark.parametrize("dim", [2, 3])
def test_sharded_forceatlas_matches_single_gpu_and_oracle(ctx, capi, oracle, graphs, ndev, dim):
    mc = _multi(capi, ndev)
    assert mc.device_count == ndev
    A = graphs.rgg(40000, 10.0, seed=3)
    n = A.shape[0]
    x0 = capi.reference_uniform(5, n * dim).reshape(n, dim)
    one = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=1))
    got = mc.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=1))
    # first step against the oracle: x1 = x0 + F * speed on sampled rows
    rows = np.random.default_rng(0).choice(n, 16, replace=False)
    p = oracle.Params()
    for r in rows:
        F, S = oracle.flat_forces(A, dim, x0, p, rows=(int(r), int(r) + 1))
        f = F[r]
        fn = float(np.sqrt((f * f).sum()))
        speed = min(p.ks / (1.0 + np.sqrt(fn)), p.ksmax / fn)
        assert np.linalg.norm(got[r] - (x0[r] + f * speed)) <= TOL_F64 * speed * S[r]
    assert np.abs(got - one).max() <= 1e-11 * np.abs(one).max()
    # a few more iterations: the map is chaotic (the first steps from a random start move every
    # vertex by the cap), so rounding-order differences grow; the bulk of the vertices stays together
    one = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=4))
    got = mc.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=4))
    assert np.isfinite(got).all()
    assert np.median(np.abs(got - one)) <= 1e-9 * np.abs(one).max()
    mc.close()


def test_small_graphs_and_other_entry_points_run_on_first_device(capi, graphs):
    mc = _multi(capi, 2)
    A = graphs.rgg(3000, 10.0, seed=1)
    x0 = capi.reference_uniform(2, A.shape[0] * 2).reshape(-1, 2)
    single = capi.Context(0)
    a = mc.flat_forceatlas(A, 2, x0, capi.flat_params(iterations=3))
    b = single.flat_forceatlas(A, 2, x0, capi.flat_params(iterations=3))
    assert np.array_equal(a, b)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=30)
    e1, _ = mc.embed(As, Ps, 2, seed=3, coarse_iterations=500)
    e2, _ = single.embed(As, Ps, 2, seed=3, coarse_iterations=500)
    assert np.array_equal(e1, e2)
    mc.close()
    single.close()


@pytest.mark.parametrize("ndev", [2, 4])
def test_embed_sharded_by_aggregates_equals_single_gpu(capi, graphs, ndev, monkeypatch):
    """partition::embed on a multi-device context: levels are cut into cost-balanced aggregate
    ranges, one per device, and summed (foreign rows are exact zeros).  Aggregates of up to 512
    members are solved by one warp / CTA each whatever device they land on: bit-identical to the
    single-GPU call for the same seed."""
    mc = _multi(capi, ndev)
    monkeypatch.setenv("GE_SHARD_MIN_MPAIRS", "0")   # shard every level, however small
    As, Ps = graphs.coarsen(graphs.rgg(30000, 10.0, seed=8), 0.1, min_coarse=40)
    assert max(int(np.diff(P.indptr).max()) for P in Ps) <= 512
    single = capi.Context(0)
    a, sa, ra, ca = mc.embed(As, Ps, 2, seed=5, coarse_iterations=2000, return_level1=True)
    b, sb, rb, cb = single.embed(As, Ps, 2, seed=5, coarse_iterations=2000, return_level1=True)
    assert np.isfinite(a).all()
    assert np.array_equal(a, b) and np.array_equal(ra, rb) and np.array_equal(ca, cb)
    assert sa["pair_interactions"] == pytest.approx(sb["pair_interactions"])
    mc.close()
    single.close()


def test_embed_sharded_large_aggregates_properties(capi, graphs, monkeypatch):
    """The reference partitioner's Delaunay hierarchy (aggregates of up to 13 689 members, the
    multi-CTA tier) on 2 GPUs: exact prolongation properties on every aggregate."""
    from helpers import load_ref_hierarchy
    mc = _multi(capi, 2)
    As, Ps, _ = load_ref_hierarchy(graphs, "delaunay1000000")
    x, st, r1, c1 = mc.embed(As, Ps, 3, seed=1, return_level1=True)
    assert np.isfinite(x).all()
    P = Ps[0]
    sizes = np.diff(P.indptr).astype(np.int64)
    v_A = capi.vertex_to_aggregate(P)
    ulp = 8 * np.finfo(float).eps * np.abs(x).max()
    dist = np.linalg.norm(x - c1[v_A], axis=1)
    assert (dist <= r1[v_A] * (1 + 1e-12) + ulp).all()
    far = np.zeros(P.shape[0])
    np.maximum.at(far, v_A, dist)
    assert np.allclose(far[sizes >= 2], r1[sizes >= 2], rtol=1e-9, atol=ulp)
    pairs = 100000.0 * As[-1].shape[0] * (As[-1].shape[0] - 1) + 100.0 * sum(
        float((np.diff(Q.indptr).astype(np.int64) * (np.diff(Q.indptr) - 1)).sum()) for Q in Ps)
    assert st["pair_interactions"] == pytest.approx(pairs)
    mc.close()
