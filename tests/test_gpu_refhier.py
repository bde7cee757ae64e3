"""BASELINE configs 3 and 5 on hierarchies produced by the REFERENCE's own partitioner
(src/partitioner.cpp:1550-1893, called as at examples/embedder.cpp:187; cached by
tests/golden/make_ref_hierarchy.py).  These hierarchies have what the stand-in generator
(graphs.coarsen) does not: aggregates of thousands of members (R-MAT-20: 3792 at level 0;
Delaunay: 13 689 at 1M points), i.e. the multi-CTA tier of the per-aggregate solver
(include/forceatlas.hpp:340-341, 394-410) at its real size.

Per level: the forces of one iteration on sampled aggregates of every size class -- always
including the largest -- against the oracle, through the C ABI; whole embed(): the exact properties
of the prolongation (:539-569) on every aggregate and the pair count of the hierarchy."""
import numpy as np
import pytest

from helpers import TOL_F64, force_error, load_ref_hierarchy

pytestmark = pytest.mark.gpu

CASES = ["rmat20", "delaunay1000000"]


def _sample_aggregates(sizes, rng):
    """The two largest aggregates plus up to two random ones from each size class."""
    picks = list(np.argsort(sizes)[-2:])
    for lo, hi in ((1, 1), (2, 32), (33, 512), (513, 1 << 30)):
        cand = np.flatnonzero((sizes >= lo) & (sizes <= hi))
        if cand.size:
            picks += list(rng.choice(cand, size=min(2, cand.size), replace=False))
    return sorted(set(int(a) for a in picks))


@pytest.mark.parametrize("name", CASES)
def test_forces_levels_0_1_sampled_aggregates(ctx, capi, oracle, graphs, name):
    As, Ps, _ = load_ref_hierarchy(graphs, name)
    dim = 3
    rng = np.random.default_rng(5)
    for l in (0, 1):
        A, P = As[l], Ps[l]
        n, m = A.shape[0], P.shape[0]
        sizes = np.diff(P.indptr)
        cA = rng.normal(size=(m, dim))
        x = capi.reference_uniform(11 + l, n * dim).reshape(n, dim)
        F = ctx.multilevel_forces(A, P, cA, x, dim, capi.multilevel_params())
        assert np.isfinite(F).all()
        worst = 0.0
        for a in _sample_aggregates(sizes, rng):
            _, F_ref, S = oracle.multilevel_run(A, P, cA, np.ones(m), dim, x, oracle.Params(iterations=1),
                                                forces_iter=0, aggregates=(a, a + 1))
            mem = P.indices[P.indptr[a]:P.indptr[a + 1]]
            err = force_error(F[mem], F_ref[mem], S[mem]).max()
            worst = max(worst, err)
            assert err < TOL_F64, (name, l, a, int(sizes[a]), err)
        print("%s level %d: max aggregate %d, worst sampled force error %.2e" % (name, l, sizes.max(), worst))


@pytest.mark.parametrize("name", CASES)
def test_embed_exact_properties(ctx, capi, graphs, name):
    As, Ps, _ = load_ref_hierarchy(graphs, name)
    dim = 3
    x, st, r1, c1 = ctx.embed(As, Ps, dim, seed=1, return_level1=True)
    assert np.isfinite(x).all() and x.shape == (As[0].shape[0], dim)
    P = Ps[0]
    sizes = np.diff(P.indptr).astype(np.int64)
    v_A = capi.vertex_to_aggregate(P)
    assert sizes.max() >= (3000 if name == "rmat20" else 10000)
    # :565-569: x_i = c_a + r_a * u_i / max|u|  =>  every member inside its parent ball, the
    # farthest member ON it, and (the local coordinates were centred, :540-553) centroid == centre
    # (the check itself subtracts nearby numbers: balls deep in the hierarchy are ~1e-6 of the
    # layout's extent, so the subtraction carries an absolute error of a few ulps of |x|)
    ulp = 8 * np.finfo(float).eps * np.abs(x).max()
    dist = np.linalg.norm(x - c1[v_A], axis=1)
    assert (dist <= r1[v_A] * (1 + 1e-12) + ulp).all()
    far = np.zeros(P.shape[0])
    np.maximum.at(far, v_A, dist)
    multi = sizes >= 2
    assert np.allclose(far[multi], r1[multi], rtol=1e-9, atol=ulp)
    cent = np.zeros((P.shape[0], dim))
    np.add.at(cent, v_A, x)
    cent /= sizes[:, None]
    assert (np.linalg.norm(cent - c1, axis=1) <= 1e-9 * r1 + ulp * np.sqrt(sizes)).all()
    single = sizes == 1
    assert np.array_equal(x[P.indices[P.indptr[:-1][single]]], c1[single])
    pairs = 100000.0 * As[-1].shape[0] * (As[-1].shape[0] - 1) + 100.0 * sum(
        float((np.diff(Q.indptr).astype(np.int64) * (np.diff(Q.indptr) - 1)).sum()) for Q in Ps)
    assert st["pair_interactions"] == pytest.approx(pairs)
