"""Edge cases the reference code paths admit: tiny graphs, isolated vertices, tier boundaries of the
per-aggregate solver (32 / 33 members, > 512 members -> segmented multi-CTA tier), all-singleton
levels, unsupported dimensions."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import TOL_F64, force_error

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", [1, 2])
def test_single_vertex_and_isolated_vertices(ctx, capi, oracle, path):
    A1 = sp.csr_matrix((1, 1))
    x = np.array([[0.3, -0.2]])
    F_ref, S = oracle.flat_forces(A1, 2, x)
    assert force_error(ctx.flat_forces(A1, 2, x, capi.flat_params(), path=path), F_ref, S).max() < TOL_F64
    # a triangle plus two isolated vertices (degree 0 -> mass 1, no attraction)
    A = sp.csr_matrix(np.array([[0, 1, 1, 0, 0], [1, 0, 1, 0, 0], [1, 1, 0, 0, 0], [0] * 5, [0] * 5], dtype=float))
    x = capi.reference_uniform(3, 10).reshape(5, 2)
    F_ref, S = oracle.flat_forces(A, 2, x)
    assert force_error(ctx.flat_forces(A, 2, x, capi.flat_params(), path=path), F_ref, S).max() < TOL_F64
    xr, _ = oracle.flat_run(A, 2, x, oracle.Params(iterations=3))
    import os
    os.environ["GE_ONCHIP_MAX"] = "0" if path == 1 else "1024"
    try:
        xg = ctx.flat_forceatlas(A, 2, x, capi.flat_params(iterations=3))
    finally:
        os.environ.pop("GE_ONCHIP_MAX")
    assert np.abs(xg - xr).max() < 1e-11


def _custom_partition(n, sizes):
    """P_T with consecutive aggregates of the given sizes (the rest as one last aggregate)."""
    bounds = np.concatenate([[0], np.cumsum(sizes)])
    assert bounds[-1] <= n
    if bounds[-1] < n:
        bounds = np.concatenate([bounds, [n]])
    m = len(bounds) - 1
    perm = np.random.default_rng(0).permutation(n).astype(np.int32)  # members in non-sorted order
    return sp.csr_matrix((np.ones(n), perm, bounds.astype(np.int32)), shape=(m, n))


@pytest.mark.parametrize("sizes", [[32, 33, 1, 2, 31], [600, 5, 5, 40], [1] * 50])
@pytest.mark.parametrize("dim", [2, 3])
def test_tier_boundaries(ctx, capi, oracle, graphs, sizes, dim):
    """Aggregates of 32 and 33 members straddle the warp/CTA tiers; 600 members exceed one CTA and go
    through the segmented tiled kernels; an all-singleton prefix exercises the closed form."""
    A = graphs.grid2d(30, 30)
    n = A.shape[0]
    P = _custom_partition(n, sizes)
    m = P.shape[0]
    rng = np.random.default_rng(1)
    cA, rA = rng.normal(size=(m, dim)), rng.random(m) * 0.3 + 0.05
    x = capi.reference_uniform(4, n * dim).reshape(n, dim)
    _, F_ref, S = oracle.multilevel_run(A, P, cA, np.ones(m), dim, x, oracle.Params(iterations=1), forces_iter=0)
    F = ctx.multilevel_forces(A, P, cA, x, dim, capi.multilevel_params())
    assert force_error(F, F_ref, S).max() < TOL_F64
    init = oracle.multilevel_init(P, dim, 6)
    ref = oracle.multilevel_run(A, P, cA, rA, dim, init, oracle.Params(iterations=2))
    got = ctx.multilevel_forceatlas(A, P, cA, rA, dim, capi.multilevel_params(iterations=2), init=init)
    assert np.abs(got - ref).max() < 1e-10


def test_unsupported_dimension_is_an_error(ctx, capi, graphs):
    A = graphs.grid2d(4, 4)
    with pytest.raises(capi.GeError) as e:
        ctx.flat_forces(A, 5, np.zeros((16, 5)), capi.flat_params(), path=1)
    assert e.value.status == capi.GE_ERR_INVALID
    with pytest.raises(capi.GeError):
        ctx.flat_forceatlas(A, 1, np.ones((16, 1)), capi.flat_params(iterations=1))


def test_shape_mismatch_is_an_error(ctx, capi, graphs):
    """The asserts of src/embed.cpp:564-570 become GE_ERR_INVALID."""
    As, Ps = graphs.coarsen(graphs.grid2d(10, 10), 0.25, min_coarse=8)
    with pytest.raises(capi.GeError) as e:
        ctx.embed(As[:-1] + [As[0]], Ps, 2, seed=1)
    assert e.value.status == capi.GE_ERR_INVALID
