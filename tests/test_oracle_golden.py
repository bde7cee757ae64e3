"""The oracle (oracle/forceatlas_oracle.c) against the golden vectors minted from the compiled
reference (tests/golden/make_golden.py).  Bitwise: the oracle keeps the reference's operation
order and both are built with -ffp-contract=off."""
import numpy as np
import pytest

from helpers import load_flat_golden, load_hier_golden


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("k", [1, 2, 5, 25, 100])
def test_flat_positions_bitwise(oracle, dim, k):
    A, z = load_flat_golden()
    x, _ = oracle.flat_run(A, dim, z["x0_d%d" % dim], oracle.Params(iterations=k))
    assert np.array_equal(x, z["x_d%d_k%d" % (dim, k)])


@pytest.mark.parametrize("key,kw", [("linlog", dict(linlog=True)),
                                    ("nohubs_delta", dict(nohubs=True, delta=0.5)),
                                    ("normalize", dict(normalize=True))])
def test_flat_options_bitwise(oracle, key, kw):
    A, z = load_flat_golden()
    x, _ = oracle.flat_run(A, 2, z["x0_d2"], oracle.Params(iterations=7, **kw))
    assert np.array_equal(x, z["x_d2_k7_" + key])


def test_reference_random_stream(oracle):
    """mt19937 + libstdc++ uniform_real_distribution, as drawn by forceatlas.hpp:104-125."""
    _, z = load_flat_golden()
    assert np.array_equal(oracle.mt_uniform(7, z["x0_d2"].size).reshape(-1, 2), z["x0_d2"])


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("level", [0, 1])
@pytest.mark.parametrize("k", [1, 3, 100])
def test_multilevel_bitwise(oracle, dim, level, k):
    As, Ps, z = load_hier_golden()
    cA, rA = z["ml_cA_l%d_d%d" % (level, dim)], z["ml_rA_l%d_d%d" % (level, dim)]
    init = oracle.multilevel_init(Ps[level], dim, 5)
    x = oracle.multilevel_run(As[level], Ps[level], cA, rA, dim, init, oracle.Params(iterations=k))
    assert np.array_equal(x, z["ml_x_l%d_d%d_k%d" % (level, dim, k)])


@pytest.mark.parametrize("L", [1, 2, 3])
def test_radii_bitwise(oracle, L):
    As, Ps, z = load_hier_golden()
    AsL, PsL = As[-(L + 1):], Ps[-L:]
    pre = "radii_L%d_" % L
    if L == 1:
        cA, rA = oracle.radii(z[pre + "coords_A_in"], 2)
    else:
        cA, rA = oracle.radii(z[pre + "coords_A_in"], 2, AsL[1], PsL[1], z[pre + "coords_Ac"], z[pre + "r_Ac"])
    assert np.array_equal(cA, z[pre + "coords_A_out"])
    assert np.array_equal(rA, z[pre + "r_A_out"])


def test_embed_bitwise(oracle):
    """Whole embed(): 100 000-iteration coarsest solve + radii + 3 multilevel levels."""
    As, Ps, z = load_hier_golden()
    x = oracle.embed(As, Ps, 2, seed=21)
    assert np.array_equal(x, z["embed_d2_seed21"])


def test_flat_forces_match_run(oracle):
    """oracle_flat_forces (the per-iteration force hook the GPU parity tests use) is the same
    arithmetic that oracle_flat_run applies: one step from x0 reproduces golden k=1."""
    A, z = load_flat_golden()
    x0 = z["x0_d2"]
    F, S = oracle.flat_forces(A, 2, x0)
    _, F_last = oracle.flat_run(A, 2, x0, oracle.Params(iterations=1))
    assert np.array_equal(F, F_last)
    assert (S >= np.linalg.norm(F, axis=1) * (1 - 1e-12)).all()
    half = A.shape[0] // 2
    F_lo, _ = oracle.flat_forces(A, 2, x0, rows=(0, half))
    assert np.array_equal(F_lo[:half], F[:half]) and not F_lo[half:].any()


def test_edge_cases(oracle, graphs):
    import scipy.sparse as sp
    # single vertex, no edges: flat reference yields NaN only if x == 0; here x != 0
    A1 = sp.csr_matrix((1, 1))
    x, _ = oracle.flat_run(A1, 2, np.array([[0.3, -0.2]]), oracle.Params(iterations=3))
    assert np.isfinite(x).all()
    # a vertex at the origin -> NaN (include/forceatlas.hpp:205 divides by |x| unclamped)
    x, _ = oracle.flat_run(A1, 2, np.zeros((1, 2)), oracle.Params(iterations=1))
    assert np.isnan(x).all()
    # coincident points: distance clamped to eps, force direction is zero
    A2 = graphs.grid2d(1, 2)
    F, _ = oracle.flat_forces(A2, 2, np.array([[0.5, 0.5], [0.5, 0.5]]))
    assert np.isfinite(F).all()
    # singleton aggregates land exactly on the parent centre (forceatlas.hpp:539-569)
    P = sp.csr_matrix((np.ones(2), [0, 1], [0, 1, 2]), shape=(2, 2))
    cA = np.array([[1.0, 2.0], [-3.0, 0.5]])
    x = oracle.multilevel_run(A2, P, cA, np.array([0.7, 0.2]), 2, np.array([[0.1, 0.2], [0.3, -0.4]]),
                              oracle.Params(iterations=100))
    assert np.array_equal(x, cA)


def test_galerkin_oracle_matches_scipy_product(oracle, graphs):
    """oracle_galerkin (examples/embedder.cpp:213-216) against scipy's P_T @ A @ P_T.T: identical
    structure; identical sums for unit weights, 1e-13 for real weights (summation order)."""
    import numpy as np
    A = graphs.rgg(3000, 10.0, seed=1)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=20)
    for l, P in enumerate(Ps):
        C = oracle.galerkin(As[l], P)
        R = graphs.galerkin(As[l], P)
        assert np.array_equal(C.indptr, R.indptr) and np.array_equal(C.indices, R.indices)
        assert np.array_equal(C.data, R.data)
    rng = np.random.default_rng(0)
    B = As[0].copy()
    B.data = rng.uniform(0.5, 2.0, B.nnz)
    C = oracle.galerkin(B, Ps[0])
    R = graphs.galerkin(B, Ps[0])
    assert np.array_equal(C.indices, R.indices) and np.abs(C.data - R.data).max() < 1e-12


def test_galerkin_golden(oracle, graphs):
    """oracle_galerkin reproduces the committed products of `P.Mult(A).Mult(P.Transpose())`
    (examples/embedder.cpp:215, minted from the compiled reference driver) level by level, bit for
    bit on the unit-weight chain; the real-weight level to rounding (different association)."""
    from helpers import load_galerkin_golden
    A, Ps, Cs, z = load_galerkin_golden(graphs)
    cur = A
    for P, C in zip(Ps, Cs):
        got = oracle.galerkin(cur, P)
        assert np.array_equal(got.indptr, C.indptr) and np.array_equal(got.indices, C.indices)
        assert np.array_equal(got.data, C.data)
        cur = C
    B = A.copy()
    B.data = z["B_data"]
    got = oracle.galerkin(B, Ps[0])
    assert np.array_equal(got.indptr, z["CB_indptr"]) and np.array_equal(got.indices, z["CB_indices"])
    assert np.abs(got.data - z["CB_data"]).max() < 1e-12
