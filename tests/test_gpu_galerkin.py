"""GPU parity of the Galerkin coarse graph A_c = P_T A P_T^T (csrc/ge_galerkin.cu; the step
examples/embedder.cpp:213-216 runs before partition::embed) against the oracle, through the C ABI.
Index and byte work: the bar is bit-exact (indptr, indices and the sums)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def _same(C, R):
    assert C.shape == R.shape
    assert np.array_equal(C.indptr, R.indptr)
    assert np.array_equal(C.indices, R.indices)
    assert np.array_equal(C.data, R.data)


def _random_partition(n, m, rng, empty=0):
    """m aggregates (the last `empty` without members), members in random order."""
    v = rng.integers(0, m - empty, n)
    order = rng.permutation(n)
    order = order[np.argsort(v[order], kind="stable")]
    ptr = np.concatenate([[0], np.cumsum(np.bincount(v, minlength=m))]).astype(np.int32)
    return sp.csr_matrix((np.ones(n), order.astype(np.int32), ptr), shape=(m, n))


@pytest.mark.parametrize("name", ["grid", "rgg", "rmat"])
def test_hierarchy_levels_unit_weights(ctx, oracle, graphs, name):
    A = {"grid": lambda: graphs.grid2d(40, 37), "rgg": lambda: graphs.rgg(6000, 10.0, seed=2),
         "rmat": lambda: graphs.rmat(12, 8, seed=4)}[name]()
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=20)
    for l, P in enumerate(Ps):  # every level: coarse inputs carry Galerkin weights and self-loops
        C = ctx.galerkin(As[l], P)
        _same(C, oracle.galerkin(As[l], P))
        ref = graphs.galerkin(As[l], P)     # scipy P_T @ A @ P_T.T
        assert np.array_equal(C.indptr, ref.indptr) and np.array_equal(C.indices, ref.indices)
        assert np.array_equal(C.data, ref.data)   # integer-valued sums: exact in any order


def test_real_weights_bit_exact_in_oracle_order(ctx, oracle, graphs):
    rng = np.random.default_rng(3)
    A = graphs.rgg(5000, 12.0, seed=5).tocsr()
    A.data = rng.uniform(0.1, 3.0, A.nnz)
    P = _random_partition(A.shape[0], 700, rng)
    C = ctx.galerkin(A, P)
    _same(C, oracle.galerkin(A, P))
    ref = (P @ A @ P.T).tocsr()
    ref.sort_indices()
    assert np.abs(C - ref).max() < 1e-12 * abs(ref).max()


def test_segments_beyond_shared_memory_and_empty_aggregates(ctx, oracle):
    """An aggregate whose members carry more than 4096 fine entries goes through the
    global-scratch sort; aggregates without members give empty rows."""
    rng = np.random.default_rng(7)
    n = 9000
    hub = sp.random(n, n, density=0.004, random_state=11, format="csr")
    hub = ((hub + hub.T) > 0).astype(np.float64).tolil()
    hub[0, 1:6000] = 1.0   # one row with 6000 entries
    hub[1:6000, 0] = 1.0
    A = hub.tocsr()
    A.data = rng.integers(1, 5, A.nnz).astype(np.float64)
    P = _random_partition(n, 300, rng, empty=3)
    C, st = ctx.galerkin(A, P, with_stats=True)
    assert st["segments_global"] >= 1
    _same(C, oracle.galerkin(A, P))
    assert (np.diff(C.indptr)[-3:] == 0).all()


def test_no_weights_array_means_unit_weights(ctx, capi, oracle, graphs):
    import ctypes as C_
    A = graphs.grid2d(12, 12)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=10, max_levels=1)
    a, pt = capi.CsrView(A, with_data=False), capi.CsrView(Ps[0], with_data=False)
    m = Ps[0].shape[0]
    ptr, idx, val = np.zeros(m + 1, np.int32), np.zeros(A.nnz, np.int32), np.zeros(A.nnz)
    nnz = C_.c_int64()
    capi._check(capi.lib().ge_galerkin(ctx.h, a.ref(), pt.ref(), capi._ptr(ptr, capi._pi),
                                       capi._ptr(idx, capi._pi), capi._ptr(val, capi._pd),
                                       C_.c_int64(A.nnz), C_.byref(nnz), None))
    R = oracle.galerkin(A, Ps[0])
    assert nnz.value == R.nnz and np.array_equal(ptr, R.indptr)
    assert np.array_equal(idx[:R.nnz], R.indices) and np.array_equal(val[:R.nnz], R.data)


def test_capacity_too_small_reports_the_size(ctx, capi, graphs):
    import ctypes as C_
    A = graphs.grid2d(10, 10)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=10, max_levels=1)
    a, pt = capi.CsrView(A), capi.CsrView(Ps[0], with_data=False)
    m = Ps[0].shape[0]
    ptr, idx, val = np.zeros(m + 1, np.int32), np.zeros(4, np.int32), np.zeros(4)
    nnz = C_.c_int64()
    st = capi.lib().ge_galerkin(ctx.h, a.ref(), pt.ref(), capi._ptr(ptr, capi._pi), capi._ptr(idx, capi._pi),
                                capi._ptr(val, capi._pd), C_.c_int64(4), C_.byref(nnz), None)
    assert st != 0 and nnz.value == As[1].nnz and ptr[-1] == As[1].nnz


def test_embed_on_gpu_built_hierarchy_matches_host_built(ctx, capi, graphs):
    """The hierarchy a caller would build with ge_galerkin feeds embed() exactly like the
    host-built one (same graphs -> same seeded layout)."""
    A = graphs.rgg(3000, 10.0, seed=9)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=40)
    Gs = [As[0]]
    for P in Ps:
        Gs.append(ctx.galerkin(Gs[-1], P))
    for G, H in zip(Gs, As):
        _same(G, H)
    x1, _ = ctx.embed(Gs, Ps, 2, seed=5, coarse_iterations=500)
    x2, _ = ctx.embed(As, Ps, 2, seed=5, coarse_iterations=500)
    assert np.array_equal(x1, x2)


@pytest.mark.parametrize("n", [1500, 4000])   # shared-memory segments / global-scratch segments
def test_long_runs_keep_the_sequential_summation_order(ctx, oracle, graphs, n):
    """Five aggregates: every coarse row merges hundreds of entries per column.  Runs longer than
    32 entries are summed by a warp (values fetched in parallel, added in run order): the result
    must still be the sequential sum, bit for bit, with real weights."""
    rng = np.random.default_rng(21)
    A = graphs.rgg(n, 10.0, seed=23).tocsr()
    A.data = rng.uniform(0.1, 3.0, A.nnz)
    P = _random_partition(A.shape[0], 5, rng)
    C, st = ctx.galerkin(A, P, with_stats=True)
    assert (st["segments_global"] > 0) == (n > 2500)
    _same(C, oracle.galerkin(A, P))


def test_golden_products_of_the_reference_driver(ctx, graphs):
    """ge_galerkin against tests/golden/galerkin_grid30.npz (the caller's expression run through the
    compiled reference driver): exact on the unit-weight chain, rounding on the real-weight level."""
    from helpers import load_galerkin_golden
    A, Ps, Cs, z = load_galerkin_golden(graphs)
    cur = A
    for P, C in zip(Ps, Cs):
        _same(ctx.galerkin(cur, P), C)
        cur = C
    B = A.copy()
    B.data = z["B_data"]
    got = ctx.galerkin(B, Ps[0])
    assert np.array_equal(got.indptr, z["CB_indptr"]) and np.array_equal(got.indices, z["CB_indices"])
    assert np.abs(got.data - z["CB_data"]).max() < 1e-12
