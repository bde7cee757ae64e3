"""Shared test helpers: golden loaders and tolerance metrics."""
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Tolerances stated by BASELINE.json's north_star for per-iteration force vectors.
TOL_F64 = 1e-10
TOL_F32 = 1e-4


def load_flat_golden():
    z = np.load(os.path.join(GOLDEN, "flat_grid12.npz"))
    n = len(z["indptr"]) - 1
    A = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
    return A, z


def load_hier_golden():
    z = np.load(os.path.join(GOLDEN, "hier_grid30.npz"))
    L = int(z["L"])
    As, Ps = [], []
    for l in range(L + 1):
        n = len(z["A%d_indptr" % l]) - 1
        As.append(sp.csr_matrix((z["A%d_data" % l], z["A%d_indices" % l], z["A%d_indptr" % l]), shape=(n, n)))
    for l in range(L):
        m = len(z["P%d_indptr" % l]) - 1
        idx = z["P%d_indices" % l]
        Ps.append(sp.csr_matrix((np.ones(len(idx)), idx, z["P%d_indptr" % l]), shape=(m, As[l].shape[0])))
    return As, Ps, z


def load_config1_golden(graphs):
    """BASELINE config 1 (grid 100x100, coarsening 0.25): hierarchy from the reference's own
    partitioner + the compiled reference's embed() output for seed 3."""
    z = np.load(os.path.join(GOLDEN, "config1_grid100.npz"))
    A = graphs.grid2d(100, 100)
    Ps, n = [], A.shape[0]
    for l in range(int(z["L"])):
        ptr, idx = z["P%d_indptr" % l], z["P%d_indices" % l]
        Ps.append(sp.csr_matrix((np.ones(len(idx)), idx, ptr), shape=(len(ptr) - 1, n)))
        n = len(ptr) - 1
    return graphs.hierarchy_from(A, Ps), Ps, z


def force_error(F, F_ref, scale):
    """Per-vertex |F - F_ref|_2 divided by the conditioning scale the oracle reports (the sum of
    the norms of the individual terms that were added into that vertex's force)."""
    return np.linalg.norm(F - F_ref, axis=1) / np.maximum(scale, 1e-300)


def layout_stats(A, x):
    """Edge-length mean / coefficient of variation and a sampled normalised stress."""
    coo = A.tocoo()
    keep = coo.row < coo.col
    el = np.linalg.norm(x[coo.row[keep]] - x[coo.col[keep]], axis=1)
    rng = np.random.default_rng(0)
    n = A.shape[0]
    i, j = rng.integers(0, n, 4000), rng.integers(0, n, 4000)
    dist = np.linalg.norm(x[i] - x[j], axis=1)
    extent = np.linalg.norm(x - x.mean(0), axis=1).max()
    return dict(edge_mean=el.mean() / extent, edge_cv=el.std() / el.mean(),
                pair_mean=dist.mean() / extent)


def load_galerkin_golden(graphs):
    """Galerkin products of the grid-30 hierarchy minted from the compiled reference driver."""
    z = np.load(os.path.join(GOLDEN, "galerkin_grid30.npz"))
    A = graphs.canonical(graphs.grid2d(30, 30))
    Ps, Cs, n = [], [], A.shape[0]
    for l in range(int(z["L"])):
        ptr, idx = z["P%d_indptr" % l], z["P%d_indices" % l]
        m = len(ptr) - 1
        Ps.append(sp.csr_matrix((np.ones(len(idx)), idx, ptr), shape=(m, n)))
        Cs.append(sp.csr_matrix((z["C%d_data" % l], z["C%d_indices" % l], z["C%d_indptr" % l]), shape=(m, m)))
        n = m
    return A, Ps, Cs, z


_REFHIER_CACHE = {}


def load_ref_hierarchy(graphs, name):
    """A hierarchy produced by the REFERENCE's own partitioner (src/partitioner.cpp:1550-1893, call
    shape of examples/embedder.cpp:187) on a synthetic graph of a BASELINE config, cached by
    tests/golden/make_ref_hierarchy.py as the vertex->aggregate map of every level.  The graph is
    regenerated from its seed and checked against the stored digest; the coarse graphs are the
    Galerkin products examples/embedder.cpp:213-216 forms.  -> (As, P_Ts, meta)"""
    import hashlib
    if name in _REFHIER_CACHE:
        return _REFHIER_CACHE[name]
    z = np.load(os.path.join(GOLDEN, "refhier_%s.npz" % name))
    kind, arg, seed = str(z["kind"]), int(z["arg"]), int(z["seed"])
    if kind == "rmat":
        A = graphs.rmat(arg, 16, seed=seed)
    elif kind == "delaunay":
        A = graphs.delaunay3d(arg, seed=seed)
    elif kind == "rgg":
        A = graphs.rgg(arg, 10.0, seed=seed)
    else:
        raise ValueError(kind)
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(A.indptr).tobytes())
    h.update(np.ascontiguousarray(A.indices).tobytes())
    assert h.hexdigest() == str(z["digest"]), "the generator no longer reproduces the cached graph"
    Ps, n = [], A.shape[0]
    for l in range(int(z["L"])):
        agg = z["agg%d" % l]
        assert agg.shape[0] == n
        m = int(agg.max()) + 1
        Ps.append(graphs.aggregation_matrix(agg, m))
        Ps[-1] = sp.csr_matrix((Ps[-1].data, Ps[-1].indices.astype(np.int32),
                                Ps[-1].indptr.astype(np.int32)), shape=Ps[-1].shape)
        n = m
    As = graphs.hierarchy_from(A, Ps)
    meta = dict(kind=kind, arg=arg, seed=seed, cf=float(z["cf"]),
                partition_seconds=float(z["partition_seconds"]))
    _REFHIER_CACHE[name] = (As, Ps, meta)
    return As, Ps, meta
