"""CPU-side checks of the C ABI: the library loads, exports every symbol the header declares,
refuses to compute without a device, and its host-side level-driver step matches the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import load_hier_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_every_declared_symbol(capi):
    header = open(os.path.join(ROOT, "include", "graph_embed_b200.h")).read()
    declared = set(re.findall(r"\b(ge_[a-z0-9_]+)\s*\(", header))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    lib = capi.lib()
    for name in sorted(declared):
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.ge_version()


def test_defaults_match_reference(capi):
    """include/forceatlas.hpp:92-103 and :320-331 (iterations = 100 as src/embed.cpp:793 passes)."""
    p = capi.flat_params()
    assert (p.iterations, p.ks, p.ksmax, p.repel, p.attract, p.gravity, p.delta, p.tolerate) == \
        (100000, 0.1, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0)
    assert (p.use_weights, p.linlog, p.nohubs, p.normalize, p.precision) == (1, 0, 0, 0, capi.GE_F64)
    assert capi.multilevel_params().iterations == 100


def test_no_cpu_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.GeError) as e:
        capi.Context(0)
    assert e.value.status == capi.GE_ERR_NO_DEVICE


def test_reference_uniform_matches_oracle(capi, oracle):
    assert np.array_equal(capi.reference_uniform(99, 1001), oracle.mt_uniform(99, 1001))


@pytest.mark.parametrize("L", [1, 2, 3])
def test_level_radii_matches_golden(capi, L):
    """ge_level_radii (heap-based) against the reference's sort-based ball growing, bit for bit."""
    As, Ps, z = load_hier_golden()
    AsL, PsL = As[-(L + 1):], Ps[-L:]
    pre = "radii_L%d_" % L
    if L == 1:
        cA, rA = capi.level_radii(z[pre + "coords_A_in"], 2)
    else:
        cA, rA = capi.level_radii(z[pre + "coords_A_in"], 2, AsL[1], PsL[1], z[pre + "coords_Ac"], z[pre + "r_Ac"])
    assert np.array_equal(cA, z[pre + "coords_A_out"])
    assert np.array_equal(rA, z[pre + "r_A_out"])


@pytest.mark.parametrize("dim", [2, 3])
def test_level_radii_matches_oracle_random(capi, oracle, graphs, dim):
    A = graphs.rgg(1500, 9.0, seed=9)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=30)
    rng = np.random.default_rng(4)
    # base case on the coarsest level
    m = As[-1].shape[0]
    x = rng.normal(size=(m, dim))
    c1, r1 = capi.level_radii(x, dim)
    c2, r2 = oracle.radii(x, dim)
    assert np.array_equal(c1, c2) and np.array_equal(r1, r2)
    # general case one level up, including duplicate points (zero distances) and singletons
    l = len(Ps) - 1
    m = As[l].shape[0]
    x = rng.normal(size=(m, dim))
    x[3] = x[4]
    cAc, rAc = rng.normal(size=(Ps[l].shape[0], dim)), rng.random(Ps[l].shape[0]) + 0.1
    c1, r1 = capi.level_radii(x, dim, As[l], Ps[l], cAc, rAc)
    c2, r2 = oracle.radii(x, dim, As[l], Ps[l], cAc, rAc)
    assert np.array_equal(c1, c2) and np.array_equal(r1, r2)


def test_invalid_arguments(capi):
    with pytest.raises(capi.GeError) as e:
        capi.level_radii(np.zeros((3, 2)), 2, A_c=__import__("scipy.sparse").sparse.identity(4, format="csr"),
                         P_T_c=__import__("scipy.sparse").sparse.identity(4, format="csr"),
                         coords_Ac=np.zeros((4, 2)), r_Ac=np.ones(4))
    assert e.value.status == capi.GE_ERR_INVALID


def test_symmetric_share_matches_python_mirror(capi):
    """The C++ cut of the triangular unit list (csrc/ge_flat_sym.cu) == sharding.pair_share."""
    from graph_embed_b200 import sharding
    for ld in (256, 1024, 2304, 500_224):
        for world in (1, 2, 4, 8):
            for rank in range(world):
                assert capi.symmetric_share(ld, rank, world) == sharding.pair_share(ld, world, rank)


def _units(blocks):
    """(row block, column tile) units of a share, as a set of (row0, tile)."""
    out = set()
    for row0, row1, tf, nt, tsym in blocks:
        for t in range(tf, tf + nt):
            assert (row0, t) not in out
            out.add((row0, t))
    return out


def test_column_panel_passes_cover_every_unit_once(capi):
    """Large graphs cut the symmetric sweep into passes over column panels (one scratch buffer
    reused): over all passes every (row block, column tile) unit of the rank's share appears
    exactly once, tiles stay inside their panel, and panels are ordered."""
    for ld in (1024, 5120, 500_224):
        for world in (1, 4):
            for rank in range(world):
                whole = _units(capi.symmetric_share(ld, rank, world))
                for npass in (1, 2, 7, 30):
                    seen, last_hi = set(), 0
                    for q in range(npass):
                        blocks = capi.symmetric_pass_share(ld, rank, world, npass, q)
                        u = _units(blocks)
                        assert not (u & seen)
                        seen |= u
                        if u:
                            lo, hi = min(t for _, t in u), max(t for _, t in u)
                            assert lo >= last_hi          # panels do not overlap and ascend
                            last_hi = hi + 1
                    assert seen == whole, (ld, world, rank, npass)


def test_embed_aggregate_ranges_are_contiguous_and_balanced(capi, graphs):
    """How ge_embed on an N-device context shares out a level: contiguous aggregate ranges that
    cover every aggregate once, each within one aggregate's cost of the mean."""
    A = graphs.rmat(13, 16, seed=2)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=50, max_levels=2)
    for l, P in enumerate(Ps):
        s = np.diff(P.indptr).astype(np.float64)
        row_nnz = np.diff(As[l].indptr).astype(np.float64)
        member_nnz = np.add.reduceat(row_nnz[P.indices], P.indptr[:-1])
        cost = s * s + member_nnz
        for ndev in (1, 2, 3, 8):
            cuts, pairs = capi.embed_aggregate_ranges(As[l], P, ndev)
            assert cuts[0] == 0 and cuts[-1] == P.shape[0] and (np.diff(cuts) >= 0).all()
            assert pairs == float((s * (s - 1)).sum())
            per = [cost[cuts[d]:cuts[d + 1]].sum() for d in range(ndev)]
            assert max(per) <= cost.sum() / ndev + cost.max() + 1e-9


def test_multi_device_context_fails_loudly_without_gpus(capi):
    """ge_context_create_multi has no fallback either: without a CUDA device it reports
    GE_ERR_NO_DEVICE (on a GPU box with fewer devices than requested: GE_ERR_INVALID)."""
    import ctypes as C
    h = C.c_void_p()
    st = capi.lib().ge_context_create_multi(2, None, C.byref(h))
    assert st in (capi.GE_ERR_NO_DEVICE, capi.GE_ERR_INVALID, capi.GE_OK)
    if st == capi.GE_OK:
        assert capi.lib().ge_context_device_count(h) == 2
        capi.lib().ge_context_destroy(h)
    else:
        assert not h.value and len(capi.lib().ge_last_error()) > 0
    assert capi.lib().ge_context_device_count(None) == 0


def test_embed_aggregate_owners_balance_giant_aggregates(capi, graphs):
    """Multi-device embed: every aggregate has exactly one owner, and with a few giant aggregates in
    the level (the shape of the reference partitioner's hierarchies) the most loaded device is
    within one light aggregate + the list-scheduling bound of the mean."""
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    sizes = np.concatenate([rng.integers(3000, 9000, 11), rng.integers(1, 12, 4000)])
    rng.shuffle(sizes)
    n = int(sizes.sum())
    agg = np.repeat(np.arange(len(sizes)), sizes)
    perm = rng.permutation(n)
    agg = agg[np.argsort(perm)]
    P = graphs.aggregation_matrix(agg.astype(np.int64), len(sizes))
    P = sp.csr_matrix((P.data, P.indices.astype(np.int32), P.indptr.astype(np.int32)), shape=P.shape)
    A = graphs.canonical(sp.identity(n, format="csr"))
    cost = sizes.astype(np.float64) ** 2 + sizes
    for ndev in (1, 2, 4, 8):
        owner, pairs = capi.embed_aggregate_owners(A, P, ndev)
        assert owner.min() >= 0 and owner.max() < ndev
        assert pairs == float((sizes.astype(np.float64) * (sizes - 1)).sum())
        load = np.bincount(owner, weights=cost, minlength=ndev)
        # largest-first list scheduling: max load <= mean + largest job (and far better in practice)
        assert load.max() <= cost.sum() / ndev + cost.max()
        if ndev == 8:
            cuts, _ = capi.embed_aggregate_ranges(A, P, ndev)
            contiguous = max(cost[cuts[d]:cuts[d + 1]].sum() for d in range(ndev))
            assert load.max() <= contiguous + 1e-9


def test_export_header_writes_the_reference_formats(tmp_path):
    """host/include/export.hpp: writePartition / writeCoords as src/export.cpp:16-39 writes them and
    the three plot inputs of examples/embedder.cpp:230-289 (host-only, compiled here with g++)."""
    src = tmp_path / "w.cpp"
    src.write_text(r'''
#include "export.hpp"
int main(int, char** argv) {
  const std::string d = argv[1];
  std::vector<std::vector<double>> xy = {{0.5, -1.25}, {1e-7, 123456789.0}, {3.0, 1.0 / 3.0}};
  partition::writeCoords(xy, d + "/coords.txt");
  partition::writePartition({2, 0, 1}, d + "/part.txt");
  SparseMatrix A({0, 1, 3, 4}, {1, 0, 2, 1}, {1.0, 1.0, 1.0, 1.0}, 3, 3);
  SparseMatrix P({0, 2, 3}, {0, 1, 2}, {1.0, 1.0, 1.0}, 2, 3);
  ge_b200::writePlotInputs(A, {P}, xy, 2, d + "/p.temp", d + "/c.temp", d + "/m.temp");
  ge_b200::writePlotInputs(A, {}, {{1, 2, 3}, {4, 5, 6}, {7, 8, 9.5}}, 3, d + "/p0.temp", d + "/c0.temp", d + "/m0.temp");
  return 0;
}
''')
    exe = tmp_path / "w"
    host = os.path.join(ROOT, "graph-embed_b200", "host")
    subprocess.check_call(["g++", "-std=c++14", "-I", os.path.join(host, "include"), "-I", os.path.join(host, "compat"),
                           str(src), "-o", str(exe)])
    subprocess.check_call([str(exe), str(tmp_path)])
    xy = [[0.5, -1.25], [1e-7, 123456789.0], [3.0, 1.0 / 3.0]]
    assert (tmp_path / "coords.txt").read_text() == "".join("".join("%g " % v for v in r) + "\n" for r in xy)
    assert (tmp_path / "part.txt").read_text() == "2\n0\n1\n"
    assert (tmp_path / "p.temp").read_text() == "3 1\n2 \n0 1 \n2 \n"
    assert (tmp_path / "c.temp").read_text() == "".join("%g %g 0\n" % (r[0], r[1]) for r in xy)
    assert (tmp_path / "m.temp").read_text() == "0 1\n1 0\n1 2\n2 1\n"
    assert (tmp_path / "p0.temp").read_text() == "3 1\n3 \n0 \n1 \n2 \n"
    assert (tmp_path / "c0.temp").read_text() == "1 2 3\n4 5 6\n7 8 9.5\n"
