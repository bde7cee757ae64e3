"""The oracle against the compiled reference itself (oracle/_ref, built from the unmodified
sources under /root/reference).  Only runs where the reference was compiled; the committed golden
vectors (test_oracle_golden.py) carry the same pin to boxes without /root/reference."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_available("strict"):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("nthreads", [1, 3])
def test_flat_bitwise_any_thread_count(ref, graphs, dim, nthreads):
    A = graphs.rgg(300, 8.0, seed=3)
    n = A.shape[0]
    x0 = ref.mt_uniform(11, n * dim).reshape(n, dim)
    for k in (1, 10):
        p = ref.Params(iterations=k)
        xo, _ = ref.flat_run(A, dim, x0, p)
        assert np.array_equal(xo, ref.ref_flat(A, dim, x0, p, nthreads=nthreads))


@pytest.mark.parametrize("kw", [dict(useWeights=False), dict(linlog=True), dict(delta=0.0), dict(delta=2.0),
                                dict(ks=0.3, ksmax=2.0, repel=2.0, attract=0.7, gravity=1.5, tolerate=0.8)])
def test_flat_options(ref, graphs, kw):
    A = graphs.galerkin(graphs.grid2d(10, 10), graphs.coarsen(graphs.grid2d(10, 10), 0.5, 10)[1][0])
    n = A.shape[0]  # weighted, with self-loops (quirk Q5)
    x0 = ref.mt_uniform(2, n * 2).reshape(n, 2)
    p = ref.Params(iterations=9, **kw)
    xo, _ = ref.flat_run(A, 2, x0, p)
    assert np.array_equal(xo, ref.ref_flat(A, 2, x0, p))


def test_multilevel_and_embed_with_own_hierarchy(ref, graphs):
    A = graphs.rgg(800, 9.0, seed=5)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=20)
    rng = np.random.default_rng(0)
    for l in range(len(Ps)):
        m = Ps[l].shape[0]
        cA, rA = rng.normal(size=(m, 3)), rng.random(m)
        p = ref.Params(iterations=20)
        xo = ref.multilevel_run(As[l], Ps[l], cA, rA, 3, ref.multilevel_init(Ps[l], 3, 8), p)
        assert np.array_equal(xo, ref.ref_multilevel(As[l], Ps[l], cA, rA, 3, p, seed=8))
    xr, _ = ref.ref_embed(As, Ps, 2, seed=13)
    assert np.array_equal(ref.embed(As, Ps, 2, seed=13), xr)


def test_galerkin_matches_the_callers_expression(ref, graphs):
    """oracle_galerkin against `P.Mult(A).Mult(P.Transpose())` (examples/embedder.cpp:215) run
    through the reference driver with the stand-in linalgcpp container: same structure; identical
    sums on unit-weight graphs (integers) and on every coarser level; real weights agree to
    rounding (the two-stage product associates the sums differently)."""
    A = graphs.rgg(2500, 10.0, seed=6)
    As, Ps = graphs.coarsen(A, 0.25, min_coarse=20)
    for l, P in enumerate(Ps):
        C, R = ref.galerkin(As[l], P), ref.ref_galerkin(As[l], P)
        assert np.array_equal(C.indptr, R.indptr) and np.array_equal(C.indices, R.indices)
        assert np.array_equal(C.data, R.data)
    B = As[0].copy()
    B.data = np.random.default_rng(1).uniform(0.5, 2.0, B.nnz)
    C, R = ref.galerkin(B, Ps[0]), ref.ref_galerkin(B, Ps[0])
    assert np.array_equal(C.indices, R.indices)
    assert np.abs(C.data - R.data).max() < 1e-12


def test_export_writers_byte_identical(ref, tmp_path):
    """host/include/export.hpp against the reference's own writeCoords / writePartition
    (src/export.cpp:16-39, compiled into oracle/_ref): the files must be byte-identical."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    host = os.path.join(root, "graph-embed_b200", "host")
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.normal(size=(200, 3)) * 10.0 ** rng.integers(-9, 9, size=(200, 1)),
                        [[0.0, -0.0, 1.0], [1e-300, 1e300, 123456.5], [0.1, 1.0 / 3.0, 2.0 / 3.0]]])
    part = rng.integers(0, 50, size=500).astype(np.int32)
    x.tofile(tmp_path / "x.bin")
    part.tofile(tmp_path / "p.bin")
    src = tmp_path / "w.cpp"
    src.write_text(r'''
#include <cstdio>
#include <cstdlib>
#include "export.hpp"
int main(int, char** argv) {
  const std::string d = argv[1];
  const int n = std::atoi(argv[2]), np_ = std::atoi(argv[3]);
  std::vector<std::vector<double>> x(n, std::vector<double>(3));
  FILE* f = std::fopen((d + "/x.bin").c_str(), "rb");
  for (auto& r : x) if (std::fread(r.data(), 8, 3, f) != 3) return 1;
  std::fclose(f);
  std::vector<int> p(np_);
  f = std::fopen((d + "/p.bin").c_str(), "rb");
  if (std::fread(p.data(), 4, np_, f) != (size_t)np_) return 1;
  std::fclose(f);
  partition::writeCoords(x, d + "/ours_coords.txt");
  partition::writePartition(p, d + "/ours_part.txt");
  return 0;
}
''')
    subprocess.check_call(["g++", "-std=c++14", "-I", os.path.join(host, "include"), "-I", os.path.join(host, "compat"),
                           str(src), "-o", str(tmp_path / "w")])
    subprocess.check_call([str(tmp_path / "w"), str(tmp_path), str(x.shape[0]), str(part.shape[0])])
    ref.ref_write_coords(x, tmp_path / "ref_coords.txt")
    ref.ref_write_partition(part, tmp_path / "ref_part.txt")
    assert (tmp_path / "ours_coords.txt").read_bytes() == (tmp_path / "ref_coords.txt").read_bytes()
    assert (tmp_path / "ours_part.txt").read_bytes() == (tmp_path / "ref_part.txt").read_bytes()
