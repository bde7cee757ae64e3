"""GPU parity of the symmetric all-pairs sweep (csrc/ge_flat_sym.cu: every unordered pair once,
applied to both endpoints) against the oracle's ordered-pair loop
(include/forceatlas.hpp:151-167), through the C ABI.  GE_SYM_MIN_ROWS=0 routes small graphs
through the kernel that normally serves n >= 32768."""
import numpy as np
import pytest

from helpers import TOL_F32, TOL_F64, force_error

pytestmark = pytest.mark.gpu


@pytest.fixture
def sym(monkeypatch):
    monkeypatch.setenv("GE_SYM_MIN_ROWS", "0")
    monkeypatch.setenv("GE_REP_SYM", "1")
    return monkeypatch


@pytest.mark.parametrize("n", [700, 2000, 5003])   # 1 block / 2 blocks / 5 blocks, ragged tail
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("shape", [(4, 8), (4, 4), (2, 8), (2, 4)])
def test_forces_fp64(ctx, capi, oracle, graphs, sym, n, dim, shape):
    sym.setenv("GE_SYM_IPT", str(shape[0]))
    sym.setenv("GE_SYM_CG", str(shape[1]))
    A = graphs.rgg(n, 10.0, seed=3)
    n = A.shape[0]
    x0 = capi.reference_uniform(11, n * dim).reshape(n, dim)
    F_ref, S = oracle.flat_forces(A, dim, x0)
    F = ctx.flat_forces(A, dim, x0, capi.flat_params(), path=1)
    assert force_error(F, F_ref, S).max() < TOL_F64
    assert np.linalg.norm(F - F_ref) / np.linalg.norm(F_ref) < TOL_F64


@pytest.mark.parametrize("n", [2000, 5003])
@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("shape", [(4, 8), (2, 4)])
def test_forces_fp32(ctx, capi, oracle, graphs, sym, n, dim, shape):
    sym.setenv("GE_SYM_IPT", str(shape[0]))
    sym.setenv("GE_SYM_CG", str(shape[1]))
    A = graphs.rgg(n, 10.0, seed=4)
    n = A.shape[0]
    # identical positions on both sides: the FP32 kernel sees x0 rounded to float, and for close
    # pairs that rounding alone changes (xi - xj) by more than the tolerance
    x0 = capi.reference_uniform(12, n * dim).reshape(n, dim).astype(np.float32).astype(np.float64)
    F_ref, S = oracle.flat_forces(A, dim, x0)
    F = ctx.flat_forces(A, dim, x0, capi.flat_params(precision=capi.GE_F32), path=1)
    assert force_error(F, F_ref, S).max() < TOL_F32


def test_matches_ordered_sweep_and_is_reproducible(ctx, capi, graphs, sym):
    """Same forces as the ordered-pair kernel up to summation order; bit-identical run to run
    (no atomics: partial sums are combined in a fixed order)."""
    A = graphs.rgg(6000, 10.0, seed=5)
    n = A.shape[0]
    x0 = capi.reference_uniform(13, n * 2).reshape(n, 2)
    F1 = ctx.flat_forces(A, 2, x0, capi.flat_params(), path=1)
    F2 = ctx.flat_forces(A, 2, x0, capi.flat_params(), path=1)
    assert np.array_equal(F1, F2)
    sym.setenv("GE_REP_SYM", "0")
    F0 = ctx.flat_forces(A, 2, x0, capi.flat_params(), path=1)
    assert np.linalg.norm(F1 - F0) / np.linalg.norm(F0) < 1e-13


def test_coincident_points(ctx, capi, oracle, graphs, sym):
    """eps clamp (include/forceatlas.hpp:155-157): coincident and sub-eps pairs, across blocks."""
    A = graphs.rgg(2500, 10.0, seed=6)
    n = A.shape[0]
    x0 = capi.reference_uniform(14, n * 2).reshape(n, 2)
    x0[10] = x0[11]
    x0[20] = x0[2100]            # different row blocks: goes through the column side
    x0[30] = x0[1500] + 1e-7
    F_ref, S = oracle.flat_forces(A, 2, x0)
    F = ctx.flat_forces(A, 2, x0, capi.flat_params(), path=1)
    assert np.isfinite(F).all()
    assert force_error(F, F_ref, S).max() < TOL_F64


def test_momentum_conservation(ctx, capi, graphs, sym):
    """Size-independent property: without gravity the pair forces cancel, sum_i F_i = 0."""
    A = graphs.rgg(40000, 10.0, seed=7)
    n = A.shape[0]
    x0 = capi.reference_uniform(15, n * 3).reshape(n, 3)
    F = ctx.flat_forces(A, 3, x0, capi.flat_params(gravity=0.0), path=1)
    assert np.isfinite(F).all()
    assert np.abs(F.sum(0)).max() < 1e-9 * np.abs(F).sum(0).max()


@pytest.mark.parametrize("passes", [2, 3, 7])
@pytest.mark.parametrize("dim", [2, 3])
def test_column_panel_passes(ctx, capi, oracle, graphs, sym, passes, dim):
    """Large graphs cut the sweep into passes over column panels that reuse one scratch buffer (the
    column-side slabs grow with n^2); forced here on a 5-block graph: forces against the oracle,
    and the same positions as the one-pass plan up to summation order, also on rank plans and on
    a segmented (per-aggregate) sweep."""
    A = graphs.rgg(5003, 10.0, seed=3)
    n = A.shape[0]
    x0 = capi.reference_uniform(11, n * dim).reshape(n, dim)
    F_ref, S = oracle.flat_forces(A, dim, x0)
    sym.setenv("GE_SYM_PASSES", "1")
    one = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=2))
    sym.setenv("GE_SYM_PASSES", str(passes))
    F = ctx.flat_forces(A, dim, x0, capi.flat_params(), path=1)
    assert force_error(F, F_ref, S).max() < TOL_F64
    got = ctx.flat_forceatlas(A, dim, x0, capi.flat_params(iterations=2))
    assert np.abs(got - one).max() < 1e-11 * np.abs(one).max()
    # the large-aggregate tier of the per-aggregate solver runs the same plan over segments
    sym.setenv("GE_CTA_MAX", "40")
    sym.setenv("GE_ML_SYM_MIN_MPAIRS", "0")
    As, Ps = graphs.coarsen(graphs.rgg(6000, 10.0, seed=3), 0.004, min_coarse=10, max_levels=1)
    Al, P = As[0], Ps[0]
    assert np.diff(P.indptr).max() > 300
    m = P.shape[0]
    cA = np.random.default_rng(1).normal(size=(m, dim))
    x = capi.reference_uniform(3, Al.shape[0] * dim).reshape(-1, dim)
    _, Fm_ref, Sm = oracle.multilevel_run(Al, P, cA, np.ones(m), dim, x, oracle.Params(iterations=1), forces_iter=0)
    Fm = ctx.multilevel_forces(Al, P, cA, x, dim, capi.multilevel_params())
    assert force_error(Fm, Fm_ref, Sm).max() < TOL_F64


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("dim", [2, 3])
def test_multi_rank_plans_on_one_gpu(ctx, capi, graphs, sym, world, dim):
    """The `world` plans of a symmetric multi-rank solve, run one after the other on one GPU with
    the reduce-scatter replaced by a sum of their pair-sum buffers, move the vertices exactly
    like the single-rank plan (up to summation order)."""
    import torch
    A = graphs.rgg(5003, 10.0, seed=8)
    n = A.shape[0]
    x0 = capi.reference_uniform(16, n * dim).reshape(n, dim)
    one = ctx.flat_plan(A, dim, capi.flat_params())
    assert one.symmetric
    one.upload(x0)
    one.iterate(1)
    ref = one.download()
    one.close()
    plans, sums = [], []
    for r in range(world):
        p = ctx.flat_plan(A, dim, capi.flat_params(), symmetric=(r, world))
        s = torch.zeros(dim * p.ld, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        p.bind_pair_sums(s.data_ptr())
        p.upload(x0)
        p.launch_repulsion()
        p.sync()
        plans.append(p)
        sums.append(s)
    total = torch.stack(sums).sum(0)
    got = np.zeros_like(ref)
    for r, p in enumerate(plans):
        sums[r].copy_(total)
        torch.cuda.synchronize()
        p.launch_step()
        p.swap()
        x = p.download()
        got[p.rows[0]:p.rows[1]] = x[p.rows[0]:p.rows[1]]
        # rows of other ranks are untouched until the all-gather
        other = np.ones(n, bool)
        other[p.rows[0]:p.rows[1]] = False
        assert np.array_equal(x[other], x0[other])
        p.close()
    assert np.abs(got - ref).max() < 1e-12 * max(1.0, np.abs(ref).max())
