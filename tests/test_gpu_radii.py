"""The ball-radius + rescale step between levels (src/embed.cpp:615-778) on the device
(csrc/ge_radii.cu): bit-identical to the reference's sort-and-pop loop
(oracle.radii restates it; the golden vectors come from the compiled reference) and to the host
restatement ge_level_radii, including ties, coincident points, singletons and hub families; and
ge_embed with device-resident coordinates == the same call with the radii computed on the host."""
import numpy as np
import pytest

from helpers import load_hier_golden, load_ref_hierarchy

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L", [1, 2, 3])
def test_device_radii_matches_reference_golden(ctx, L):
    As, Ps, z = load_hier_golden()
    AsL, PsL = As[-(L + 1):], Ps[-L:]
    pre = "radii_L%d_" % L
    if L == 1:
        cA, rA = ctx.level_radii(z[pre + "coords_A_in"], 2)
    else:
        cA, rA = ctx.level_radii(z[pre + "coords_A_in"], 2, AsL[1], PsL[1], z[pre + "coords_Ac"], z[pre + "r_Ac"])
    assert np.array_equal(cA, z[pre + "coords_A_out"])
    assert np.array_equal(rA, z[pre + "r_A_out"])


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("m", [1, 2, 3, 34, 97, 300])
def test_base_case_vs_oracle(ctx, oracle, dim, m):
    rng = np.random.default_rng(m * 10 + dim)
    for kind in ("random", "lattice", "duplicates"):
        if kind == "random":
            x = rng.normal(size=(m, dim))
        elif kind == "lattice":   # many exactly equal distances: the (t, i, j) tie-break decides
            x = rng.integers(0, 4, size=(m, dim)).astype(float) + np.arange(m)[:, None] * 1e-3 * (dim == 3)
        else:                     # coincident points: zero-radius balls stay "growing" (r <= 0)
            x = rng.normal(size=(m, dim))
            x[m // 2:] = x[:m - m // 2]
        c_ref, r_ref = oracle.radii(x, dim)
        c, r = ctx.level_radii(x, dim)
        assert np.array_equal(c, c_ref), kind
        assert np.array_equal(r, r_ref), kind


@pytest.mark.parametrize("dim", [2, 3])
def test_general_case_vs_oracle_and_host(ctx, capi, oracle, graphs, dim):
    A = graphs.rgg(6000, 10.0, seed=9)
    As, Ps = graphs.coarsen(A, 0.1, min_coarse=30)
    rng = np.random.default_rng(4)
    for l in range(len(Ps)):
        m, mc = As[l].shape[0], Ps[l].shape[0]
        x = rng.normal(size=(m, dim))
        if m > 8:
            x[3] = x[4]
        cAc, rAc = rng.normal(size=(mc, dim)), rng.random(mc) + 0.1
        c, r = ctx.level_radii(x, dim, As[l], Ps[l], cAc, rAc)
        ch, rh = capi.level_radii(x, dim, As[l], Ps[l], cAc, rAc)
        assert np.array_equal(c, ch) and np.array_equal(r, rh), l
        if m <= 2000:   # the reference's loop re-sorts after every pop: small levels only
            co, ro = oracle.radii(x, dim, As[l], Ps[l], cAc, rAc)
            assert np.array_equal(c, co) and np.array_equal(r, ro), l


def test_hub_families_vs_host(ctx, capi, graphs):
    """Families of thousands of members with 1e5+ intra-family edges (the reference partitioner's
    Delaunay hierarchy, levels 1 and 2): the wide (1024-thread) event loop."""
    As, Ps, _ = load_ref_hierarchy(graphs, "delaunay1000000")
    rng = np.random.default_rng(2)
    for l in (1, 2):
        m, mc = As[l].shape[0], Ps[l].shape[0]
        x = rng.normal(size=(m, 3))
        cAc, rAc = rng.normal(size=(mc, 3)), rng.random(mc) + 0.1
        c, r = ctx.level_radii(x, 3, As[l], Ps[l], cAc, rAc)
        ch, rh = capi.level_radii(x, 3, As[l], Ps[l], cAc, rAc)
        assert np.array_equal(r, rh) and np.array_equal(c, ch), l


def test_embed_device_radii_equals_host_radii(ctx, capi, graphs, monkeypatch):
    """ge_embed keeps the coordinates on the device between the levels; with GE_HOST_RADII the same
    call round-trips them through the host restatement.  Same seed -> same bits, including the
    out-parameters of embedMultilevel (level 1's radii and rescaled coordinates)."""
    As, Ps = graphs.coarsen(graphs.rgg(20000, 10.0, seed=5), 0.25, min_coarse=40)
    monkeypatch.delenv("GE_HOST_RADII", raising=False)
    x1, st1, r1, c1 = ctx.embed(As, Ps, 2, seed=7, coarse_iterations=3000, return_level1=True)
    monkeypatch.setenv("GE_HOST_RADII", "1")
    x2, st2, r2, c2 = ctx.embed(As, Ps, 2, seed=7, coarse_iterations=3000, return_level1=True)
    assert np.isfinite(x1).all()
    assert np.array_equal(x1, x2) and np.array_equal(r1, r2) and np.array_equal(c1, c2)
    assert st1["host_radii_ms"] == 0.0 and st2["host_radii_ms"] > 0.0
    # device-resident: the only device->host traffic is the result (+ a few bytes of counters)
    n = As[0].shape[0]
    assert st1["d2h_bytes"] <= n * 2 * 8 + (As[1].shape[0] * 3 * 8) + 4096
    x3, st3 = ctx.embed(As, Ps, 2, seed=7, coarse_iterations=3000)
    monkeypatch.delenv("GE_HOST_RADII", raising=False)
    x4, st4 = ctx.embed(As, Ps, 2, seed=7, coarse_iterations=3000)
    assert np.array_equal(x3, x4)
    assert st4["d2h_bytes"] <= n * 2 * 8 + 4096
