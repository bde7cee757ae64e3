"""Mints tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libge_ref_strict.so, built by
oracle/Makefile from the unmodified sources under /root/reference).  Run in the build container
(the GPU box has no /root/reference):  python tests/golden/make_golden.py

The reference itself has no golden vectors for this path, so these files are the
pin: inputs + the reference's own outputs, bit for bit.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
O = entry.load_oracle()
from graph_embed_b200 import graphs as G  # noqa: E402


def pack_levels(As, Ps):
    d = {"L": np.int32(len(Ps))}
    for l, A in enumerate(As):
        d["A%d_indptr" % l], d["A%d_indices" % l], d["A%d_data" % l] = A.indptr, A.indices, A.data
    for l, P in enumerate(Ps):
        d["P%d_indptr" % l], d["P%d_indices" % l] = P.indptr, P.indices
    return d


def mint_galerkin():
    """galerkin_grid30.npz: `P.Mult(A).Mult(P.Transpose())` (examples/embedder.cpp:215) of every level of
    the grid-30 hierarchy (aggregation from the reference's own partitioner), evaluated by the
    compiled reference driver with the stand-in linalgcpp container; plus one real-weight level."""
    O.build(ref=True)
    assert O.ref_available("strict"), "oracle/_ref missing (needs /root/reference)"
    A = G.grid2d(30, 30)
    Ps = O.ref_partition(A, 0.25, matching_iterations=2, nthreads=1)
    out = {"L": np.int32(len(Ps))}
    cur = G.canonical(A)
    for l, P in enumerate(Ps):
        out["P%d_indptr" % l], out["P%d_indices" % l] = P.indptr, P.indices
        C = O.ref_galerkin(cur, P)
        out["C%d_indptr" % l], out["C%d_indices" % l], out["C%d_data" % l] = C.indptr, C.indices, C.data
        cur = C
    B = G.canonical(A).copy()
    B.data = np.random.default_rng(4).uniform(0.5, 2.0, B.nnz)
    C = O.ref_galerkin(B, Ps[0])
    out["B_data"], out["CB_indptr"], out["CB_indices"], out["CB_data"] = B.data, C.indptr, C.indices, C.data
    np.savez_compressed(os.path.join(HERE, "galerkin_grid30.npz"), **out)
    print("wrote galerkin_grid30.npz")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "galerkin":
        return mint_galerkin()
    O.build(ref=True)
    assert O.ref_available("strict"), "oracle/_ref missing (needs /root/reference)"
    # ---- flat kernel: positions after k iterations from fixed initial coordinates ---------------
    A = G.grid2d(12, 12)
    out = {"indptr": A.indptr, "indices": A.indices, "data": A.data}
    for dim in (2, 3):
        x0 = O.mt_uniform(7, A.shape[0] * dim).reshape(-1, dim)
        out["x0_d%d" % dim] = x0
        for k in (1, 2, 5, 25, 100):
            out["x_d%d_k%d" % (dim, k)] = O.ref_flat(A, dim, x0, O.Params(iterations=k), nthreads=1)
    out["x_d2_k7_linlog"] = O.ref_flat(A, 2, out["x0_d2"], O.Params(iterations=7, linlog=True))
    out["x_d2_k7_nohubs_delta"] = O.ref_flat(A, 2, out["x0_d2"], O.Params(iterations=7, nohubs=True, delta=0.5))
    out["x_d2_k7_normalize"] = O.ref_flat(A, 2, out["x0_d2"], O.Params(iterations=7, normalize=True))
    np.savez_compressed(os.path.join(HERE, "flat_grid12.npz"), **out)

    # ---- hierarchy from the reference's own partitioner; multilevel kernel, radii, embed --------
    A = G.grid2d(30, 30)
    Ps = O.ref_partition(A, 0.25, matching_iterations=2, nthreads=1)
    As = G.hierarchy_from(A, Ps)
    out = pack_levels(As, Ps)
    rng = np.random.default_rng(1)
    for dim in (2, 3):
        for l in (0, 1):
            m = Ps[l].shape[0]
            cA, rA = rng.normal(size=(m, dim)), rng.random(m) * 0.3 + 0.05
            out["ml_cA_l%d_d%d" % (l, dim)], out["ml_rA_l%d_d%d" % (l, dim)] = cA, rA
            for k in (1, 3, 100):
                out["ml_x_l%d_d%d_k%d" % (l, dim, k)] = O.ref_multilevel(
                    As[l], Ps[l], cA, rA, dim, O.Params(iterations=k), seed=5, nthreads=1)
    for L in (1, 2, 3):
        r = O.ref_radii_case(As[-(L + 1):], Ps[-L:], 2, seed=4)
        for key, val in r.items():
            out["radii_L%d_%s" % (L, key)] = val
    x, _ = O.ref_embed(As, Ps, 2, seed=21, nthreads=1)
    out["embed_d2_seed21"] = x
    np.savez_compressed(os.path.join(HERE, "hier_grid30.npz"), **out)
    # ---- BASELINE config 1: 100 x 100 grid, coarsening 0.25, d = 2, the reference's own partition ---
    A = G.grid2d(100, 100)
    Ps = O.ref_partition(A, 0.25, matching_iterations=2, nthreads=1)
    As = G.hierarchy_from(A, Ps)
    out = {"L": np.int32(len(Ps))}
    for l, P in enumerate(Ps):
        out["P%d_indptr" % l], out["P%d_indices" % l] = P.indptr.astype(np.int32), P.indices.astype(np.int32)
    x, secs = O.ref_embed(As, Ps, 2, seed=3, nthreads=1)
    out["embed_d2_seed3"] = x.astype(np.float32)   # statistics only: single precision keeps the file small
    out["ref_embed_seconds_1thread"] = np.float64(secs)
    np.savez_compressed(os.path.join(HERE, "config1_grid100.npz"), **out)
    for f in ("flat_grid12.npz", "hier_grid30.npz", "config1_grid100.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
