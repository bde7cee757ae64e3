"""ORACLE -- TEST INFRASTRUCTURE ONLY (ctypes bindings).

Binds oracle/liboracle.so (the plain-C restatement, forceatlas_oracle.c) and, when
present, oracle/_ref/libge_ref_{strict,fast}.so (the reference's unmodified sources
compiled by oracle/Makefile).  Importable only from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; nothing in graph-embed_b200/
may import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"


class Params(C.Structure):
    """fa_params of forceatlas_oracle.c; defaults = include/forceatlas.hpp:92-103."""
    _fields_ = [("iterations", C.c_int), ("ks", C.c_double), ("ksmax", C.c_double),
                ("repel", C.c_double), ("attract", C.c_double), ("gravity", C.c_double),
                ("delta", C.c_double), ("tolerate", C.c_double), ("useWeights", C.c_int),
                ("linlog", C.c_int), ("nohubs", C.c_int), ("normalize", C.c_int)]

    def __init__(self, iterations=100000, ks=0.1, ksmax=1.0, repel=1.0, attract=1.0, gravity=1.0,
                 delta=1.0, tolerate=1.0, useWeights=True, linlog=False, nohubs=False,
                 normalize=False):
        super().__init__(int(iterations), ks, ksmax, repel, attract, gravity, delta, tolerate,
                         int(useWeights), int(linlog), int(nohubs), int(normalize))


def build(ref=True):
    """Compile liboracle.so and (where /root/reference exists) oracle/_ref/*.so."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_pi = C.POINTER(C.c_int)
_pd = C.POINTER(C.c_double)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    if a is None:
        return None
    if a.dtype == np.int32:
        return a.ctypes.data_as(_pi)
    return a.ctypes.data_as(_pd)


def csr_arrays(A):
    """scipy CSR -> (n, indptr int32, indices int32, data float64)."""
    return A.shape[0], _i32(A.indptr), _i32(A.indices), _f64(A.data)


_oracle = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        path = os.path.join(HERE, "liboracle.so")
        src = os.path.join(HERE, "forceatlas_oracle.c")
        if not os.path.exists(path) or os.path.getmtime(src) > os.path.getmtime(path):
            build(ref=False)
        _oracle = C.CDLL(path)
        _oracle.oracle_embed.restype = C.c_int
    return _oracle


def flat_forces(A, dim, coords, params=None, rows=None):
    """Forces of one flat iteration (include/forceatlas.hpp:148-212) -> (forces, fscale)."""
    params = params or Params()
    n, I, J, D = csr_arrays(A)
    x = _f64(coords).reshape(n, dim)
    r0, r1 = rows if rows is not None else (0, n)
    F = np.zeros((n, dim))
    S = np.zeros(n)
    oracle_lib().oracle_flat_forces(n, _p(I), _p(J), _p(D), dim, _p(x), C.byref(params),
                                    int(r0), int(r1), _p(F), _p(S))
    return F, S


def flat_run(A, dim, coords, params):
    """`iterations` flat iterations from `coords` -> (coords, forces_of_last_iteration)."""
    n, I, J, D = csr_arrays(A)
    x = _f64(coords).reshape(n, dim).copy()
    F = np.zeros((n, dim))
    oracle_lib().oracle_flat_run(n, _p(I), _p(J), _p(D), dim, _p(x), C.byref(params), _p(F))
    return x, F


def vertex_to_aggregate(P_T):
    """P_T.Transpose().GetIndices() of src/embed.cpp:605."""
    v_A = np.empty(P_T.shape[1], dtype=np.int32)
    v_A[P_T.indices] = np.repeat(np.arange(P_T.shape[0], dtype=np.int32), np.diff(P_T.indptr))
    return v_A


def multilevel_run(A, P_T, coords_A, r_A, dim, init, params, forces_iter=None, aggregates=None):
    """forceAtlasMultilevel (include/forceatlas.hpp:314-574) from caller-supplied init.

    init: n x dim by global vertex id.  Returns coords, or (coords, forces, fscale) of
    iteration `forces_iter` when requested."""
    n, I, J, D = csr_arrays(A)
    m = P_T.shape[0]
    PI, PJ = _i32(P_T.indptr), _i32(P_T.indices)
    v_A = vertex_to_aggregate(P_T)
    cA, rA, x0 = _f64(coords_A).reshape(m, dim), _f64(r_A), _f64(init).reshape(n, dim)
    out = np.zeros((n, dim))
    F = S = None
    if forces_iter is not None:
        F, S = np.zeros((n, dim)), np.zeros(n)
    a0, a1 = aggregates if aggregates is not None else (0, m)   # rows of other aggregates stay 0
    oracle_lib().oracle_multilevel_run_range(n, _p(I), _p(J), _p(D), m, _p(PI), _p(PJ), _p(v_A), _p(cA),
                                             _p(rA), dim, _p(x0), C.byref(params), _p(out),
                                             int(forces_iter or 0), _p(F), _p(S), int(a0), int(a1))
    return out if forces_iter is None else (out, F, S)


def radii(coords_A, dim, A_c=None, P_T_c=None, coords_Ac=None, r_Ac=None):
    """src/embed.cpp:615-778 -> (rescaled coords_A, r_A).  Base case when P_T_c is None."""
    cA = _f64(coords_A).reshape(-1, dim).copy()
    m = cA.shape[0]
    rA = np.zeros(m)
    if P_T_c is None:
        oracle_lib().oracle_radii(m, dim, _p(cA), _p(rA), None, None, 0, None, None, None, None)
    else:
        AcI, AcJ = _i32(A_c.indptr), _i32(A_c.indices)
        PcI, PcJ = _i32(P_T_c.indptr), _i32(P_T_c.indices)
        cAc, rAc = _f64(coords_Ac).reshape(-1, dim), _f64(r_Ac)
        oracle_lib().oracle_radii(m, dim, _p(cA), _p(rA), _p(AcI), _p(AcJ), P_T_c.shape[0],
                                  _p(PcI), _p(PcJ), _p(cAc), _p(rAc))
    return cA, rA


def galerkin(A, P_T):
    """oracle_galerkin: A_c = P_T A P_T^T (examples/embedder.cpp:213-216) -> scipy CSR."""
    import scipy.sparse as sp
    L = oracle_lib()
    L.oracle_galerkin.restype = C.c_long
    n, I, J, D = csr_arrays(A)
    PI, PJ = _i32(P_T.indptr), _i32(P_T.indices)
    m = P_T.shape[0]
    ptr = np.zeros(m + 1, dtype=np.int32)
    idx = np.zeros(max(A.nnz, 1), dtype=np.int32)
    val = np.zeros(max(A.nnz, 1))
    nnz = L.oracle_galerkin(n, m, _p(I), _p(J), _p(D), _p(PI), _p(PJ), _p(ptr), _p(idx), _p(val))
    return sp.csr_matrix((val[:nnz].copy(), idx[:nnz].copy(), ptr), shape=(m, m))


def mt_uniform(seed, count):
    """std::mt19937(seed) through libstdc++'s uniform_real_distribution<double>(-1,1)."""
    out = np.zeros(int(count))
    fn = oracle_lib().oracle_mt19937_uniform
    fn.argtypes = [C.c_uint32, C.c_size_t, _pd]
    fn(int(seed), int(count), _p(out))
    return out


def multilevel_init(P_T, dim, seed):
    """Initial local coordinates in the reference's draw order (forceatlas.hpp:341,356-358):
    aggregate-major, member-major, k-minor; returned n x dim by global vertex id."""
    n = P_T.shape[1]
    stream = mt_uniform(seed, n * dim).reshape(n, dim)
    init = np.empty((n, dim))
    init[P_T.indices] = stream
    return init


class _Levels:
    """Marshals (As, P_Ts) into the pointer arrays the C entry points take."""

    def __init__(self, As, P_Ts):
        self.keep = []
        L = len(P_Ts)
        assert len(As) == L + 1
        self.L = L
        self.An = (C.c_int * (L + 1))(*[A.shape[0] for A in As])
        self.Pm = (C.c_int * max(L, 1))(*([P.shape[0] for P in P_Ts] or [0]))
        self.AI, self.AJ, self.AD = (_pi * (L + 1))(), (_pi * (L + 1))(), (_pd * (L + 1))()
        self.PI, self.PJ = (_pi * max(L, 1))(), (_pi * max(L, 1))()
        for l, A in enumerate(As):
            _, I, J, D = csr_arrays(A)
            self.keep += [I, J, D]
            self.AI[l], self.AJ[l], self.AD[l] = _p(I), _p(J), _p(D)
        for l, P in enumerate(P_Ts):
            I, J = _i32(P.indptr), _i32(P.indices)
            self.keep += [I, J]
            self.PI[l], self.PJ[l] = _p(I), _p(J)

    def args(self):
        return (self.L, self.An, self.AI, self.AJ, self.AD, self.Pm, self.PI, self.PJ)


def embed(As, P_Ts, dim, seed, coarse_iterations=100000, level_iterations=100):
    """src/embed.cpp:561-796 restated, every stream = mt19937(seed)."""
    lv = _Levels(As, P_Ts)
    out = np.zeros((As[0].shape[0], dim))
    oracle_lib().oracle_embed(*lv.args(), dim, C.c_uint32(int(seed)), int(coarse_iterations),
                              int(level_iterations), _p(out))
    return out


# --------------------------------------------------------------------------- #
# the compiled reference (oracle/_ref)                                         #
# --------------------------------------------------------------------------- #
_ref = {}


def ref_available(kind="strict"):
    return os.path.exists(os.path.join(HERE, "_ref", "libge_ref_%s.so" % kind))


def ref_lib(kind="strict"):
    if kind not in _ref:
        lib = C.CDLL(os.path.join(HERE, "_ref", "libge_ref_%s.so" % kind))
        lib.ref_embed.restype = C.c_double
        lib.ref_max_threads.restype = C.c_int
        _ref[kind] = lib
    return _ref[kind]


def ref_flat(A, dim, coords, params, seed=0, nthreads=1, kind="strict"):
    """partition::forceAtlas (15-arg).  coords=None -> the reference draws its own init."""
    lib = ref_lib(kind)
    n, I, J, D = csr_arrays(A)
    lib.ref_set_seed(C.c_uint(int(seed)))
    random_init = coords is None
    x = np.zeros((n, dim)) if random_init else _f64(coords).reshape(n, dim).copy()
    lib.ref_flat_forceatlas(n, _p(I), _p(J), _p(D), dim, _p(x), int(random_init),
                            C.byref(params), int(nthreads))
    return x


def ref_multilevel(A, P_T, coords_A, r_A, dim, params, seed, nthreads=1, kind="strict"):
    """partition::forceAtlasMultilevel (18-arg), init drawn from mt19937(seed)."""
    lib = ref_lib(kind)
    n, I, J, D = csr_arrays(A)
    m = P_T.shape[0]
    PI, PJ = _i32(P_T.indptr), _i32(P_T.indices)
    v_A = vertex_to_aggregate(P_T)
    cA, rA = _f64(coords_A).reshape(m, dim), _f64(r_A)
    out = np.zeros((n, dim))
    lib.ref_set_seed(C.c_uint(int(seed)))
    lib.ref_multilevel_forceatlas(n, _p(I), _p(J), _p(D), m, _p(PI), _p(PJ), _p(v_A), _p(cA),
                                  _p(rA), dim, C.byref(params), _p(out), int(nthreads))
    return out


def ref_embed(As, P_Ts, dim, seed, nthreads=1, kind="strict"):
    """partition::embed -> (coords, wall seconds bracketed as examples/embedder.cpp:219-222)."""
    lib = ref_lib(kind)
    lv = _Levels(As, P_Ts)
    out = np.zeros((As[0].shape[0], dim))
    lib.ref_set_seed(C.c_uint(int(seed)))
    secs = lib.ref_embed(*lv.args(), dim, _p(out), int(nthreads))
    return out, secs


def ref_radii_case(As, P_Ts, dim, seed, kind="strict"):
    """Inputs and outputs of the reference's radii/rescale step at level 0."""
    lib = ref_lib(kind)
    lv = _Levels(As, P_Ts)
    m = As[1].shape[0]
    mc = As[2].shape[0] if len(As) > 2 else 0
    cA_in, cA_out, rA_out = np.zeros((m, dim)), np.zeros((m, dim)), np.zeros(m)
    rAc, cAc = np.zeros(max(mc, 1)), np.zeros((max(mc, 1), dim))
    lib.ref_set_seed(C.c_uint(int(seed)))
    nrc = lib.ref_radii_case(*lv.args(), dim, _p(cA_in), _p(rAc), _p(cAc), _p(cA_out), _p(rA_out))
    return dict(coords_A_in=cA_in, r_Ac=rAc[:nrc], coords_Ac=cAc[:nrc] if nrc else cAc[:0],
                coords_A_out=cA_out, r_A_out=rA_out)


def ref_galerkin(A, P_T, kind="strict"):
    """examples/embedder.cpp:215 through the compiled reference driver (stand-in linalgcpp)."""
    import scipy.sparse as sp
    L = ref_lib(kind)
    L.ref_galerkin.restype = C.c_long
    n, I, J, D = csr_arrays(A)
    PI, PJ = _i32(P_T.indptr), _i32(P_T.indices)
    m = P_T.shape[0]
    ptr = np.zeros(m + 1, dtype=np.int32)
    idx = np.zeros(max(A.nnz, 1), dtype=np.int32)
    val = np.zeros(max(A.nnz, 1))
    nnz = L.ref_galerkin(n, _p(I), _p(J), _p(D), m, _p(PI), _p(PJ), _p(ptr), _p(idx), _p(val))
    return sp.csr_matrix((val[:nnz].copy(), idx[:nnz].copy(), ptr), shape=(m, m))


def ref_write_coords(coords, path, kind="strict"):
    """partition::writeCoords (src/export.cpp:27-39) of the compiled reference."""
    x = _f64(coords)
    ref_lib(kind).ref_write_coords(_p(x), int(x.shape[0]), int(x.shape[1]), str(path).encode())


def ref_write_partition(part, path, kind="strict"):
    """partition::writePartition (src/export.cpp:16-25) of the compiled reference."""
    q = _i32(part)
    ref_lib(kind).ref_write_partition(_p(q), int(q.shape[0]), str(path).encode())


def ref_partition(A, coarsening_factor, matching_iterations=2, nthreads=8, kind="fast"):
    """partition::partition(A, cf, false, true, 1.0, matchingIterations, false) -> [P_T csr]."""
    import scipy.sparse as sp
    lib = ref_lib(kind)
    n, I, J, D = csr_arrays(A)
    L = lib.ref_partition(n, _p(I), _p(J), _p(D), C.c_double(coarsening_factor),
                          int(matching_iterations), int(nthreads))
    out = []
    for l in range(L):
        rows, cols = lib.ref_hierarchy_rows(l), lib.ref_hierarchy_cols(l)
        PI, PJ = np.zeros(rows + 1, dtype=np.int32), np.zeros(cols, dtype=np.int32)
        lib.ref_hierarchy_get(l, _p(PI), _p(PJ))
        out.append(sp.csr_matrix((np.ones(cols), PJ, PI), shape=(rows, cols)))
    return out
