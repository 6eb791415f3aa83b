"""CPU: the stand-ins for Intel MKL and CSparse (oracle/shim) against scipy.

The reference programs behind tests/golden/ref_*.npz are linked with these routines instead of the real libraries
(oracle/Makefile), so the fixtures are only as good as the stand-ins: each routine the reference calls is checked here on
random matrices, with the calling conventions the reference uses (BCH.cpp:3006-3022, 3746-3758, 3865-3868, 4006-4043; 3349-3472)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I, D = C.POINTER(C.c_int), C.POINTER(C.c_double)


def ip(a):
    return a.ctypes.data_as(I)


def dp(a):
    return a.ctypes.data_as(D)


@pytest.fixture(scope="module")
def L():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libshim.so"], stdout=subprocess.DEVNULL)
    return C.CDLL(os.path.join(ROOT, "oracle", "libshim.so"))


class cs(C.Structure):
    _fields_ = [("nzmax", C.c_int), ("m", C.c_int), ("n", C.c_int), ("p", I), ("i", I), ("x", D), ("nz", C.c_int)]


def rand_csr(m, n, density, seed):
    A = sp.random(m, n, density=density, format="csr", random_state=seed, data_rvs=np.random.default_rng(seed).standard_normal)
    A.sort_indices()
    return A


@pytest.mark.parametrize("trans", [b"n", b"t"])
@pytest.mark.parametrize("beta", [0.0, 1.0])
def test_mkl_dcsrmv_zero_based_four_array(L, trans, beta):
    A = rand_csr(57, 31, 0.2, 1)
    rs = A.indptr.astype(np.int32)
    pb, pe = rs[:-1].copy(), rs[1:].copy()          # pointerB / pointerE = rowStarts, rowStartsMod (BCH.cpp:1814-1816)
    col, val = A.indices.astype(np.int32), A.data.copy()
    rng = np.random.default_rng(2)
    x = rng.standard_normal(31 if trans == b"n" else 57)
    y = rng.standard_normal(57 if trans == b"n" else 31)
    want = -0.7 * ((A @ x) if trans == b"n" else (A.T @ x)) + beta * y
    m, k, alpha, b = C.c_int(57), C.c_int(31), C.c_double(-0.7), C.c_double(beta)
    L.mkl_dcsrmv(trans, C.byref(m), C.byref(k), C.byref(alpha), b"gunc", dp(val), ip(col), ip(pb), ip(pe), dp(x), C.byref(b), dp(y))
    assert np.allclose(y, want, rtol=1e-13, atol=1e-13)


def test_mkl_dcsrsymv_and_jacobi_dcsrsv_one_based_upper(L):
    n = 64
    B = rand_csr(n, n, 0.15, 3)
    A = (B + B.T + sp.diags(np.full(n, 9.0))).tocsr()
    U = sp.triu(A, format="csr")
    U.sort_indices()
    ia, ja, a = (U.indptr + 1).astype(np.int32), (U.indices + 1).astype(np.int32), U.data.copy()
    x = np.random.default_rng(4).standard_normal(n)
    y = np.zeros(n)
    nn = C.c_int(n)
    for _ in range(2):                               # second call goes through the cached expansion
        L.mkl_dcsrsymv(b"u", C.byref(nn), dp(a), ip(ia), ip(ja), dp(x), dp(y))
        assert np.allclose(y, A @ x, rtol=1e-13, atol=1e-13)
    a *= 2.0                                         # edited in place: the cache must notice
    L.mkl_dcsrsymv(b"u", C.byref(nn), dp(a), ip(ia), ip(ja), dp(x), dp(y))
    assert np.allclose(y, 2.0 * (A @ x), rtol=1e-13, atol=1e-13)
    # diagonal solve with the reference's 3-character descriptor (BCH.cpp:3993-3996, 4039)
    z = np.zeros(n)
    one = C.c_double(1.0)
    L.mkl_dcsrsv(b"n", C.byref(nn), C.byref(one), b"dln", dp(a), ip(ja), ip(ia), ip(ia[1:].copy()), dp(x), dp(z))
    assert np.allclose(z, x / (2.0 * A.diagonal()), rtol=1e-14)


@pytest.mark.parametrize("precond", [0, 1])
def test_rci_cg_driver_like_the_reference(L, precond):
    """The MKL_CG_solver loop of the reference (BCH.cpp:3981-4055) around the stand-in: converges to the scipy solution,
    stops by the squared-norm rule, reports the iteration count."""
    n = 200
    B = rand_csr(n, n, 0.03, 5)
    A = (B @ B.T + sp.diags(np.linspace(1.0, 50.0, n))).tocsr()
    b = np.random.default_rng(6).standard_normal(n)
    x = np.zeros(n)
    ipar, dpar, tmp = np.zeros(128, np.int32), np.zeros(128), np.zeros(4 * n)
    nn, rci = C.c_int(n), C.c_int(0)
    L.dcg_init(C.byref(nn), dp(x), dp(b), C.byref(rci), ip(ipar), dp(dpar), dp(tmp))
    assert rci.value == 0 and ipar[4] == 150 and dpar[0] == 1e-6
    ipar[4], ipar[7], ipar[8], ipar[9], ipar[10] = 1000, 1, 1, 0, precond
    dpar[0] = 1e-12
    L.dcg_check(C.byref(nn), dp(x), dp(b), C.byref(rci), ip(ipar), dp(dpar), dp(tmp))
    assert rci.value == 0
    dinv = 1.0 / A.diagonal()
    requests = []
    while True:
        L.dcg(C.byref(nn), dp(x), dp(b), C.byref(rci), ip(ipar), dp(dpar), dp(tmp))
        requests.append(rci.value)
        if rci.value == 1:
            tmp[n:2 * n] = A @ tmp[:n]
        elif rci.value == 3:
            tmp[3 * n:4 * n] = dinv * tmp[2 * n:3 * n]
        else:
            break
    assert rci.value == 0 and (3 in requests) == bool(precond) and 2 not in requests
    its = C.c_int(0)
    L.dcg_get(C.byref(nn), dp(x), dp(b), C.byref(rci), ip(ipar), dp(dpar), dp(tmp), C.byref(its))
    r = b - A @ x
    assert 0 < its.value < 1000 and r @ r <= 1e-12 * (b @ b) * 1.01
    xs = spla.spsolve(A.tocsc(), b)
    assert np.linalg.norm(x - xs) <= 1e-4 * np.linalg.norm(xs)     # residual 1e-6 x conditioning
    # iteration limit: -1, like MKL, and the current iterate is left in x
    x2 = np.zeros(n)
    L.dcg_init(C.byref(nn), dp(x2), dp(b), C.byref(rci), ip(ipar), dp(dpar), dp(tmp))
    ipar[4], ipar[8], ipar[9] = 3, 1, 0
    dpar[0] = 1e-30
    while True:
        L.dcg(C.byref(nn), dp(x2), dp(b), C.byref(rci), ip(ipar), dp(dpar), dp(tmp))
        if rci.value == 1:
            tmp[n:2 * n] = A @ tmp[:n]
        else:
            break
    assert rci.value == -1 and ipar[3] == 3 and np.linalg.norm(b - A @ x2) < np.linalg.norm(b)


def test_csparse_compress_transpose_multiply(L):
    """Z = G^T diag(d) G the way calculateZ() forms it (BCH.cpp:3349-3472): triplets with caller-owned arrays ->
    cs_compress -> cs_transpose -> cs_multiply -> cs_transpose (sorts the columns)."""
    for f in ("cs_spalloc", "cs_compress", "cs_transpose", "cs_multiply"):
        getattr(L, f).restype = C.POINTER(cs)
    L.cs_spalloc.argtypes = [C.c_int] * 5
    L.cs_compress.argtypes = [C.POINTER(cs)]
    L.cs_transpose.argtypes = [C.POINTER(cs), C.c_int]
    L.cs_multiply.argtypes = [C.POINTER(cs), C.POINTER(cs)]
    G = rand_csr(40, 17, 0.25, 7).tocoo()
    d = np.random.default_rng(8).uniform(0.5, 2.0, 40)
    rows, cols = G.row.astype(np.int32), G.col.astype(np.int32)
    vals, scaled = G.data.copy(), G.data * d[G.row]

    def triplet(v):
        T = L.cs_spalloc(40, 17, len(v), 1, 1)
        T.contents.i, T.contents.p, T.contents.x, T.contents.nz = ip(rows), ip(cols), dp(v), len(v)
        return T

    Gc = L.cs_compress(triplet(vals))
    Gt = L.cs_transpose(Gc, 1)
    Dc = L.cs_compress(triplet(scaled))
    Z = L.cs_multiply(Gt, Dc)
    Zs = L.cs_transpose(Z, 1)

    def to_scipy(M, sort_check=False):
        c = M.contents
        assert c.nz == -1
        nnz = c.p[c.n]
        assert c.nzmax == max(nnz, 1)                 # the reference loops to nzmax (BCH.cpp:3427, 3494)
        ind = np.array(c.i[:nnz]); ptr = np.array(c.p[:c.n + 1]); x = np.array(c.x[:nnz])
        if sort_check:
            assert all(np.all(np.diff(ind[ptr[j]:ptr[j + 1]]) > 0) for j in range(c.n))
        return sp.csc_matrix((x, ind, ptr), shape=(c.m, c.n))

    Gs = G.tocsr()
    assert abs(to_scipy(Gc) - Gs).max() == 0 and abs(to_scipy(Gt, True) - Gs.T).max() == 0
    want = (Gs.T @ sp.diags(d) @ Gs).toarray()
    assert np.allclose(to_scipy(Z).toarray(), want, rtol=1e-13, atol=1e-14)
    assert np.allclose(to_scipy(Zs, True).toarray(), want.T, rtol=1e-13, atol=1e-14)
    # same pattern, same entry order for a second product with other values (the reference adds x arrays entry by entry, :3427)
    doubled = 2.0 * scaled                             # the triplet only borrows the array
    Z2 = L.cs_multiply(Gt, L.cs_compress(triplet(doubled)))
    n1 = Z.contents.p[Z.contents.n]
    assert np.array_equal(np.array(Z.contents.i[:n1]), np.array(Z2.contents.i[:n1]))
    assert np.allclose(np.array(Z2.contents.x[:n1]), 2.0 * np.array(Z.contents.x[:n1]))


def test_paralution_bicgstab_standin(L, tmp_path, monkeypatch):
    """LocalMatrix / LocalVector / Jacobi / BiCGStab with the reference's call sequence (GQ.cpp:4183-4231) on a
    non-symmetric, diagonally dominant matrix: scipy's solution, residual below rel_tol, iteration count logged."""
    n = 300
    B = rand_csr(n, n, 0.02, 9)
    A = (B + sp.diags(np.linspace(4.0, 9.0, n))).tocsr()
    A.sort_indices()
    rs, col, val = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    b = np.random.default_rng(10).standard_normal(n)
    x = np.full(n, 7.0)                                # Zeros() must reset the initial guess (GQ.cpp:4199)
    log = tmp_path / "shim.log"
    monkeypatch.setenv("REF_SHIM_LOG", str(log))
    L.shim_paralution_bicgstab.argtypes = [C.c_int, C.c_int, I, I, D, D, D, C.c_double, C.c_double, C.c_double, C.c_int]
    rc = L.shim_paralution_bicgstab(n, len(val), ip(rs), ip(col), dp(val), dp(b), dp(x), 1e-15, 1e-6, 1e8, 1000)
    assert rc == 0
    assert np.linalg.norm(b - A @ x) <= 1e-6 * np.linalg.norm(b)
    xs = spla.spsolve(A.tocsc(), b)
    assert np.linalg.norm(x - xs) <= 1e-5 * np.linalg.norm(xs)
    kind, its, ok = log.read_text().split()
    assert kind == "bicgstab" and 0 < int(its) < 100 and ok == "1"
    # the tolerance override used for the "tight" fixtures
    monkeypatch.setenv("REF_SHIM_BICG_TOL", "1e-12")
    x[:] = 0
    assert L.shim_paralution_bicgstab(n, len(val), ip(rs), ip(col), dp(val), dp(b), dp(x), 1e-15, 1e-6, 1e8, 1000) == 0
    assert np.linalg.norm(b - A @ x) <= 1e-11 * np.linalg.norm(b)
    # iteration limit reached: not converged, logged as such
    monkeypatch.setenv("REF_SHIM_BICG_TOL", "1e-300")
    assert L.shim_paralution_bicgstab(n, len(val), ip(rs), ip(col), dp(val), dp(b), dp(x), 0.0, 1e-6, 1e8, 5) == 0
    assert log.read_text().strip().split("\n")[-1] == "bicgstab 5 0"


@pytest.mark.parametrize("mk", [dict(n=4, cluster=2.0), dict(n=3, ny=4, nz=5, pin=1)])
def test_integration_shim_explicit_Z(L, mesh, mk):
    """integration/cfdb200_explicit_z.hpp (what calculateZ_CUSP() of the replacement shim runs) against the product's own step0
    on the same G, MdOrigInv: same pattern, same values, same pin rule (node id > 0)."""
    import fracstep
    s = fracstep.Solver(mesh("shimz_%d_%s" % (mk["n"], mk.get("pin", "d")), **mk), "explicit", init_gpu=False)
    NN, NNp = s.i("NN"), s.i("NNp")
    grs, gcol = s.arr("sparseGrowStarts"), s.arr("sparseGcol")
    G = [np.ascontiguousarray(s.arr("sparseG%dvalue" % c)) for c in (1, 2, 3)]
    mdo = np.ascontiguousarray(s.arr("MdOrigInv"))
    zrs_p, zcol_p, zval_p = s.arr("sparseZrowStarts"), s.arr("sparseZcol"), s.arr("sparseZvalue")
    cap = len(zcol_p) + 16
    zrs, zcol, zval = np.zeros(NNp + 1, np.int32), np.zeros(cap, np.int32), np.zeros(cap)
    L.shim_form_explicit_Z.argtypes = [C.c_int, C.c_int, I, I, D, D, D, D, C.c_int, I, I, D, C.c_int]
    n = L.shim_form_explicit_Z(NN, NNp, ip(np.ascontiguousarray(grs, np.int32)), ip(np.ascontiguousarray(gcol, np.int32)), dp(G[0]), dp(G[1]),
                               dp(G[2]), dp(mdo), s.i("zeroPressureNode"), ip(zrs), ip(zcol), dp(zval), cap)
    assert n == len(zcol_p)
    assert np.array_equal(zrs, zrs_p) and np.array_equal(zcol[:n], zcol_p)
    assert np.allclose(zval[:n], zval_p, rtol=1e-12, atol=1e-14 * np.abs(zval_p).max())
