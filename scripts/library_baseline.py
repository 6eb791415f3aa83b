#!/usr/bin/env python
"""Pressure solve: fused persistent kernel (k_pcg) vs the reference's GPU call sequence re-created on cuSPARSE + cuBLAS
(baseline/lib_pcg.cu) -- same matrix, same right-hand side (R2 of a real time step), same stopping rule, same B200."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cfd-with-cuda_b200"))
import fracstep  # noqa: E402
import meshgen  # noqa: E402

subprocess.check_call(["make", "-C", os.path.join(ROOT, "baseline")], stdout=subprocess.DEVNULL)
LB = C.CDLL(os.path.join(ROOT, "baseline", "liblibpcg.so"))
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
LB.lib_pcg.argtypes = [C.c_int, C.c_int, ip, ip, dp, dp, dp, dp, C.c_double, C.c_int, ip, dp]


def lib_pcg(rs, cl, v, b, tol=1e-6, maxit=1000):
    n = len(rs) - 1
    dinv = np.ones(n)
    rows = np.repeat(np.arange(n), np.diff(rs))
    d = rows == cl
    dinv[rows[d]] = 1.0 / v[d]
    x = np.zeros(n)
    its, ms = C.c_int(), C.c_double()
    best = 1e30
    for _ in range(3):
        rc = LB.lib_pcg(n, len(cl), rs.ctypes.data_as(ip), cl.ctypes.data_as(ip), v.ctypes.data_as(dp), dinv.ctypes.data_as(dp),
                        b.ctypes.data_as(dp), x.ctypes.data_as(dp), tol, maxit, C.byref(its), C.byref(ms))
        assert rc == 0
        best = min(best, ms.value)
    return x, its.value, best


def case(name, solver, steps=2):
    c = solver.ctx
    c.time_loop(steps)
    c.profile(True)
    c.time_loop(1)
    p = c.profile_get("pcg")
    c.profile(False)
    R2 = c.download("R2")
    x_ours = c.download("Pdot" if solver.scheme == fracstep.EXPLICIT else "Pdiff")
    its_ours = c.last_cg()
    rs = np.ascontiguousarray(solver.arr("sparseZrowStarts"), np.int32)
    cl = np.ascontiguousarray(solver.arr("sparseZcol"), np.int32)
    v = np.ascontiguousarray(solver.arr("sparseZvalue"), np.float64)
    x_lib, its_lib, ms_lib = lib_pcg(rs, cl, v, np.ascontiguousarray(R2))
    ms_ours = p["ms"] / p["launches"]          # the last solve of the step has the same R2 (explicit: last inner iteration)
    err = np.linalg.norm(x_ours - x_lib) / max(np.linalg.norm(x_lib), 1e-300)
    print("%-28s n=%9d nnz=%11d | k_pcg %8.3f ms (%4d it, %6.2f us/it) | cuSPARSE+cuBLAS %9.3f ms (%4d it, %7.2f us/it) | speed-up %5.1fx | rel diff %.1e"
          % (name, len(rs) - 1, len(cl), ms_ours, its_ours, 1e3 * ms_ours / max(its_ours, 1), ms_lib, its_lib, 1e3 * ms_lib / max(its_lib, 1),
             (ms_lib / max(its_lib, 1)) / (ms_ours / max(its_ours, 1)), err))


if __name__ == "__main__":
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ne27000.inp")
        meshgen.cube(path, 30, cluster=2.0)
        case("NE27000 explicit (Z 125-pt)", fracstep.Solver(path, "explicit"))
        case("NE27000 implicit (Z 27-pt)", fracstep.Solver(path, "implicit"))
    case("128^3 Q1 implicit (Z 27-pt)", fracstep.Solver(scheme="implicit", cavity=dict(nx=128, NENv=8)))
