"""ctypes binding of libcfdb200.so (include/cfdb200.h) -- the Python face of the product.

The hot path is the CUDA library; this module only marshals arrays.  There is no CPU
fallback: every compute entry point raises if the library or a CUDA device is missing.

Classes
  Problem  -- the `cfd_problem` POD built from numpy arrays carrying the reference's global names
  Context  -- `cfd_ctx`: the `...GPU()` seam (bch_* explicit, gq_* implicit) + cfd_timeLoop
  Solver   -- `cfd_solver`: .inp / in-memory cavity -> preprocessing -> step0 -> Context
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.environ.get("CFD_LIBPATH") or os.path.join(_HERE, "libcfdb200.so")   # CFD_LIBPATH: an instrumented build (diagnostics)
_lib = None

EXPLICIT, IMPLICIT = 0, 1
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class CfdError(RuntimeError):
    pass


class cfd_problem(C.Structure):
    _fields_ = [
        ("scheme", C.c_int), ("NE", C.c_int), ("NN", C.c_int), ("NNp", C.c_int),
        ("NENv", C.c_int), ("NENp", C.c_int), ("NGP", C.c_int),
        ("dt", C.c_double), ("density", C.c_double), ("viscosity", C.c_double),
        ("maxIter", C.c_int), ("tolerance", C.c_double), ("convergenceCriteria", C.c_double),
        ("coord", _dp), ("LtoGnode", _ip),
        ("sparseM_nnz", C.c_int), ("sparseMrowStarts", _ip), ("sparseMcol", _ip),
        ("sparseKvalue", _dp), ("sparseMvalue", _dp),
        ("sparseG_nnz", C.c_int), ("sparseGrowStarts", _ip), ("sparseGcol", _ip),
        ("sparseG1value", _dp), ("sparseG2value", _dp), ("sparseG3value", _dp),
        ("sparseZ_nnz", C.c_int), ("sparseZrowStarts", _ip), ("sparseZcol", _ip), ("sparseZvalue", _dp),
        ("sparseMapM_1d", _ip), ("MdInv", _dp), ("MdOrigInv", _dp),
        ("BCnVelNodes", C.c_int), ("BCvelNodes", _ip), ("nBC", C.c_int), ("BCstr", _dp),
        ("zeroPressureNode", C.c_int), ("monPoint", C.c_int), ("U0", _dp), ("P0", _dp),
        ("world", C.c_int), ("rank", C.c_int), ("NN_owned", C.c_int), ("NNp_owned", C.c_int), ("NE_apply", C.c_int),
        ("pinLocal", C.c_int), ("monOwned", C.c_int), ("nPeers", C.c_int),
        ("peerRank", _ip), ("sendVptr", _ip), ("sendVidx", _ip), ("recvVoff", _ip), ("recvVcnt", _ip),
        ("sendPptr", _ip), ("sendPidx", _ip), ("recvPoff", _ip), ("recvPcnt", _ip),
        ("remoteVoff", _ip), ("remotePoff", _ip),
    ]


class cfd_options(C.Structure):
    _fields_ = [
        ("cg_tol", C.c_double), ("cg_maxit", C.c_int), ("cg_warm_start", C.c_int),
        ("bicg_tol", C.c_double), ("bicg_maxit", C.c_int), ("maxacc_rule", C.c_int),
        ("device", C.c_int), ("ebe_scatter", C.c_int), ("cg_no_precond", C.c_int), ("deterministic", C.c_int), ("bicg_no_halfstep", C.c_int), ("bicg_no_verify", C.c_int), ("cg_variant", C.c_int),
        ("operator_path", C.c_int), ("step0_device", C.c_int), ("cuda_stream", C.c_void_p),
    ]


class cfd_step_report(C.Structure):
    _fields_ = [
        ("timeN", C.c_int), ("iter", C.c_int), ("cg_iterations", C.c_int), ("cg_solves", C.c_int),
        ("bicg_iterations", C.c_int * 3), ("maxAcc", C.c_double), ("steady", C.c_int), ("monitor", C.c_double * 4),
        ("cg_converged", C.c_int), ("bicg_converged", C.c_int * 3), ("bicg_restarts", C.c_int * 3),
        ("cg_residual", C.c_double), ("bicg_true_residual", C.c_double * 3),
    ]


def build(force=False):
    """Compile libcfdb200.so in-tree (nvcc, sm_100a)."""
    src = os.path.join(_HERE, "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src)
                 if f.endswith((".cu", ".cuh", ".cpp", ".hpp")))
    hdr = os.path.getmtime(os.path.join(_HERE, "..", "include", "cfdb200.h"))
    if force or not os.path.exists(_LIBPATH) or os.path.getmtime(_LIBPATH) < max(newest, hdr):
        subprocess.check_call(["make", "-C", src], stdout=subprocess.DEVNULL)
    return _LIBPATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise CfdError("libcfdb200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(there is no CPU fallback)")
    L = C.CDLL(_LIBPATH)
    vp = C.c_void_p
    L.cfd_last_error.restype = C.c_char_p
    L.cfd_version.restype = C.c_char_p
    L.cfd_default_options.argtypes = [C.POINTER(cfd_options)]
    L.cfd_selectCUDAdevice.argtypes = [C.c_int]
    L.cfd_initializeAndAllocateGPU.argtypes = [C.POINTER(cfd_problem), C.POINTER(cfd_options), C.POINTER(vp)]
    L.cfd_destroy.argtypes = [vp]
    L.cfd_download.argtypes = [vp, C.c_char_p, _dp, C.c_longlong]
    L.cfd_upload.argtypes = [vp, C.c_char_p, _dp, C.c_longlong]
    L.cfd_get_fields.argtypes = [vp, _dp, _dp]
    for n in ("bch_beginTimeStep", "bch_calculateMatrixAGPU", "bch_rotateIterates", "bch_calculate_KtimesAcc_prevGPU",
              "bch_endTimeStep", "gq_calculateMatrixAGPU", "gq_step1GPU", "gq_step2GPU", "gq_endTimeStep"):
        getattr(L, n).argtypes = [vp]
    for n in ("bch_step1GPUpart", "bch_step2GPU", "bch_step3GPU"):
        getattr(L, n).argtypes = [vp, C.c_int]
    L.bch_checkConvergenceGPU.argtypes = [vp, _ip]
    L.bch_checkConvergenceInTimeGPU.argtypes = [vp, _ip, _dp]
    L.gq_checkConvergenceInTimeGPU.argtypes = [vp, _ip, _dp]
    L.cfd_printMonitorDataGPU.argtypes = [vp, _dp]
    L.cfd_last_cg_iterations.argtypes = [vp]
    L.cfd_last_bicgstab_iterations.argtypes = [vp, _ip]
    L.cfd_last_krylov_status.argtypes = [vp, _ip, _dp, _ip, _dp, _ip]
    L.cfd_timeLoop.argtypes = [vp, C.c_int, C.c_int, C.POINTER(cfd_step_report), _ip]
    L.cfd_timeLoop_timed.argtypes = [vp, C.c_int, C.POINTER(cfd_step_report), _dp]
    L.cfd_timeStep_host.argtypes = [vp, _dp, _dp, C.POINTER(cfd_step_report), _dp]
    L.cfd_profile.argtypes = [vp, C.c_int]
    L.cfd_profile_get.argtypes = [vp, C.c_char_p, _dp, C.POINTER(C.c_longlong), _dp]
    L.cfd_profile_get_layout_bytes.argtypes = [vp, C.c_char_p, _dp]
    L.cfd_kernel_launches.argtypes = [vp]
    L.cfd_kernel_launches.restype = C.c_longlong
    L.cfd_matrix_stats.argtypes = [vp, C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _ip]
    L.cfd_csr_spmv.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, _dp, _dp, C.c_double, C.c_double, C.c_int, _dp]
    L.cfd_csr_pcg.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_double, C.c_int, _ip, _dp]
    L.cfd_csr_pcg_variant.argtypes = [C.c_int, _ip, _ip, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int, _ip, _dp]
    L.cfd_p2p_reset.argtypes = [vp]
    L.cfd_p2p_detach.argtypes = [vp]
    L.cfd_operator_apply.argtypes = [vp, C.c_char_p, C.c_int, _dp, _dp, C.c_int, _dp, _dp]
    L.cfd_solver_create_from_inp.argtypes = [C.c_char_p, C.c_int, C.POINTER(cfd_options), C.c_int, C.POINTER(vp)]
    L.cfd_solver_create_cavity.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                           C.c_int, C.c_double, C.POINTER(cfd_options), C.c_int, C.POINTER(vp)]
    L.cfd_solver_create_from_inp_part.argtypes = [C.c_char_p, C.c_int, C.POINTER(cfd_options), C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.cfd_solver_create_cavity_part.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double,
                                                C.c_int, C.c_double, C.POINTER(cfd_options), C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.cfd_nccl_unique_id.argtypes = [C.c_char_p]
    L.cfd_comm_init.argtypes = [vp, C.c_char_p]
    L.cfd_p2p_export.argtypes = [vp, C.c_char_p]
    L.cfd_p2p_attach.argtypes = [vp, C.c_char_p]
    L.cfd_solver_init_gpu.argtypes = [vp]
    L.cfd_solver_destroy.argtypes = [vp]
    L.cfd_solver_ctx.argtypes = [vp]
    L.cfd_solver_ctx.restype = vp
    L.cfd_solver_problem.argtypes = [vp]
    L.cfd_solver_problem.restype = C.POINTER(cfd_problem)
    L.cfd_solver_int.argtypes = [vp, C.c_char_p]
    L.cfd_solver_int.restype = C.c_longlong
    L.cfd_solver_array.argtypes = [vp, C.c_char_p, C.POINTER(C.c_longlong), _ip]
    L.cfd_solver_array.restype = vp
    L.cfd_solver_write_tecplot.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.cfd_solver_read_restart.argtypes = [vp, C.c_char_p]
    L.cfd_solver_write_tecplot_fields.argtypes = [vp, C.c_char_p, C.c_char_p, _dp, _dp]
    L.cfd_solver_read_restart_fields.argtypes = [vp, C.c_char_p, _dp, _dp]
    L.cfd_diag_sync.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    L.cfd_sell_layout_check.argtypes = [C.c_int, C.c_int, _ip, _ip, _dp, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), _dp]
    L.cfd_sell_layout_check.restype = C.c_longlong
    L.cfd_solver_main.argtypes = [C.c_char_p, C.c_int, C.c_int]
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise CfdError("cfdb200 error %d: %s" % (rc, lib().cfd_last_error().decode()))


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def options(**kw):
    o = cfd_options()
    lib().cfd_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise KeyError(k)
        setattr(o, k, v)
    return o


class Problem:
    """Builds a `cfd_problem` from arrays named like the reference's globals.

    `src` is any object with `.i(name)`, `.d(name)`, `.arr(name)` (the oracle wrapper in tests, or a
    Solver): this is the role the reference's host code plays at the seam."""

    def __init__(self, src, scheme):
        self.keep = {}
        p = cfd_problem()
        p.scheme = EXPLICIT if scheme in (0, "explicit") else IMPLICIT
        for k in ("NE", "NN", "NNp", "NENv", "NENp", "NGP", "maxIter", "BCnVelNodes", "nBC", "zeroPressureNode", "monPoint"):
            setattr(p, k, src.i(k))
        for k in ("dt", "density", "viscosity", "tolerance", "convergenceCriteria"):
            setattr(p, k, src.d(k))

        def put(field, name, dtype):
            a = np.ascontiguousarray(src.arr(name), dtype=dtype)
            self.keep[field] = a
            setattr(p, field, _d(a) if dtype == np.float64 else _i(a))
            return a

        put("coord", "coord", np.float64)
        put("LtoGnode", "LtoGnode", np.int32)
        put("sparseMrowStarts", "sparseMrowStarts", np.int32)
        p.sparseM_nnz = len(put("sparseMcol", "sparseMcol", np.int32))
        put("sparseKvalue", "sparseKvalue", np.float64)
        put("sparseGrowStarts", "sparseGrowStarts", np.int32)
        p.sparseG_nnz = len(put("sparseGcol", "sparseGcol", np.int32))
        for g in ("sparseG1value", "sparseG2value", "sparseG3value"):
            put(g, g, np.float64)
        put("sparseZrowStarts", "sparseZrowStarts", np.int32)
        p.sparseZ_nnz = len(put("sparseZcol", "sparseZcol", np.int32))
        put("sparseZvalue", "sparseZvalue", np.float64)
        put("BCvelNodes", "BCvelNodes", np.int32)
        put("BCstr", "BCstr", np.float64)
        if p.scheme == EXPLICIT:
            put("MdInv", "MdInv", np.float64)
            put("MdOrigInv", "MdOrigInv", np.float64)
        else:
            put("sparseMvalue", "sparseMvalue", np.float64)
            put("sparseMapM_1d", "sparseMapM_1d", np.int32)
        self.c = p


class Context:
    """`cfd_ctx`: device state + the seam functions."""

    def __init__(self, problem=None, opts=None, _handle=None, _owner=None):
        self.L = lib()
        self._owner = _owner
        if _handle is not None:
            self.h = _handle
            self.owned = False
            self._prob = _owner.problem_struct()
        else:
            h = C.c_void_p()
            self._keep = problem
            _check(self.L.cfd_initializeAndAllocateGPU(C.byref(problem.c), C.byref(opts) if opts is not None else None, C.byref(h)))
            self.h = h
            self.owned = True
            self._prob = problem.c
        self.NN, self.NNp, self.scheme = self._prob.NN, self._prob.NNp, self._prob.scheme

    def close(self):
        if getattr(self, "owned", False) and self.h:
            self.L.cfd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data movement
    def _len(self, name):
        return self.NNp if name[0] == "P" or name == "R2" else 3 * self.NN

    def download(self, name):
        out = np.empty(self._len(name))
        _check(self.L.cfd_download(self.h, name.encode(), _d(out), out.size))
        return out

    def upload(self, name, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        _check(self.L.cfd_upload(self.h, name.encode(), _d(a), a.size))

    def fields(self):
        U = np.empty(3 * self.NN)
        P = np.empty(self.NNp)
        _check(self.L.cfd_get_fields(self.h, _d(U), _d(P)))
        return U, P

    # ---- seam
    def call(self, name, *args):
        _check(getattr(self.L, name)(self.h, *args))

    def bch_checkConvergenceGPU(self):
        c = C.c_int()
        _check(self.L.bch_checkConvergenceGPU(self.h, C.byref(c)))
        return bool(c.value)

    def checkConvergenceInTimeGPU(self):
        s = C.c_int()
        m = C.c_double()
        fn = self.L.bch_checkConvergenceInTimeGPU if self.scheme == EXPLICIT else self.L.gq_checkConvergenceInTimeGPU
        _check(fn(self.h, C.byref(s), C.byref(m)))
        return bool(s.value), m.value

    def monitor(self):
        out = (C.c_double * 4)()
        _check(self.L.cfd_printMonitorDataGPU(self.h, out))
        return list(out)

    def last_cg(self):
        return int(self.L.cfd_last_cg_iterations(self.h))

    def last_bicg(self):
        its = (C.c_int * 3)()
        _check(self.L.cfd_last_bicgstab_iterations(self.h, its))
        return list(its)

    def krylov_status(self):
        """Status of the most recent solves: dict(cg_converged, cg_residual, bicg_converged[3], bicg_true_residual[3], bicg_restarts[3])."""
        cc, cr = C.c_int(), C.c_double()
        bc, bt, br = (C.c_int * 3)(), (C.c_double * 3)(), (C.c_int * 3)()
        _check(self.L.cfd_last_krylov_status(self.h, C.byref(cc), C.byref(cr), bc, bt, br))
        return dict(cg_converged=bool(cc.value), cg_residual=cr.value, bicg_converged=[bool(v) for v in bc],
                    bicg_true_residual=list(bt), bicg_restarts=list(br))

    def time_loop(self, nsteps, stop_on_steady=False):
        reps = (cfd_step_report * nsteps)()
        done = C.c_int()
        _check(self.L.cfd_timeLoop(self.h, nsteps, int(stop_on_steady), reps, C.byref(done)))
        return [self._rep(r) for r in reps[:done.value]]

    @staticmethod
    def _rep(r):
        return dict(timeN=r.timeN, iter=r.iter, cg_iterations=r.cg_iterations, cg_solves=r.cg_solves,
                    bicg_iterations=list(r.bicg_iterations), maxAcc=r.maxAcc, steady=bool(r.steady),
                    monitor=list(r.monitor), cg_converged=bool(r.cg_converged), cg_residual=r.cg_residual,
                    bicg_converged=[bool(v) for v in r.bicg_converged], bicg_restarts=list(r.bicg_restarts),
                    bicg_true_residual=list(r.bicg_true_residual))

    def time_loop_timed(self, nsteps):
        """nsteps time steps; returns (reports, device milliseconds by CUDA events on the launch stream)."""
        reps = (cfd_step_report * nsteps)()
        ms = C.c_double()
        _check(self.L.cfd_timeLoop_timed(self.h, nsteps, reps, C.byref(ms)))
        return [self._rep(r) for r in reps], ms.value

    def time_step_host(self, U, P):
        """One step from/to host buffers U [3NN], P [NNp] (in place); returns (report, device ms)."""
        assert U.dtype == np.float64 and P.dtype == np.float64 and U.size == 3 * self.NN and P.size == self.NNp
        rep = cfd_step_report()
        ms = C.c_double()
        _check(self.L.cfd_timeStep_host(self.h, _d(U), _d(P), C.byref(rep), C.byref(ms)))
        return self._rep(rep), ms.value

    def operator_apply(self, op, x, path=0, repeat=1):
        """y = Op x through the assembled (path 0) or the matrix-free element (path 1) kernels; returns (y, best ms, algorithmic bytes)."""
        x = np.ascontiguousarray(x, np.float64)
        y = np.zeros(self.NNp if op in ("GT", "Z") else 3 * self.NN)
        ms, nb = C.c_double(), C.c_double()
        _check(self.L.cfd_operator_apply(self.h, op.encode(), path, _d(x), _d(y), repeat, C.byref(ms), C.byref(nb)))
        return y, ms.value, nb.value

    def profile(self, enable=True):
        _check(self.L.cfd_profile(self.h, int(enable)))

    def profile_get(self, name):
        ms = C.c_double()
        n = C.c_longlong()
        b = C.c_double()
        lb = C.c_double()
        self.L.cfd_profile_get(self.h, name.encode(), C.byref(ms), C.byref(n), C.byref(b))
        self.L.cfd_profile_get_layout_bytes(self.h, name.encode(), C.byref(lb))
        return dict(ms=ms.value, launches=n.value, bytes=b.value, layout_bytes=lb.value)

    def matrix_stats(self, name):
        a, b, c_, t = C.c_longlong(), C.c_longlong(), C.c_longlong(), C.c_int()
        _check(self.L.cfd_matrix_stats(self.h, name.encode(), C.byref(a), C.byref(b), C.byref(c_), C.byref(t)))
        return dict(nnz=a.value, stored_values=b.value, stored_indices=c_.value, lanes_per_row=t.value)

    def kernel_launches(self):
        return int(self.L.cfd_kernel_launches(self.h))


class Solver:
    """`cfd_solver`: the reference's solverGPU program as an object."""

    def __init__(self, inp_path=None, scheme="explicit", cavity=None, opts=None, init_gpu=True, world=1, rank=0):
        self.L = lib()
        self.scheme = EXPLICIT if scheme in (0, "explicit") else IMPLICIT
        self.world, self.rank = world, rank
        h = C.c_void_p()
        o = C.byref(opts) if opts is not None else None
        if inp_path is not None and world > 1:
            _check(self.L.cfd_solver_create_from_inp_part(os.fsencode(inp_path), self.scheme, o, int(init_gpu), world, rank, C.byref(h)))
        elif inp_path is not None:
            _check(self.L.cfd_solver_create_from_inp(os.fsencode(inp_path), self.scheme, o, int(init_gpu), C.byref(h)))
        elif world > 1:
            cv = dict(nx=8, ny=None, nz=None, cluster=0.0, NENv=27, dt=1e-3, viscosity=0.01, maxIter=4, tolerance=1e-3)
            cv.update(cavity or {})
            _check(self.L.cfd_solver_create_cavity_part(cv["nx"], cv["ny"] or cv["nx"], cv["nz"] or cv["nx"], cv["cluster"], cv["NENv"],
                                                        self.scheme, cv["dt"], cv["viscosity"], cv["maxIter"], cv["tolerance"], o,
                                                        int(init_gpu), world, rank, C.byref(h)))
        else:
            cv = dict(nx=8, ny=None, nz=None, cluster=0.0, NENv=27, dt=1e-3, viscosity=0.01, maxIter=4, tolerance=1e-3)
            cv.update(cavity or {})
            ny = cv["ny"] or cv["nx"]
            nz = cv["nz"] or cv["nx"]
            _check(self.L.cfd_solver_create_cavity(cv["nx"], ny, nz, cv["cluster"], cv["NENv"], self.scheme, cv["dt"],
                                                   cv["viscosity"], cv["maxIter"], cv["tolerance"], o, int(init_gpu), C.byref(h)))
        self.h = h
        self._ctx = None

    def close(self):
        if self.h:
            self.L.cfd_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_gpu(self):
        _check(self.L.cfd_solver_init_gpu(self.h))

    def comm_init(self, unique_id):
        """Attach the NCCL communicator (partitioned runs): `unique_id` = the 128 bytes from nccl_unique_id() on rank 0."""
        ch = self.L.cfd_solver_ctx(self.h)
        if not ch:
            raise CfdError("solver has no GPU context")
        _check(self.L.cfd_comm_init(C.c_void_p(ch), bytes(unique_id)))

    def p2p_export(self):
        buf = C.create_string_buffer(96)
        _check(self.L.cfd_p2p_export(C.c_void_p(self.L.cfd_solver_ctx(self.h)), buf))
        return buf.raw

    def p2p_attach(self, all_desc):
        _check(self.L.cfd_p2p_attach(C.c_void_p(self.L.cfd_solver_ctx(self.h)), bytes(all_desc)))

    def owned_fields(self):
        """(global node ids, U[3, n_owned], global pressure ids, P[n_owned]) of the entries this rank owns."""
        U, P = self.ctx.fields()
        NN, NNo, NPo = self.i("NN"), self.i("NN_owned"), self.i("NNp_owned")
        if self.world > 1:
            ng, pg = self.arr("part_nodeGlobal")[:NNo], self.arr("part_presGlobal")[:NPo]
        else:
            ng, pg = np.arange(NNo), np.arange(NPo)
        return ng, U.reshape(3, NN)[:, :NNo].copy(), pg, P[:NPo].copy()

    def problem_struct(self):
        return self.L.cfd_solver_problem(self.h).contents

    @property
    def ctx(self):
        if self._ctx is None:
            ch = self.L.cfd_solver_ctx(self.h)
            if not ch:
                raise CfdError("solver has no GPU context (created with init_gpu=False)")
            self._ctx = Context(_handle=C.c_void_p(ch), _owner=self)
        return self._ctx

    def i(self, name):
        p = self.problem_struct()
        if hasattr(p, name) and name not in ("coord", "NN_owned", "NNp_owned", "NE_apply", "pinLocal", "nPeers"):
            return int(getattr(p, name))
        v = self.L.cfd_solver_int(self.h, name.encode())
        if v == -999999:
            raise KeyError(name)
        return int(v)

    def d(self, name):
        return float(getattr(self.problem_struct(), name))

    def arr(self, name):
        n = C.c_longlong()
        isint = C.c_int()
        p = self.L.cfd_solver_array(self.h, name.encode(), C.byref(n), C.byref(isint))
        if n.value < 0:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0, dtype=np.int32 if isint.value else np.float64)
        ct = C.c_int if isint.value else C.c_double
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(n.value,)).copy()

    def write_tecplot(self, path, title):
        _check(self.L.cfd_solver_write_tecplot(self.h, os.fsencode(path), title.encode()))

    def write_tecplot_fields(self, path, title, U, P):
        U = np.ascontiguousarray(U, np.float64)
        P = np.ascontiguousarray(P, np.float64)
        _check(self.L.cfd_solver_write_tecplot_fields(self.h, os.fsencode(path), title.encode(), _d(U), _d(P)))

    def read_restart_fields(self, path):
        U = np.zeros(3 * self.i("NN"))
        P = np.zeros(self.i("NNp"))
        _check(self.L.cfd_solver_read_restart_fields(self.h, os.fsencode(path), _d(U), _d(P)))
        return U, P

    def read_restart(self, path):
        _check(self.L.cfd_solver_read_restart(self.h, os.fsencode(path)))


def csr_spmv(rowStarts, col, val, x, y=None, alpha=1.0, beta=0.0, ncols=None, repeat=1):
    """y = alpha*A*x + beta*y on the GPU for x of shape [nvec, ncols]; returns (y, best ms)."""
    rs = np.ascontiguousarray(rowStarts, np.int32)
    cl = np.ascontiguousarray(col, np.int32)
    vl = np.ascontiguousarray(val, np.float64)
    x = np.ascontiguousarray(np.atleast_2d(x), np.float64)
    nrows = len(rs) - 1
    ncols = ncols or x.shape[1]
    yy = np.zeros((x.shape[0], nrows)) if y is None else np.ascontiguousarray(np.atleast_2d(y), np.float64).copy()
    ms = C.c_double()
    _check(lib().cfd_csr_spmv(nrows, ncols, _i(rs), _i(cl), _d(vl), x.shape[0], _d(x), _d(yy), alpha, beta, repeat, C.byref(ms)))
    return yy, ms.value


def sell_layout_check(rowStarts, col, val, ncols, T=0):
    """Host-side round trip CSR -> device layout -> CSR.  Returns (mismatches, slices, compressed slices, stored/nnz)."""
    rs = np.ascontiguousarray(rowStarts, np.int32)
    cl = np.ascontiguousarray(col, np.int32)
    vl = np.ascontiguousarray(val, np.float64)
    ns, nc, ratio = C.c_longlong(), C.c_longlong(), C.c_double()
    bad = lib().cfd_sell_layout_check(len(rs) - 1, ncols, _i(rs), _i(cl), _d(vl), T, C.byref(ns), C.byref(nc), C.byref(ratio))
    return int(bad), int(ns.value), int(nc.value), float(ratio.value)


def csr_pcg(rowStarts, col, val, b, x0=None, tol=1e-6, maxit=1000, variant=1):
    """Jacobi-PCG on the GPU; variant 1 = classical recurrences, 2 = Chronopoulos-Gear single-reduction form."""
    rs = np.ascontiguousarray(rowStarts, np.int32)
    cl = np.ascontiguousarray(col, np.int32)
    vl = np.ascontiguousarray(val, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    x = np.zeros_like(b) if x0 is None else np.ascontiguousarray(x0, np.float64).copy()
    its = C.c_int()
    ms = C.c_double()
    _check(lib().cfd_csr_pcg_variant(len(b), _i(rs), _i(cl), _d(vl), _d(b), _d(x), tol, maxit, variant, C.byref(its), C.byref(ms)))
    return x, its.value, ms.value


def setup_distributed(solver, dist, torch, peer_memory=True):
    """Attach communicators to a partitioned Solver using an initialised torch.distributed (NCCL) group:
    broadcast of the NCCL unique id, then (optionally) all-gather of the CUDA-IPC window descriptors."""
    rank = dist.get_rank()
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    solver.comm_init(idt.cpu().numpy().tobytes())
    if peer_memory:
        mine = torch.frombuffer(bytearray(solver.p2p_export()), dtype=torch.uint8).cuda()
        allb = [torch.zeros(96, dtype=torch.uint8, device="cuda") for _ in range(dist.get_world_size())]
        dist.all_gather(allb, mine)
        solver.p2p_attach(b"".join(t.cpu().numpy().tobytes() for t in allb))
        dist.barrier()


def teardown_distributed(solver, dist):
    """Orderly end of a partitioned run: every rank unmaps its peers' exchange windows, the ranks meet, then the solver (and with
    it this rank's own window) is destroyed -- exported memory must not be freed while a peer still has it mapped."""
    ch = solver.L.cfd_solver_ctx(solver.h)
    if ch:
        _check(solver.L.cfd_p2p_detach(C.c_void_p(ch)))
    dist.barrier()
    solver.close()
    dist.barrier()


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _check(lib().cfd_nccl_unique_id(buf))
    return buf.raw


def solver_main(workdir, scheme="explicit", max_steps=0):
    _check(lib().cfd_solver_main(os.fsencode(workdir), EXPLICIT if scheme in (0, "explicit") else IMPLICIT, max_steps))
