"""ctypes wrapper around the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the product.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "fracstep_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_error.restype = ctypes.c_char_p
        L.orc_error.argtypes = [ctypes.c_void_p]
        L.orc_int.restype = ctypes.c_longlong
        L.orc_int.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.orc_double.restype = ctypes.c_double
        L.orc_double.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.orc_set.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]
        L.orc_array.restype = ctypes.c_void_p
        L.orc_array.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_longlong),
                                ctypes.POINTER(ctypes.c_int)]
        L.orc_run.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.orc_phase.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]
        L.orc_write_tecplot.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
        L.orc_read_restart.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
        L.orc_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


class Oracle:
    """One solver instance: reads <inp>, preprocesses, runs step0.  scheme: 'explicit' | 'implicit'."""

    def __init__(self, inp_path, scheme="explicit", nthreads=0, **options):
        self.L = lib()
        self.scheme = scheme
        self.h = self.L.orc_create(os.fsencode(inp_path), 0 if scheme == "explicit" else 1, nthreads)
        err = self.L.orc_error(self.h).decode()
        if err:
            raise RuntimeError("oracle: " + err)
        for k, v in options.items():
            self.set(k, v)

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def i(self, name):
        v = self.L.orc_int(self.h, name.encode())
        if v == -999999:
            raise KeyError(name)
        return int(v)

    def d(self, name):
        return float(self.L.orc_double(self.h, name.encode()))

    def set(self, name, value):
        if self.L.orc_set(self.h, name.encode(), float(value)) != 0:
            raise KeyError(name)

    def arr(self, name, copy=True):
        n = ctypes.c_longlong()
        isint = ctypes.c_int()
        p = self.L.orc_array(self.h, name.encode(), ctypes.byref(n), ctypes.byref(isint))
        if n.value < 0:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0, dtype=np.int32 if isint.value else np.float64)
        ct = ctypes.c_int if isint.value else ctypes.c_double
        a = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ct)), shape=(n.value,))
        return a.copy() if copy else a

    def run(self, nsteps):
        self.L.orc_run(self.h, int(nsteps))

    def phase(self, name, it=1):
        if self.L.orc_phase(self.h, name.encode(), int(it)) != 0:
            raise KeyError(name)

    def write_tecplot(self, path, title):
        if self.L.orc_write_tecplot(self.h, os.fsencode(path), title.encode()) != 0:
            raise RuntimeError(self.L.orc_error(self.h).decode())

    def read_restart(self, path):
        if self.L.orc_read_restart(self.h, os.fsencode(path)) != 0:
            raise RuntimeError(self.L.orc_error(self.h).decode())

    @property
    def U(self):
        return self.arr("Un" if self.scheme == "explicit" else "Uk_prev")

    @property
    def P(self):
        return self.arr("Pn" if self.scheme == "explicit" else "Pk_prev")


def threads():
    return int(lib().orc_threads())
