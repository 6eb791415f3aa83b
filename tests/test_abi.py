"""CPU: the C-ABI library loads and exports every symbol include/cfdb200.h declares; argument checks
that need no GPU behave; the product never imports the oracle."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "cfdb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([a-z][A-Za-z0-9_]*)\s*\(", src)
    return sorted({n for n in names if n.startswith(("cfd_", "bch_", "gq_"))})


def test_every_declared_symbol_is_exported():
    import fracstep
    L = ctypes.CDLL(fracstep.build())
    fns = declared_functions()
    assert len(fns) >= 40
    missing = [f for f in fns if not hasattr(L, f)]
    assert not missing, missing


def test_struct_layouts_match_header():
    """ctypes mirrors of cfd_problem / cfd_options / cfd_step_report have the C sizes (checked through
    cfd_default_options writing the whole struct and a round trip of a report)."""
    import fracstep
    o = fracstep.options()
    assert o.cg_tol == 1e-6 and o.cg_maxit == 1000 and o.bicg_tol == 1e-6 and o.bicg_maxit == 1000
    assert o.cg_warm_start == 0 and o.maxacc_rule == 0 and o.device == -1 and not o.cuda_stream
    sz = (ctypes.c_int * 3)()
    fracstep.lib().cfd_struct_sizes(sz)      # the sizes the C side was compiled with
    assert [ctypes.sizeof(fracstep.cfd_problem), ctypes.sizeof(fracstep.cfd_options), ctypes.sizeof(fracstep.cfd_step_report)] == list(sz)
    assert o.deterministic == 0 and o.bicg_no_halfstep == 0 and o.bicg_no_verify == 0 and o.cg_variant == 0


def test_argument_errors_without_gpu():
    import fracstep
    L = fracstep.lib()
    h = ctypes.c_void_p()
    assert L.cfd_initializeAndAllocateGPU(None, None, ctypes.byref(h)) == -1
    p = fracstep.cfd_problem()
    p.NENv, p.NENp, p.NGP = 10, 4, 4     # tetrahedra: unsupported, like the reference's "not implemented" branches
    assert L.cfd_initializeAndAllocateGPU(ctypes.byref(p), None, ctypes.byref(h)) == -1
    assert b"unsupported element" in L.cfd_last_error()
    assert L.cfd_version().startswith(b"cfdb200")


def test_no_cpu_fallback_without_device(mesh):
    """On a box without a GPU the compute entry points fail loudly (no silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import fracstep
    s = fracstep.Solver(mesh("abi_q1", n=3, NENv=8), "explicit", init_gpu=False)
    with pytest.raises(fracstep.CfdError):
        s.init_gpu()
    with pytest.raises(fracstep.CfdError):
        s.ctx


def test_solverGPU_executables_fail_loudly_without_device(mesh, tmp_path):
    """bin/solverGPU_explicit / _implicit (the reference's executables, no arguments) are built with the library; without a
    GPU they must stop with an error code and a message instead of computing anything on the CPU."""
    import shutil
    import subprocess
    import torch
    for scheme in ("explicit", "implicit"):
        exe = os.path.join(ROOT, "cfd-with-cuda_b200", "bin", "solverGPU_" + scheme)
        assert os.access(exe, os.X_OK), exe
        # no ProblemName.txt: an I/O error, not a crash
        out = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, stdin=subprocess.DEVNULL, timeout=120)
        assert out.returncode == 1 and "ProblemName.txt" in out.stderr
    if torch.cuda.is_available():
        return
    shutil.copy(mesh("abi_q2", n=2), str(tmp_path / "c.inp"))
    (tmp_path / "ProblemName.txt").write_text("c\n")
    out = subprocess.run([exe], cwd=str(tmp_path), capture_output=True, text=True, stdin=subprocess.DEVNULL, timeout=120)
    assert out.returncode == 1 and "solverGPU: error -2" in out.stderr
    assert not os.path.exists(str(tmp_path / "c.dat"))


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "cfd-with-cuda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in txt and "import orc" not in txt and "fracstep_oracle" not in txt, f
