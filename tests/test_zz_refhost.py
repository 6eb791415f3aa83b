"""The reference's own HOST programs driving libcfdb200.so through the seam.

`make -C oracle refhost` compiles blascoCodinaHuerta.cpp / guermondQuartapelle.cpp unmodified with -DUSECUDA and links them
with integration/cfdb200_shim_bch.cpp / cfdb200_shim_gq.cpp in place of their CUDAcodes.cu (the binding INTEGRATION.md
describes, as real code).  Everything outside the eleven / six seam functions is then the reference: main(), reader,
preprocessing, step0(), timeLoop() with its own cudaMemcpy rotations on the `*_d` globals, createTecplot().

CPU part: the programs exist (where the reference tree was mounted at build time) and stop with an error code without a GPU.
GPU part: they must write the .dat the reference's CPU build wrote for the same input deck (tests/golden/ref_*.npz).
The file sorts last on purpose (the programs are the slowest to start)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from test_reference_pin import ROOT, errors, load

MOCK = {"explicit": os.path.join(ROOT, "oracle", "_ref", "mock", "bch_mock"), "implicit": os.path.join(ROOT, "oracle", "_ref", "mock", "gq_mock")}
EXE = {"explicit": os.path.join(ROOT, "oracle", "_ref", "bch_gpu_cfdb200"), "implicit": os.path.join(ROOT, "oracle", "_ref", "gq_gpu_cfdb200")}


def deck(tmp_path, meta):
    import meshgen
    kw = dict(meta["kw"])
    getattr(meshgen, meta["kind"])(str(tmp_path / "case.inp"), t_final=(meta["steps"] - 0.5) * kw.get("dt", 0.001), **kw)
    (tmp_path / "ProblemName.txt").write_text("case\n")


@pytest.mark.parametrize("scheme", ["explicit", "implicit"])
def test_reference_host_program_links_against_the_library(scheme, tmp_path):
    if not os.path.exists(EXE[scheme]):
        pytest.skip("oracle/_ref/*_gpu_cfdb200 not built (needs the reference tree: make -C oracle refhost)")
    out = subprocess.run(["ldd", EXE[scheme]], capture_output=True, text=True).stdout
    assert "libcfdb200.so" in out and "not found" not in out
    assert "cusparse" not in out and "cublas" not in out        # CUDAcodes.cu's libraries are gone from the link line
    import torch
    if torch.cuda.is_available():
        return
    _, meta = load("ref_bch_cav4" if scheme == "explicit" else "ref_gq_cav4")
    deck(tmp_path, meta)
    r = subprocess.run([EXE[scheme]], cwd=str(tmp_path), capture_output=True, text=True, stdin=subprocess.DEVNULL, timeout=120)
    assert r.returncode == 1 and "cfdb200: error -2" in r.stderr   # selectCUDAdevice(): no device, no CPU fallback, no `cin >>`


@pytest.mark.parametrize("name,env", [("ref_bch_cav4", {"CFDB200_CPU_SEMANTICS": "1"}), ("ref_bch_cav6c", {"CFDB200_CPU_SEMANTICS": "1"}),
                                      ("ref_bch_chan644", {"CFDB200_CPU_SEMANTICS": "1"}), ("ref_bch_box345_nopin", {"CFDB200_CPU_SEMANTICS": "1"}),
                                      ("ref_gq_cav4_tight", {"CFD_CG_TOL": "1e-12", "CFD_BICG_TOL": "1e-12", "OMP_NUM_THREADS": "1"}),
                                      ("ref_gq_chan644_tight", {"CFD_CG_TOL": "1e-12", "CFD_BICG_TOL": "1e-12", "OMP_NUM_THREADS": "1"})])
def test_reference_host_program_on_the_cpu_mock(name, env, tmp_path):
    """No GPU needed: the reference's -DUSECUDA host program + the replacement shim, linked against a CPU mock of the library's
    C ABI that serves every seam call from the oracle (oracle/mock/, `make -C oracle refhost-mock`).  Checks
      * array by array, that what the shim marshals into cfd_problem -- the REFERENCE's own preprocessing and step0 results
        (coordinates, connectivity, CSR patterns, K, M, G1-3, MdInv, the element map, BC lists, initial state) and the shim's
        Z -- equals what the oracle derives from the same input file (the mock prints one line per array);
      * that the reference's timeLoop() driving the seam in its own order, with its own rotations on the pointers
        cfd_device_pointer hands out and the per-step read-back, ends in the .dat file its CPU build writes."""
    g, meta = load(name)
    exe = MOCK[meta["scheme"]]
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/mock not built (needs the reference tree: make -C oracle refhost-mock)")
    deck(tmp_path, meta)
    r = subprocess.run([exe], cwd=str(tmp_path), env=dict(os.environ, **env), capture_output=True, text=True, stdin=subprocess.DEVNULL,
                       timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stderr.split("\n") if ln.startswith("mock: cfd_problem.")]
    assert len(lines) >= 19 and all(ln.endswith(" ok") for ln in lines), "\n".join(lines)
    # the reference's own step0 arrays are not just close to the oracle's but identical (same formulas, same summation order)
    for key in ("sparseKvalue", "sparseG1value", "sparseG2value", "sparseG3value"):
        assert any(key in ln and "= 0.00e+00 ok" in ln for ln in lines), key
    NN = int(g["NN"])
    t = np.loadtxt(str(tmp_path / "case.dat"), skiprows=3, max_rows=NN)
    assert np.array_equal(t[:, :3], g["coord"])
    eu, ep = errors(np.concatenate([t[:, 3], t[:, 4], t[:, 5]]), t[:int(g["NNp"]), 6], g, meta, NN)
    bar = 1e-7 if (meta["env"] or meta["kw"].get("pin", 2) <= 1) else 1e-10
    assert eu <= bar and ep <= bar, (eu, ep)
    ncol = 9 if meta["scheme"] == "explicit" else 8
    rows = [x for x in (ln.split() for ln in r.stdout.split("\n")) if len(x) == ncol and x[0].isdigit()]
    assert len(rows) == meta["steps"]
    if meta["scheme"] == "explicit":
        assert [int(x[1]) for x in rows] == [int(v) for v in g["monitor"][:, 1]]
        assert np.allclose([float(x[8]) for x in rows], g["monitor"][:, 8], rtol=1e-4, atol=2e-5)     # maxAcc column


def test_reference_step0_arrays_on_the_shipped_mesh(tmp_path):
    """BASELINE.json configs[0] at the array level: on lidDrivenCavity_NE27000.inp (regenerated byte for byte) the reference's
    own preprocessing and step0 -- 226 981 generated node coordinates, connectivity, the 14.0 M-entry K, G1-3, the lumped
    masses -- are bitwise what the oracle computes; Z (3.3 M entries, the shim's product) to rounding."""
    if not os.path.exists(MOCK["explicit"]):
        pytest.skip("oracle/_ref/mock not built (needs the reference tree: make -C oracle refhost-mock)")
    import meshgen
    meshgen.cube(str(tmp_path / "case.inp"), 30, cluster=2.0, t_final=0.0005)       # one time step
    (tmp_path / "ProblemName.txt").write_text("case\n")
    r = subprocess.run([MOCK["explicit"]], cwd=str(tmp_path), env=dict(os.environ, CFDB200_CPU_SEMANTICS="1"), capture_output=True, text=True,
                       stdin=subprocess.DEVNULL, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = {ln.split()[1].split(".")[1]: ln for ln in r.stderr.split("\n") if ln.startswith("mock: cfd_problem.")}
    assert len(lines) == 19 and all(ln.endswith(" ok") for ln in lines.values()), "\n".join(lines.values())
    assert " 13997521 values" in lines["sparseKvalue"] and " 3442951 values" in lines["sparseG1value"] and " 3307949 values" in lines["sparseZvalue"]
    for key in ("coord", "sparseKvalue", "sparseG1value", "sparseG2value", "sparseG3value", "MdInv", "MdOrigInv"):
        assert "= 0.00e+00 ok" in lines[key], lines[key]


@pytest.mark.gpu
@pytest.mark.parametrize("name,env", [("ref_bch_cav4", {"CFDB200_CPU_SEMANTICS": "1"}), ("ref_bch_cav6c", {"CFDB200_CPU_SEMANTICS": "1"}),
                                      ("ref_gq_cav4_tight", {"CFD_CG_TOL": "1e-12", "CFD_BICG_TOL": "1e-12", "CFD_EBE_SCATTER": "1",
                                                             "CFD_DETERMINISTIC": "1"})])
def test_reference_host_program_on_the_gpu_library(name, env, tmp_path):
    g, meta = load(name)
    exe = EXE[meta["scheme"]]
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/*_gpu_cfdb200 not built")
    deck(tmp_path, meta)
    r = subprocess.run([exe], cwd=str(tmp_path), env=dict(os.environ, **env), capture_output=True, text=True, stdin=subprocess.DEVNULL,
                       timeout=60)
    assert r.returncode == 0, (r.stderr[-1500:], r.stdout[-1500:])
    NN = int(g["NN"])
    t = np.loadtxt(str(tmp_path / "case.dat"), skiprows=3, max_rows=NN)
    assert np.array_equal(t[:, :3], g["coord"])
    eu, ep = errors(np.concatenate([t[:, 3], t[:, 4], t[:, 5]]), t[:int(g["NNp"]), 6], g, meta, NN)
    # explicit: the CPU build's settings selected; implicit: both sides iterated to 1e-12 -> the parity bar in both cases
    assert eu <= 1e-6 and ep <= 1e-6, (eu, ep)
    ncol = 9 if meta["scheme"] == "explicit" else 8
    rows = [ln.split() for ln in r.stdout.split("\n")]
    rows = [x for x in rows if len(x) == ncol and x[0].isdigit()]
    assert len(rows) == meta["steps"]
    if meta["scheme"] == "explicit":
        assert [int(x[1]) for x in rows] == [int(v) for v in g["monitor"][:, 1]]   # inner iterations per step
