"""CPU: the product's own host code (mesh reader, hashed node generation, sparsity, step0 assembly,
Tecplot / restart I/O) against the oracle's line-faithful restatement -- no GPU involved."""
import os

import numpy as np
import pytest

MESHES = {
    "h_q2_c6": dict(n=6, cluster=2.0),
    "h_q2_u5": dict(n=5, full_precision=True),
    "h_q1_u7": dict(n=7, NENv=8, full_precision=True),
    "h_q2_343": dict(n=3, ny=4, nz=3),
}
INT_ARRAYS = ["LtoGnode", "BCvelNodes", "sparseMrowStarts", "sparseMcol", "sparseGrowStarts", "sparseGcol", "sparseZrowStarts",
              "sparseZcol", "sparseMapM_1d", "meshColors", "NmeshColors", "elementsOfColor"]


@pytest.mark.parametrize("scheme", ["explicit", "implicit"])
@pytest.mark.parametrize("name", list(MESHES))
def test_preprocessing_and_step0_match_oracle(mesh, name, scheme):
    import fracstep
    import orc
    path = mesh(name, **MESHES[name])
    o = orc.Oracle(path, scheme)
    s = fracstep.Solver(path, scheme, init_gpu=False)
    for k in ("NE", "NN", "NNp", "BCnVelNodes", "monPoint", "zeroPressureNode", "maxIter"):
        assert s.i(k) == o.i(k), k
    for k in ("nnzM", "nnzG", "nnzZ", "nActiveColors"):
        assert s.L.cfd_solver_int(s.h, k.encode()) == o.i(k), k
    assert np.array_equal(s.arr("coord"), o.arr("coord"))   # same midpoint arithmetic, bit for bit
    for k in INT_ARRAYS:
        assert np.array_equal(s.arr(k), o.arr(k)), k
    vals = ["sparseKvalue", "sparseMvalue", "sparseG1value", "sparseG2value", "sparseG3value", "sparseZvalue"]
    if scheme == "explicit":
        vals += ["MdInv", "MdOrigInv"]
    for k in vals:
        a, b = s.arr(k), o.arr(k)
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-12 * np.abs(b).max(), k
    for k in ("dt", "viscosity", "density", "tolerance", "convergenceCriteria"):
        assert s.d(k) == o.d(k)


def test_in_memory_cavity_equals_inp(mesh):
    """cfd_solver_create_cavity numbers nodes / elements / BC faces exactly like the .inp written by the cube generator."""
    import fracstep
    for kw in (dict(n=4, cluster=2.0, full_precision=True), dict(n=5, NENv=8, full_precision=True)):
        path = mesh("mem_%d_%d" % (kw["n"], kw.get("NENv", 27)), **kw)
        a = fracstep.Solver(path, "explicit", init_gpu=False)
        b = fracstep.Solver(scheme="explicit", init_gpu=False,
                            cavity=dict(nx=kw["n"], cluster=kw.get("cluster", 0.0), NENv=kw.get("NENv", 27)))
        for k in ("LtoGnode", "BCvelNodes", "sparseMcol", "sparseGcol", "sparseZcol"):
            assert np.array_equal(a.arr(k), b.arr(k)), k
        assert np.abs(a.arr("coord") - b.arr("coord")).max() < 1e-15
        assert a.i("zeroPressureNode") == b.i("zeroPressureNode") and a.i("monPoint") == b.i("monPoint")


def test_reader_rejects_bad_input(tmp_path, mesh):
    import fracstep
    with pytest.raises(fracstep.CfdError):
        fracstep.Solver(str(tmp_path / "missing.inp"), "explicit", init_gpu=False)
    good = open(mesh("h_q2_c6", **MESHES["h_q2_c6"])).read()
    bad = tmp_path / "tet.inp"
    bad.write_text(good.replace("eType    : 1", "eType    : 2"))
    with pytest.raises(fracstep.CfdError):
        fracstep.Solver(str(bad), "explicit", init_gpu=False)
    trunc = tmp_path / "trunc.inp"
    trunc.write_text(good[: len(good) // 3])
    with pytest.raises(fracstep.CfdError):
        fracstep.Solver(str(trunc), "explicit", init_gpu=False)


def test_restart_reader_rejects_bad_files(tmp_path, mesh):
    """readRestartFile's input checked row by row (the rows are parsed in parallel: the FIRST bad row is reported)."""
    import fracstep
    s = fracstep.Solver(mesh("h_q2_c6", **MESHES["h_q2_c6"]), "explicit", init_gpu=False)
    NN, NNp = s.i("NN"), s.i("NNp")
    rng = np.random.default_rng(1)
    U, P = rng.standard_normal(3 * NN), rng.standard_normal(NNp)
    good = str(tmp_path / "good.dat")
    s.write_tecplot_fields(good, "t", U, P)
    lines = open(good).read().split("\n")
    U2, P2 = s.read_restart_fields(good)
    assert np.allclose(U2, U, rtol=2e-11, atol=0) and np.allclose(P2, P, rtol=2e-11, atol=0)
    with pytest.raises(fracstep.CfdError, match="cannot open"):
        s.read_restart_fields(str(tmp_path / "missing.dat"))
    (tmp_path / "short.dat").write_text("\n".join(lines[:3 + NN // 2]) + "\n")
    with pytest.raises(fracstep.CfdError, match="malformed row %d" % (NN // 2)):
        s.read_restart_fields(str(tmp_path / "short.dat"))
    broken = list(lines)
    broken[3 + 1234] = "1.0 2.0 three 4.0 5.0 6.0 7.0"
    broken[3 + 2000] = "1.0 2.0"
    (tmp_path / "broken.dat").write_text("\n".join(broken))
    with pytest.raises(fracstep.CfdError, match="malformed row 1234"):
        s.read_restart_fields(str(tmp_path / "broken.dat"))
    (tmp_path / "hdr.dat").write_text("TITLE = t\n")
    with pytest.raises(fracstep.CfdError, match="truncated header"):
        s.read_restart_fields(str(tmp_path / "hdr.dat"))
    # no final newline, Windows line ends: still readable (operator>> in the reference skips any white space)
    (tmp_path / "crlf.dat").write_text("\r\n".join(lines[:3 + NN]))
    U3, P3 = s.read_restart_fields(str(tmp_path / "crlf.dat"))
    assert np.array_equal(U3, U2) and np.array_equal(P3, P2)


@pytest.mark.parametrize("name", ["h_q2_c6", "h_q1_u7"])
def test_tecplot_and_restart_formats(mesh, tmp_path, name):
    """createTecplot bytes (BCH.cpp:4249-4482) and readRestartFile (BCH.cpp:4214-4242): product == oracle."""
    import fracstep
    import orc
    path = mesh(name, **MESHES[name])
    o = orc.Oracle(path, "explicit")
    o.run(2)
    s = fracstep.Solver(path, "explicit", init_gpu=False)
    f_o, f_p = str(tmp_path / "o.dat"), str(tmp_path / "p.dat")
    o.write_tecplot(f_o, "lidDrivenCavity")
    s.write_tecplot_fields(f_p, "lidDrivenCavity", o.U, o.P)
    a, b = open(f_o).read(), open(f_p).read()
    assert a == b
    lines = a.split("\n")
    assert lines[0] == "TITLE = lidDrivenCavity" and lines[1] == "VARIABLES = x,  y,  z,  u, v, w, p"
    NN, NE = o.i("NN"), o.i("NE")
    nsub = 8 if o.i("NENv") == 27 else 1
    assert lines[2] == "ZONE N=%d, E=%d, F=FEPOINT, ET=BRICK" % (NN, nsub * NE)
    assert len(lines) == 3 + NN + nsub * NE + 1
    # restart: 11 significant digits survive
    U, P = s.read_restart_fields(f_p)
    assert np.allclose(U, o.U, rtol=2e-11, atol=1e-300) and np.allclose(P, o.P, rtol=2e-11, atol=1e-300)
    o2 = orc.Oracle(path, "explicit")
    o2.read_restart(f_o)
    assert np.array_equal(o2.U, U) and np.array_equal(o2.P, P)


def test_device_matrix_layout_round_trip(mesh):
    """CSR -> sliced layout -> CSR on the host (no GPU): every lane-per-row setting, explicit slices (Q2 numbering) and
    offset-compressed slices (structured Q1 numbering: most slices must compress, boundary rows zero-padded onto the
    shared offset list)."""
    import fracstep
    s = fracstep.Solver(mesh("h_q2_c6", **MESHES["h_q2_c6"]), "implicit", init_gpu=False)
    NN, NNp = s.i("NN"), s.i("NNp")
    for (rs, cl, v, ncols) in ((s.arr("sparseMrowStarts"), s.arr("sparseMcol"), s.arr("sparseKvalue"), NN),
                               (s.arr("sparseGrowStarts"), s.arr("sparseGcol"), s.arr("sparseG2value"), NNp),
                               (s.arr("sparseZrowStarts"), s.arr("sparseZcol"), s.arr("sparseZvalue"), NNp)):
        for T in (0, 1, 2, 4, 8):
            bad, ns, nc, ratio = fracstep.sell_layout_check(rs, cl, v, ncols, T)
            assert bad == 0 and 1.0 <= ratio < 2.0   # (T = 8 on 15-entry rows of a 6^3 mesh pads heavily; never chosen by the library)
    q1 = fracstep.Solver(scheme="implicit", cavity=dict(nx=40, NENv=8), init_gpu=False)
    n = q1.i("NN")
    bad, ns, nc, ratio = fracstep.sell_layout_check(q1.arr("sparseZrowStarts"), q1.arr("sparseZcol"), q1.arr("sparseZvalue"), n, 1)
    assert bad == 0 and nc >= 0.95 * ns and ratio < 1.12        # structured numbering: (nearly) every slice shares its offsets
    for T in (2, 4):                                              # several lanes per row: offsets stored [slot][lane-in-row]
        bad, ns, nc, ratio = fracstep.sell_layout_check(q1.arr("sparseZrowStarts"), q1.arr("sparseZcol"), q1.arr("sparseZvalue"), n, T)
        assert bad == 0 and nc >= 0.9 * ns
    e = fracstep.Solver(mesh("h_q2_c6", **MESHES["h_q2_c6"]), "explicit", init_gpu=False)   # 125-point Z = G^T Md^-1 G on corner nodes
    for T in (1, 4, 8):
        bad, ns, nc, ratio = fracstep.sell_layout_check(e.arr("sparseZrowStarts"), e.arr("sparseZcol"), e.arr("sparseZvalue"), e.i("NNp"), T)
        assert bad == 0 and nc >= 0.8 * ns
    # a random unstructured matrix never compresses and still round-trips
    import scipy.sparse as sp
    A = sp.random(500, 300, density=0.03, format="csr", random_state=1)
    A.sort_indices()
    bad, ns, nc, ratio = fracstep.sell_layout_check(A.indptr, A.indices, A.data, 300, 1)
    assert bad == 0 and nc == 0
