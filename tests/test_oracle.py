"""CPU: pins the oracle against everything the reference offers for this path.

The reference has no tests and no golden outputs.  What can be pinned:
closed-form counts of the one shipped input (lidDrivenCavity_NE27000.inp), the colouring the
reference's greedy loop produces on it, analytic FE invariants of the assembled operators, and
the committed golden fixtures (regression of the oracle itself)."""
import ast
import os

import numpy as np
import pytest
import scipy.sparse as sp

REF_INP = "/root/reference/inputFiles/lidDrivenCavity/lidDrivenCavity_NE27000.inp"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def csr(o, vals, which):
    rows, cols = {"M": ("NN", "NN"), "G": ("NN", "NNp"), "Z": ("NNp", "NNp")}[which]
    return sp.csr_matrix((o.arr(vals), o.arr("sparse%scol" % which), o.arr("sparse%srowStarts" % which)),
                         shape=(o.i(rows), o.i(cols)))


@pytest.fixture(scope="module")
def ne27000(mesh):
    import orc
    return orc.Oracle(mesh("ne27000", n=30, cluster=2.0), "explicit")


def test_generated_mesh_is_the_shipped_file(mesh):
    """meshgen.cube(30, cluster=2.0) reproduces the reference's only shipped input byte for byte."""
    if not os.path.exists(REF_INP):
        pytest.skip("reference tree not mounted (GPU box)")
    assert open(mesh("ne27000", n=30, cluster=2.0), "rb").read() == open(REF_INP, "rb").read()


def test_ne27000_closed_form_counts(ne27000):
    """NN=(2n+1)^3, NNp=(n+1)^3, nnzM=(8n+1)^3, nnzG=(5n+1)^3, nnzZ=(5n-1)^3, n=30."""
    o, n = ne27000, 30
    assert o.i("NE") == 27000 and o.i("NCN") == 29791
    assert o.i("NN") == (2 * n + 1) ** 3 == 226981
    assert o.i("NNp") == (n + 1) ** 3
    assert o.i("nnzM") == (8 * n + 1) ** 3 == 13997521
    assert o.i("nnzG") == (5 * n + 1) ** 3 == 3442951
    assert o.i("nnzZ") == (5 * n - 1) ** 3 == 3307949
    assert o.i("zeroPressureNode") == 480          # 1-based 481 in the file
    assert o.i("monPoint") == 14895                # centre of the 31^3 corner grid
    assert o.i("BCnVelFaces") == 5400
    assert o.i("BCnVelNodes") == 61 ** 3 - 59 ** 3


def test_ne27000_colouring(ne27000):
    """The reference's greedy colouring gives 8 colours x 3375 elements, no same-coloured neighbours."""
    o = ne27000
    assert o.i("nActiveColors") == 8 and o.i("colorConflicts") == 0
    assert list(o.arr("NmeshColors")[:8]) == [3375] * 8
    assert sorted(o.arr("elementsOfColor")) == list(range(27000))


def test_implicit_z_count(mesh):
    import orc
    o = orc.Oracle(mesh("q2_c6", n=6, cluster=2.0), "implicit")
    assert o.i("nnzZ") == (3 * 6 + 1) ** 3


@pytest.mark.parametrize("kw", [dict(n=4, cluster=2.0), dict(n=5, full_precision=True), dict(n=6, NENv=8, full_precision=True)])
def test_fe_invariants(mesh, kw):
    """Partition of unity, zero row sums of K, lumped mass = volume, G*1 = 0 on interior nodes,
    implicit Z symmetric with zero row sums before the pin."""
    import orc
    name = "inv_%d_%d" % (kw["n"], kw.get("NENv", 27))
    oe = orc.Oracle(mesh(name, **kw), "explicit")
    NN, NNp = oe.i("NN"), oe.i("NNp")
    nv = oe.i("NENv")
    assert np.allclose(oe.arr("Sv_1d").reshape(8, nv).sum(1), 1.0, atol=1e-14)
    assert np.allclose(oe.arr("Sp").reshape(8, 8).sum(1), 1.0, atol=1e-14)
    assert np.allclose(oe.arr("dSv").reshape(3, nv, 8).sum(1), 0.0, atol=1e-13)
    K = csr(oe, "sparseKvalue", "M")
    assert abs(K.sum(1)).max() < 1e-12 * abs(K).max()
    assert abs(K - K.T).max() < 1e-13 * abs(K).max()
    assert abs(1.0 / oe.arr("MdOrigInv")[:NN]).sum() == pytest.approx(1.0, rel=1e-12)   # unit cube
    bc = np.zeros(NN, bool)
    bc[oe.arr("BCvelNodes")[0::2]] = True
    for g in ("sparseG1value", "sparseG2value", "sparseG3value"):
        G = csr(oe, g, "G")
        assert abs((G @ np.ones(NNp))[~bc]).max() < 1e-13
    md = oe.arr("MdInv")
    assert np.all(md[:NN][bc] == 1.0)
    Ze = csr(oe, "sparseZvalue", "Z")
    assert abs(Ze - Ze.T).max() < 1e-12 * abs(Ze).max()
    oi = orc.Oracle(mesh(name, **kw), "implicit")
    Zi = csr(oi, "sparseZvalue", "Z").tolil()
    pin = oi.i("zeroPressureNode")
    if pin > 0:
        Zi[pin, pin] = Zi[pin, pin] / 1000.0
    Zi = Zi.tocsr()
    assert abs(Zi.sum(1)).max() < 1e-12 * abs(Zi).max()
    assert abs(Zi - Zi.T).max() < 1e-13 * abs(Zi).max()
    assert Zi.diagonal().max() < 0          # negative semi-definite FE Laplacian (GQ.cpp:3609)
    M = csr(oi, "sparseMvalue", "M")
    assert M.sum() == pytest.approx(1.0 / oi.d("dt"), rel=1e-12)   # implicit M carries 1/dt (GQ.cpp:3495)


def test_invalid_colouring_detected(mesh):
    """Odd element counts expose the reference's colouring defect (restart skips neighbour 0,
    BCH.cpp:880-886): the oracle counts the conflicts and serialises its element loops."""
    import orc
    o = orc.Oracle(mesh("q2_u5", n=5, full_precision=True), "explicit")
    assert o.i("colorConflicts") > 0
    o.run(1)
    r1 = o.arr("Un")
    o2 = orc.Oracle(mesh("q2_u5", n=5, full_precision=True), "explicit")
    o2.run(1)
    # no race: two runs agree to OpenMP reduction-order rounding (a lost update would show at ~1e-6)
    assert np.linalg.norm(r1 - o2.arr("Un")) <= 1e-12 * np.linalg.norm(r1)


@pytest.mark.parametrize("name", ["q2_c4_explicit", "q2_c6_explicit", "q1_u8_explicit", "q1_u8_implicit", "q2_c4_implicit_tight"])
def test_oracle_matches_golden_fixture(mesh, name):
    """The committed golden vectors (tests/golden/make_golden.py) are reproduced by the oracle built here."""
    import orc
    g = np.load(os.path.join(GOLD, name + ".npz"))
    mk, opt = ast.literal_eval(str(g["mesh"])), ast.literal_eval(str(g["options"]))
    o = orc.Oracle(mesh("gold_" + name, **mk), str(g["scheme"]), **opt)
    assert (o.i("NN"), o.i("NNp"), o.i("nnzM"), o.i("nnzG"), o.i("nnzZ")) == tuple(int(g[k]) for k in ("NN", "NNp", "nnzM", "nnzG", "nnzZ"))
    o.run(int(g["steps"]))
    tight = "tight" in name
    tol = 1e-9 if str(g["scheme"]) == "explicit" else 1e-7   # OpenMP reduction order differs with thread count / run
    assert np.linalg.norm(o.U - g["U"]) <= tol * np.linalg.norm(g["U"])
    # pressure is only determined to the CG tolerance (1e-6) times the conditioning of Z
    assert np.linalg.norm(o.P - g["P"]) <= (1e-6 if tight else 1e-4) * np.linalg.norm(g["P"])
    assert np.all(np.abs(o.arr("cgIters") - g["cgIters"]) <= 2)
    if str(g["scheme"]) == "explicit":
        assert list(o.arr("innerIters")) == list(g["innerIters"])
