"""CPU: the .inp writers follow the reference generators' grammar and numbering."""
import numpy as np


def test_channel_standin(mesh):
    """Stand-in for the missing bendingSquareDuct_49x33x33.inp: 48x32x32 straight channel, inlet faces last."""
    import fracstep
    path = mesh("chan", kind="channel", nx=12, ny=8, nz=8)
    s = fracstep.Solver(path, "implicit", init_gpu=False)
    assert s.i("NE") == 12 * 8 * 8 and s.i("NN") == 25 * 17 * 17 and s.i("NNp") == 13 * 9 * 9
    coord = s.arr("coord").reshape(-1, 3)
    bc = s.arr("BCvelNodes").reshape(-1, 2)
    inlet = np.isclose(coord[bc[:, 0], 0], 0.0)
    assert np.all(bc[inlet, 1] == 1)                 # BC 2 (uniform inlet) wins on the inlet edges
    outlet_interior = np.isclose(coord[:, 0], 10.0) & (coord[:, 1] > 0) & (coord[:, 1] < 1) & (coord[:, 2] > 0) & (coord[:, 2] < 1)
    assert not np.isin(np.nonzero(outlet_interior)[0], bc[:, 0]).any()   # outlet left free
    pin = s.i("zeroPressureNode")
    assert np.allclose(coord[pin], [10.0, 0.5, 0.5])
    assert np.allclose(coord[s.i("monPoint")], [10.0, 0.5, 0.5])


def test_header_grammar(mesh):
    lines = open(mesh("hdr", n=2)).read().split("\n")
    keys = [l.split(":")[0].strip() for l in lines[2:20]]
    assert keys == ["eType", "NE", "NCN", "NENv", "NENp", "NGP", "alpha", "dt", "t_ini", "t_final", "maxIter", "tolerance",
                    "converge", "isRestart", "density", "viscosity", "fx", "fy"]


def test_duct_standin_has_a_prescribed_outflow_and_the_scheme_converges_on_it(mesh):
    """The well-posed configs[2] stand-in: the plug profile on the outlet face as well.  On it the reference's implicit scheme (oracle)
    settles -- BiCGStab converges well below its cap and the fields stay bounded -- where the free-outlet channel blows up by step 3."""
    import fracstep
    import orc
    path = mesh("duct", kind="duct", nx=12, ny=8, nz=8)
    s = fracstep.Solver(path, "implicit", init_gpu=False)
    coord = s.arr("coord").reshape(-1, 3)
    bc = s.arr("BCvelNodes").reshape(-1, 2)
    outlet_interior = np.isclose(coord[:, 0], 10.0) & (coord[:, 1] > 0) & (coord[:, 1] < 1) & (coord[:, 2] > 0) & (coord[:, 2] < 1)
    on = np.isin(np.nonzero(outlet_interior)[0], bc[:, 0])
    assert on.all()                                                       # every outlet node carries a velocity condition ...
    assert np.all(bc[np.isin(bc[:, 0], np.nonzero(outlet_interior)[0]), 1] == 1)   # ... the plug profile (BC 2)
    s.close()
    norms = {}
    for kind, p in (("duct", path), ("channel", mesh("chan", kind="channel", nx=12, ny=8, nz=8))):
        o = orc.Oracle(p, "implicit", 1)
        o.run(4)
        norms[kind] = float(np.linalg.norm(o.U))
        if kind == "duct":
            assert o.arr("bicgIters").max() < 1000 and o.arr("cgIters").max() < 1000
        o.close()
    assert np.isfinite(norms["duct"]) and norms["duct"] < 1e3
    assert not (np.isfinite(norms["channel"]) and norms["channel"] < 1e3)   # the free-outlet channel: NaN or astronomically large
