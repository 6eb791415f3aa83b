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
