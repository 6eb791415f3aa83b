#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot be compiled here:
MKL / CSparse / Paralution / CUSP are absent; the reference ships no golden
vectors of its own).  Each fixture holds the fields and Krylov iteration counts
after N time steps of a small generated mesh, plus the mesh parameters needed to regenerate the
.inp with cfd-with-cuda_b200/meshgen.py.  Run:  python tests/golden/make_golden.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "cfd-with-cuda_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = {
    # name: (mesh kwargs, scheme, steps, oracle options)
    "q2_c4_explicit": (dict(n=4, cluster=2.0), "explicit", 6, {}),
    "q2_c6_explicit": (dict(n=6, cluster=2.0), "explicit", 4, {}),
    "q1_u8_explicit": (dict(n=8, NENv=8, full_precision=True), "explicit", 4, {}),
    "q1_u8_implicit": (dict(n=8, NENv=8, full_precision=True), "implicit", 5, {}),
    "q2_c4_implicit_tight": (dict(n=4, cluster=2.0, viscosity=0.001), "implicit", 4,
                             dict(cg_tol=1e-12, bicg_tol=1e-12, cg_maxit=5000, bicg_maxit=5000)),
}


def main():
    import meshgen
    import orc
    for name, (mk, scheme, steps, opt) in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, name + ".inp")
            meshgen.cube(path, **mk)
            o = orc.Oracle(path, scheme, **opt)
            o.run(steps)
            np.savez_compressed(os.path.join(HERE, name + ".npz"), U=o.U, P=o.P, cgIters=o.arr("cgIters"),
                                bicgIters=o.arr("bicgIters"), innerIters=o.arr("innerIters"), maxAcc=o.d("maxAcc"),
                                steps=steps, scheme=scheme, mesh=repr(mk), options=repr(opt),
                                NN=o.i("NN"), NNp=o.i("NNp"), nnzM=o.i("nnzM"), nnzG=o.i("nnzG"), nnzZ=o.i("nnzZ"))
            print(name, "NN", o.i("NN"), "cg", list(o.arr("cgIters"))[:8], "|U|", float(np.linalg.norm(o.U)))
            o.close()


if __name__ == "__main__":
    main()
