#!/usr/bin/env python
"""Measured floor of the implicit comparison AT THE REFERENCE'S OWN TOLERANCES (BiCGStab 1e-6, plain CG 1e-3).

The implicit Q2/Q1 momentum matrix is under-integrated (27-node mass matrix, 8 Gauss points, GQ.cpp:3490-3500; cond ~ 1e4), so two
correct solves that both meet |b - A u| <= 1e-6 |b| differ by ~1e-6 * cond in the fields.  This script prints, for the as-is
reference fixtures (tests/golden/ref_gq_cav4.npz, ref_gq_ne27000.npz):
  * oracle (1 thread) vs oracle (all threads)        -- the same code, another summation order of its dot products
  * oracle vs the reference program's output          -- for both placements of BiCGStab's residual test
  * with a GPU: the CUDA path vs the reference program's output
The bars of tests/test_reference_pin.py::bars for these fixtures are 3x the largest value printed here.
Run: python scripts/asis_spread.py > profiles/r02_asis_spread.txt"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("tests", "oracle", "cfd-with-cuda_b200"):
    sys.path.insert(0, os.path.join(ROOT, p))
import meshgen   # noqa: E402
import orc       # noqa: E402
from test_reference_pin import errors, load   # noqa: E402


def main():
    try:
        import torch
        gpu = torch.cuda.is_available()
    except Exception:
        gpu = False
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    with tempfile.TemporaryDirectory() as d:
        for name in ("ref_gq_cav4", "ref_gq_ne27000"):
            g, meta = load(name)
            path = os.path.join(d, name + ".inp")
            getattr(meshgen, meta["kind"])(path, **meta["kw"])
            runs = {}
            for hs in (1, 0):
                for thr in (1, 0):
                    o = orc.Oracle(path, "implicit", thr, cg_precond=0, cg_tol=1e-3, bicg_halfstep=hs)
                    o.run(meta["steps"])
                    runs[(hs, thr)] = (o.U, o.P)
                    eu, ep = errors(o.U, o.P, g, meta, o.i("NN"))
                    print("%-16s oracle halfstep=%d threads=%-3s vs reference: U %.2e  P %.2e   cg %s (reference %s)  bicgstab %s" % (
                        name, hs, thr or orc.threads(), eu, ep, list(map(int, o.arr("cgIters"))), list(map(int, g["cg"][:, 0])),
                        list(map(int, o.arr("bicgIters")))))
                    o.close()
                a, b = runs[(hs, 1)], runs[(hs, 0)]
                print("%-16s oracle halfstep=%d 1 thread vs all threads:        U %.2e  P %.2e" % (name, hs, rel(a[0], b[0]), rel(a[1], b[1])))
            if gpu:
                import fracstep
                for hs in (1, 0):
                    for k in range(3):
                        s = fracstep.Solver(path, "implicit", opts=fracstep.options(cg_no_precond=1, cg_tol=1e-3, bicg_no_halfstep=1 - hs))
                        reps = s.ctx.time_loop(meta["steps"])
                        U, P = s.ctx.fields()
                        eu, ep = errors(U, P, g, meta, s.i("NN"))
                        print("%-16s CUDA   halfstep=%d run %d       vs reference: U %.2e  P %.2e   cg %s  bicgstab %s  true residuals <= %.1e" % (
                            name, hs, k, eu, ep, [r["cg_iterations"] for r in reps], [r["bicg_iterations"] for r in reps],
                            max(v for r in reps for v in r["bicg_true_residual"])))
                        s.close()


if __name__ == "__main__":
    main()
