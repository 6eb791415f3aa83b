#!/usr/bin/env python
"""Prints, for every fixture written by the reference itself (tests/golden/ref_*.npz), the rel-L2 difference of the CUDA
path's fields (through the C ABI, cuda:0) and the Krylov iteration counts on both sides.  Output kept in profiles/."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("cfd-with-cuda_b200", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import fracstep  # noqa: E402
import meshgen  # noqa: E402
import test_reference_pin as T  # noqa: E402

print("%-24s %-8s %5s  %-9s %-9s  %s" % ("fixture", "scheme", "steps", "relL2(U)", "relL2(P)", "CG per step: CUDA | reference   (BiCGStab: CUDA | reference)"))
for name in T.SMALL + T.FULL:
    g, meta = T.load(name)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "case.inp")
        getattr(meshgen, meta["kind"])(path, **meta["kw"])
        popt = T.settings(meta)[1]
        s = fracstep.Solver(path, meta["scheme"], opts=fracstep.options(device=0, **popt))
        reps = s.ctx.time_loop(meta["steps"])
        U, P = s.ctx.fields()
        eu, ep = T.errors(U, P, g, meta, s.i("NN"))
        cg_ref = [int(v) for v in g["cg"][:, 0]]
        per = len(cg_ref) // meta["steps"] if meta["scheme"] == "implicit" else None
        if meta["scheme"] == "explicit":
            line = "%s | %s" % ([r["cg_iterations"] for r in reps], "sum %d over %d solves" % (sum(cg_ref), len(cg_ref)))
            line = "sum %d | %s" % (sum(r["cg_iterations"] for r in reps), "sum %d over %d solves" % (sum(cg_ref), len(cg_ref)))
        else:
            line = "%s | %s   (%s | %s)" % ([r["cg_iterations"] for r in reps], cg_ref, [list(r["bicg_iterations"]) for r in reps],
                                            [int(v) for v in g["bicg"][:, 0]])
        print("%-24s %-8s %5d  %-9.1e %-9.1e  %s" % (name, meta["scheme"], meta["steps"], eu, ep, line))
        s.close()
