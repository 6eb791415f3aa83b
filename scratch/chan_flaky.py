import sys, os
sys.path.insert(0, 'cfd-with-cuda_b200'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np, fracstep, orc, meshgen
TIGHT = dict(cg_tol=1e-12, bicg_tol=1e-9, cg_maxit=5000, bicg_maxit=5000)
meshgen.channel('/tmp/ch.inp', nx=12, ny=8, nz=8, full_precision=True)
o = orc.Oracle('/tmp/ch.inp', "implicit", **TIGHT); o.run(3)
print("oracle bicg", list(o.arr("bicgIters")), "cg", list(o.arr("cgIters")))
rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
for k in range(4):
    s = fracstep.Solver('/tmp/ch.inp', "implicit", opts=fracstep.options(**TIGHT))
    reps = s.ctx.time_loop(3)
    U, P = s.ctx.fields()
    print(k, "relU %.2e relP %.2e" % (rel(U, o.U), rel(P, o.P)), [list(r["bicg_iterations"]) for r in reps], [r["cg_iterations"] for r in reps])
    s.close()
