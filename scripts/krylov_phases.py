#!/usr/bin/env python
"""Where the time of a Krylov iteration goes: per-phase %globaltimer clocks of block 0 inside the persistent kernels.

Needs the diagnostics build (`make -C cfd-with-cuda_b200/csrc timers` -> cfd-with-cuda_b200/libcfdb200_timers.so); the clocks are printed
to stderr by the library.  Usage (on a GPU box):
    python scripts/krylov_phases.py pcg  ne27000_explicit|ne27000_implicit|q1_64|q1_128   # both CG recurrences on that pressure matrix
    python scripts/krylov_phases.py bicg ne27000|q2_64                                    # the lock-step momentum BiCGStab
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ.setdefault("CFD_LIBPATH", os.path.join(ROOT, "cfd-with-cuda_b200", "libcfdb200_timers.so"))
os.environ["CFD_PCG_PHASES"] = "1"
sys.path.insert(0, os.path.join(ROOT, "cfd-with-cuda_b200"))
import numpy as np   # noqa: E402

import fracstep      # noqa: E402
import meshgen       # noqa: E402


def pcg(which):
    if which.startswith("ne27000"):
        scheme = which.split("_")[1]
        meshgen.cube("/tmp/kp.inp", 30, cluster=2.0)
        s = fracstep.Solver("/tmp/kp.inp", scheme, init_gpu=False)
    else:
        s = fracstep.Solver(scheme="implicit", cavity=dict(nx=int(which.split("_")[1]), NENv=8), init_gpu=False)
    rs, cl, v = s.arr("sparseZrowStarts"), s.arr("sparseZcol"), s.arr("sparseZvalue")
    N, nnz = len(rs) - 1, len(cl)
    b = np.random.default_rng(0).standard_normal(N)
    b -= b.mean()
    for variant in (1, 2):
        for rep in range(2):
            x, its, ms = fracstep.csr_pcg(rs, cl, v, b, tol=1e-6, maxit=300, variant=variant)
        print("%s variant %d: rows %d nnz %d its %d  %.2f us/iteration  (%.0f GB/s algorithmic)" % (
            which, variant, N, nnz, its, 1e3 * ms / max(its, 1), (12.0 * nnz + 4.0 * N + 104.0 * N) * its / ms / 1e6), flush=True)


def bicg(which):
    if which == "ne27000":
        meshgen.cube("/tmp/kp.inp", 30, cluster=2.0, viscosity=0.001)
    else:
        meshgen.cube("/tmp/kp.inp", 64, cluster=0.0, viscosity=0.01, full_precision=True)
    s = fracstep.Solver("/tmp/kp.inp", "implicit", opts=fracstep.options(device=0))
    reps, ms = s.ctx.time_loop_timed(4)
    print("ms/step", ms / 4, [r["bicg_iterations"] for r in reps], [r["cg_iterations"] for r in reps], flush=True)


if __name__ == "__main__":
    (pcg if sys.argv[1] == "pcg" else bicg)(sys.argv[2])
