#!/usr/bin/env python
"""EBE vs CSR, operator by operator (BASELINE.json configs[3]): K, G, G^T and the FE Laplacian applied through the assembled Sell
matrices (path 0) and through the matrix-free element kernels (path 1) on the same mesh; time per application, GB/s in each path's
own byte model (assembled: bytes the layout stores + vectors; matrix-free: node ids + corner coordinates + gathered / scattered values),
and the relative difference of the two results.
    python scripts/operator_bench.py 128 8      # 128^3 Q1/Q1
    python scripts/operator_bench.py 64 27      # 64^3 Q2/Q1"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cfd-with-cuda_b200"))
import numpy as np   # noqa: E402

import fracstep      # noqa: E402


def main():
    n, nenv = int(sys.argv[1]), int(sys.argv[2])
    rng = np.random.default_rng(1)
    for scheme, ops in (("explicit", ("K", "G", "GT")), ("implicit", ("Z",))):
        s = fracstep.Solver(scheme=scheme, cavity=dict(nx=n, NENv=nenv))
        c = s.ctx
        for op in ops:
            x = rng.standard_normal(c.NNp if op in ("G", "Z") else 3 * c.NN)
            y0, ms0, b0 = c.operator_apply(op, x, path=0, repeat=5)
            y1, ms1, b1 = c.operator_apply(op, x, path=1, repeat=5)
            print(json.dumps(dict(mesh="%d^3 %s" % (n, "Q2/Q1" if nenv == 27 else "Q1/Q1"), operator=op,
                                  csr=dict(ms=round(ms0, 4), gbs=round(b0 / ms0 / 1e6, 1), bytes=b0),
                                  ebe=dict(ms=round(ms1, 4), gbs=round(b1 / ms1 / 1e6, 1), bytes=b1),
                                  ebe_over_csr_time=round(ms1 / ms0, 3), rel_diff=float(np.linalg.norm(y1 - y0) / np.linalg.norm(y0)))), flush=True)
        s.close()


if __name__ == "__main__":
    main()
