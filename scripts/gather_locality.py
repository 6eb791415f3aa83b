"""How local are the x-gathers of one slice (warp) of the Q2 momentum matrix?  Distinct 32-byte sectors / 128-byte lines of x touched
per slice of R rows, for the reference's node numbering, reverse Cuthill-McKee and a geometric (z, y, x) sort.  CPU only."""
import sys, numpy as np, scipy.sparse as sp
from scipy.sparse.csgraph import reverse_cuthill_mckee
sys.path.insert(0, "cfd-with-cuda_b200")
import fracstep
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
s = fracstep.Solver(scheme="implicit", opts=fracstep.options(), init_gpu=False, cavity=dict(nx=n, cluster=2.0 if n % 2 == 0 else 0.0, NENv=27, dt=1e-3, viscosity=0.001))
NN = s.i("NN")
A = sp.csr_matrix((np.ones(len(s.arr("sparseMcol"))), s.arr("sparseMcol"), s.arr("sparseMrowStarts")), shape=(NN, NN))
xyz = s.arr("coord").reshape(-1, 3)[:NN]

def stats(P, R=16):
    B = A[P][:, P].tocsr()
    rs, col = B.indptr, B.indices
    sec, lin, nnz = 0, 0, 0
    for s0 in range(0, NN, R):
        c = col[rs[s0]:rs[min(NN, s0 + R)]]
        sec += len(np.unique(c >> 2)); lin += len(np.unique(c >> 4)); nnz += len(c)
    return nnz / (NN / R), sec / (NN / R), lin / (NN / R)

orders = {"reference numbering": np.arange(NN), "RCM": reverse_cuthill_mckee(A.tocsr(), symmetric_mode=True),
          "sort by (z, y, x)": np.lexsort((xyz[:, 0], xyz[:, 1], xyz[:, 2]))}
print("NN %d, nnz %d, slices of 16 rows (T = 2)" % (NN, A.nnz))
for k, P in orders.items():
    a, b, c = stats(np.asarray(P))
    print("%-22s nnz/slice %7.1f   32B sectors/slice %7.1f   128B lines/slice %6.1f   -> gathered bytes per non-zero (sector granularity) %.2f" % (k, a, b, c, 32 * b / a))

# constrained variants: the first NCN numbers must stay with the corner (pressure) nodes
NCN = s.i("NNp")
def corners_first(order):
    order = np.asarray(order)
    return np.concatenate([order[order < NCN], order[order >= NCN]])
for k in ("RCM", "sort by (z, y, x)"):
    P = corners_first(orders[k])
    a, b, c = stats(P)
    print("%-22s nnz/slice %7.1f   32B sectors/slice %7.1f   128B lines/slice %6.1f   -> gathered bytes per non-zero (sector granularity) %.2f" % (k + ", corners first", a, b, c, 32 * b / a))
# interleaved by element: nodes in order of first appearance in the element list (x-fastest element order), corners first
L = s.arr("LtoGnode").reshape(-1, 27)
seen = np.zeros(NN, bool); order = []
for e in range(L.shape[0]):
    for v in L[e]:
        if not seen[v]: seen[v] = True; order.append(v)
P = corners_first(np.array(order))
a, b, c = stats(P)
print("%-22s nnz/slice %7.1f   32B sectors/slice %7.1f   128B lines/slice %6.1f   -> %.2f" % ("element order, corners first", a, b, c, 32 * b / a))
