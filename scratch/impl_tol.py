import sys, numpy as np
sys.path.insert(0,'cfd-with-cuda_b200'); sys.path.insert(0,'oracle')
import meshgen, orc, fracstep
def rel(a,b): return np.linalg.norm(a-b)/np.linalg.norm(b)
meshgen.cube('/tmp/q2_c4.inp', 4, cluster=2.0)
meshgen.cube('/tmp/q2_u6.inp', 6, full_precision=True)
for inp in ('/tmp/q2_c4.inp','/tmp/q2_u6.inp'):
  for tol,mx in ((1e-6,1000),(1e-9,3000),(1e-12,5000)):
    o=orc.Oracle(inp,'implicit', bicg_tol=tol, cg_tol=tol, bicg_maxit=mx, cg_maxit=mx)
    c=fracstep.Context(fracstep.Problem(o,'implicit'), fracstep.options(bicg_tol=tol, cg_tol=tol, bicg_maxit=mx, cg_maxit=mx))
    o.run(5); reps=c.time_loop(5); U,P=c.fields()
    print(inp, tol, 'U',rel(U,o.U),'P',rel(P,o.P))
    print('   oracle bicg', o.arr('bicgIters').reshape(-1,3).tolist(), 'cg', o.arr('cgIters').tolist())
    print('   gpu    bicg', [r['bicg_iterations'] for r in reps], 'cg', [r['cg_iterations'] for r in reps])
