import sys, os, numpy as np
sys.path.insert(0,'cfd-with-cuda_b200')
import fracstep, meshgen
os.environ['CFD_PCG_PHASES']='1'
meshgen.cube('/tmp/c30.inp', 30, cluster=2.0)
s=fracstep.Solver('/tmp/c30.inp','explicit',init_gpu=False)
rs,cl,v=s.arr('sparseZrowStarts'),s.arr('sparseZcol'),s.arr('sparseZvalue')
b=np.random.default_rng(0).standard_normal(len(rs)-1)
for T in (2,4,8):
    os.environ['CFD_Z_T']=str(T)
    for rep in range(2):
        x,its,ms=fracstep.csr_pcg(rs,cl,v,b,tol=1e-6,maxit=1000)
    print('T',T,'its',its,'ms',ms,'us/it',1e3*ms/its)
