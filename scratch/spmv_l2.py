import sys, os, numpy as np
sys.path.insert(0,'cfd-with-cuda_b200')
import fracstep, meshgen
meshgen.cube('/tmp/c30.inp', 30, cluster=2.0)
s=fracstep.Solver('/tmp/c30.inp','explicit',init_gpu=False)
for nm,which in (('Z','Z'),('K','M')):
    rs,cl=s.arr('sparse%srowStarts'%which),s.arr('sparse%scol'%which); v=s.arr('sparseZvalue' if nm=='Z' else 'sparseKvalue')
    n=len(rs)-1; x=np.random.default_rng(0).standard_normal((1,n))
    byts=12*len(cl)+4*(n+1)+16*n
    for T in (1,2,4,8,16,32):
        os.environ['CFD_Z_T']=str(T)
        y,ms=fracstep.csr_spmv(rs,cl,v,x,repeat=30)
        print(nm,'T',T,'best ms %.4f'%ms,'GB/s %.0f'%(byts/ms/1e6), 'MB %.1f'%(byts/1e6))
