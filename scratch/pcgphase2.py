import sys, os, numpy as np, shutil
root=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# use the timer build
shutil.copy(os.path.join(root,'scratch','libcfdb200_timers.so'), '/tmp/libcfdb200.so')
sys.path.insert(0, os.path.join(root,'cfd-with-cuda_b200'))
import fracstep
fracstep._LIBPATH='/tmp/libcfdb200.so'
os.environ['CFD_PCG_PHASES']='1'
for n in (64,128):
    s=fracstep.Solver(scheme='implicit', cavity=dict(nx=n,NENv=8), init_gpu=False)
    rs,cl,v=s.arr('sparseZrowStarts'),s.arr('sparseZcol'),s.arr('sparseZvalue')
    b=np.random.default_rng(0).standard_normal(len(rs)-1); b-=b.mean()
    for rep in range(2):
        x,its,ms=fracstep.csr_pcg(rs,cl,v,b,tol=1e-6,maxit=200)
    N=len(rs)-1; nnz=len(cl)
    print('n',n,'its',its,'ms',ms,'us/it %.1f'%(1e3*ms/its), 'alg GB/s %.0f'%((12*nnz+4*N+104*N)*its/ms/1e6), 'spmv bytes MB %.0f'%((8.13*nnz+16*N)/1e6), 'upd MB %.0f'%(56*N/1e6), 'pupd MB %.0f'%(32*N/1e6))
