import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0,'cfd-with-cuda_b200')
import fracstep, meshgen
def csr(s, v, which, shape): return sp.csr_matrix((s.arr(v), s.arr('sparse%scol'%which), s.arr('sparse%srowStarts'%which)), shape=shape)
def check(cav, scheme, W):
    g = fracstep.Solver(scheme=scheme, cavity=cav, init_gpu=False)
    NN, NNp = g.i('NN'), g.i('NNp')
    Kg = csr(g,'sparseKvalue','M',(NN,NN)); Zg = csr(g,'sparseZvalue','Z',(NNp,NNp)); G1g = csr(g,'sparseG1value','G',(NN,NNp))
    locs=[fracstep.Solver(scheme=scheme, cavity=cav, init_gpu=False, world=W, rank=r) for r in range(W)]
    ownedV=np.zeros(NN,int); ownedP=np.zeros(NNp,int)
    for r,s in enumerate(locs):
        ng=s.arr('part_nodeGlobal'); pg=s.arr('part_presGlobal'); NNl,NPl=s.i('NN'),s.i('NNp'); NNo,NPo=s.i('NN_owned'),s.i('NNp_owned')
        assert len(ng)==NNl and len(pg)==NPl
        ownedV[ng[:NNo]]+=1; ownedP[pg[:NPo]]+=1
        Kl=csr(s,'sparseKvalue','M',(NNl,NNl)).tocsr()[:NNo]
        # map to global
        Pn=sp.csr_matrix((np.ones(NNl),(np.arange(NNl),ng)),shape=(NNl,NN)); Pp=sp.csr_matrix((np.ones(NPl),(np.arange(NPl),pg)),shape=(NPl,NNp))
        d=abs(Kl@Pn - Kg[ng[:NNo]]).max(); assert d<1e-13*abs(Kg).max(), ('K',d)
        Zl=csr(s,'sparseZvalue','Z',(NPl,NPl)).tocsr()[:NPo]
        d=abs(Zl@Pp - Zg[pg[:NPo]]).max(); assert d<1e-12*abs(Zg).max(), ('Z',d, abs(Zg).max())
        Gl=csr(s,'sparseG1value','G',(NNl,NPl)).tocsr()[:NNo]
        d=abs(Gl@Pp - G1g[ng[:NNo]]).max(); assert d<1e-13*abs(G1g).max(), ('G',d)
        if scheme=='explicit':
            assert np.allclose(s.arr('MdInv')[:NNo], g.arr('MdInv')[ng[:NNo]], rtol=1e-13)
        # BC nodes
        bl=s.arr('BCvelNodes').reshape(-1,2); bg=g.arr('BCvelNodes').reshape(-1,2)
        mask=bl[:,0]<NNo
        gb=dict(zip(bg[:,0],bg[:,1]))
        assert all(gb.get(ng[n],-1)==w for n,w in bl[mask]) and mask.sum()==sum(1 for n in bg[:,0] if n in set(ng[:NNo]))
        # plan consistency
        peers=s.arr('peerRank'); svp=s.arr('sendVptr'); svi=s.arr('sendVidx'); spp=s.arr('sendPptr'); spi=s.arr('sendPidx')
        for k,q in enumerate(peers):
            t=locs[q]; tp=list(t.arr('peerRank')); assert r in tp, (r,q)
            kk=tp.index(r)
            off,cnt=t.arr('recvVoff')[kk],t.arr('recvVcnt')[kk]
            assert np.array_equal(ng[svi[svp[k]:svp[k+1]]], t.arr('part_nodeGlobal')[off:off+cnt]), 'V plan'
            off,cnt=t.arr('recvPoff')[kk],t.arr('recvPcnt')[kk]
            assert np.array_equal(pg[spi[spp[k]:spp[k+1]]], t.arr('part_presGlobal')[off:off+cnt]), 'P plan'
        print('  rank',r,'NE',s.i('NE'),'NEa',s.i('NE_apply'),'NN',NNl,'owned',NNo,'NNp',NPl,'owned',NPo,'peers',list(peers),'pin',s.i('pinLocal'),'mon',s.i('monLocal'))
    assert (ownedV==1).all() and (ownedP==1).all()
for cav,scheme,W in ((dict(nx=6,NENv=27),'explicit',2),(dict(nx=6,NENv=27),'implicit',3),(dict(nx=8,NENv=8),'explicit',4),(dict(nx=5,ny=4,nz=7,NENv=27,cluster=0.0),'explicit',3),(dict(nx=8,NENv=8),'implicit',8)):
    print(cav,scheme,W); check(cav,scheme,W)
print('OK')
