import sys, ctypes as C
sys.path.insert(0,'cfd-with-cuda_b200')
import fracstep
L=fracstep.lib()
L.cfd_diag_sync.argtypes=[C.c_int,C.c_int,C.c_int,C.c_int,C.POINTER(C.c_double)]
for blocks,threads in ((148,256),(148,512),(148,1024),(296,512),(444,512),(592,256),(74,1024),(37,1024)):
    r=[]
    for mode in (0,1,2):
        us=C.c_double(); rc=L.cfd_diag_sync(blocks,threads,2000,mode,C.byref(us)); r.append(us.value if rc==0 else None)
    print(blocks,threads,'sync %.2f us  gsum1 %.2f us  gsum3 %.2f us'%tuple(r))
