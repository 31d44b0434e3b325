import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'oracle')
import numpy as np, ttn_b200 as t
from ttn_b200 import _lib
lib=_lib.lib()
rng=np.random.default_rng(0)
for shape in [(128,128),(128,1024),(64,64),(256,256),(128,512)]:
    A=np.asfortranarray(rng.standard_normal(shape))
    t0=time.time(); U,s,Vt=t.svdtrunc(A); dt=time.time()-t0
    sref=np.linalg.svd(A,compute_uv=False)
    print(shape,"sweeps",lib.ttn_last_jacobi_sweeps(),"sigma err",np.abs(s-sref).max()/sref[0],"orth",np.abs(U.T@U-np.eye(len(s))).max(),"ms",dt*1e3)
Ac=np.asfortranarray(rng.standard_normal((128,512))+1j*rng.standard_normal((128,512)))
U,s,Vt=t.svdtrunc(Ac); print("complex sweeps",lib.ttn_last_jacobi_sweeps(), np.abs(s-np.linalg.svd(Ac,compute_uv=False)).max()/s[0])
