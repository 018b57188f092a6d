import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import oracle
from cholla_b200 import initial_conditions as ic
from helpers import ghosted, real_view
n=(40,32,24); flags=[3,3,2,2,1,1]; gamma=1.66666667; g=4
fields = ic.spherical_overpressure(n)
flat, dims = ghosted(fields, g, flags)
dx,dy,dz=[1.0/v for v in n]
a=flat.copy(); b=flat.copy()
oracle.advance(a,*dims,g,dx,dy,dz,gamma,2,0,flags,1)
b[:]=a
rng=np.random.default_rng(0)
b*= (1.0 + (rng.integers(-1,2,size=b.shape))*1.1e-16)   # +-1 ulp noise after step 1
for s in range(3):
    oracle.advance(a,*dims,g,dx,dy,dz,gamma,2,0,flags,1)
    oracle.advance(b,*dims,g,dx,dy,dz,gamma,2,0,flags,1)
    d=np.abs(real_view(a,dims,g)-real_view(b,dims,g))
    print("after",s+2,"steps: max abs diff/field",d.reshape(5,-1).max(axis=1),"n>1e-10",(d>1e-10).sum())
