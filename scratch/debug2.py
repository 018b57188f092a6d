import sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import cholla_b200 as cb
from cholla_b200 import initial_conditions as ic
from cholla_b200.grid import Grid3D, n_ghost_for
from helpers import ghosted, real_view

n=(40,32,24); flags=[3,3,2,2,1,1]; gamma=1.66666667
fields = ic.spherical_overpressure(n)
g=4
flat, dims = ghosted(fields, g, flags)
Gs = Grid3D(n,(1.,1.,1.),recon=2,solver=0,gamma=gamma,bc=flags,strict=True); Gs.upload(flat)
Gf = Grid3D(n,(1.,1.,1.),recon=2,solver=0,gamma=gamma,bc=flags,strict=False); Gf.upload(flat)
for step in range(5):
    Gs.advance(1); Gf.advance(1)
    a=real_view(Gs.download(),dims,g); b=real_view(Gf.download(),dims,g)
    d=np.abs(a-b)
    print("step",step+1,"dt",Gs.H.dt,Gf.H.dt,"max abs diff per field",d.reshape(5,-1).max(axis=1), "n>1e-10", (d>1e-10).sum(), "n>1e-14",(d>1e-14).sum())
    if step==0:
        idx=np.argwhere(d[0]>1e-10)
        print(" cells (z,y,x) with rho diff>1e-10:", idx[:10].tolist())
        c=np.array([12,16,20])
        if len(idx): print(" radius of those cells (in cells):", np.sqrt(((idx-c+0.5)**2).sum(axis=1))[:10])
