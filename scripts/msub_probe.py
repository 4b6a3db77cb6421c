import sys; sys.path.insert(0,".")
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H
def run(name, ttype, grid, pts, T, **kw):
  M=pts.shape[0]; N=int(np.prod(grid))
  plan=_lib.Plan(ttype,grid[::-1],-1,T,1e-6,0,profile=1,**kw)
  dp=torch.from_numpy(pts).cuda(); c=torch.from_numpy(H.random_complex((T,M),1)).cuda(); f=torch.from_numpy(H.random_complex((T,N),2)).cuda()
  best=None
  for _ in range(6):
    plan.set_points_interleaved(M,dp.data_ptr(),None); plan.execute(c.data_ptr(),f.data_ptr(),None); torch.cuda.synchronize(); t=plan.timings()
    if best is None or t["spread_interp_ms"]<best["spread_interp_ms"]: best=t
  print(name, kw, {k:round(v,4) for k,v in best.items()})
  plan.close()
r=H.radial_points(200,500)
for tt in (2,1):
  for ms in (0,512,256,128,64):
    run(f"cfg1 type{tt}", tt,(256,256),r,1,max_subproblem_size=ms)
  run(f"cfg1 type{tt} global", tt,(256,256),r,1,spread_method=1,interp_method=1)
sp=H.spiral_points(32,62500)
for ms in (0,512,256):
  run("cfg2 type1", 1,(512,512),sp,8,max_subproblem_size=ms)
  run("cfg2 type2", 2,(512,512),sp,8,max_subproblem_size=ms)
