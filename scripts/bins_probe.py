import sys; sys.path.insert(0,".")
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H
def run(name, ttype, grid, pts, T, **kw):
  M=pts.shape[0]; N=int(np.prod(grid))
  try:
    plan=_lib.Plan(ttype,grid[::-1],-1,T,1e-6,0,profile=1,**kw)
  except Exception as e:
    print(name, kw, "ERR", str(e)[:80]); return
  dp=torch.from_numpy(pts).cuda(); c=torch.from_numpy(H.random_complex((T,M),1)).cuda(); f=torch.from_numpy(H.random_complex((T,N),2)).cuda()
  best=None
  for _ in range(5):
    plan.set_points_interleaved(M,dp.data_ptr(),None); plan.execute(c.data_ptr(),f.data_ptr(),None); torch.cuda.synchronize(); t=plan.timings()
    if best is None or t["spread_interp_ms"]<best["spread_interp_ms"]: best=t
  print(name, kw, {k:round(v,4) for k,v in best.items()}, flush=True)
  plan.close()
sp=H.spiral_points(32,62500)
for nc in (2,4,8):
  for b in [(16,16),(16,8),(16,24),(32,8)]:
    run("cfg2 type1 ws nc", 1,(512,512),sp,8,bin_dims=b,coils_per_cta=nc)
