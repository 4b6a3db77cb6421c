import sys; sys.path.insert(0,".")
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H
sp=H.spiral_points(32,62500); M=sp.shape[0]; T=32; grid=(512,512); N=512*512
dp=torch.from_numpy(sp).cuda(); c=torch.from_numpy(H.random_complex((T,M),1)).cuda(); f=torch.empty((T,N),dtype=torch.complex64,device="cuda")
for mb in (8,16,32):
  plan=_lib.Plan(1,grid[::-1],1,T,1e-6,0,profile=1,max_batch_size=mb)
  st=torch.cuda.current_stream().cuda_stream
  for _ in range(3):
    plan.set_points_interleaved(M,dp.data_ptr(),st); plan.execute(c.data_ptr(),f.data_ptr(),st)
  torch.cuda.synchronize()
  e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(10):
    plan.set_points_interleaved(M,dp.data_ptr(),st); plan.execute(c.data_ptr(),f.data_ptr(),st)
  e1.record(); torch.cuda.synchronize()
  print("max_batch",mb,"ms/step",e0.elapsed_time(e1)/10, plan.timings())
  plan.close()
