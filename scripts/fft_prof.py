"""Scratch for ncu captures of the own FFT passes: one cfg4-shaped type-2 plan (2 x 512^3 fine grid)
and one cfg2-shaped type-1 plan (32 x 1024^2), few points, two executes each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H
def one(ttype, grid, pts, T, sign):
  M = pts.shape[0]; N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  src = torch.view_as_complex(torch.rand((T, N, 2) if ttype == 2 else (T, M, 2), device="cuda") - 0.5)
  dst = torch.zeros((T, M) if ttype == 2 else (T, N), dtype=torch.complex64, device="cuda")
  plan = _lib.Plan(ttype, grid[::-1], sign, T, 1e-6, 0, device=0)
  st = torch.cuda.current_stream().cuda_stream
  plan.set_points_interleaved(M, dp.data_ptr(), st)
  for _ in range(2):
    if ttype == 2: plan.execute(dst.data_ptr(), src.data_ptr(), st)
    else: plan.execute(src.data_ptr(), dst.data_ptr(), st)
  torch.cuda.synchronize(); plan.close()
one(2, (256, 256, 256), H.stack_of_stars_points(125, 125, 256)[:200000], 2, -1)
one(1, (512, 512), H.spiral_points()[:200000], 32, 1)
