"""Scratch: run one plan configuration a few times with the engine's default kernels (for ncu
captures). Usage: prof_one.py cfg{1,2,3,4} <transform type 1|2> [coils]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H
cfg, ttype = sys.argv[1], int(sys.argv[2])
if cfg == "cfg2":
  grid, pts, T = (512, 512), H.spiral_points(32, 62500), 32
elif cfg == "cfg3":
  grid, pts, T = (128, 128, 128), H.uniform_points(8000000, 3, 3), 1
elif cfg == "cfg4":
  grid, pts, T = (256, 256, 256), H.stack_of_stars_points(125, 125, 256), 2
else:
  grid, pts, T = (256, 256), H.radial_points(200, 500), 1
if len(sys.argv) > 3: T = int(sys.argv[3])
M = pts.shape[0]; N = int(np.prod(grid))
plan = _lib.Plan(ttype, grid[::-1], -1, T, 1e-6, 0, device=0)
dp = torch.from_numpy(pts).cuda()
c = torch.from_numpy(H.random_complex((T, M), 1)).cuda()
f = torch.from_numpy(H.random_complex((T, N), 2)).cuda()
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
  plan.set_points_interleaved(M, dp.data_ptr(), st)
  plan.execute(c.data_ptr(), f.data_ptr(), st)
torch.cuda.synchronize()
