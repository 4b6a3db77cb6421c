"""Small end-to-end runs of every kernel variant, for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200.python.ops import nufft_ops
from tests import helpers as H
for rank, grid in ((2, (48, 40)), (3, (24, 20, 16)), (1, (64,))):
  pts = H.uniform_points(3000, rank, 5)
  pts[:4] = np.pi * np.array([[1] * rank, [-1] * rank, [0] * rank, [1] + [-1] * (rank - 1)], np.float32)
  for ttype in (1, 2):
    src = H.random_complex((8, 3000) if ttype == 1 else (8,) + grid, 6)
    outs = []
    for meth in (1, 2, 3, 4):
      out = nufft_ops._run_op(torch.from_numpy(src).cuda(), torch.from_numpy(pts).cuda(), grid, f"type_{ttype}",
                              "forward", 1e-6, None, "nufft",
                              engine_kwargs={"spread_method": meth, "interp_method": meth, "bin_dims": (16, 16, 4) if rank == 3 and meth == 3 and ttype == 1 else (0, 0, 0)})
      outs.append(out.cpu().numpy())
    for o in outs[1:]:
      assert np.linalg.norm(o - outs[0]) / np.linalg.norm(outs[0]) < 1e-6
  # complex128 generic kernels
  src = H.random_complex((2, 3000), 7, np.complex128)
  nufft_ops._run_op(torch.from_numpy(src).cuda(), torch.from_numpy(pts.astype(np.float64)).cuda(), grid, "type_1",
                    "backward", 1e-10, None, "nufft")
torch.cuda.synchronize()
print("sanitize_small: ok")
