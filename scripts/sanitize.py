"""Small transforms through every default kernel family (sweep spreaders 2D / 3D, quarter-warp
interpolators, row-lane kernels 2D / 3D, generic 1D, set_points incl. the fingerprint, the own FFT
passes of every length 64 .. 1024 on both kinds of axis) for
compute-sanitizer (memcheck / racecheck / initcheck run this script)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tensorflow_nufft_b200 as tfft
from tests import helpers as H

tfft.set_points_reuse(len(sys.argv) > 1 and sys.argv[1] == "reuse")
cases = [((40, 36), 6000, 8, np.complex64, 1e-6), ((40, 36), 3000, 5, np.complex64, 1e-4), ((20, 24, 28), 8000, 1, np.complex64, 1e-6),
         ((20, 24, 28), 4000, 2, np.complex64, 1e-3), ((24, 20), 2000, 2, np.complex128, 1e-12), ((12, 16, 10), 1500, 1, np.complex128, 1e-9),
         ((16, 12, 14), 1500, 1, np.complex64, 1e-7), ((64,), 800, 2, np.complex64, 1e-6),
         # power-of-two fine grids: the own pruned FFT passes (rows: last axis, strided: the others)
         ((32, 64), 2000, 3, np.complex64, 1e-6), ((128, 32), 2000, 1, np.complex64, 1e-6), ((32, 512), 1000, 1, np.complex64, 1e-6),
         ((512, 32), 1000, 2, np.complex64, 1e-6), ((256, 256), 3000, 1, np.complex64, 1e-6),
         ((32, 32, 32), 3000, 2, np.complex64, 1e-6), ((64, 32, 128), 3000, 1, np.complex64, 1e-6)]
for grid, M, T, cd, tol in cases:
  rank = len(grid)
  rd = np.float32 if cd == np.complex64 else np.float64
  pts = torch.from_numpy(H.uniform_points(M, rank, 3, rd)).cuda()
  for tt in (1, 2):
    src = torch.from_numpy(H.random_complex((T, M) if tt == 1 else (T,) + grid, 4, cd)).cuda()
    for _ in range(2):
      out = tfft.nufft(src, pts, grid_shape=grid, transform_type=f"type_{tt}", fft_direction="forward", tol=tol)
    torch.cuda.synchronize()
    assert torch.isfinite(torch.view_as_real(out)).all()
# dense 3D set (narrow sweep bins) and the opt-in ring interpolator
from tensorflow_nufft_b200.python.ops import nufft_ops
pd = torch.from_numpy(H.uniform_points(30000, 3, 9)).cuda()
cd_ = torch.from_numpy(H.random_complex((1, 30000), 10)).cuda()
tfft.nufft(cd_, pd, grid_shape=(16, 16, 16), transform_type="type_1")
gd = torch.from_numpy(H.random_complex((2, 16, 20, 24), 11)).cuda()
nufft_ops._run_op(gd, pd, (16, 20, 24), "type_2", "forward", 1e-6, None, "nufft", engine_kwargs={"interp_method": 7})
torch.cuda.synchronize()
f = torch.ones((2, 32, 48), dtype=torch.complex64, device="cuda")
p2 = torch.from_numpy(H.uniform_points(3000, 2, 5) * np.float32(0.99)).cuda()
tfft.interp(f, p2)
tfft.spread(torch.ones((2, 3000), dtype=torch.complex64, device="cuda"), p2, (32, 48))
torch.cuda.synchronize()
print("sanitize script done")
