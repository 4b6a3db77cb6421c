"""GPU probe: complex128 path (cfg5-class: 2D 256x256, tol 1e-12 -> ns 14, 100k radial points)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H

def run(name, ttype, grid, pts, T, tol, reps=5, **kw):
  M = pts.shape[0]; N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  c = torch.from_numpy(H.random_complex((T, M), 1, np.complex128)).cuda()
  f = torch.from_numpy(H.random_complex((T, N), 2, np.complex128)).cuda()
  plan = _lib.Plan(ttype, grid[::-1], -1, T, tol, 1, device=0, profile=1, **kw)
  st = torch.cuda.current_stream().cuda_stream
  best = None
  for r in range(reps):
    plan.set_points_interleaved(M, dp.data_ptr(), st)
    plan.execute(c.data_ptr(), f.data_ptr(), st)
    torch.cuda.synchronize()
    t = plan.timings()
    if best is None or t["spread_interp_ms"] < best["spread_interp_ms"]: best = t
  inf = plan.info()
  print(json.dumps({"case": name, "type": ttype, "T": T, "M": M, "ns": inf.kernel_width, **{k: round(v, 4) for k, v in best.items()}}), flush=True)
  plan.close()

if __name__ == "__main__":
  p = H.radial_points(200, 500, np.float64)
  tol = float(np.float32(1e-12))
  run("cfg5-fwd", 2, (256, 256), p, 1, tol)
  run("cfg5-adj", 1, (256, 256), p, 1, tol)
  run("cfg5-dpts", 2, (256, 256), p, 2, tol)
  p2 = H.spiral_points(32, 62500, dtype=np.float64)
  run("cfg2-f64-type1-T8", 1, (512, 512), p2, 8, tol)
  run("cfg2-f64-type2-T8", 2, (512, 512), p2, 8, tol)
  run("cfg2-f64-type1-T8-tol1e-6", 1, (512, 512), p2, 8, 1e-6)
  run("cfg2-f64-type2-T8-tol1e-6", 2, (512, 512), p2, 8, 1e-6)
