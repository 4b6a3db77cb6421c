"""GPU probe: 3D row-lane tile kernels (complex128, wide complex64, sigma = 1.25) against the
point-driven global kernels (method 1): time of the spread / interp stage and agreement."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H


def run(name, grid, M, T, cd, tol, ups=0, reps=3):
  rd = np.float32 if cd == np.complex64 else np.float64
  pts = H.uniform_points(M, 3, 5, rd)
  dp = torch.from_numpy(pts).cuda()
  N = int(np.prod(grid))
  tdt = torch.complex64 if cd == np.complex64 else torch.complex128
  code = _lib.COMPLEX64 if cd == np.complex64 else _lib.COMPLEX128
  for ttype in (1, 2):
    src = torch.from_numpy(H.random_complex((T, M) if ttype == 1 else (T, N), 1, cd)).cuda()
    ref = None
    for meth in (1, 0):
      out = torch.zeros((T, N) if ttype == 1 else (T, M), dtype=tdt, device="cuda")
      plan = _lib.Plan(ttype, grid[::-1], -1, T, float(np.float32(tol)), code, device=0, profile=1,
                       spread_method=meth, interp_method=meth, upsampling=ups)
      st = torch.cuda.current_stream().cuda_stream
      best = None
      for r in range(reps):
        plan.set_points_interleaved(M, dp.data_ptr(), st)
        if ttype == 1: plan.execute(src.data_ptr(), out.data_ptr(), st)
        else: plan.execute(out.data_ptr(), src.data_ptr(), st)
        torch.cuda.synchronize()
        t = plan.timings()
        if best is None or t["spread_interp_ms"] < best["spread_interp_ms"]: best = t
      o = out.cpu().numpy()
      err = 0.0 if ref is None else H.rel_l2(o, ref)
      if ref is None: ref = o
      inf = plan.info()
      print(json.dumps({"case": name, "type": ttype, "requested": meth, "spread_method": inf.spread_method, "interp_method": inf.interp_method,
                        "ns": inf.kernel_width, "sigma": inf.upsampling_factor, "bins": list(inf.bin_dims), "M": M, "T": T,
                        **{k: round(x, 4) for k, x in best.items()}, "rel_l2_vs_generic": err, "finite": bool(np.isfinite(o).all())}), flush=True)
      plan.close()


if __name__ == "__main__":
  run("c128-tol1e-12", (64, 64, 64), 500000, 1, np.complex128, 1e-12)
  run("c128-tol1e-6", (64, 64, 64), 1000000, 2, np.complex128, 1e-6)
  run("c128-tol1e-9", (48, 40, 36), 300000, 1, np.complex128, 1e-9)
  run("c64-tol1e-7", (96, 96, 96), 2000000, 1, np.complex64, 1e-7)
  run("c64-sigma1.25", (128, 128, 128), 2000000, 2, np.complex64, 1e-6, ups=1)
  run("c64-sigma1.25-small", (24, 30, 20), 20000, 1, np.complex64, 1e-4, ups=1)
