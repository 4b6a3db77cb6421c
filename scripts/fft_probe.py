"""GPU probe: the plan's own pruned + fused FFT passes (fft_pruned.cuh, fft_mode 0) against cuFFT
(fft_mode 2: cuFFT + amplify / deconvolve kernels; 3D: the three-plan pruned cuFFT scheme) on the
BASELINE configs and on small / odd-mode shapes: stage times and agreement."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H


def run(name, ttype, grid, pts, T, sign=-1, reps=5, tol=1e-6, modes=(2, 0)):
  M = pts.shape[0]
  N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  gen = torch.Generator(device="cuda").manual_seed(1)
  src_shape = (T, N, 2) if ttype == 2 else (T, M, 2)
  src = torch.view_as_complex(torch.rand(src_shape, generator=gen, device="cuda") - 0.5)
  ref = None
  for fm in modes:
    dst = torch.zeros((T, M) if ttype == 2 else (T, N), dtype=torch.complex64, device="cuda")
    plan = _lib.Plan(ttype, grid[::-1], sign, T, tol, 0, device=0, profile=1, fft_mode=fm)
    st = torch.cuda.current_stream().cuda_stream
    best = None
    for r in range(reps):
      plan.set_points_interleaved(M, dp.data_ptr(), st)
      if ttype == 2: plan.execute(dst.data_ptr(), src.data_ptr(), st)
      else: plan.execute(src.data_ptr(), dst.data_ptr(), st)
      torch.cuda.synchronize()
      t = plan.timings()
      t["fft_plus_deconv_ms"] = t["fft_ms"] + t["deconv_ms"]
      if best is None or t["fft_plus_deconv_ms"] < best["fft_plus_deconv_ms"]: best = t
    out = dst.cpu().numpy()
    err = 0.0 if ref is None else H.rel_l2(out, ref)
    if ref is None: ref = out
    print(json.dumps({"case": name, "type": ttype, "fft_mode": fm, "T": T, "M": M, "grid": list(grid),
                      **{k: round(x, 4) for k, x in best.items()}, "rel_l2_vs_cufft": err, "finite": bool(np.isfinite(out).all())}), flush=True)
    plan.close()


if __name__ == "__main__":
  quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
  run("cfg4-sos-256^3-4M-T2", 2, (256, 256, 256), H.stack_of_stars_points(125, 125, 256), 2)
  run("cfg2-spiral-512^2-2M-T32", 1, (512, 512), H.spiral_points(), 32, sign=1)
  run("cfg3-uniform-128^3-8M", 1, (128, 128, 128), H.uniform_points(8000000, 3, 3), 1, sign=1)
  run("cfg1-radial-256^2-100k", 2, (256, 256), H.radial_points(), 1)
  run("cfg2-as-type2", 2, (512, 512), H.spiral_points(), 32)
  run("cfg4-as-type1", 1, (256, 256, 256), H.stack_of_stars_points(125, 125, 256), 2, sign=1)
  if not quick:
    q = H.uniform_points(50000, 3, 9)
    for ttype in (1, 2):
      for sign in (-1, 1):
        run("small-32^3", ttype, (32, 32, 32), q, 3, sign=sign, reps=1)
        run("mixed-64x32x128", ttype, (128, 32, 64), q, 2, sign=sign, reps=1)
        run("2d-32x64", ttype, (64, 32), q[:, :2].copy(), 5, sign=sign, reps=1)
        run("2d-512x32", ttype, (32, 512), q[:, :2].copy(), 2, sign=sign, reps=1)
    run("tol1e-3-64^3", 2, (64, 64, 64), q, 2, reps=1, tol=1e-3)
    run("batch40-64^2", 1, (64, 64), q[:, :2].copy(), 40, sign=1, reps=1)
