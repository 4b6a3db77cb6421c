"""GPU probe: sweep spreader (method 6) against the window-sorted one (method 4) on cfg2, plus
agreement on awkward shapes (odd grids, narrow kernels, fold-boundary points, coil counts that are
not multiples of the coil group)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H


def run(name, grid, pts, T, variants, reps=6, tol=1e-6):
  M = pts.shape[0]
  N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  c = torch.from_numpy(H.random_complex((T, M), 1)).cuda()
  ref = None
  for v in variants:
    f = torch.zeros((T, N), dtype=torch.complex64, device="cuda")
    kw = dict(spread_method=v["method"], profile=1)
    if "bins" in v: kw["bin_dims"] = v["bins"]
    for k in ("coils_per_cta", "max_subproblem_size", "no_pack", "no_tma_flush", "no_point_major"):
      if k in v: kw[k] = v[k]
    plan = _lib.Plan(1, grid[::-1], 1, T, tol, 0, device=0, **kw)
    st = torch.cuda.current_stream().cuda_stream
    best = None
    for r in range(reps):
      plan.set_points_interleaved(M, dp.data_ptr(), st)
      plan.execute(c.data_ptr(), f.data_ptr(), st)
      torch.cuda.synchronize()
      t = plan.timings()
      if best is None or t["spread_interp_ms"] < best["spread_interp_ms"]:
        best = t
    out = f.cpu().numpy()
    if ref is None:
      ref = out
      err = 0.0
    else:
      err = H.rel_l2(out, ref)
    inf = plan.info()
    print(json.dumps({"case": name, **v, "bins_used": list(inf.bin_dims)[:len(grid)], "T": T, "M": M,
                      **{k: round(x, 4) for k, x in best.items()}, "rel_l2_vs_first": err,
                      "finite": bool(np.isfinite(out).all())}), flush=True)
    plan.close()


if __name__ == "__main__":
  quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
  p = H.spiral_points(32, 62500)
  if len(sys.argv) > 1 and sys.argv[1] == "tune":
    V = [dict(method=6), dict(method=6, bins=(8, 8)), dict(method=6, bins=(8, 16)), dict(method=6, bins=(24, 8)),
         dict(method=6, bins=(8, 8), coils_per_cta=16), dict(method=6, bins=(8, 4)), dict(method=6, bins=(8, 8), coils_per_cta=4)]
    run("cfg2-spiral-512-T32", (512, 512), p, 32, V, reps=5)
    run("uniform-512-T32", (512, 512), H.uniform_points(2000000, 2, 7), 32, V[:5], reps=3)
    sys.exit(0)
  if len(sys.argv) > 1 and sys.argv[1] == "prof":   # one variant, for ncu
    run("cfg2-spiral-512-T32", (512, 512), p, 32, [dict(method=6)], reps=3)
    sys.exit(0)
  V = [dict(method=4), dict(method=6), dict(method=6, no_point_major=1), dict(method=6, coils_per_cta=16),
       dict(method=6, coils_per_cta=4), dict(method=6, no_pack=1), dict(method=6, coils_per_cta=16, no_point_major=1)]
  if not quick:
    V += [dict(method=6, bins=(16, 16)), dict(method=6, bins=(32, 8)), dict(method=6, coils_per_cta=16, bins=(32, 8)),
          dict(method=6, max_subproblem_size=256), dict(method=6, max_subproblem_size=4096)]
  run("cfg2-spiral-512-T32", (512, 512), p, 32, V)
  run("uniform-512-T32", (512, 512), H.uniform_points(2000000, 2, 7), 32, V[:4], reps=3)
  run("radial-256-T8", (256, 256), H.radial_points(200, 500), 8, V[:4], reps=3)
  # agreement on awkward shapes: odd grid, narrow kernels, points on the fold boundaries
  rng = np.random.default_rng(5)
  q = rng.uniform(-np.pi, np.pi, (50000, 2)).astype(np.float32)
  q[:64, 0] = np.float32(np.pi); q[64:128, 1] = -np.float32(np.pi); q[128:160] = 0
  for tol in (1e-6, 1e-4, 1e-3, 1e-2):
    for T in (4, 5, 1, 19):
      run(f"odd-130x94-tol{tol}", (94, 130), q, T, [dict(method=2), dict(method=4), dict(method=6), dict(method=6, coils_per_cta=16)],
          reps=1, tol=tol)
  run("ext-range", (64, 64), (q * 2.9).astype(np.float32), 8, [dict(method=2), dict(method=4), dict(method=6)], reps=1)
  run("tiny-grid", (16, 14), q[:3000], 8, [dict(method=2), dict(method=6)], reps=1)
  run("sparse", (512, 512), q[:2000], 8, [dict(method=2), dict(method=6)], reps=1)
