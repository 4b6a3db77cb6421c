"""GPU probe: 3D ring interpolator (method 7, z-slab streaming) against the quarter-warp tile
interpolator (method 3) on cfg4, cfg3's point set as type 2, 800k points; agreement on awkward shapes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tensorflow_nufft_b200 import _lib
from tests import helpers as H


def run(name, grid, pts, T, variants, reps=4, tol=1e-6):
  M = pts.shape[0]
  N = int(np.prod(grid))
  dp = torch.from_numpy(pts).cuda()
  gen = torch.Generator(device="cuda").manual_seed(1)
  f = torch.view_as_complex(torch.rand((T, N, 2), generator=gen, device="cuda") - 0.5)
  ref = None
  for v in variants:
    c = torch.zeros((T, M), dtype=torch.complex64, device="cuda")
    kw = dict(interp_method=v["method"], profile=1)
    if "bins" in v: kw["bin_dims"] = v["bins"]
    for k in ("max_subproblem_size", "no_tma"):
      if k in v: kw[k] = v[k]
    plan = _lib.Plan(2, grid[::-1], -1, T, tol, 0, device=0, **kw)
    st = torch.cuda.current_stream().cuda_stream
    best = None
    for r in range(reps):
      plan.set_points_interleaved(M, dp.data_ptr(), st)
      plan.execute(c.data_ptr(), f.data_ptr(), st)
      torch.cuda.synchronize()
      t = plan.timings()
      if best is None or t["spread_interp_ms"] < best["spread_interp_ms"]: best = t
    out = c.cpu().numpy()
    err = 0.0 if ref is None else H.rel_l2(out, ref)
    if ref is None: ref = out
    inf = plan.info()
    print(json.dumps({"case": name, **v, "method_used": inf.interp_method, "bins_used": list(inf.bin_dims)[:3], "T": T, "M": M,
                      **{k: round(x, 4) for k, x in best.items()}, "rel_l2_vs_first": err, "finite": bool(np.isfinite(out).all())}), flush=True)
    plan.close()


if __name__ == "__main__":
  V = [dict(method=3), dict(method=7), dict(method=7, bins=(16, 8, 8)), dict(method=7, bins=(16, 8, 32)), dict(method=7, bins=(16, 16, 16)),
       dict(method=7, max_subproblem_size=256), dict(method=7, no_tma=1)]
  run("cfg4-sos-256^3-4M-T2", (256, 256, 256), H.stack_of_stars_points(125, 125, 256), 2, V)
  run("cfg3pts-uniform-128^3-8M", (128, 128, 128), H.uniform_points(8000000, 3, 3), 1, V[:5])
  run("ref7-uniform-128^3-800k", (128, 128, 128), H.uniform_points(800000, 3, 17), 1, V[:4], reps=3)
  rng = np.random.default_rng(5)
  q = rng.uniform(-np.pi, np.pi, (60000, 3)).astype(np.float32)
  q[:64, 0] = np.float32(np.pi); q[64:128, 1] = -np.float32(np.pi); q[128:192, 2] = np.float32(np.pi); q[192:224] = 0
  for tol in (1e-6, 1e-4, 1e-3, 1e-2):
    for T in (1, 3):
      run(f"odd-34x26x30-tol{tol}", (30, 26, 34), q, T, [dict(method=1), dict(method=3), dict(method=7), dict(method=7, bins=(16, 8, 4)), dict(method=7, no_tma=1)], reps=1, tol=tol)
  run("ext-range", (32, 32, 32), (q * 2.9).astype(np.float32), 2, [dict(method=1), dict(method=7)], reps=1)
  run("tiny-grid", (8, 10, 12), q[:3000], 2, [dict(method=1), dict(method=7)], reps=1)
  run("sparse", (128, 128, 128), q[:2000], 2, [dict(method=1), dict(method=7)], reps=1)
